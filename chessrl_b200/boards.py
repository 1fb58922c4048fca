"""Host-side data formats for the device board record and move word (format conversion only -- no chess
rules live here; rules run in the CUDA kernels behind libchessrl_b200.so).

Board record = 9 x uint64 (72 B, include/chessrl_b200.h):
  [0..5] pawns, knights, bishops, rooks, queens, kings   [6] white occupancy   [7] black occupancy
  [8]    meta: bit0 turn(1=white) | bits1-4 castling K,Q,k,q | bits5-11 ep+1 | bits12-23 halfmove |
               bits24-37 fullmove | bits38-51 ply (len(move_stack)) | bits52-59 reversible-run length
Move word  = from | to<<6 | promo<<12   (promo 0 = none, 1 N, 2 B, 3 R, 4 Q);  0xFFFF = none.
UCI strings follow python-chess Move.uci() as used by game.py:49 (queen promotions carry the 'q').
"""

from __future__ import annotations

import numpy as np

STARTING_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
NULL_MOVE = "00000"          # game.py:13
MOVE_NONE = 0xFFFF
RECORD_WORDS = 9
MAX_MOVES = 256
RESULT_NONE = 2

_PIECE_INDEX = {"p": 0, "n": 1, "b": 2, "r": 3, "q": 4, "k": 5}
_PIECE_CHARS = "pnbrqk"
_PROMO_CHARS = {1: "n", 2: "b", 3: "r", 4: "q"}
_PROMO_CODES = {v: k for k, v in _PROMO_CHARS.items()}
_FILES = "abcdefgh"
_MASK64 = (1 << 64) - 1


def pack_meta(turn, castle, ep, half, full, ply=0, rev=0):
    half = min(int(half), 4095)
    full = min(int(full), 16383)
    ply = min(int(ply), 16383)
    rev = min(int(rev), 255)
    return (int(bool(turn)) | (castle & 15) << 1 | ((ep + 1) & 127) << 5 | half << 12 | full << 24
            | ply << 38 | rev << 52)


def meta_fields(meta):
    meta = int(meta)
    return {"turn": bool(meta & 1), "castle": (meta >> 1) & 15, "ep": ((meta >> 5) & 127) - 1,
            "halfmove": (meta >> 12) & 4095, "fullmove": (meta >> 24) & 16383,
            "ply": (meta >> 38) & 16383, "rev": (meta >> 52) & 255}


def board_fen_to_bitboards(board_fen):
    """Piece-placement field -> 8 bitboards (python ints)."""
    bbs = [0] * 8
    rows = board_fen.strip().split("/")
    if len(rows) != 8:
        raise ValueError("expected 8 rows in position part of fen: %r" % (board_fen,))
    for ri, row in enumerate(rows):
        f = 0
        for ch in row:
            if ch.isdigit():
                f += int(ch)
                continue
            k = _PIECE_INDEX.get(ch.lower())
            if k is None or f > 7:
                raise ValueError("invalid fen row %r" % (row,))
            b = 1 << ((7 - ri) * 8 + f)
            bbs[k] |= b
            bbs[6 if ch.isupper() else 7] |= b
            f += 1
        if f != 8:
            raise ValueError("expected 8 columns per row in position part of fen: %r" % (board_fen,))
    return bbs


def record_from_fen(fen=STARTING_FEN):
    """Full FEN -> uint64[9].  Castling rights are cleaned like Board.clean_castling_rights()."""
    parts = fen.split()
    bbs = board_fen_to_bitboards(parts[0])
    turn = len(parts) < 2 or parts[1] == "w"
    cf = parts[2] if len(parts) > 2 else "-"
    castle = 0
    rooks, kings, w, b = bbs[3], bbs[5], bbs[6], bbs[7]
    if "K" in cf and rooks & w & (1 << 7) and kings & w & (1 << 4):
        castle |= 1
    if "Q" in cf and rooks & w & (1 << 0) and kings & w & (1 << 4):
        castle |= 2
    if "k" in cf and rooks & b & (1 << 63) and kings & b & (1 << 60):
        castle |= 4
    if "q" in cf and rooks & b & (1 << 56) and kings & b & (1 << 60):
        castle |= 8
    ep = -1
    if len(parts) > 3 and parts[3] != "-":
        ep = _FILES.index(parts[3][0]) + 8 * (int(parts[3][1]) - 1)
    half = int(parts[4]) if len(parts) > 4 else 0
    full = max(int(parts[5]), 1) if len(parts) > 5 else 1
    rec = np.zeros(RECORD_WORDS, dtype=np.uint64)
    for k in range(8):
        rec[k] = bbs[k]
    rec[8] = pack_meta(turn, castle, ep, half, full, 0, 0)
    return rec


def board_fen_from_record(rec):
    bbs = [int(x) for x in rec[:8]]
    out = []
    for r in range(7, -1, -1):
        empty = 0
        for f in range(8):
            b = 1 << (r * 8 + f)
            ch = None
            for k in range(6):
                if bbs[k] & b:
                    ch = _PIECE_CHARS[k]
            if ch is None:
                empty += 1
                continue
            if empty:
                out.append(str(empty))
                empty = 0
            out.append(ch.upper() if bbs[6] & b else ch)
        if empty:
            out.append(str(empty))
        if r:
            out.append("/")
    return "".join(out)


def fen_from_record(rec, ep_legal=True):
    m = meta_fields(rec[8])
    c = "".join(ch for ch, bit in (("K", 1), ("Q", 2), ("k", 4), ("q", 8)) if m["castle"] & bit) or "-"
    ep = "-"
    if m["ep"] >= 0 and ep_legal:
        ep = _FILES[m["ep"] & 7] + str((m["ep"] >> 3) + 1)
    return "%s %s %s %s %d %d" % (board_fen_from_record(rec), "w" if m["turn"] else "b", c, ep,
                                   m["halfmove"], m["fullmove"])


def move_to_uci(mv):
    mv = int(mv)
    if mv == MOVE_NONE:
        return NULL_MOVE
    f, t, p = mv & 63, (mv >> 6) & 63, (mv >> 12) & 7
    s = _FILES[f & 7] + str((f >> 3) + 1) + _FILES[t & 7] + str((t >> 3) + 1)
    return s + _PROMO_CHARS[p] if p else s


def uci_to_move(uci):
    """UCI string -> move word, or MOVE_NONE for anything that is not a well-formed move."""
    if not isinstance(uci, str) or len(uci) not in (4, 5):
        return MOVE_NONE
    try:
        f = _FILES.index(uci[0]) + 8 * (int(uci[1]) - 1)
        t = _FILES.index(uci[2]) + 8 * (int(uci[3]) - 1)
        p = _PROMO_CODES[uci[4]] if len(uci) == 5 else 0
    except (ValueError, KeyError):
        return MOVE_NONE
    if not (0 <= f < 64 and 0 <= t < 64) or f == t:
        return MOVE_NONE
    return f | t << 6 | p << 12

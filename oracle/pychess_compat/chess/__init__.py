"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A clean-room CPU restatement of the slice of **python-chess 0.28.3**
(`requirements.txt:9` of the reference pins it; its source is NOT under
/root/reference and the wheel is not installable here) that the reference's
hot path touches:

  * game.py:1-2,19,39,49,60,69,72,77,80,83,99-102   Board(), push, legal_moves,
    move_stack, board_fen, set_board_fen, turn, copy, reset,
    can_claim_fifty_moves, is_game_over, result
  * netencoder.py:8,25-27,58,63                    PIECE_TYPES, pieces().mirror().tolist(),
    copy(), pop() raising IndexError on an empty stack
  * mctree.py:186-188,308                          move_stack[-2], str(move)

The module is importable as ``chess`` (put ``oracle/pychess_compat`` on
``sys.path``) so that the reference's own ``game.py`` / ``netencoder.py`` /
``mctree.py`` run UNMODIFIED on top of it ("reference-on-shims").

What is restated is the library's *published algorithm*: bitboard pseudo-legal
generation in the documented order (non-pawn pieces scanned from the most
significant square down, then castling, pawn captures, single pushes, double
pushes, en passant), evasion generation when in check, the slider-blocker pin
test, the en-passant skewer test, and the draw rules.  Parity pins: the public
perft table (start, Kiwipete, positions 3-6) and the start-position move order
printed in python-chess's README; see tests/test_oracle_chess.py.

Version-sensitive corners are switches at the top of the file
and default to the 0.28.3 behaviour.
"""

from __future__ import annotations

# --------------------------------------------------------------------------
# version-sensitive switches (defaults = python-chess 0.28.3)
# --------------------------------------------------------------------------
FIFTY_MOVE_CLAIM_LOOKAHEAD = False   # >=1.0 also claims at clock 99 if a quiet move exists
REPETITION_STOPS_ON_LEGAL_EP = False  # >=1.0 treats legal-ep positions as irreversible

WHITE, BLACK = True, False
COLORS = [WHITE, BLACK]
PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = range(1, 7)
PIECE_TYPES = range(1, 7)
PIECE_SYMBOLS = [None, "p", "n", "b", "r", "q", "k"]
FILE_NAMES = "abcdefgh"
RANK_NAMES = "12345678"
STARTING_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
STARTING_BOARD_FEN = STARTING_FEN.split()[0]

SQUARES = list(range(64))
SQUARE_NAMES = [f + r for r in RANK_NAMES for f in FILE_NAMES]
(A1, B1, C1, D1, E1, F1, G1, H1) = range(0, 8)
(A8, B8, C8, D8, E8, F8, G8, H8) = range(56, 64)

BB_ALL = 0xFFFF_FFFF_FFFF_FFFF
BB_EMPTY = 0
BB_SQUARES = [1 << s for s in SQUARES]
BB_FILES = [0x0101_0101_0101_0101 << f for f in range(8)]
BB_RANKS = [0xFF << (8 * r) for r in range(8)]
BB_FILE_A, BB_FILE_C, BB_FILE_D, BB_FILE_F, BB_FILE_G, BB_FILE_H = (
    BB_FILES[0], BB_FILES[2], BB_FILES[3], BB_FILES[5], BB_FILES[6], BB_FILES[7])
BB_RANK_1, BB_RANK_3, BB_RANK_4, BB_RANK_5, BB_RANK_6, BB_RANK_8 = (
    BB_RANKS[0], BB_RANKS[2], BB_RANKS[3], BB_RANKS[4], BB_RANKS[5], BB_RANKS[7])
BB_LIGHT_SQUARES = 0x55AA_55AA_55AA_55AA
BB_DARK_SQUARES = 0xAA55_AA55_AA55_AA55
BB_A1, BB_H1, BB_A8, BB_H8 = BB_SQUARES[A1], BB_SQUARES[H1], BB_SQUARES[A8], BB_SQUARES[H8]
BB_E1, BB_E8 = BB_SQUARES[E1], BB_SQUARES[E8]


def square_file(sq):
    return sq & 7


def square_rank(sq):
    return sq >> 3


def square(file_index, rank_index):
    return rank_index * 8 + file_index


def square_distance(a, b):
    return max(abs(square_file(a) - square_file(b)), abs(square_rank(a) - square_rank(b)))


def msb(bb):
    return bb.bit_length() - 1


def lsb(bb):
    return (bb & -bb).bit_length() - 1


def popcount(bb):
    return bin(bb).count("1")


def scan_reversed(bb):
    """Squares of a bitboard, most significant first (h8 -> a1)."""
    while bb:
        s = bb.bit_length() - 1
        yield s
        bb ^= 1 << s


def scan_forward(bb):
    while bb:
        low = bb & -bb
        yield low.bit_length() - 1
        bb ^= low


def flip_vertical(bb):
    out = 0
    for r in range(8):
        out |= ((bb >> (8 * r)) & 0xFF) << (8 * (7 - r))
    return out


# --------------------------------------------------------------------------
# attack tables, built by walking the board (no magics needed on the host)
# --------------------------------------------------------------------------

def _step_attacks(sq, deltas):
    out = 0
    for d in deltas:
        t = sq + d
        if 0 <= t < 64 and square_distance(sq, t) <= 2:
            out |= 1 << t
    return out


BB_KNIGHT_ATTACKS = [_step_attacks(s, (17, 15, 10, 6, -17, -15, -10, -6)) for s in SQUARES]
BB_KING_ATTACKS = [_step_attacks(s, (9, 8, 7, 1, -9, -8, -7, -1)) for s in SQUARES]
BB_PAWN_ATTACKS = {
    BLACK: [_step_attacks(s, (-7, -9)) for s in SQUARES],
    WHITE: [_step_attacks(s, (7, 9)) for s in SQUARES],
}

_ROOK_DIRS = ((1, 0), (-1, 0), (0, 1), (0, -1))
_BISHOP_DIRS = ((1, 1), (1, -1), (-1, 1), (-1, -1))


def _walk(sq, dirs, occupied):
    """Sliding attack set of `sq` along `dirs`, stopping at (and including) blockers."""
    f0, r0 = square_file(sq), square_rank(sq)
    out = 0
    for df, dr in dirs:
        f, r = f0 + df, r0 + dr
        while 0 <= f < 8 and 0 <= r < 8:
            b = 1 << (r * 8 + f)
            out |= b
            if occupied & b:
                break
            f, r = f + df, r + dr
    return out


def rook_attacks(sq, occupied):
    return _walk(sq, _ROOK_DIRS, occupied)


def bishop_attacks(sq, occupied):
    return _walk(sq, _BISHOP_DIRS, occupied)


def rank_attacks(sq, occupied):
    return _walk(sq, ((1, 0), (-1, 0)), occupied)


def file_attacks(sq, occupied):
    return _walk(sq, ((0, 1), (0, -1)), occupied)


def diag_attacks(sq, occupied):
    return _walk(sq, _BISHOP_DIRS, occupied)


# empty-board reach, full lines through two squares, and the open segment between them
_RANK_REACH = [rank_attacks(s, 0) for s in SQUARES]
_FILE_REACH = [file_attacks(s, 0) for s in SQUARES]
_DIAG_REACH = [diag_attacks(s, 0) for s in SQUARES]


def _build_rays():
    rays = [[0] * 64 for _ in SQUARES]
    between = [[0] * 64 for _ in SQUARES]
    for a in SQUARES:
        for b in SQUARES:
            if a == b:
                continue
            bb_b = 1 << b
            if _DIAG_REACH[a] & bb_b:
                rays[a][b] = (_DIAG_REACH[a] & _DIAG_REACH[b]) | (1 << a) | bb_b
                between[a][b] = diag_attacks(a, bb_b) & diag_attacks(b, 1 << a)
            elif _RANK_REACH[a] & bb_b:
                rays[a][b] = _RANK_REACH[a] | (1 << a)
                between[a][b] = rank_attacks(a, bb_b) & rank_attacks(b, 1 << a)
            elif _FILE_REACH[a] & bb_b:
                rays[a][b] = _FILE_REACH[a] | (1 << a)
                between[a][b] = file_attacks(a, bb_b) & file_attacks(b, 1 << a)
    return rays, between


BB_RAYS, BB_BETWEEN = _build_rays()


# --------------------------------------------------------------------------
# Move / SquareSet
# --------------------------------------------------------------------------

class Move:
    __slots__ = ("from_square", "to_square", "promotion", "drop")

    def __init__(self, from_square, to_square, promotion=None, drop=None):
        self.from_square = from_square
        self.to_square = to_square
        self.promotion = promotion
        self.drop = drop

    def uci(self):
        if self.promotion:
            return SQUARE_NAMES[self.from_square] + SQUARE_NAMES[self.to_square] + PIECE_SYMBOLS[self.promotion]
        if self:
            return SQUARE_NAMES[self.from_square] + SQUARE_NAMES[self.to_square]
        return "0000"

    def __bool__(self):
        return bool(self.from_square or self.to_square or self.promotion)

    def __eq__(self, other):
        return (isinstance(other, Move) and self.from_square == other.from_square
                and self.to_square == other.to_square and self.promotion == other.promotion)

    def __hash__(self):
        return hash((self.from_square, self.to_square, self.promotion))

    def __str__(self):
        return self.uci()

    def __repr__(self):
        return "Move.from_uci({!r})".format(self.uci())

    @classmethod
    def from_uci(cls, uci):
        if uci == "0000":
            return cls.null()
        if len(uci) == 4 or len(uci) == 5:
            try:
                f = SQUARE_NAMES.index(uci[0:2])
                t = SQUARE_NAMES.index(uci[2:4])
                promo = PIECE_SYMBOLS.index(uci[4]) if len(uci) == 5 else None
            except ValueError:
                raise ValueError("invalid uci: {!r}".format(uci))
            if f == t:
                raise ValueError("invalid uci (use 0000 for null moves): {!r}".format(uci))
            return cls(f, t, promotion=promo)
        raise ValueError("expected uci string to be of length 4 or 5: {!r}".format(uci))

    @classmethod
    def null(cls):
        return cls(0, 0)


class SquareSet:
    def __init__(self, mask=0):
        self.mask = int(mask) & BB_ALL

    def mirror(self):
        """Vertical flip (rank 1 <-> rank 8), as python-chess SquareSet.mirror()."""
        return SquareSet(flip_vertical(self.mask))

    def tolist(self):
        return [bool(self.mask >> s & 1) for s in SQUARES]

    def __iter__(self):
        return scan_forward(self.mask)

    def __len__(self):
        return popcount(self.mask)

    def __int__(self):
        return self.mask

    def __bool__(self):
        return bool(self.mask)

    def __contains__(self, sq):
        return bool(self.mask >> sq & 1)


class LegalMoveGenerator:
    def __init__(self, board):
        self.board = board

    def __iter__(self):
        return self.board.generate_legal_moves()

    def __bool__(self):
        return any(self.board.generate_legal_moves())

    def count(self):
        return sum(1 for _ in self)

    def __len__(self):
        return self.count()

    def __contains__(self, move):
        return any(m == move for m in self)


class _Snapshot:
    __slots__ = ("pawns", "knights", "bishops", "rooks", "queens", "kings", "occ_w", "occ_b",
                 "occupied", "promoted", "turn", "castling_rights", "ep_square",
                 "halfmove_clock", "fullmove_number")

    def __init__(self, b):
        self.pawns, self.knights, self.bishops = b.pawns, b.knights, b.bishops
        self.rooks, self.queens, self.kings = b.rooks, b.queens, b.kings
        self.occ_w, self.occ_b, self.occupied = b.occupied_co[WHITE], b.occupied_co[BLACK], b.occupied
        self.promoted = b.promoted
        self.turn, self.castling_rights, self.ep_square = b.turn, b.castling_rights, b.ep_square
        self.halfmove_clock, self.fullmove_number = b.halfmove_clock, b.fullmove_number

    def restore(self, b):
        b.pawns, b.knights, b.bishops = self.pawns, self.knights, self.bishops
        b.rooks, b.queens, b.kings = self.rooks, self.queens, self.kings
        b.occupied_co = {WHITE: self.occ_w, BLACK: self.occ_b}
        b.occupied, b.promoted = self.occupied, self.promoted
        b.turn, b.castling_rights, b.ep_square = self.turn, self.castling_rights, self.ep_square
        b.halfmove_clock, b.fullmove_number = self.halfmove_clock, self.fullmove_number


# --------------------------------------------------------------------------
# Board
# --------------------------------------------------------------------------

class Board:
    chess960 = False

    def __init__(self, fen=STARTING_FEN):
        self.move_stack = []
        self._stack = []
        self.occupied_co = {WHITE: 0, BLACK: 0}
        if fen is None:
            self.clear()
        elif fen == STARTING_FEN:
            self.reset()
        else:
            self.set_fen(fen)

    # ---- setup -----------------------------------------------------------
    def _clear_pieces(self):
        self.pawns = self.knights = self.bishops = self.rooks = self.queens = self.kings = 0
        self.promoted = 0
        self.occupied_co = {WHITE: 0, BLACK: 0}
        self.occupied = 0

    def clear(self):
        self.turn = WHITE
        self.castling_rights = 0
        self.ep_square = None
        self.halfmove_clock = 0
        self.fullmove_number = 1
        self._clear_pieces()
        self.clear_stack()

    def clear_stack(self):
        del self.move_stack[:]
        del self._stack[:]

    def reset(self):
        self.turn = WHITE
        self.castling_rights = BB_A1 | BB_H1 | BB_A8 | BB_H8
        self.ep_square = None
        self.halfmove_clock = 0
        self.fullmove_number = 1
        self.pawns = BB_RANKS[1] | BB_RANKS[6]
        self.knights = BB_SQUARES[B1] | BB_SQUARES[G1] | BB_SQUARES[B8] | BB_SQUARES[G8]
        self.bishops = BB_SQUARES[C1] | BB_SQUARES[F1] | BB_SQUARES[C8] | BB_SQUARES[F8]
        self.rooks = BB_A1 | BB_H1 | BB_A8 | BB_H8
        self.queens = BB_SQUARES[D1] | BB_SQUARES[D8]
        self.kings = BB_E1 | BB_E8
        self.promoted = 0
        self.occupied_co = {WHITE: BB_RANK_1 | BB_RANKS[1], BLACK: BB_RANKS[6] | BB_RANK_8}
        self.occupied = self.occupied_co[WHITE] | self.occupied_co[BLACK]
        self.clear_stack()

    def set_board_fen(self, fen):
        """Piece placement only; leaves turn/castling/ep/clocks/stack untouched (game.py:71-72)."""
        rows = fen.strip().split("/")
        if len(rows) != 8:
            raise ValueError("expected 8 rows in position part of fen: {!r}".format(fen))
        self._clear_pieces()
        for ri, row in enumerate(rows):
            f = 0
            for ch in row:
                if ch.isdigit():
                    f += int(ch)
                elif ch == "~":
                    continue
                else:
                    pt = PIECE_SYMBOLS.index(ch.lower())
                    self._set_piece_at((7 - ri) * 8 + f, pt, ch.isupper())
                    f += 1
            if f != 8:
                raise ValueError("expected 8 columns per row in position part of fen: {!r}".format(fen))

    def set_fen(self, fen):
        parts = fen.split()
        board_part = parts[0]
        turn = WHITE if len(parts) < 2 or parts[1] == "w" else BLACK
        castling = parts[2] if len(parts) > 2 else "-"
        ep = parts[3] if len(parts) > 3 else "-"
        half = int(parts[4]) if len(parts) > 4 else 0
        full = max(int(parts[5]), 1) if len(parts) > 5 else 1
        self.set_board_fen(board_part)
        self.turn = turn
        cr = 0
        for ch in castling:
            if ch == "K":
                cr |= BB_H1
            elif ch == "Q":
                cr |= BB_A1
            elif ch == "k":
                cr |= BB_H8
            elif ch == "q":
                cr |= BB_A8
        self.castling_rights = cr
        self.ep_square = None if ep == "-" else SQUARE_NAMES.index(ep)
        self.halfmove_clock = half
        self.fullmove_number = full
        self.clear_stack()

    def board_fen(self):
        out = []
        for r in range(7, -1, -1):
            empty = 0
            for f in range(8):
                sq = r * 8 + f
                pt = self.piece_type_at(sq)
                if pt is None:
                    empty += 1
                    continue
                if empty:
                    out.append(str(empty))
                    empty = 0
                sym = PIECE_SYMBOLS[pt]
                out.append(sym.upper() if self.occupied_co[WHITE] >> sq & 1 else sym)
            if empty:
                out.append(str(empty))
            if r:
                out.append("/")
        return "".join(out)

    def fen(self):
        cr = self.clean_castling_rights()
        c = "".join(ch for ch, bb in (("K", BB_H1), ("Q", BB_A1), ("k", BB_H8), ("q", BB_A8)) if cr & bb) or "-"
        ep = SQUARE_NAMES[self.ep_square] if self.has_legal_en_passant() else "-"
        return "{} {} {} {} {} {}".format(self.board_fen(), "w" if self.turn else "b", c, ep,
                                          self.halfmove_clock, self.fullmove_number)

    # ---- piece access ----------------------------------------------------
    def pieces_mask(self, piece_type, color):
        bb = (None, self.pawns, self.knights, self.bishops, self.rooks, self.queens, self.kings)[piece_type]
        return bb & self.occupied_co[color]

    def pieces(self, piece_type, color):
        return SquareSet(self.pieces_mask(piece_type, color))

    def piece_type_at(self, sq):
        m = 1 << sq
        if not self.occupied & m:
            return None
        if self.pawns & m:
            return PAWN
        if self.knights & m:
            return KNIGHT
        if self.bishops & m:
            return BISHOP
        if self.rooks & m:
            return ROOK
        if self.queens & m:
            return QUEEN
        return KING

    def king(self, color):
        k = self.occupied_co[color] & self.kings & ~self.promoted
        return msb(k) if k else None

    def _remove_piece_at(self, sq):
        pt = self.piece_type_at(sq)
        if pt is None:
            return None
        keep = ~(1 << sq)
        self.pawns &= keep
        self.knights &= keep
        self.bishops &= keep
        self.rooks &= keep
        self.queens &= keep
        self.kings &= keep
        self.occupied &= keep
        self.occupied_co[WHITE] &= keep
        self.occupied_co[BLACK] &= keep
        self.promoted &= keep
        return pt

    def _set_piece_at(self, sq, piece_type, color, promoted=False):
        self._remove_piece_at(sq)
        m = 1 << sq
        if piece_type == PAWN:
            self.pawns |= m
        elif piece_type == KNIGHT:
            self.knights |= m
        elif piece_type == BISHOP:
            self.bishops |= m
        elif piece_type == ROOK:
            self.rooks |= m
        elif piece_type == QUEEN:
            self.queens |= m
        else:
            self.kings |= m
        self.occupied |= m
        self.occupied_co[color] |= m
        if promoted:
            self.promoted |= m

    # ---- attacks ---------------------------------------------------------
    def attacks_mask(self, sq):
        m = 1 << sq
        if m & self.pawns:
            return BB_PAWN_ATTACKS[bool(m & self.occupied_co[WHITE])][sq]
        if m & self.knights:
            return BB_KNIGHT_ATTACKS[sq]
        if m & self.kings:
            return BB_KING_ATTACKS[sq]
        out = 0
        if m & (self.bishops | self.queens):
            out = bishop_attacks(sq, self.occupied)
        if m & (self.rooks | self.queens):
            out |= rook_attacks(sq, self.occupied)
        return out

    def _attackers_mask(self, color, sq, occupied):
        rq = self.rooks | self.queens
        bq = self.bishops | self.queens
        att = ((BB_KING_ATTACKS[sq] & self.kings) |
               (BB_KNIGHT_ATTACKS[sq] & self.knights) |
               (rook_attacks(sq, occupied) & rq) |
               (bishop_attacks(sq, occupied) & bq) |
               (BB_PAWN_ATTACKS[not color][sq] & self.pawns))
        return att & self.occupied_co[color]

    def attackers_mask(self, color, sq):
        return self._attackers_mask(color, sq, self.occupied)

    def is_attacked_by(self, color, sq):
        return bool(self.attackers_mask(color, sq))

    def _attacked_for_king(self, path, occupied):
        return any(self._attackers_mask(not self.turn, s, occupied) for s in scan_reversed(path))

    def checkers_mask(self):
        k = self.king(self.turn)
        return 0 if k is None else self.attackers_mask(not self.turn, k)

    def is_check(self):
        return bool(self.checkers_mask())

    def pin_mask(self, color, sq):
        k = self.king(color)
        if k is None:
            return BB_ALL
        sq_mask = 1 << sq
        for reach, sliders in ((_FILE_REACH, self.rooks | self.queens),
                               (_RANK_REACH, self.rooks | self.queens),
                               (_DIAG_REACH, self.bishops | self.queens)):
            line = reach[k]
            if line & sq_mask:
                snipers = line & sliders & self.occupied_co[not color]
                for sniper in scan_reversed(snipers):
                    if BB_BETWEEN[sniper][k] & (self.occupied | sq_mask) == sq_mask:
                        return BB_RAYS[k][sniper]
                break
        return BB_ALL

    # ---- castling --------------------------------------------------------
    def clean_castling_rights(self):
        castling = self.castling_rights & self.rooks
        w = castling & BB_RANK_1 & self.occupied_co[WHITE] & (BB_A1 | BB_H1)
        b = castling & BB_RANK_8 & self.occupied_co[BLACK] & (BB_A8 | BB_H8)
        if not self.occupied_co[WHITE] & self.kings & ~self.promoted & BB_E1:
            w = 0
        if not self.occupied_co[BLACK] & self.kings & ~self.promoted & BB_E8:
            b = 0
        return w | b

    def generate_castling_moves(self, from_mask=BB_ALL, to_mask=BB_ALL):
        backrank = BB_RANK_1 if self.turn == WHITE else BB_RANK_8
        king = self.occupied_co[self.turn] & self.kings & ~self.promoted & backrank & from_mask
        king &= -king
        if not king or self._attacked_for_king(king, self.occupied):
            return
        bb_c, bb_d = BB_FILE_C & backrank, BB_FILE_D & backrank
        bb_f, bb_g = BB_FILE_F & backrank, BB_FILE_G & backrank
        for candidate in scan_reversed(self.clean_castling_rights() & backrank & to_mask):
            rook = 1 << candidate
            a_side = rook < king
            king_to = bb_c if a_side else bb_g
            rook_to = bb_d if a_side else bb_f
            king_path = BB_BETWEEN[msb(king)][msb(king_to)]
            rook_path = BB_BETWEEN[candidate][msb(rook_to)]
            if (self.occupied ^ king ^ rook) & (king_path | rook_path | king_to | rook_to):
                continue
            if self._attacked_for_king(king_path | king_to, self.occupied ^ king):
                continue
            if self._castling_uncovers_rank_attack(rook, king_to):
                continue
            # standard (non-960) mode emits castling as the king's two-square move
            yield Move(msb(king), msb(king_to))

    def _castling_uncovers_rank_attack(self, rook_bb, king_to_bb):
        king_to = msb(king_to_bb)
        sliders = (self.queens | self.rooks) & self.occupied_co[not self.turn]
        return bool(rank_attacks(king_to, self.occupied ^ rook_bb) & sliders)

    def is_castling(self, move):
        if self.kings & (1 << move.from_square):
            diff = square_file(move.from_square) - square_file(move.to_square)
            return abs(diff) > 1 or bool(self.rooks & self.occupied_co[self.turn] & (1 << move.to_square))
        return False

    # ---- move generation (python-chess order) ------------------------------
    def generate_pseudo_legal_moves(self, from_mask=BB_ALL, to_mask=BB_ALL):
        ours = self.occupied_co[self.turn]

        # (1) officers and king, source squares from h8 down to a1
        for frm in scan_reversed(ours & ~self.pawns & from_mask):
            for to in scan_reversed(self.attacks_mask(frm) & ~ours & to_mask):
                yield Move(frm, to)

        # (2) castling
        if from_mask & self.kings:
            yield from self.generate_castling_moves(from_mask, to_mask)

        pawns = self.pawns & ours & from_mask
        if not pawns:
            return

        # (3) pawn captures
        theirs = self.occupied_co[not self.turn]
        for frm in scan_reversed(pawns):
            for to in scan_reversed(BB_PAWN_ATTACKS[self.turn][frm] & theirs & to_mask):
                if square_rank(to) in (0, 7):
                    for promo in (QUEEN, ROOK, BISHOP, KNIGHT):
                        yield Move(frm, to, promo)
                else:
                    yield Move(frm, to)

        # (4)/(5) pushes, ordered by destination
        if self.turn == WHITE:
            single = (pawns << 8) & ~self.occupied & BB_ALL
            double = (single << 8) & ~self.occupied & (BB_RANK_3 | BB_RANK_4) & BB_ALL
            back1, back2 = -8, -16
        else:
            single = (pawns >> 8) & ~self.occupied
            double = (single >> 8) & ~self.occupied & (BB_RANK_6 | BB_RANK_5)
            back1, back2 = 8, 16
        single &= to_mask
        double &= to_mask
        for to in scan_reversed(single):
            if square_rank(to) in (0, 7):
                for promo in (QUEEN, ROOK, BISHOP, KNIGHT):
                    yield Move(to + back1, to, promo)
            else:
                yield Move(to + back1, to)
        for to in scan_reversed(double):
            yield Move(to + back2, to)

        # (6) en passant
        if self.ep_square:
            yield from self.generate_pseudo_legal_ep(from_mask, to_mask)

    def generate_pseudo_legal_ep(self, from_mask=BB_ALL, to_mask=BB_ALL):
        if not self.ep_square or not BB_SQUARES[self.ep_square] & to_mask:
            return
        if BB_SQUARES[self.ep_square] & self.occupied:
            return
        capturers = (self.pawns & self.occupied_co[self.turn] & from_mask &
                     BB_PAWN_ATTACKS[not self.turn][self.ep_square] &
                     BB_RANKS[4 if self.turn else 3])
        for c in scan_reversed(capturers):
            yield Move(c, self.ep_square)

    def _slider_blockers(self, king):
        rq = self.rooks | self.queens
        bq = self.bishops | self.queens
        snipers = ((_RANK_REACH[king] & rq) | (_FILE_REACH[king] & rq) | (_DIAG_REACH[king] & bq))
        blockers = 0
        for sniper in scan_reversed(snipers & self.occupied_co[not self.turn]):
            b = BB_BETWEEN[king][sniper] & self.occupied
            if b and (b & (b - 1)) == 0:
                blockers |= b
        return blockers & self.occupied_co[self.turn]

    def is_en_passant(self, move):
        return (self.ep_square == move.to_square and
                bool(self.pawns & (1 << move.from_square)) and
                abs(move.to_square - move.from_square) in (7, 9) and
                not self.occupied & (1 << move.to_square))

    def _ep_skewered(self, king, capturer):
        last_double = self.ep_square + (-8 if self.turn == WHITE else 8)
        occ = (self.occupied & ~(1 << last_double) & ~(1 << capturer)) | (1 << self.ep_square)
        theirs = self.occupied_co[not self.turn]
        if rank_attacks(king, occ) & theirs & (self.rooks | self.queens):
            return True
        if diag_attacks(king, occ) & theirs & (self.bishops | self.queens):
            return True
        return False

    def _is_safe(self, king, blockers, move):
        if move.from_square == king:
            if self.is_castling(move):
                return True
            return not self.is_attacked_by(not self.turn, move.to_square)
        if self.is_en_passant(move):
            return bool(self.pin_mask(self.turn, move.from_square) & (1 << move.to_square)
                        and not self._ep_skewered(king, move.from_square))
        return bool(not blockers & (1 << move.from_square)
                    or BB_RAYS[move.from_square][move.to_square] & (1 << king))

    def _generate_evasions(self, king, checkers, from_mask=BB_ALL, to_mask=BB_ALL):
        sliders = checkers & (self.bishops | self.rooks | self.queens)
        attacked = 0
        for checker in scan_reversed(sliders):
            attacked |= BB_RAYS[king][checker] & ~(1 << checker)
        if (1 << king) & from_mask:
            for to in scan_reversed(BB_KING_ATTACKS[king] & ~self.occupied_co[self.turn] & ~attacked & to_mask):
                yield Move(king, to)
        checker = msb(checkers)
        if (1 << checker) == checkers:
            target = BB_BETWEEN[king][checker] | checkers
            yield from self.generate_pseudo_legal_moves(~self.kings & from_mask, target & to_mask)
            if self.ep_square and not (1 << self.ep_square) & target:
                last_double = self.ep_square + (-8 if self.turn == WHITE else 8)
                if last_double == checker:
                    yield from self.generate_pseudo_legal_ep(from_mask, to_mask)

    def generate_legal_moves(self, from_mask=BB_ALL, to_mask=BB_ALL):
        king_mask = self.kings & self.occupied_co[self.turn]
        if not king_mask:
            yield from self.generate_pseudo_legal_moves(from_mask, to_mask)
            return
        king = msb(king_mask)
        blockers = self._slider_blockers(king)
        checkers = self.attackers_mask(not self.turn, king)
        gen = (self._generate_evasions(king, checkers, from_mask, to_mask) if checkers
               else self.generate_pseudo_legal_moves(from_mask, to_mask))
        for move in gen:
            if self._is_safe(king, blockers, move):
                yield move

    def generate_legal_ep(self, from_mask=BB_ALL, to_mask=BB_ALL):
        king_mask = self.kings & self.occupied_co[self.turn]
        for move in self.generate_pseudo_legal_ep(from_mask, to_mask):
            if not king_mask:
                yield move
                continue
            king = msb(king_mask)
            if self._ep_is_legal(king, move):
                yield move

    def _ep_is_legal(self, king, move):
        """python-chess: `not self.is_into_check(move)` for ep = evasions-aware safety test."""
        checkers = self.attackers_mask(not self.turn, king)
        if checkers:
            if not any(m == move for m in self._generate_evasions(
                    king, checkers, 1 << move.from_square, 1 << move.to_square)):
                return False
        return self._is_safe(king, self._slider_blockers(king), move)

    def has_legal_en_passant(self):
        return self.ep_square is not None and any(self.generate_legal_ep())

    @property
    def legal_moves(self):
        return LegalMoveGenerator(self)

    def is_legal(self, move):
        return any(m == move for m in self.generate_legal_moves())

    # ---- make / unmake ---------------------------------------------------
    def is_zeroing(self, move):
        return bool((1 << move.from_square) & self.pawns or
                    (1 << move.to_square) & self.occupied_co[not self.turn])

    def is_irreversible(self, move):
        backrank = BB_RANK_1 if self.turn == WHITE else BB_RANK_8
        cr = self.clean_castling_rights() & backrank
        if REPETITION_STOPS_ON_LEGAL_EP and self.has_legal_en_passant():
            return True
        return bool(self.is_zeroing(move) or
                    (cr and (1 << move.from_square) & self.kings & ~self.promoted) or
                    cr & (1 << move.from_square) or
                    cr & (1 << move.to_square))

    def push(self, move):
        self.move_stack.append(Move(move.from_square, move.to_square, move.promotion))
        self._stack.append(_Snapshot(self))

        ep_square = self.ep_square
        self.ep_square = None
        self.halfmove_clock += 1
        if self.turn == BLACK:
            self.fullmove_number += 1

        if not move:  # null move
            self.turn = not self.turn
            return

        if self.is_zeroing(move):
            self.halfmove_clock = 0

        from_bb, to_bb = 1 << move.from_square, 1 << move.to_square
        to_square = move.to_square
        promoted = bool(self.promoted & from_bb)
        piece_type = self._remove_piece_at(move.from_square)
        assert piece_type is not None, "push() expects move to be pseudo-legal"

        # standard-mode castling arrives as e1g1/e1c1/e8g8/e8c8: translate to king-takes-rook
        if piece_type == KING and abs(square_file(move.from_square) - square_file(to_square)) > 1:
            to_square = (H1 if square_file(to_square) == 6 else A1) + (0 if self.turn == WHITE else 56)
            to_bb = 1 << to_square

        captured = self.piece_type_at(to_square)

        self.castling_rights = self.clean_castling_rights() & ~to_bb & ~from_bb
        if piece_type == KING and not promoted:
            self.castling_rights &= ~(BB_RANK_1 if self.turn == WHITE else BB_RANK_8)

        if piece_type == PAWN:
            diff = to_square - move.from_square
            if diff == 16 and square_rank(move.from_square) == 1:
                self.ep_square = move.from_square + 8
            elif diff == -16 and square_rank(move.from_square) == 6:
                self.ep_square = move.from_square - 8
            elif to_square == ep_square and abs(diff) in (7, 9) and not captured:
                self._remove_piece_at(ep_square + (-8 if self.turn == WHITE else 8))

        if move.promotion:
            promoted = True
            piece_type = move.promotion

        castling = piece_type == KING and bool(self.occupied_co[self.turn] & to_bb)
        if castling:
            a_side = square_file(to_square) < square_file(move.from_square)
            self._remove_piece_at(move.from_square)
            self._remove_piece_at(to_square)
            base = 0 if self.turn == WHITE else 56
            if a_side:
                self._set_piece_at(base + C1, KING, self.turn)
                self._set_piece_at(base + D1, ROOK, self.turn)
            else:
                self._set_piece_at(base + G1, KING, self.turn)
                self._set_piece_at(base + F1, ROOK, self.turn)
        else:
            self._set_piece_at(to_square, piece_type, self.turn, promoted)

        self.turn = not self.turn

    def pop(self):
        move = self.move_stack.pop()        # IndexError on an empty stack (netencoder.py:62-65)
        self._stack.pop().restore(self)
        return move

    def peek(self):
        return self.move_stack[-1]

    def copy(self, stack=True):
        b = Board(None)
        b.pawns, b.knights, b.bishops = self.pawns, self.knights, self.bishops
        b.rooks, b.queens, b.kings = self.rooks, self.queens, self.kings
        b.occupied_co = {WHITE: self.occupied_co[WHITE], BLACK: self.occupied_co[BLACK]}
        b.occupied, b.promoted = self.occupied, self.promoted
        b.turn, b.castling_rights, b.ep_square = self.turn, self.castling_rights, self.ep_square
        b.halfmove_clock, b.fullmove_number = self.halfmove_clock, self.fullmove_number
        if stack:
            b.move_stack = [Move(m.from_square, m.to_square, m.promotion) for m in self.move_stack]
            b._stack = list(self._stack)   # snapshots are immutable once taken
        return b

    # ---- game end --------------------------------------------------------
    def is_checkmate(self):
        return self.is_check() and not any(self.generate_legal_moves())

    def is_stalemate(self):
        return not self.is_check() and not any(self.generate_legal_moves())

    def has_insufficient_material(self, color):
        ours = self.occupied_co[color]
        if ours & (self.pawns | self.rooks | self.queens):
            return False
        if ours & self.knights:
            return (popcount(ours) <= 2 and
                    not (self.occupied_co[not color] & ~self.kings & ~self.queens))
        if ours & self.bishops:
            same_colour = (not self.bishops & BB_DARK_SQUARES) or (not self.bishops & BB_LIGHT_SQUARES)
            return bool(same_colour and not self.pawns and not self.knights)
        return True

    def is_insufficient_material(self):
        return all(self.has_insufficient_material(c) for c in COLORS)

    def is_seventyfive_moves(self):
        return self.halfmove_clock >= 150 and any(self.generate_legal_moves())

    def _transposition_key(self):
        return (self.pawns, self.knights, self.bishops, self.rooks, self.queens, self.kings,
                self.occupied_co[WHITE], self.occupied_co[BLACK], self.turn,
                self.clean_castling_rights(),
                self.ep_square if self.has_legal_en_passant() else None)

    def is_repetition(self, count=3):
        key = self._transposition_key()
        seen = 1
        undone = []
        try:
            while self.move_stack and seen < count:
                move = self.pop()
                undone.append(move)
                if self.is_irreversible(move):
                    break
                if self._transposition_key() == key:
                    seen += 1
        finally:
            while undone:
                self.push(undone.pop())
        return seen >= count

    def is_fivefold_repetition(self):
        return self.is_repetition(5)

    def can_claim_fifty_moves(self):
        if self.halfmove_clock >= 100:
            if any(self.generate_legal_moves()):
                return True
        if FIFTY_MOVE_CLAIM_LOOKAHEAD and self.halfmove_clock >= 99:
            for move in self.generate_legal_moves():
                if not self.is_zeroing(move):
                    self.push(move)
                    try:
                        if self.can_claim_fifty_moves():
                            return True
                    finally:
                        self.pop()
        return False

    def is_game_over(self, claim_draw=False):
        if self.is_seventyfive_moves():
            return True
        if self.is_insufficient_material():
            return True
        if not any(self.generate_legal_moves()):
            return True
        if self.is_fivefold_repetition():
            return True
        return False

    def result(self, claim_draw=False):
        if self.is_checkmate():
            return "0-1" if self.turn == WHITE else "1-0"
        if self.is_seventyfive_moves() or self.is_fivefold_repetition():
            return "1/2-1/2"
        if self.is_insufficient_material():
            return "1/2-1/2"
        if not any(self.generate_legal_moves()):
            return "1/2-1/2"
        return "*"

    def __repr__(self):
        return "Board({!r})".format(self.fen())

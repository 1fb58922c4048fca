"""Pins the python-chess 0.28.3 restatement (oracle/pychess_compat/chess) with PUBLIC known answers:
the chessprogramming.org perft table and the start-position move order printed in
python-chess's README (KAT-1), plus the rule corners Game.get_result depends on (game.py:92-109)."""
import json
import os

import pytest

import chessrl_oracle as O
import perft_kats

chess = O.chess

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
PERFT = [
    (chess.STARTING_FEN, [20, 400, 8902, 197281]),
    (KIWI, [48, 2039, 97862]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", [14, 191, 2812, 43238]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", [6, 264, 9467]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890]),
]


def perft(b, d):
    if d == 1:
        return sum(1 for _ in b.generate_legal_moves())
    n = 0
    for m in list(b.generate_legal_moves()):
        b.push(m)
        n += perft(b, d - 1)
        b.pop()
    return n


@pytest.mark.parametrize("fen,expected", PERFT)
def test_perft_table(fen, expected):
    b = chess.Board(fen)
    for d, e in enumerate(expected, 1):
        assert perft(b, d) == e


@pytest.mark.parametrize("fen,expected", perft_kats.EDGE)
def test_perft_rule_corner_positions(fen, expected):
    """En passant pins, castling through / into check, promotions in and out of check, stalemate traps."""
    b = chess.Board(fen)
    for d, e in enumerate(expected, 1):
        if e > perft_kats.ORACLE_NODE_CAP:
            break
        assert perft(b, d) == e, (fen, d)


@pytest.mark.parametrize("fen", perft_kats.MAX_MOVES)
def test_most_legal_moves_of_any_position(fen):
    ms = [m.uci() for m in chess.Board(fen).legal_moves]
    assert len(ms) == 218 and len(set(ms)) == 218
    idx = O.label_index()
    assert all(m in idx for m in ms)                       # every one has a policy label


def test_start_move_order_kat1():
    want = ("g1h3 g1f3 b1c3 b1a3 h2h3 g2g3 f2f3 e2e3 d2d3 c2c3 b2b3 a2a3 "
            "h2h4 g2g4 f2f4 e2e4 d2d4 c2c4 b2b4 a2a4").split()
    assert [m.uci() for m in chess.Board().legal_moves] == want
    idx = O.label_index()
    assert [idx[m] for m in want[:4]] == [1403, 1402, 217, 215]


def test_castling_order_and_uci():
    b = chess.Board(KIWI)
    ms = [m.uci() for m in b.legal_moves]
    assert ms.index("e1g1") + 1 == ms.index("e1c1")          # king side first, both as king moves
    assert chess.Move.null().uci() == "0000"


def test_promotion_order_q_r_b_n():
    b = chess.Board("8/P6k/8/8/8/8/8/K7 w - - 0 1")
    ms = [m.uci() for m in b.legal_moves]
    i = ms.index("a7a8q")
    assert ms[i:i + 4] == ["a7a8q", "a7a8r", "a7a8b", "a7a8n"]


def test_ep_square_set_after_every_double_push_but_fen_only_if_legal():
    b = chess.Board()
    b.push(chess.Move.from_uci("e2e4"))
    assert b.ep_square == 20                                   # e3, although no black pawn can capture
    assert b.fen().split()[3] == "-"
    assert not b.has_legal_en_passant()


def test_ep_pin_is_illegal():
    b = chess.Board("8/8/8/K2pP2r/8/8/8/4k3 w - d6 0 2")       # capturing e5xd6 would expose the king on the rank
    assert "e5d6" not in [m.uci() for m in b.legal_moves]


def test_insufficient_material_cases():
    cases = {"8/8/8/8/8/8/8/K6k w - - 0 1": True, "8/8/8/8/8/8/8/KN5k w - - 0 1": True,
             "8/8/8/8/8/8/8/KB5k w - - 0 1": True, "8/8/8/8/8/8/8/KNN4k w - - 0 1": False,
             "8/8/8/8/8/8/8/KN4nk w - - 0 1": False, "8/8/8/8/8/8/8/KB4bk w - - 0 1": False,   # b1/g1 opposite colours
             "8/8/8/8/8/8/8/KB3b1k w - - 0 1": True,                                            # b1/f1 same colour
             "8/8/8/8/8/8/P7/K6k w - - 0 1": False, "8/8/8/8/8/8/8/KR5k w - - 0 1": False}
    for fen, want in cases.items():
        assert chess.Board(fen).is_insufficient_material() is want, fen


def test_fifty_move_claim_needs_a_legal_move():
    g = O.OGame(board=chess.Board("8/8/4k3/8/8/3K4/R7/8 w - - 100 80"))
    assert g.get_result() == 0
    g = O.OGame(board=chess.Board("8/8/4k3/8/8/3K4/R7/8 w - - 99 80"))
    assert g.get_result() is None
    mate = O.OGame(board=chess.Board("R5k1/5ppp/8/8/8/8/8/6K1 b - - 120 90"))    # mated: no claim, white wins
    assert mate.get_result() == 1


def test_fivefold_but_not_threefold():
    g = O.OGame()
    cycle = ["g1f3", "g8f6", "f3g1", "f6g8"]
    for rep in range(4):
        for m in cycle:
            assert g.get_result() is None
            assert g.move(m)
    assert g.get_result() == 0          # start position seen five times
    g2 = O.OGame()
    for m in cycle * 2:
        g2.move(m)
    assert g2.get_result() is None      # threefold alone does not end the game (game.py:95-96)


def test_irreversible_move_cuts_repetition_window():
    g = O.OGame()
    for m in ["g1f3", "g8f6", "f3g1", "f6g8"] * 3:
        g.move(m)
    g.move("h1g1")                       # gives up a castling right: irreversible, though not zeroing
    g.move("g8f6")
    g.move("g1h1")
    g.move("f6g8")
    assert g.get_result() is None


def test_game_move_rejects_illegal_and_null():
    g = O.OGame()
    assert g.move("e2e5") is False and g.move("00000") is False and len(g) == 0
    assert g.move("e2e4") is True and len(g) == 1


def test_golden_rules(golden_dir):
    data = json.load(open(os.path.join(golden_dir, "rules.json")))["cases"]
    n = 0
    for c in data:
        if c["fen"]:
            g = O.OGame(board=chess.Board(c["fen"]))
            assert g.get_legal_moves() == c["legal"] and g.get_result() == c["result"]
            continue
        g = O.OGame()
        trace = {t["ply"]: t for t in c["trace"]}
        for i, m in enumerate(c["moves"]):
            assert g.move(m)
            t = trace.get(i + 1)
            if t:
                assert g.get_legal_moves() == t["legal"] and g.get_result() == t["result"]
                assert g.board.fen() == t["fen"]
                n += 1
        assert g.get_result() == c["final_result"]
    assert n > 500


# ---- vectors from outside this repository (tests/pychess_kats.py: python-chess README / docs, and its test-suite) ------
import random  # noqa: E402

import move_order_rules  # noqa: E402
import position_fuzz  # noqa: E402
import pychess_kats as K  # noqa: E402
from chessrl_b200 import boards as B  # noqa: E402


def test_readme_scholars_mate_walkthrough():
    b = chess.Board()
    assert [m.uci() for m in b.legal_moves] == K.README_START_LEGAL_MOVES
    assert b.legal_moves.count() == 20 and bool(b.legal_moves) and chess.Move.from_uci("g1f3") in b.legal_moves
    for m in K.README_SCHOLARS_MATE:
        mv = chess.Move.from_uci(m)
        assert mv in b.legal_moves
        b.push(mv)
    assert b.fen() == K.README_SCHOLARS_MATE_FEN
    for name, want in K.README_SCHOLARS_MATE_FLAGS.items():
        assert getattr(b, name)() is want, name
    assert b.halfmove_clock == K.README_SCHOLARS_MATE_HALFMOVE_CLOCK
    assert b.is_attacked_by(chess.WHITE, 60) is K.README_ATTACKED_E8_BY_WHITE
    assert b.attackers_mask(chess.WHITE, 21) == K.README_ATTACKERS_OF_F3_BY_WHITE
    assert b.result() == "1-0" and O.OGame(board=b).get_result() == 1
    assert b.pop().uci() == "h5f7" and not b.is_checkmate()               # "make and unmake moves"
    r = chess.Board(K.README_FEN_ROUND_TRIP)
    assert r.fen() == K.README_FEN_ROUND_TRIP and r.piece_type_at(34) == chess.KING and not r.occupied_co[chess.WHITE] & (1 << 34)


def test_squareset_semantics_used_by_netencoder():
    s = chess.SquareSet(K.SQUARESET_DOC_MASK)
    assert len(s) == 9 and bool(s) and 1 in s and list(s) == K.SQUARESET_DOC_LIST
    bools = s.tolist()
    assert len(bools) == 64 and [i for i, x in enumerate(bools) if x] == K.SQUARESET_DOC_LIST
    assert list(s.mirror()) == [0, 56, 57, 58, 59, 60, 61, 62, 63]        # vertical mirror: rank 1 <-> rank 8


@pytest.mark.parametrize("fen,white,black", K.INSUFFICIENT)
def test_suite_insufficient_material_per_colour(fen, white, black):
    b = chess.Board(fen)
    assert b.has_insufficient_material(chess.WHITE) is white and b.has_insufficient_material(chess.BLACK) is black
    assert b.is_insufficient_material() is (white and black)


def test_suite_fivefold_repetition_need_not_be_consecutive():
    b = chess.Board(K.FIVEFOLD_FEN)
    g = O.OGame(board=chess.Board(K.FIVEFOLD_FEN))
    for cycle in range(4):
        for m in K.FIVEFOLD_CYCLE:
            assert not b.is_fivefold_repetition() and not b.is_game_over() and g.get_result() is None
            b.push(chess.Move.from_uci(m))
            assert g.move(m)
    assert b.is_fivefold_repetition() and b.is_game_over() and g.get_result() == 0
    assert b.is_repetition(3)
    for m in K.FIVEFOLD_DETOUR:
        b.push(chess.Move.from_uci(m))
        assert not b.is_fivefold_repetition() and not b.is_game_over()
    b.push(chess.Move.from_uci(K.FIVEFOLD_RETURN))
    assert b.is_fivefold_repetition() and b.fen().split()[0] == K.FIVEFOLD_FEN.split()[0]


@pytest.mark.parametrize("fen,claim,seventyfive", K.FIFTY_MOVES)
def test_suite_fifty_move_claim(fen, claim, seventyfive):
    b = chess.Board(fen)
    assert b.can_claim_fifty_moves() is claim and b.is_seventyfive_moves() is seventyfive
    res = O.OGame(board=chess.Board(fen)).get_result()
    if claim:
        assert res == 0                                                  # game.py:95-96: the claim ends the game
    elif b.is_checkmate():
        assert res in (1, -1)
    elif b.is_stalemate():
        assert res == 0


def test_version_switches_stay_at_0_28_3():
    assert chess.FIFTY_MOVE_CLAIM_LOOKAHEAD is False and chess.REPETITION_STOPS_ON_LEGAL_EP is False


def _order_violations(board):
    rec = B.record_from_fen(board.fen())                                  # bitboards only: the validator generates nothing
    words = [B.uci_to_move(m.uci()) for m in board.legal_moves]
    return move_order_rules.violations(rec, words, board.is_check())


def test_move_order_invariants_on_fuzzed_positions_and_games():
    """Class order 1-6, from / to squares descending, q r b n, king evasions first: the structural rules of
    SURVEY.md 8c as an independent validator over positions with promoted material, checks, castling and ep."""
    rng = random.Random(2024)
    n_check = n_castle = n_promo = n_ep = 0
    for _ in range(1500):
        fen, b = position_fuzz.random_fen(rng)
        bad = _order_violations(b)
        assert not bad, (fen, bad)
        ms = [m.uci() for m in b.legal_moves]
        n_check += b.is_check()
        n_castle += any(m in ("e1g1", "e1c1", "e8g8", "e8c8") and b.piece_type_at(chess.Move.from_uci(m).from_square) == chess.KING for m in ms)
        n_promo += any(len(m) == 5 for m in ms)
        n_ep += b.has_legal_en_passant()
    assert min(n_check, n_castle, n_promo, n_ep) >= 25, (n_check, n_castle, n_promo, n_ep)
    for fen, _ in PERFT + [(f, None) for f in perft_kats.MAX_MOVES]:
        assert not _order_violations(chess.Board(fen))
    b = chess.Board()
    for _ in range(300):                                                  # along a random game from the start
        ms = list(b.legal_moves)
        if not ms:
            break
        assert not _order_violations(b)
        b.push(rng.choice(ms))


def test_move_order_validator_rejects_wrong_orders():
    """The validator bites: swapping two moves, or reordering a promotion group, is reported."""
    b = chess.Board(KIWI)
    rec = B.record_from_fen(KIWI)
    words = [B.uci_to_move(m.uci()) for m in b.legal_moves]
    assert not move_order_rules.violations(rec, words, False)
    for i in range(len(words) - 1):
        sw = list(words)
        sw[i], sw[i + 1] = sw[i + 1], sw[i]
        assert move_order_rules.violations(rec, sw, False), i
    p = chess.Board("1n2k3/P7/8/8/8/8/8/4K3 w - - 0 1")
    words = [B.uci_to_move(m.uci()) for m in p.legal_moves]
    rec = B.record_from_fen(p.fen())
    assert not move_order_rules.violations(rec, words, False)
    qs = [i for i, w in enumerate(words) if (w >> 12) == 4]
    sw = list(words)
    sw[qs[0]], sw[qs[0] + 3] = sw[qs[0] + 3], sw[qs[0]]                   # n, r, b, q instead of q, r, b, n
    assert move_order_rules.violations(rec, sw, False)

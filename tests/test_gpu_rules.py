"""GPU parity, rules: the CUDA movegen / make-move / perft kernels (through the C ABI) against the oracle,
the golden move lists and the public perft table -- bit-exact, python-chess move ORDER included."""
import json
import os
import random

import numpy as np
import pytest
import torch

import chessrl_oracle as O
import perft_kats
import position_fuzz
from chessrl_b200 import boards as B

pytestmark = pytest.mark.gpu
chess = O.chess

PAIR = os.environ.get("CRL_PERFT_PAIR", "0") in ("5", "6")     # the last two plies as one pass (k_perft_pair; optional)
KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
PERFT5 = {B.STARTING_FEN: [20, 400, 8902, 197281, 4865609], KIWI: [48, 2039, 97862, 4085603, 193690690]}
EXTRA = {
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1": [14, 191, 2812, 43238, 674624],
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1": [6, 264, 9467, 422333],
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8": [44, 1486, 62379, 2103487],
    "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10": [46, 2079, 89890, 3894594],
}


def _moves_host(moves_t, counts_t):
    mv = moves_t.cpu().numpy().view(np.uint16)
    cn = counts_t.cpu().numpy()
    return [[B.move_to_uci(m) for m in mv[i, :cn[i]]] for i in range(len(cn))]


def _random_positions(n_games, plies, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n_games):
        b = chess.Board()
        for _ in range(rng.randrange(1, plies)):
            ms = list(b.generate_legal_moves())
            if not ms:
                break
            b.push(rng.choice(ms))
        out.append(b)
    return out


def test_movegen_golden_positions(engine1, golden_dir):
    cases = [c for c in json.load(open(os.path.join(golden_dir, "rules.json")))["cases"] if c["fen"]]
    recs = np.stack([B.record_from_fen(c["fen"]) for c in cases])
    mv, cn, fl = engine1.movegen(engine1.boards_to_device(recs))
    got = _moves_host(mv, cn)
    for c, g in zip(cases, got):
        assert g == c["legal"], c["name"]


def test_movegen_matches_oracle_on_random_positions(engine1):
    boards = _random_positions(400, 120, seed=5)
    recs = np.stack([B.record_from_fen(b.fen()) for b in boards])
    # the FEN drops an ep square that is not capturable; restore the raw python-chess ep square in the record
    for r, b in zip(recs, boards):
        m = B.meta_fields(r[8])
        r[8] = B.pack_meta(m["turn"], m["castle"], -1 if b.ep_square is None else b.ep_square, m["halfmove"], m["fullmove"])
    mv, cn, fl = engine1.movegen(engine1.boards_to_device(recs))
    got = _moves_host(mv, cn)
    flags = fl.cpu().numpy()
    for b, g, f in zip(boards, got, flags):
        assert g == [m.uci() for m in b.generate_legal_moves()], b.fen()
        assert bool(f & 1) == b.is_check() and bool(f & 2) == b.has_legal_en_passant()


def test_make_moves_matches_oracle(engine1):
    boards = _random_positions(200, 100, seed=6)
    rng = random.Random(7)
    recs, moves, after = [], [], []
    for b in boards:
        ms = list(b.generate_legal_moves())
        if not ms:
            continue
        m = rng.choice(ms)
        r = B.record_from_fen(b.fen())
        f = B.meta_fields(r[8])
        r[8] = B.pack_meta(f["turn"], f["castle"], -1 if b.ep_square is None else b.ep_square, f["halfmove"], f["fullmove"])
        recs.append(r)
        moves.append(B.uci_to_move(m.uci()))
        b.push(m)
        after.append(b)
    t = engine1.boards_to_device(np.stack(recs))
    mt = torch.from_numpy(np.array(moves, dtype=np.uint16).view(np.int16)).to(engine1.device)
    engine1.make_moves(t, mt)
    out = engine1.boards_to_host(t)
    for r, b in zip(out, after):
        f = B.meta_fields(r[8])
        assert B.board_fen_from_record(r) == b.board_fen()
        assert f["turn"] == b.turn and f["halfmove"] == b.halfmove_clock and f["fullmove"] == b.fullmove_number
        assert f["ep"] == (-1 if b.ep_square is None else b.ep_square)
        cr = b.clean_castling_rights()
        want = (1 if cr & chess.BB_H1 else 0) | (2 if cr & chess.BB_A1 else 0) | (4 if cr & chess.BB_H8 else 0) | (8 if cr & chess.BB_A8 else 0)
        assert f["castle"] == want and f["ply"] == 1


@pytest.mark.parametrize("fen,expected", list(PERFT5.items()) + list(EXTRA.items()))
def test_perft_single_lane(engine1, fen, expected):
    t = engine1.boards_to_device(B.record_from_fen(fen)[None, :])
    for d, e in enumerate(expected[:4], 1):
        assert int(engine1.perft(t, d, bulk=True)[0]) == e
    assert int(engine1.perft(t, 3, bulk=False)[0]) == expected[2]


@pytest.mark.parametrize("fen", [B.STARTING_FEN, KIWI])
def test_perft_depth5_lockstep_frontier(engine1, fen):
    """BASELINE config 2: breadth-first to a frontier of >= 65,536 boards on the device, then every lane finishes
    the remaining plies depth-first in lockstep; totals must equal the public perft(5) numbers."""
    expected = PERFT5[fen]
    frontier = engine1.boards_to_device(B.record_from_fen(fen)[None, :])
    depth = 0
    while frontier.shape[1] < 65536:
        frontier, counts = engine1.expand_frontier(frontier)
        depth += 1
        assert frontier.shape[1] == expected[depth - 1]
    nodes = engine1.perft(frontier, 5 - depth, bulk=True)
    assert int(nodes.sum()) == expected[4]
    nodes2 = engine1.perft(frontier, 5 - depth, bulk=False)
    assert int(nodes2.sum()) == expected[4]


def test_perft_replicated_65536_lanes(engine1):
    """65,536 copies of the start position and Kiwipete run perft(3) in lockstep (pure lockstep throughput case)."""
    for fen, want in ((B.STARTING_FEN, 8902), (KIWI, 97862)):
        t = engine1.boards_to_device(np.tile(B.record_from_fen(fen), (65536, 1)))
        nodes = engine1.perft(t, 3, bulk=True)
        assert bool((nodes == want).all())


def test_game_replay_semantics(engine1):
    start = B.record_from_fen()
    moves = [B.uci_to_move(m) for m in ["e2e4", "e7e5", "e2e5", "00000", "g1f3"]]
    r = engine1.game_replay(start, moves)
    assert list(r["accepted"]) == [True, True, False, False, True]          # Game.move rejects illegal / null moves
    g = O.OGame()
    for m in ["e2e4", "e7e5", "g1f3"]:
        g.move(m)
    assert [B.move_to_uci(m) for m in r["legal"]] == g.get_legal_moves() and r["result"] is None
    # fool's mate: black wins -> -1
    r = engine1.game_replay(start, [B.uci_to_move(m) for m in ["f2f3", "e7e5", "g2g4", "d8h4"]])
    assert r["result"] == -1 and len(r["legal"]) == 0
    # fivefold repetition through the device key ring
    cyc = [B.uci_to_move(m) for m in ["g1f3", "g8f6", "f3g1", "f6g8"]]
    assert engine1.game_replay(start, cyc * 3)["result"] is None
    assert engine1.game_replay(start, cyc * 4)["result"] == 0


def test_game_replay_records_one_round_trip(engine1):
    """crl_game_replay_records_host: the record after every ACCEPTED ply equals replaying the prefixes one by one,
    and the drop-in Game builds its history from it (bulk load = one device round trip)."""
    from chessrl_b200.game import Game
    start = B.record_from_fen()
    ucis = ["e2e4", "e7e5", "e2e5", "g1f3", "00000", "b8c6", "f1b5", "a7a6", "b5c6", "d7c6", "e1g1"]
    words = [B.uci_to_move(m) for m in ucis]
    r = engine1.game_replay(start, words, records=True)
    kept = [m for m, ok in zip(ucis, r["accepted"]) if ok]
    assert kept == [m for m in ucis if m not in ("e2e5", "00000")]
    assert r["records"].shape == (len(kept) + 1, 9) and (r["records"][0] == start).all()
    for j in range(len(kept) + 1):
        one = engine1.game_replay(start, [B.uci_to_move(m) for m in kept[:j]])
        assert (one["record"] == r["records"][j]).all(), j
    assert (r["record"] == r["records"][-1]).all()
    empty = engine1.game_replay(start, [], records=True)
    assert empty["records"].shape == (1, 9) and (empty["records"][0] == start).all()
    g = Game()
    g._sync(extra=ucis)
    og = O.OGame()
    for m in kept:
        og.move(m)
    assert g.board.move_stack == kept and g.get_legal_moves() == og.get_legal_moves()
    assert [tuple(x) for x in g.history_records()] == [tuple(x) for x in r["records"][::-1][:9]]


def test_game_results_along_golden_games(engine1, golden_dir):
    cases = [c for c in json.load(open(os.path.join(golden_dir, "rules.json")))["cases"] if not c["fen"]]
    for c in cases:
        mv = [B.uci_to_move(m) for m in c["moves"]]
        r = engine1.game_replay(B.record_from_fen(), mv)
        assert r["accepted"].all() and r["result"] == c["final_result"], c["name"]
        last = c["trace"][-1]
        assert [B.move_to_uci(m) for m in r["legal"]] == last["legal"]


@pytest.mark.parametrize("fen,expected", list(PERFT5.items()) + list(EXTRA.items()))
def test_perft_root_device_driven(engine1, fen, expected):
    """crl_perft_root_host: breadth-first plies on the device (atomic placement, no host round trip) + per-lane walk
    give the published totals for every split between breadth-first plies and walk depth, bulk or not, including the
    capacity-overflow retry (Kiwipete: 2,039 boards < 4,096 -> 97,862 > the initial 65,536-board buffers)."""
    rec = B.record_from_fen(fen)
    for depth, want in enumerate(expected[:5], start=1):
        for min_frontier, bulk in ((1, True), (300, False), (4096, True), (1 << 20, True)):
            if (not bulk and want > 5_000_000) or (min_frontier == 1 and want > 200_000):
                continue                      # a single lane walking millions of nodes only tests patience
            total, lanes, plies = engine1.perft_root(rec, depth, bulk=bulk, min_frontier=min_frontier)
            assert total == want, (fen, depth, min_frontier, bulk, total, want)
            assert 0 <= plies <= depth - 1 and lanes >= 1
            if min_frontier == 1:                 # nothing expanded into HBM; depth 2 = the two-ply pass on the root itself
                assert plies == 0 and lanes == (expected[0] if depth == 2 and PAIR else 1)
            elif depth >= 2 and PAIR and min_frontier == 1 << 20:
                assert plies == depth - 2 and lanes == expected[depth - 2]     # last-but-one ply dealt, never stored
    assert engine1.perft_root(rec, 0)[0] == 1


def test_perft_rule_corner_positions(engine1):
    """perft_kats.EDGE (ep pins, castling into / through check, promotions in and out of check, stalemate traps):
    shallow depths with one lane per root, every depth through the device-driven breadth-first + walk path."""
    recs = np.stack([B.record_from_fen(fen) for fen, _ in perft_kats.EDGE])
    t = engine1.boards_to_device(recs)
    for d in range(1, 5):
        got = engine1.perft(t, d, bulk=bool(d & 1)).cpu().numpy()
        for (fen, exp), g in zip(perft_kats.EDGE, got):
            assert int(g) == exp[d - 1], (fen, d)
    for fen, exp in perft_kats.EDGE:
        rec = B.record_from_fen(fen)
        for depth, want in enumerate(exp, start=1):
            total, lanes, plies = engine1.perft_root(rec, depth, bulk=True, min_frontier=4096)
            assert total == want, (fen, depth, total, want)
        total, _, _ = engine1.perft_root(rec, len(exp) - 1, bulk=False, min_frontier=256)
        assert total == exp[-2], fen


def test_movegen_most_legal_moves(engine1):
    """The two 218-move positions: list and python-chess ORDER, policy indices, one breadth-first ply."""
    recs = np.stack([B.record_from_fen(f) for f in perft_kats.MAX_MOVES])
    t = engine1.boards_to_device(recs)
    mv, cn, fl = engine1.movegen(t)
    got = _moves_host(mv, cn)
    idx = engine1.policy_index(mv, cn).cpu().numpy()
    lab = O.label_index()
    for fen, g, row in zip(perft_kats.MAX_MOVES, got, idx):
        want = [m.uci() for m in chess.Board(fen).generate_legal_moves()]
        assert len(g) == 218 and g == want
        assert [int(x) for x in row[:218]] == [lab[m] for m in want]
    kids, counts = engine1.expand_frontier(t)
    assert counts.cpu().tolist() == [218, 218] and kids.shape[1] == 436
    assert engine1.perft(t, 1).cpu().tolist() == [218, 218]


def test_rules_on_unreachable_random_positions(engine1):
    """800 seeded positions no game from the start reaches (position_fuzz: promoted material, castling rights, raw ep
    squares): k_movegen lists in python-chess ORDER + flags, and every child of one breadth-first ply (k_bfs / make-move)
    against the restatement."""
    rng = random.Random(12)
    cases = [position_fuzz.random_fen(rng) for _ in range(800)]
    t = engine1.boards_to_device(np.stack([B.record_from_fen(fen) for fen, _ in cases]))
    mv, cn, fl = engine1.movegen(t)
    got, flags = _moves_host(mv, cn), fl.cpu().numpy()
    for (fen, b), g, f in zip(cases, got, flags):
        assert g == [m.uci() for m in b.generate_legal_moves()], fen
        assert bool(f & 1) == b.is_check() and bool(f & 2) == b.has_legal_en_passant(), fen
    kids, counts = engine1.expand_frontier(t)
    assert counts.cpu().tolist() == [len(g) for g in got]
    kid_recs = engine1.boards_to_host(kids)
    _, _, kfl = engine1.movegen(kids)
    kflags = kfl.cpu().numpy()
    k = 0
    for (fen, b), g in zip(cases, got):
        for m in g:
            b.push(chess.Move.from_uci(m))
            assert B.fen_from_record(kid_recs[k], bool(kflags[k] & 2)) == b.fen(), (fen, m)
            b.pop()
            k += 1
    assert k == kids.shape[1]


@pytest.mark.parametrize("fen,depth,want", [(B.STARTING_FEN, 5, 4865609), (KIWI, 4, 4085603), (KIWI, 5, 193690690),
                                            ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 6, 11030083)])
def test_perft_root_sharded_totals_add_up(engine1, fen, depth, want):
    """crl_perft_root_shard_host: the shards' totals add up to the published perft, wherever the split happens (in a
    breadth-first ply, or only at the walk because the frontier never reached shard_min) and however the lanes divide."""
    rec = B.record_from_fen(fen)
    for n_shards, shard_min, min_frontier in ((2, 1 << 16, 1 << 18), (3, 200, 5000), (8, 1 << 12, 1 << 16), (5, 1 << 40, 1 << 12)):
        parts = [engine1.perft_root(rec, depth, bulk=(s % 2 == 0), min_frontier=min_frontier, shard=s, n_shards=n_shards,
                                    shard_min_frontier=shard_min) for s in range(n_shards)]
        assert sum(p[0] for p in parts) == want, (n_shards, shard_min, [p[0] for p in parts])
        assert all(p[0] > 0 for p in parts)
    assert engine1.perft_root(rec, depth)[0] == want                      # the unsharded call is shard 0 of 1


def test_warp_cooperative_generator_equals_k_movegen(engine1):
    """crl_debug_movegen_warp (warp_gen.cuh, the generator behind the tree expansions and the small perft plies) against
    k_movegen, list for list in order, with the in-check / legal-ep flags: 6,000 fuzzed positions, 40,000 of their
    children, the perft suite and the 218-move positions."""
    rng = random.Random(5)
    fens = [position_fuzz.random_fen(rng)[0] for _ in range(6000)] + [f for f, _ in perft_kats.EDGE] + perft_kats.MAX_MOVES
    recs = np.stack([B.record_from_fen(f) for f in fens])
    t = engine1.boards_to_device(recs)
    kids, _ = engine1.expand_frontier(t[:, :1500].contiguous())
    t = torch.cat([t, kids[:, :40000]], dim=1).contiguous()
    mv, cn, fl = engine1.movegen(t)
    mw, cw, fw = engine1.movegen_warp(t)
    assert torch.equal(cn, cw) and torch.equal(fl, fw)
    idx = torch.arange(B.MAX_MOVES, device=mv.device)[None, :] < cn[:, None]
    assert torch.equal(torch.where(idx, mv, torch.zeros_like(mv)), torch.where(idx, mw, torch.zeros_like(mw)))
    assert int(cn.max()) == 218 and int((fl & 1).sum()) > 1000 and int((fl & 2).sum()) > 100


def test_movegen_order_invariants_on_device(engine1):
    """k_movegen's lists against the structural order rules of python-chess (tests/move_order_rules.py: class order
    1-6, from / to descending, q r b n, king evasions first) -- a validator that generates no move itself -- over
    4,000 fuzzed positions and the children of 300 of them; the in-check flag comes from the kernel as well."""
    import move_order_rules
    rng = random.Random(77)
    cases = [position_fuzz.random_fen(rng)[0] for _ in range(4000)]
    recs = np.stack([B.record_from_fen(fen) for fen in cases])
    t = engine1.boards_to_device(recs)
    kids, _ = engine1.expand_frontier(t[:, :300].contiguous())
    recs = np.concatenate([recs, engine1.boards_to_host(kids)])
    t = engine1.boards_to_device(recs)
    mv, cn, fl = engine1.movegen(t)
    moves = mv.cpu().numpy().view(np.uint16)
    counts, flags = cn.cpu().numpy(), fl.cpu().numpy()
    seen = {"check": 0, "castle": 0, "promo": 0, "ep": 0}
    for i in range(recs.shape[0]):
        words = [int(m) for m in moves[i, :counts[i]]]
        bad = move_order_rules.violations(recs[i], words, bool(flags[i] & 1))
        assert not bad, (B.fen_from_record(recs[i]), bad)
        cls = [move_order_rules.classify(recs[i], w) for w in words]
        seen["check"] += int(flags[i] & 1)
        seen["castle"] += any(c[0] == 2 for c in cls)
        seen["promo"] += any(c[3] >= 0 for c in cls)
        seen["ep"] += any(c[0] == 6 for c in cls)
    assert min(seen.values()) >= 100, seen


@pytest.mark.parametrize("mode", ["5", "6"])
def test_perft_root_two_ply_pass(mode, monkeypatch):
    """The optional fused pass over the last two plies (CRL_PERFT_PAIR=5 / 6, read at engine creation): same totals as
    the default path for every depth and frontier target; the last-but-one ply is dealt to the lanes but never stored."""
    from chessrl_b200.engine import Engine
    monkeypatch.setenv("CRL_PERFT_PAIR", mode)
    e = Engine(max_games=1, max_nodes=8)
    for fen, exp in list(PERFT5.items()) + perft_kats.EDGE[:4] + perft_kats.EDGE[6:9]:
        rec = B.record_from_fen(fen)
        for depth, want in enumerate(exp, start=1):
            for min_frontier, bulk in ((1, True), (300, False), (1 << 20, True)):
                if (not bulk and want > 5_000_000) or (min_frontier == 1 and want > 200_000):
                    continue
                total, lanes, plies = e.perft_root(rec, depth, bulk=bulk, min_frontier=min_frontier)
                assert total == want, (fen, depth, min_frontier, bulk, total, want)
                if depth >= 2 and min_frontier == 1 << 20:
                    assert plies == depth - 2 and lanes == exp[depth - 2]
    e.close()

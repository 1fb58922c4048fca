"""GPU parity at BASELINE.json's FULL sizes (4,096 games x 200 simulations; 65,536 perft lanes) through
size-independent properties, plus sampled lanes against the oracle.  The deterministic evaluator keeps the runs short
and makes "identical network outputs" hold by construction."""
import functools
import random

import numpy as np
import pytest

import chessrl_oracle as O
from chessrl_b200 import boards as B
from chessrl_b200._lib import EVAL_HASH

pytestmark = pytest.mark.gpu

G, S = 4096, 200


@functools.lru_cache(maxsize=None)
def _games(n, seed):
    """n seeded random openings (0-40 plies); every 8th lane repeats lane 0 so identical games sit far apart."""
    rng = random.Random(seed)
    lines = []
    for i in range(n // 8):
        g = O.OGame()
        for _ in range(rng.randrange(0, 41)):
            if g.get_result() is not None:
                break
            legal = g.get_legal_moves()
            g.move(legal[rng.randrange(len(legal))])
        if g.get_result() is not None:
            g = O.OGame()
        lines.append([str(m) for m in g.board.move_stack])
    out = []
    for i in range(n):
        out.append(lines[0] if i % 8 == 0 else lines[(i * 7919) % len(lines)])
    return out


@pytest.mark.parametrize("inflight", [1, 6])
def test_full_size_search_properties(inflight):
    from chessrl_b200.engine import Engine
    lines = _games(G, 5)
    e = Engine(max_games=G, max_nodes=S + 1, avg_moves=64, max_inflight=inflight)
    e.set_evaluator(EVAL_HASH, 77, 24)
    e.games_set(np.tile(B.record_from_fen(), (G, 1)), [[B.uci_to_move(m) for m in ln] for ln in lines])
    c0 = e.counters()
    e.mcts_begin_move()
    e.mcts_simulate(S, inflight=inflight)
    st = e.root_stats()
    c1 = e.counters()
    assert c1["simulations"] - c0["simulations"] == G * S
    # every game: S simulations below a root that starts with one visit (mctree.py:111, 278-296)
    assert (st["root_visits"] == S + 1).all()
    assert (st["visits"].sum(axis=1) == S).all()
    # the root's value is the sum of what was backed up through its children (float64, different summation order)
    assert np.allclose(st["values"].sum(axis=1), st["root_values"], rtol=0, atol=1e-9)
    # children beyond n_children are empty, priors are probabilities, results are in {-1, 0, 1, none}
    k = st["n_children"]
    col = np.arange(B.MAX_MOVES)[None, :]
    assert (st["visits"][col >= k[:, None]] == 0).all()
    assert ((st["priors"] >= 0) & (st["priors"] <= 1)).all()
    assert np.isin(st["results"], [-1, 0, 1, B.RESULT_NONE]).all()
    # identical games in different lanes -> identical trees (no cross-lane state at full width)
    by_line = {}
    for g, ln in enumerate(lines):
        by_line.setdefault(tuple(ln), []).append(g)
    checked = 0
    for lanes in by_line.values():
        for g in lanes[1:]:
            assert (st["visits"][g] == st["visits"][lanes[0]]).all() and (st["values"][g] == st["values"][lanes[0]]).all()
            checked += 1
    assert checked >= G // 8
    # sampled lanes against the oracle's tree at the full simulation count
    rng = random.Random(1)
    for g in rng.sample(range(G), 6):
        og = O.OGame()
        for m in lines[g]:
            og.move(m)
        ot = O.OSelfPlayTree(og, threads=inflight)
        ot.search_move(O.OAgent(O.hash_evaluator(77, 24)), max_iters=S, noise=False)
        n = len(ot.root.children)
        assert int(k[g]) == n
        assert list(st["visits"][g, :n]) == [c.visits for c in ot.root.children], g
        assert [float(x) for x in st["values"][g, :n]] == [float(c.value) for c in ot.root.children], g
    e.close()


def test_full_size_perft_lanes():
    """65,536 lockstep lanes: every replicated lane reports the known count; bulk and plain leaf counting agree on a
    frontier of distinct boards, and the total is the published perft number."""
    from chessrl_b200.engine import Engine
    e = Engine(max_games=1, max_nodes=8)
    kiwi = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
    for fen, d3 in ((B.STARTING_FEN, 8902), (kiwi, 97862)):
        rep = e.boards_to_device(np.tile(B.record_from_fen(fen), (65536, 1)))
        assert bool((e.perft(rep, 3, bulk=True) == d3).all())
    fr = e.boards_to_device(B.record_from_fen(kiwi)[None, :])
    for _ in range(3):
        fr, _ = e.expand_frontier(fr)
    assert fr.shape[1] == 97862
    a, b = e.perft(fr, 2, bulk=True), e.perft(fr, 2, bulk=False)
    assert bool((a == b).all()) and int(a.sum().item()) == 193690690
    e.close()


def test_encode_full_batch_properties():
    """4,096-position batch: planes are 0/1, channel 127 is zero padding, each square has exactly one of
    {no black piece, black piece type} and {no white piece, white piece type} set in the current-position block."""
    import torch
    from chessrl_b200.engine import Engine
    e = Engine(max_games=1, max_nodes=8)
    fr = e.boards_to_device(B.record_from_fen()[None, :])
    for _ in range(3):
        fr, _ = e.expand_frontier(fr)
    boards = fr[:, :4096].contiguous()
    p = e.encode(boards).float()
    assert p.shape == (4096, 8, 8, 128)
    assert bool(((p == 0) | (p == 1)).all()) and float(p[..., 127].abs().sum()) == 0.0
    assert bool((p[..., 0:7].sum(-1) == 1).all()) and bool((p[..., 7:14].sum(-1) == 1).all())
    assert float(p[..., 14:126].abs().sum()) == 0.0                     # no history given: all-zero blocks
    turn = (boards[8] & 1).float()
    assert bool((p[..., 126] == turn[:, None, None]).all())
    e.close()


def test_full_size_evaluation_reuse_is_invisible():
    """BASELINE configs[3] scale: 4,096 lanes, real network, three lockstep moves with the most-visited child played --
    evaluation reuse on and off give identical root statistics for every lane, and with reuse on the second and third
    searches run a fraction of the evaluations."""
    import netpacks
    from chessrl_b200._lib import EVAL_NET
    from chessrl_b200.engine import Engine
    from chessrl_b200.lockstep import LockstepSelfPlay
    G, sims = 4096, 48
    e = Engine(max_games=G, max_nodes=sims + 1, avg_moves=120)
    e.load_weights(netpacks.lively_pack())
    e.set_evaluator(EVAL_NET)
    rng = random.Random(5)
    openings = [["e2e4", "e7e5"], ["d2d4", "d7d5"], ["g1f3", "g8f6"], ["c2c4", "e7e6", "b1c3", "d7d5"], []]
    mls = [[B.uci_to_move(m) for m in rng.choice(openings)] for _ in range(G)]
    recs = np.tile(B.record_from_fen(), (G, 1))

    def run(reuse):
        sp = LockstepSelfPlay(e, n_games=G, sims=sims, noise=False, reuse=reuse)
        sp.start(start_records=recs, move_lists=mls)
        out, evals = [], []
        for _ in range(3):
            c0 = e.counters()
            e.mcts_begin_move()
            e.mcts_simulate(sims, 1)
            st = e.root_stats(want=("visits", "values", "priors"))
            out.append(st)
            picks = np.where(st["n_children"] > 0, np.argmax(st["visits"], axis=1), -1).astype(np.int32)
            e.commit(picks, apply=True)
            sp._read_status()
            c1 = e.counters()
            evals.append((c1["evaluations"] - c0["evaluations"], c1["reused_evaluations"] - c0["reused_evaluations"]))
        return out, evals

    try:
        off, ev_off = run(False)
        on, ev_on = run(True)
        for a, b in zip(off, on):
            for k in a:
                assert np.array_equal(a[k], b[k]), k
        assert ev_on[0] == ev_off[0] and ev_on[0][1] == 0
        for (run_on, reused), (run_off, _) in zip(ev_on[1:], ev_off[1:]):
            assert run_on + reused == run_off and run_on < 0.7 * run_off, (ev_on, ev_off)
    finally:
        e.set_reuse(False)
        e.close()

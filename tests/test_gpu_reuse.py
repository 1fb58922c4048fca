"""Evaluation reuse across consecutive move searches (include/chessrl_b200.h crl_set_reuse; csrc/tree_core.cuh
"evaluation reuse"): with reuse on, an expansion whose node already exists in the previous move's tree takes the
opponent's reply, the value and the children's priors from there instead of running the two network evaluations.
The searches must stay bit-identical -- the reference builds every tree from scratch (agentdistributed.py:61-63,
mctree.py:104-111) and so does the engine; only evaluations are looked up -- so every test here runs the same
lockstep games twice, reuse off and on, and compares root statistics move for move."""
import numpy as np
import pytest

from chessrl_b200 import boards as B
from chessrl_b200._lib import EVAL_HASH, EVAL_NET

pytestmark = pytest.mark.gpu

WANT = ("visits", "values", "priors", "moves", "replies", "results")


def _play(engine, lanes, sims, n_moves, reuse, noise, seed, refill=True, picker=None):
    """n_moves lockstep moves; returns (per-move root statistics, finished games, counters delta)."""
    from chessrl_b200.lockstep import LockstepSelfPlay, pick_moves
    np.random.seed(seed)
    sp = LockstepSelfPlay(engine, n_games=lanes, sims=sims, noise=noise, refill=refill, reuse=reuse)
    colors = [(g % 3) != 0 for g in range(lanes)]
    sp.start(colors=colors)
    c0 = engine.counters()
    trace = []
    for mv in range(n_moves):
        sp.harvest(max_plies=400)
        if not sp.running().any():
            break
        engine.mcts_begin_move()
        engine.mcts_simulate(sims, 1)
        st = engine.root_stats(want=WANT)
        live = (sp._results == B.RESULT_NONE) & ~sp._retired
        assert (st["n_children"][:lanes][~live] == 0).all()      # a lane without a running game has no tree this move
        trace.append({k: st[k][:lanes].copy() for k in st})
        picks = np.full(engine.max_games, -1, dtype=np.int32)
        if picker is None:
            picks[:lanes] = pick_moves(st["visits"][:lanes], st["n_children"][:lanes], st["root_visits"][:lanes],
                                       sp._plies, live, noise)
        else:
            picks[:lanes] = picker(mv, st, live)
        engine.commit(picks, apply=True)
        sp._read_status()
    c1 = engine.counters()
    return trace, list(sp.finished), {k: c1[k] - c0[k] for k in c1}


def _same(ta, tb):
    assert len(ta) == len(tb)
    for mv, (a, b) in enumerate(zip(ta, tb)):
        for k in a:
            assert np.array_equal(a[k], b[k]), (mv, k)


def test_reuse_hash_evaluator_same_trees_fewer_evaluations():
    """64 lanes x 48 simulations x 40 moves with refill and both colours, Dirichlet noise on (the pick is often a barely
    visited child, as in the reference): identical root statistics and finished games, fewer evaluations."""
    from chessrl_b200.engine import Engine
    e = Engine(max_games=64, max_nodes=49, avg_moves=218)
    e.set_evaluator(EVAL_HASH, 7, 24)
    try:
        off = _play(e, 64, 48, 40, False, True, 123)
        on = _play(e, 64, 48, 40, True, True, 123)
        _same(off[0], on[0])
        assert len(off[1]) == len(on[1])
        for (ma, ra, ca), (mb, rb, cb) in zip(off[1], on[1]):
            assert list(ma) == list(mb) and ra == rb and ca == cb
        assert off[2]["simulations"] == on[2]["simulations"]
        assert off[2]["reused_evaluations"] == 0 and on[2]["reused_evaluations"] > 0
        # every evaluation is either run or taken from the previous tree
        assert on[2]["evaluations"] + on[2]["reused_evaluations"] == off[2]["evaluations"]
        # (with the reference's unscaled Dirichlet noise the pick is mostly a barely visited child: few twins -- measured
        # 3.7 % here; the next test follows the most-visited child)
        assert on[2]["evaluations"] < off[2]["evaluations"]
    finally:
        e.set_reuse(False)
        e.close()


def test_reuse_follows_the_most_visited_child_and_survives_foreign_moves():
    """The pick is always the most-visited child (the case reuse is made for: most of the new tree already exists), with
    every fifth move a move from OUTSIDE the tree played through crl_games_play_host -- which must unlink the lane."""
    from chessrl_b200.engine import Engine
    e = Engine(max_games=32, max_nodes=101, avg_moves=218)
    e.set_evaluator(EVAL_HASH, 3, 11)           # 11-bit policy: ties exercise the first-maximum rule on both paths

    def run(reuse):
        rng = np.random.default_rng(9)

        def picker(mv, st, live):
            p = np.where(live & (st["n_children"][:32] > 0), np.argmax(st["visits"][:32], axis=1), -1).astype(np.int32)
            return p

        from chessrl_b200.lockstep import LockstepSelfPlay
        sp = LockstepSelfPlay(e, n_games=32, sims=100, noise=False, reuse=reuse)
        sp.start()
        c0 = e.counters()
        trace = []
        for mv in range(14):
            if not sp.running().any():
                break
            e.mcts_begin_move()
            e.mcts_simulate(100, 1)
            st = e.root_stats(want=WANT)
            assert (st["n_children"][~sp.running()] == 0).all()  # a finished game has no tree this move
            trace.append({k: st[k].copy() for k in st})
            e.commit(picker(mv, st, sp.running()), apply=True)
            sp._read_status()
            if mv % 5 == 4:                     # two foreign plies for half of the lanes
                for _ in range(2):
                    legal, cnt = e.games_legal()
                    mvz = np.full(32, 0xFFFF, dtype=np.uint16)
                    for g in range(0, 32, 2):
                        if cnt[g] > 0 and sp._results[g] == B.RESULT_NONE:
                            mvz[g] = legal[g, int(rng.integers(cnt[g]))]
                    e.games_play(mvz)
                    sp._read_status()
        c1 = e.counters()
        return trace, {k: c1[k] - c0[k] for k in c1}

    try:
        off, on = run(False), run(True)
        _same(off[0], on[0])
        assert on[1]["evaluations"] + on[1]["reused_evaluations"] == off[1]["evaluations"]
        assert on[1]["evaluations"] < 0.6 * off[1]["evaluations"], (on[1], off[1])
    finally:
        e.set_reuse(False)
        e.close()


def test_reuse_real_network_same_trees():
    """The real network (history planes matter here: a twin is the same position reached by the same plies, so its
    input planes are the same bits): 48 lanes x 32 simulations x 12 moves, lively weights, noise on."""
    import netpacks
    from chessrl_b200.engine import Engine
    e = Engine(max_games=48, max_nodes=33, avg_moves=218)
    e.load_weights(netpacks.lively_pack())
    e.set_evaluator(EVAL_NET)
    try:
        off = _play(e, 48, 32, 12, False, True, 77, refill=False)
        on = _play(e, 48, 32, 12, True, True, 77, refill=False)
        _same(off[0], on[0])
        spread = np.concatenate([t["values"][t["visits"] > 0] / t["visits"][t["visits"] > 0] for t in off[0]])
        assert spread.max() - spread.min() > 0.1            # a live value head: Q really steers these searches
        assert on[2]["reused_evaluations"] > 0
        assert on[2]["evaluations"] + on[2]["reused_evaluations"] == off[2]["evaluations"]
    finally:
        e.set_reuse(False)
        e.close()


def test_reuse_is_not_applied_without_a_commit():
    """Searching the same positions again (commit with apply = 0, or no commit at all) must evaluate everything again:
    bench.py's per-move steps re-search unchanged positions and may not be served from the previous tree."""
    from chessrl_b200.engine import Engine
    e = Engine(max_games=16, max_nodes=41, avg_moves=218)
    e.set_evaluator(EVAL_HASH, 5, 24)
    e.set_reuse(True)
    try:
        e.games_set(np.tile(B.record_from_fen(), (16, 1)))
        counts = []
        for rep in range(3):
            c0 = e.counters()
            e.mcts_begin_move()
            e.mcts_simulate(40, 1)
            st = e.root_stats(want=("visits",))
            picks = np.argmax(st["visits"], axis=1).astype(np.int32)
            e.commit(picks, apply=False)
            c1 = e.counters()
            counts.append((c1["evaluations"] - c0["evaluations"], c1["reused_evaluations"] - c0["reused_evaluations"]))
        assert counts[0] == counts[1] == counts[2] and counts[0][1] == 0, counts
    finally:
        e.set_reuse(False)
        e.close()

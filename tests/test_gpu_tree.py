"""GPU parity, search: the lockstep PUCT kernels against golden vectors produced by the reference's own
mctree.py (threads=1) -- visit counts, value sums, priors, PUCT scores, replies and returned moves, bit-exact,
with the shared deterministic evaluator."""
import json
import os

import numpy as np
import pytest

import chessrl_oracle as O
from chessrl_b200 import boards as B
from chessrl_b200._lib import EVAL_HASH
from chessrl_b200.lockstep import LockstepSelfPlay, compute_policy

pytestmark = pytest.mark.gpu
chess = O.chess


def _check_case(c, st, g, out_moves):
    kids = c["children"]
    n = int(st["n_children"][g])
    assert n == len(kids), (c["name"], c["sims"], n, len(kids))
    assert int(st["root_visits"][g]) == c["root_visits"]
    assert float(st["root_values"][g]) == float.fromhex(c["root_value"])
    for k, kid in enumerate(kids):
        line = [B.move_to_uci(st["moves"][g, k])]
        if st["replies"][g, k] != B.MOVE_NONE:
            line.append(B.move_to_uci(st["replies"][g, k]))
        assert line == kid["line"], (c["name"], k)
        assert int(st["visits"][g, k]) == kid["visits"]
        assert float(st["values"][g, k]) == float.fromhex(kid["value"])
        r = int(st["results"][g, k])
        assert (None if r == B.RESULT_NONE else r) == kid["result"]
        if float.fromhex(kid["prior"]) != 1.0:
            assert float(st["priors"][g, k]) == float.fromhex(kid["prior"])
    pi = compute_policy(st["visits"][g, :n], st["root_visits"][g], len(c["moves"]), noise=False)
    assert [float(x) for x in pi] == [float.fromhex(x) for x in c["policy_no_noise"]]
    pick = int(np.argmax(pi))
    return pick


def test_mcts_golden_all_cases_in_lockstep(golden_dir):
    """Every golden case sharing (evaluator seed, sims) runs in its own lane of ONE lockstep engine."""
    from chessrl_b200.engine import Engine
    cases = json.load(open(os.path.join(golden_dir, "mcts_chess.json")))["cases"]
    groups = {}
    for c in cases:
        groups.setdefault((c["eval_seed"], c["policy_bits"], c["sims"]), []).append(c)
    for (seed, bits, sims), cs in groups.items():
        e = Engine(max_games=len(cs), max_nodes=sims + 1, avg_moves=96)
        e.set_evaluator(EVAL_HASH, seed, bits)
        recs = np.stack([B.record_from_fen(c["fen"] or B.STARTING_FEN) for c in cs])
        mls = [[B.uci_to_move(m) for m in c["moves"]] for c in cs]
        e.games_set(recs, mls)
        e.mcts_begin_move()
        e.mcts_simulate(sims)
        st = e.root_stats()
        picks = np.full(len(cs), -1, dtype=np.int32)
        for g, c in enumerate(cs):
            picks[g] = _check_case(c, st, g, None)
        out = e.commit(picks, apply=False)
        for g, c in enumerate(cs):
            assert [B.move_to_uci(out[g, 0]), B.move_to_uci(out[g, 1])] == c["returned"], (c["name"], sims)
        cnt = e.counters()
        assert cnt["simulations"] == sims * len(cs)
        e.close()


def test_mcts_wave_golden_all_cases_in_lockstep(golden_dir):
    """threads=K > 1 (wave schedule with virtual loss): bit-exact against the reference's own select / simulate /
    backprop driven in that schedule; cases sharing (evaluator, sims, K) run as lanes of one engine, so lanes whose
    waves get cut (narrow trees) run next to lanes that never are."""
    from chessrl_b200.engine import Engine
    cases = json.load(open(os.path.join(golden_dir, "mcts_wave.json")))["cases"]
    groups = {}
    for c in cases:
        groups.setdefault((c["eval_seed"], c["policy_bits"], c["sims"], c["threads"]), []).append(c)
    assert len(cases) >= 100
    for (seed, bits, sims, K), cs in groups.items():
        e = Engine(max_games=len(cs), max_nodes=sims + 1, avg_moves=96, max_inflight=K)
        e.set_evaluator(EVAL_HASH, seed, bits)
        recs = np.stack([B.record_from_fen(c["fen"] or B.STARTING_FEN) for c in cs])
        mls = [[B.uci_to_move(m) for m in c["moves"]] for c in cs]
        e.games_set(recs, mls)
        e.mcts_begin_move()
        e.mcts_simulate(sims, inflight=K)
        st = e.root_stats()
        picks = np.full(len(cs), -1, dtype=np.int32)
        for g, c in enumerate(cs):
            picks[g] = _check_case(c, st, g, None)
        out = e.commit(picks, apply=False)
        for g, c in enumerate(cs):
            assert [B.move_to_uci(out[g, 0]), B.move_to_uci(out[g, 1])] == c["returned"], (c["name"], sims, K)
        assert e.counters()["simulations"] == sims * len(cs)
        # the same engine still runs the exact schedule afterwards (scratch indexing switches back to K = 1)
        e.mcts_begin_move()
        e.mcts_simulate(min(sims, 30))
        assert (e.root_stats(want=("visits",))["root_visits"] == min(sims, 30) + 1).all()
        e.close()


def test_wave_mode_against_oracle_many_lanes():
    """64 different midgame positions x K = 6 (the reference's default thread count) against the oracle's wave
    schedule, all lanes in one engine."""
    import random
    from chessrl_b200.engine import Engine
    rng = random.Random(7)
    games = []
    for _ in range(64):
        g = O.OGame()
        for _ in range(rng.randrange(4, 40)):
            if g.get_result() is not None:
                break
            legal = g.get_legal_moves()
            g.move(legal[rng.randrange(len(legal))])
        if g.get_result() is None:
            games.append(g)
    K, sims = 6, 50
    e = Engine(max_games=len(games), max_nodes=sims + 1, avg_moves=96, max_inflight=K)
    e.set_evaluator(EVAL_HASH, 21, 24)
    e.games_set(np.tile(B.record_from_fen(), (len(games), 1)),
                [[B.uci_to_move(str(m)) for m in g.board.move_stack] for g in games])
    e.mcts_begin_move()
    e.mcts_simulate(sims, inflight=K)
    st = e.root_stats()
    for lane, g in enumerate(games):
        ot = O.OSelfPlayTree(g, threads=K)
        ot.search_move(O.OAgent(O.hash_evaluator(21, 24)), max_iters=sims, noise=False)
        n = len(ot.root.children)
        assert int(st["n_children"][lane]) == n
        assert list(st["visits"][lane, :n]) == [c.visits for c in ot.root.children], lane
        assert [float(x) for x in st["values"][lane, :n]] == [float(c.value) for c in ot.root.children], lane
    e.close()


def test_node_dump_matches_oracle_tree(engine1):
    g = O.OGame()
    for m in ["d2d4", "g8f6", "c2c4", "e7e6"]:
        g.move(m)
    ot = O.OSelfPlayTree(g)
    ot.search_move(O.OAgent(O.hash_evaluator(9, 24)), max_iters=150, noise=False)
    engine1.set_evaluator(EVAL_HASH, 9, 24)
    engine1.games_set(B.record_from_fen()[None, :], [[B.uci_to_move(m) for m in ["d2d4", "g8f6", "c2c4", "e7e6"]]])
    engine1.mcts_begin_move()
    engine1.mcts_simulate(150)
    nodes = engine1.node_dump(0)
    assert len(nodes) == 151
    # rebuild parent -> children (creation order) and compare the whole tree shape with the oracle's
    kids = {}
    for i, n in enumerate(nodes):
        if n.parent >= 0:
            kids.setdefault(n.parent, []).append((n.slot, i))

    def walk(onode, idx):
        n = nodes[idx]
        assert n.visits == onode.visits and n.value == float(onode.value)
        ch = [i for _, i in sorted(kids.get(idx, []))]
        assert len(ch) == len(onode.children)
        for oc, ci in zip(onode.children, ch):
            stack = [str(m) for m in oc.state.board.move_stack]
            assert B.move_to_uci(nodes[ci].move) == stack[len(onode.state.board.move_stack)]
            walk(oc, ci)
    walk(ot.root, 0)


def test_selfplay_golden_with_noise(golden_dir):
    """selfplay.play_game loop incl. the Dirichlet draw (numpy legacy RNG, KAT-7) and the black-opens case."""
    from chessrl_b200.engine import Engine
    for run in json.load(open(os.path.join(golden_dir, "selfplay.json")))["runs"]:
        e = Engine(max_games=1, max_nodes=run["sims"] + 1, avg_moves=96)
        e.set_evaluator(EVAL_HASH, run["eval_seed"], 24)
        sp = LockstepSelfPlay(e, sims=run["sims"], noise=True)
        sp.start(colors=[run["player_color"]])
        for k, (bm, am) in enumerate(run["picks"]):
            np.random.seed(run["noise_seed_base"] + k)
            out = sp.step()
            assert [B.move_to_uci(out[0, 0]), B.move_to_uci(out[0, 1])] == [bm, am], (k, run["player_color"])
        assert [B.move_to_uci(m) for m in e.game_moves(0)] == run["history"]["moves"]
        e.close()


def test_selfplay_slot_refill_plays_the_same_games():
    """play_games_lockstep with fewer lanes than games (finished lanes are refilled at once, parked at the end) returns,
    game for game, what one lane per game returns; parked lanes do no work."""
    from chessrl_b200 import selfplay
    kw = dict(sims=12, noise=False, seed=3, max_moves=70, evaluator=("hash", 9, 24))
    s_wide, s_narrow = {}, {}
    wide = selfplay.play_games_lockstep(None, 10, lanes=10, stats=s_wide, **kw)
    narrow = selfplay.play_games_lockstep(None, 10, lanes=3, stats=s_narrow, **kw)
    assert len(wide) == len(narrow) == 10
    for a, b in zip(wide.games, narrow.games):
        assert a.get_history()["moves"] == b.get_history()["moves"] and a.player_color == b.player_color
        assert a.get_result() == b.get_result()
        assert len(a) > 0
    assert {g.player_color for g in wide.games} == {True, False}                 # both openings are exercised
    assert s_narrow["moves"] == s_wide["moves"]                                  # the same agent moves were searched
    assert s_narrow["simulations"] == s_wide["simulations"]                      # parked / finished lanes did none
    assert s_narrow["steps"] > s_wide["steps"]


def test_lockstep_lanes_are_independent():
    """The same game in 64 lanes (plus different games in between) gives identical trees: no cross-lane state."""
    from chessrl_b200.engine import Engine
    e = Engine(max_games=64, max_nodes=81, avg_moves=96)
    e.set_evaluator(EVAL_HASH, 4, 24)
    recs = np.tile(B.record_from_fen(), (64, 1))
    mls = [[B.uci_to_move(m) for m in (["e2e4", "c7c5"] if g % 2 else ["d2d4", "d7d5", "c2c4"])] for g in range(64)]
    e.games_set(recs, mls)
    e.mcts_begin_move()
    e.mcts_simulate(80)
    st = e.root_stats()
    for g in range(2, 64):
        assert (st["visits"][g] == st["visits"][g % 2]).all() and (st["values"][g] == st["values"][g % 2]).all()
    e.close()


def test_search_from_unreachable_roots_matches_oracle():
    """Lockstep search rooted at position_fuzz positions (terminal children, promotions, ep and castling inside the
    tree, roots with up to ~90 legal moves), one position per lane: visit counts, value sums, results and
    (move, reply) lines against the restated SelfPlayTree with the shared evaluator."""
    import random
    import position_fuzz
    from chessrl_b200.engine import Engine
    rng = random.Random(22)
    for sims, seed, bits in ((33, 7, 24), (90, 401, 3), (150, 12, 11)):
        fens = []
        while len(fens) < 24:
            fen, _ = position_fuzz.random_fen(rng)
            if O.OGame(board=chess.Board(fen)).get_result() is None:
                fens.append(fen)
        e = Engine(max_games=len(fens), max_nodes=sims + 1, avg_moves=218)
        e.set_evaluator(EVAL_HASH, seed, bits)
        e.games_set(np.stack([B.record_from_fen(f) for f in fens]))
        e.mcts_begin_move()
        e.mcts_simulate(sims)
        st = e.root_stats()
        for g, fen in enumerate(fens):
            ot = O.OSelfPlayTree(O.OGame(board=chess.Board(fen)))
            ot.search_move(O.OAgent(O.hash_evaluator(seed, bits)), max_iters=sims, noise=False)
            kids = ot.root.children
            assert int(st["n_children"][g]) == len(kids) and int(st["root_visits"][g]) == ot.root.visits, fen
            assert float(st["root_values"][g]) == float(ot.root.value), fen
            for k, c in enumerate(kids):
                assert int(st["visits"][g, k]) == c.visits and float(st["values"][g, k]) == float(c.value), (fen, k)
                r = int(st["results"][g, k])
                assert (None if r == B.RESULT_NONE else r) == c.state.get_result(), (fen, k)
                line = [B.move_to_uci(st["moves"][g, k])]
                if st["replies"][g, k] != B.MOVE_NONE:
                    line.append(B.move_to_uci(st["replies"][g, k]))
                assert line == [str(x) for x in c.state.board.move_stack], (fen, k)
        e.close()


def _search_all(engine, recs, mls, sims, inflight=1):
    engine.games_set(recs, mls)
    engine.mcts_begin_move()
    engine.mcts_simulate(sims, inflight)
    st = engine.root_stats()
    dumps = [[(n.parent, n.slot, n.n_legal, n.n_children, n.result, n.move, n.reply, n.visits, float(n.value), float(n.prior))
              for n in engine.node_dump(g)] for g in (0, 1, len(recs) - 1)]
    return st, dumps


def test_level_parallel_backup_equals_the_parent_chase(monkeypatch):
    """The backup of a simulation runs at the head of the next k_select_expand, one lane per level of the path the select
    recorded; paths deeper than the recorded levels fall back to the parent-pointer chase.  CRL_PATH_CAP lowers the number
    of recorded levels: 32 (default), 2 (both forms inside one search) and 0 (chase only) must give the same trees."""
    from chessrl_b200.engine import Engine
    recs = np.tile(B.record_from_fen(), (24, 1))
    mls = [[B.uci_to_move(m) for m in (["e2e4", "e7e5", "g1f3"][:g % 4])] for g in range(24)]
    out = []
    for cap in ("32", "2", "0"):
        monkeypatch.setenv("CRL_PATH_CAP", cap)
        e = Engine(max_games=24, max_nodes=161, avg_moves=96)
        e.set_evaluator(EVAL_HASH, 17, 7)          # 7-bit priors, values in a narrow band: the search digs deep lines
        out.append(_search_all(e, recs, mls, 160))
        e.close()
    depth = max(len([1 for n in d if n[3] > 0]) for d in out[0][1])
    assert depth >= 3
    for st, dumps in out[1:]:
        for k in ("visits", "values", "priors", "moves", "replies", "results", "n_children", "root_visits", "root_values"):
            assert np.array_equal(st[k], out[0][0][k]), k
        assert dumps == out[0][1]


def test_wave_call_followed_by_exact_call_on_the_same_tree():
    """crl_mcts_simulate may be called several times on one tree, with different numbers of simulations in flight: a wave
    call leaves nothing in flight (every slot is finished by k_finalize_wave), so the exact-schedule call that follows
    starts clean, and its own last simulation is finished before the call returns (k_finalize)."""
    from chessrl_b200.engine import Engine
    e = Engine(max_games=16, max_nodes=200, avg_moves=96, max_inflight=8)
    e.set_evaluator(EVAL_HASH, 2, 24)
    e.games_set(np.tile(B.record_from_fen(), (16, 1)))
    e.mcts_begin_move()
    total = 0
    for sims, k in ((40, 8), (30, 1), (24, 4), (1, 1), (25, 1)):
        e.mcts_simulate(sims, k)
        total += sims
        st = e.root_stats(want=("visits",))
        assert (st["root_visits"] == 1 + total).all()
        assert (st["visits"].sum(axis=1) == total).all()
        nodes = e.node_dump(3)
        assert len(nodes) == 1 + total and sum(n.visits for n in nodes if n.parent == 0) == total
    c = e.counters()
    assert c["simulations"] == 16 * total
    e.close()


def test_row_bound_shrinks_the_launches_and_a_broken_promise_is_reported():
    """crl_mcts_set_row_bound: with 100 of 400 lanes running, a bound of 296 rows makes every launch cover 296 rows (and the
    tower take its single-tile path) -- same trees as without the bound, hash evaluator and real network; promising 50
    while 100 games run is reported, not evaluated silently wrong; the engine works again after the report."""
    import netpacks
    from chessrl_b200._lib import EVAL_NET, CrlError
    from chessrl_b200.engine import Engine
    e = Engine(max_games=400, max_nodes=41, avg_moves=96)
    recs = np.tile(B.record_from_fen(), (400, 1))
    mls = [[B.uci_to_move(m) for m in (["e2e4", "e7e5", "g1f3", "b8c6"][:g % 5])] for g in range(400)]
    active = np.zeros(400, dtype=np.uint8)
    active[::4] = 1                                            # 100 running games, scattered over the lanes

    def search(bound):
        e.games_set(recs, mls)
        e.games_set_active(active)
        e.set_row_bound(bound)
        e.mcts_begin_move()
        e.mcts_simulate(40)
        return e.root_stats()

    try:
        for kind in ("hash", "net"):
            if kind == "hash":
                e.set_evaluator(EVAL_HASH, 11, 24)
            else:
                e.load_weights(netpacks.lively_pack())
                e.set_evaluator(EVAL_NET)
            free, bounded = search(0), search(296)
            for k in free:
                assert np.array_equal(free[k], bounded[k]), (kind, k)
            assert (bounded["n_children"][::4] > 0).all() and (bounded["n_children"][1::4] == 0).all()
        with pytest.raises(CrlError):
            search(50)
        again = search(0)                                      # the error was reported once; the engine is usable again
        for k in again:
            assert np.array_equal(again[k], free[k]), k
    finally:
        e.set_row_bound(0)
        e.close()

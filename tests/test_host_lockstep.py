"""Host logic of the lockstep driver on the CPU: chessrl_b200.lockstep.LockstepSelfPlay / selfplay.LockstepRun and
benchmark.play_policy_games run on tests/fake_engine.FakeEngine (the oracle behind the Engine surface), so lane
bookkeeping -- cached status, batched harvest / refill / retire, game order, colours, caps -- is checked without a GPU.
The GPU suite runs the same drivers on the real engine (tests/test_gpu_tree.py, tests/test_gpu_api.py)."""
import random

import numpy as np

import chessrl_oracle as O
from chessrl_b200 import boards as B
from chessrl_b200 import selfplay
from fake_engine import FakeEngine


def _oracle_game(seed_agent, color, sims, max_agent_moves):
    agent = O.OAgent(O.hash_evaluator(*seed_agent))
    return O.play_game(agent, max_iters=sims, noise=False, player_color=color, max_agent_moves=max_agent_moves)


def test_lockstep_run_plays_every_game_like_the_reference_loop():
    """7 games on 3 lanes, refill as games are harvested, lanes parked at the end: every game's record equals the
    oracle's selfplay.play_game for the same colour, in game-start order, and each harvest / refill is ONE engine call."""
    sims, cap = 5, 3
    eng = FakeEngine(3, O.hash_evaluator(4, 24))
    run = selfplay.LockstepRun(None, 7, sims=sims, lanes=3, noise=False, seed=11, max_moves=cap, engine=eng)
    steps = 0
    while run.advance():
        steps += 1
        assert steps < 100
    st = run.stats()
    assert st["games_finished"] == 7 and st["refills"] == 4 and st["steps"] == steps
    rng = random.Random(11)
    colors = [rng.random() >= 0.5 for _ in range(7)]
    assert len({c for c in colors}) == 2
    for i, (moves, color) in enumerate(run.records):
        assert color == colors[i]
        want = _oracle_game((4, 24), color, sims, cap)
        assert [B.move_to_uci(m) for m in moves] == [m.uci() for m in want.board.move_stack], i
    # all lanes take the same number of steps here (every game is capped at `cap` agent moves): three rounds of games
    assert steps == 3 * cap
    # batched calls: one games_moves per harvest, one games_restart per refill, never one per lane
    assert eng.calls["games_moves"] == 3 and eng.calls["games_restart"] == 2
    assert eng.calls["games_set_active"] == 2            # two lanes parked together when game 6 starts alone, then its lane
    assert st["simulations"] == eng.n_sims == 7 * cap * sims
    assert abs(st["lane_occupancy"] - 7 / 9) < 1e-9


def test_lockstep_run_unbounded_supply_never_parks_a_lane():
    eng = FakeEngine(2, O.hash_evaluator(2, 24))
    run = selfplay.LockstepRun(None, None, sims=3, lanes=2, noise=False, seed=5, max_moves=2, engine=eng)
    for _ in range(7):
        assert run.advance()
    assert run.finished_games == run.refills == 6 and "games_set_active" not in eng.calls
    assert run.sp.running().all() and run.stats()["lane_occupancy"] == 1.0
    assert len([r for r in run.records if r is not None]) == 6


def test_lockstep_status_cache_follows_restart_and_commit():
    """The status read after each commit is what harvest(), running() and the next move pick use: no stale lane."""
    from chessrl_b200.lockstep import LockstepSelfPlay
    eng = FakeEngine(2, O.hash_evaluator(9, 24))
    sp = LockstepSelfPlay(eng, sims=3, noise=False)
    sp.start(colors=[True, False])
    assert list(sp._plies) == [0, 1] and sp.running().all()          # the black game got its opening reply
    sp.step()
    assert list(sp._plies) == [2, 3]
    got = sp.harvest(max_plies=2)                                     # both lanes reached the cap: unfinished games
    assert [(g, len(m), r) for g, m, r, _ in got] == [(0, 2, None), (1, 3, None)]
    sp.restart([1], [True])
    assert list(sp._plies) == [2, 0] and not sp._harvested[1] and sp._harvested[0]
    sp.retire([0])
    assert list(sp.running()) == [False, True] and list(eng.active) == [False, True]
    n = eng.calls["games_get"]
    sp.step()
    assert eng.calls["games_get"] == n + 1 and list(sp._plies) == [2, 2]   # one status read per step, lane 0 untouched


def test_policy_only_benchmark_loop_on_the_fake_engine(monkeypatch):
    """benchmark.play_policy_games against a seeded random mover: lanes are refilled until `games` are played, colours
    follow the seed, the agent's plies are its policy argmax, results are the oracle's."""
    from chessrl_b200 import benchmark, engine as engine_mod

    made = []

    def fake_engine(max_games=1, **kw):
        e = FakeEngine(max_games, O.hash_evaluator(6, 24))
        made.append(e)
        return e

    monkeypatch.setattr(engine_mod, "Engine", fake_engine)

    class M:
        weights = []

    monkeypatch.setattr(benchmark, "_as_model", lambda x: M() if x is not None else None)
    games = benchmark.play_policy_games(M(), opponent="random", games=5, lanes=2, seed=3, max_plies=12)
    assert len(made) == 1 and len(games) == 5 and all(g is not None for g in games)
    rng = random.Random(3)
    assert [g["color"] for g in games] == [rng.random() <= .5 for _ in range(5)]
    agent = O.OAgent(O.hash_evaluator(6, 24))
    for rec in games:
        og = O.OGame()
        for ply, m in enumerate(rec["moves"]):
            if (ply % 2 == 0) == rec["color"]:
                assert agent.best_move(og, real_game=True) == m
            assert og.move(m)
        assert rec["result"] == og.get_result() and (rec["result"] is not None or len(rec["moves"]) >= 12)

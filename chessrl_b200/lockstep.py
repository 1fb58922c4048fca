"""Lockstep self-play driver: selfplay.play_game (selfplay.py:59-84) for thousands of games at once.

Every running game does the same thing at the same time -- build a fresh tree (agentdistributed.py:61-63),
run `sims` simulations (mctree.py:173-176), pick the move from the root visit counts (mctree.py:178, 305-322),
play (our move, opponent reply) (selfplay.py:77-78) -- so one kernel sequence advances all of them.  The only
host work per move is what the reference also does on the host with numpy: the temperature / Dirichlet-noise
policy and its argmax (negligible, and bit-identical because it is the same numpy call).
"""

from __future__ import annotations

import numpy as np

from . import boards as B
from ._lib import EVAL_NET


def compute_policy(child_visits, root_visits, n_plies, noise=True):
    """SelfPlayTree.compute_policy (mctree.py:305-322) from the root statistics of one game."""
    if n_plies < 30:
        # tau = 1: np.power(v, 1.0) is float(v) exactly, so the vectorised division is bit-identical
        policy = np.asarray(child_visits, dtype=np.float64) / np.float64(int(root_visits))
    else:
        tau = n_plies / (1 + np.power(n_plies, 1.3))
        memo = {}                       # the same scalar np.power call as the reference, once per distinct count
        for v in child_visits:
            v = int(v)
            if v not in memo:
                memo[v] = np.power(v, 1 / tau)
        policy = np.array([memo[int(v)] for v in child_visits]) / np.power(int(root_visits), 1 / tau)
    if noise:
        epsilon = 0.25
        policy = (1 - epsilon) * policy + np.random.dirichlet([0.03] * len(child_visits))
    return policy


class LockstepSelfPlay:
    """Plays `n_games` games in lockstep on one Engine.

    colors: per-game player colour (True = the agent plays white).  When False the opponent opens with its
    policy-argmax move (selfplay.py:68-70).  noise=True draws Dirichlet noise from numpy's legacy global RNG in
    game-index order, once per game per move (the reference draws once per move of its single game).
    inflight: simulations in flight per game (the reference's `threads`; needs Engine(max_inflight >= inflight)).
    """

    def __init__(self, engine, n_games=None, sims=900, noise=True, refill=False, inflight=1):
        self.e = engine
        self.n = engine.max_games if n_games is None else n_games
        self.sims = sims
        self.inflight = int(inflight)
        self.noise = noise
        self.refill = refill
        self.colors = np.ones(self.n, dtype=bool)
        self.finished = []          # (moves u16[], result, player_color)
        self.moves_played = 0
        self._harvested = np.zeros(self.n, dtype=bool)

    def start(self, colors=None, start_records=None, move_lists=None):
        if colors is not None:
            self.colors = np.asarray(colors, dtype=bool)
        if start_records is None:
            start_records = np.tile(B.record_from_fen(), (self.n, 1))
        self.e.games_set(start_records, move_lists)
        if not self.colors.all():
            self.e.policy_move(mask=(~self.colors).astype(np.uint8))

    def step(self):
        """One agent move (+ reply) for every running game.  Returns the (our move, reply) words [n, 2]."""
        e = self.e
        e.mcts_begin_move()
        e.mcts_simulate(self.sims, self.inflight)
        st = e.root_stats(want=("visits",))
        _, plies, results = e.games_get(0, self.n)
        picks = np.full(e.max_games, -1, dtype=np.int32)
        for g in range(self.n):
            k = int(st["n_children"][g])
            if results[g] != B.RESULT_NONE or k == 0:
                continue
            pi = compute_policy(st["visits"][g, :k], st["root_visits"][g], int(plies[g]), self.noise)
            picks[g] = int(np.argmax(pi))
        out = e.commit(picks, apply=True)
        self.moves_played += int((picks >= 0).sum())
        return out[:self.n]

    def running(self):
        _, _, results = self.e.games_get(0, self.n)
        return results == B.RESULT_NONE

    def harvest(self):
        """Collects finished games (and restarts their lanes when refill=True)."""
        rec, plies, results = self.e.games_get(0, self.n)
        done = np.nonzero((results != B.RESULT_NONE) & ~self._harvested)[0]
        for g in done:
            self.finished.append((self.e.game_moves(int(g)), int(results[g]), bool(self.colors[g])))
            self._harvested[g] = True
        if self.refill and len(done):
            start = B.record_from_fen()
            mask = np.zeros(self.e.max_games, dtype=np.uint8)
            for g in done:
                self.e.games_set(start[None, :], None, first=int(g))
                self._harvested[g] = False
                mask[g] = 0 if self.colors[g] else 1
            if mask.any():
                self.e.policy_move(mask=mask)
        return len(done)

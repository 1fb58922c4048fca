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


SMALL_BATCH_ROWS = 296      # 74 CTA pairs x 4 boards: the largest batch the tower runs one tile per pair


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


_POW_TABLES = {}


def _tempered_table(n_plies, upto):
    """table[c] = np.power(c, 1 / tau) for c = 0..upto, every entry computed by the very scalar call compute_policy makes
    (a vectorised np.power may round differently from the scalar one), cached per ply count for the whole run."""
    n = int(n_plies)
    t = _POW_TABLES.get(n)
    if t is None or len(t) <= upto:
        if len(_POW_TABLES) > 4096:
            _POW_TABLES.clear()
        tau = n / (1 + np.power(n, 1.3))
        have = 0 if t is None else len(t)
        more = np.array([np.power(c, 1 / tau) for c in range(have, max(upto + 1, 2 * have, 64))], dtype=np.float64)
        t = _POW_TABLES[n] = more if t is None else np.concatenate([t, more])
    return t


def _tempered(n_plies, count):
    """np.power(count, 1 / tau) exactly as compute_policy evaluates it (same scalar call), memoised."""
    return _tempered_table(n_plies, int(count))[int(count)]


def pick_moves(child_visits, n_children, root_visits, n_plies, live, noise=True):
    """argmax(compute_policy(...)) for every live game of a lockstep batch in one pass -> int32 picks (-1 = skip).

    Bit-identical to calling compute_policy game by game in lane order, including the position of numpy's legacy
    global RNG afterwards: np.random.dirichlet([a] * k) draws k standard gammas one after the other, sums them in
    order and multiplies by the reciprocal, so one standard_gamma call for all live games, a per-row cumulative sum
    and one multiply reproduce every draw (tests/test_host_policy.py checks this against the per-game calls).
    """
    child_visits = np.asarray(child_visits)
    G = child_visits.shape[0]
    picks = np.full(G, -1, dtype=np.int32)
    k = np.where(np.asarray(live, dtype=bool), np.asarray(n_children, dtype=np.int64), 0)
    rows = np.nonzero(k > 0)[0]
    if rows.size == 0:
        return picks
    kr = k[rows]
    kmax = int(kr.max())
    mask = np.arange(kmax)[None, :] < kr[:, None]
    vis = child_visits[rows, :kmax]
    rv = np.asarray(root_visits)[rows]
    pl = np.asarray(n_plies)[rows]
    policy = np.zeros((rows.size, kmax), dtype=np.float64)
    early = pl < 30
    if early.any():                      # tau = 1
        policy[early] = vis[early].astype(np.float64) / rv[early].astype(np.float64)[:, None]
    late = np.nonzero(~early)[0]
    if late.size:                        # tau = n / (1 + n^1.3): the reference's scalar np.power results, from per-ply tables
        uniq, inv = np.unique(pl[late], return_inverse=True)
        top = int(max(vis[late].max(), rv[late].max()))
        tables = np.stack([_tempered_table(n, top)[:top + 1] for n in uniq])
        policy[late] = tables[inv[:, None], vis[late]] / tables[inv, rv[late]][:, None]
    if noise:
        g = np.random.standard_gamma(0.03, size=int(kr.sum()))
        pad = np.zeros((rows.size, kmax), dtype=np.float64)
        pad[mask] = g
        inv = 1 / np.cumsum(pad, axis=1)[:, -1]
        policy = (1 - 0.25) * policy + pad * inv[:, None]
    policy[~mask] = -np.inf
    # np.argmax returns the first maximum and the first NaN if there is one, row-wise as for the 1-D call
    picks[rows] = np.argmax(policy, axis=1).astype(np.int32)
    return picks


class LockstepSelfPlay:
    """Plays `n_games` games in lockstep on one Engine.

    colors: per-game player colour (True = the agent plays white).  When False the opponent opens with its
    policy-argmax move (selfplay.py:68-70).  noise=True draws Dirichlet noise from numpy's legacy global RNG in
    game-index order, once per game per move (the reference draws once per move of its single game).
    inflight: simulations in flight per game (the reference's `threads`; needs Engine(max_inflight >= inflight)).
    reuse: None leaves the engine as it is; True / False switches its evaluation reuse (Engine.set_reuse): a search
    takes the evaluations of nodes the previous move's search of the same game already ran -- same games, fewer
    network evaluations (exact schedule only; ignored for inflight > 1).

    Host round trips per move: root statistics, the commit, and one status read (plies + results) that harvest /
    running / the next step's move pick all reuse; harvest and refill are one batched call each whatever the number
    of lanes involved.
    """

    def __init__(self, engine, n_games=None, sims=900, noise=True, refill=False, inflight=1, reuse=None):
        self.e = engine
        if reuse is not None:
            engine.set_reuse(bool(reuse) and int(inflight) == 1)
        self.n = engine.max_games if n_games is None else n_games
        self.sims = sims
        self.inflight = int(inflight)
        self.noise = noise
        self.refill = refill
        self.colors = np.ones(self.n, dtype=bool)
        self.finished = []          # (moves u16[], result, player_color)
        self.moves_played = 0
        self._harvested = np.zeros(self.n, dtype=bool)
        self._retired = np.zeros(self.n, dtype=bool)
        self._plies = np.zeros(self.n, dtype=np.int32)
        self._results = np.full(self.n, B.RESULT_NONE, dtype=np.int8)

    def _read_status(self):
        _, plies, results = self.e.games_get(0, self.n)
        self._plies, self._results = plies, results

    def start(self, colors=None, start_records=None, move_lists=None):
        if colors is not None:
            self.colors = np.asarray(colors, dtype=bool)
        if start_records is None:
            start_records = np.tile(B.record_from_fen(), (self.n, 1))
        self.e.games_set(start_records, move_lists)
        if not self.colors.all():
            self.e.policy_move(mask=self._lane_mask(~self.colors))
        self._read_status()

    def _lane_mask(self, flags):
        mask = np.zeros(self.e.max_games, dtype=np.uint8)
        mask[:self.n] = np.asarray(flags, dtype=np.uint8)
        return mask

    def step(self):
        """One agent move (+ reply) for every running game.  Returns the (our move, reply) words [n, 2]."""
        e = self.e
        # the drain of a finite run: once few lanes still hold a game, tell the engine, so that its launches cover those
        # rows only and the evaluations take the tower's single-tile path (two regimes only: the simulation graph is
        # re-captured when the bound changes)
        n_run = int(self.running().sum())
        small = SMALL_BATCH_ROWS // max(1, self.inflight)
        e.set_row_bound(small if 0 < n_run <= small < e.max_games else 0)
        e.mcts_begin_move()
        e.mcts_simulate(self.sims, self.inflight)
        st = e.root_stats(want=("visits",))
        picks = np.full(e.max_games, -1, dtype=np.int32)
        picks[:self.n] = pick_moves(st["visits"][:self.n], st["n_children"][:self.n], st["root_visits"][:self.n],
                                    self._plies, (self._results == B.RESULT_NONE) & ~self._retired, self.noise)
        out = e.commit(picks, apply=True)
        self.moves_played += int((picks >= 0).sum())
        self._read_status()
        return out[:self.n]

    def running(self):
        return (self._results == B.RESULT_NONE) & ~self._retired

    def harvest(self, max_plies=None):
        """Finished games since the last call: list of (lane, moves u16[], result, player_color); also kept in
        self.finished as (moves, result, player_color).  max_plies: games that reached that many plies are
        harvested unfinished (result None), like a capped play_game run.  With refill=True the lanes are restarted."""
        over = self._results != B.RESULT_NONE
        if max_plies is not None:
            over = over | (self._plies >= max_plies)
        done = np.nonzero(over & ~self._harvested & ~self._retired)[0]
        out = []
        if len(done):
            lists = self.e.games_moves(done)                   # one round trip for all of them
            for g, moves in zip(done, lists):
                res = None if self._results[g] == B.RESULT_NONE else int(self._results[g])
                out.append((int(g), moves, res, bool(self.colors[g])))
                self._harvested[g] = True
        self.finished.extend((m, r, c) for _, m, r, c in out)
        if self.refill and len(done):
            self.restart(done, [self.colors[g] for g in done])
        return out

    def restart(self, lanes, colors):
        """Per-GPU slot refill: a new game from the start position in every listed lane (one batched
        call); where the agent plays black the opponent opens with its policy-argmax move (selfplay.py:68-70) -- one
        evaluation batch for all of them."""
        lanes = np.asarray([int(g) for g in lanes], dtype=np.int32)
        if lanes.size == 0:
            return
        self.e.games_restart(lanes)
        flags = np.zeros(self.n, dtype=bool)
        for i, g in enumerate(lanes):
            self.colors[g] = bool(colors[i])
            self._harvested[g] = False
            self._retired[g] = False
            flags[g] = not self.colors[g]
        if flags.any():
            self.e.policy_move(mask=self._lane_mask(flags))
        self._plies[lanes] = flags[lanes].astype(np.int32)     # 0 plies, or 1 after the opponent's opening move
        self._results[lanes] = B.RESULT_NONE

    def retire(self, lanes):
        """Lanes that stay empty from now on (no game left to start): step() skips them."""
        lanes = [int(g) for g in lanes]
        if not lanes:
            return
        self._retired[lanes] = True
        active = np.zeros(self.e.max_games, dtype=np.uint8)
        active[:self.n] = ~self._retired
        self.e.games_set_active(active, first=0)               # the whole mask in one call

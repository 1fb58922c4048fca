"""TEST-ONLY stand-in for chessrl_b200.engine.Engine built on the oracle (oracle/chessrl_oracle.py), so that the HOST logic
of the lockstep driver -- lane status cache, batched harvest / refill / retire, game accounting (chessrl_b200/lockstep.py,
chessrl_b200/selfplay.LockstepRun) and the policy-only evaluation loop (chessrl_b200/benchmark.py) -- runs in the GPU-less
CPU suite.  It implements exactly the Engine methods those modules call; every chess / search answer comes from the
oracle.  The product never imports it."""
import numpy as np

import chessrl_oracle as O
from chessrl_b200 import boards as B

chess = O.chess


class FakeEngine:
    def __init__(self, max_games, evaluator=None):
        self.max_games = int(max_games)
        self.agent = O.OAgent(evaluator or O.hash_evaluator(1, 24))
        self.games = [O.OGame() for _ in range(self.max_games)]
        self.active = np.ones(self.max_games, dtype=bool)
        self.trees = [None] * self.max_games
        self.n_sims = 0
        self.n_evals = 0
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    # ---- configuration (no-ops here) ------------------------------------------------------------------------
    def load_weights(self, tensors):
        pass

    def set_evaluator(self, kind, seed=0, policy_bits=24):
        pass

    def set_reuse(self, enable=True):
        pass

    def set_row_bound(self, max_running_games=0):
        pass

    def close(self):
        pass

    def counters(self):
        return {"simulations": self.n_sims, "evaluations": self.agent.n_evals, "launches": 0}

    @staticmethod
    def pack_move_lists(move_lists):
        from chessrl_b200.engine import Engine
        return Engine.pack_move_lists(move_lists)

    # ---- games ----------------------------------------------------------------------------------------------
    def _running(self, g):
        return self.active[g] and self.games[g].get_result() is None

    def games_set(self, start_records, move_lists=None, first=0):
        self._count("games_set")
        rec = np.asarray(start_records, dtype=np.uint64).reshape(-1, 9)
        if isinstance(move_lists, tuple):
            move_lists = [[int(m) for m in row[:c]] for row, c in zip(*move_lists)]
        for i in range(rec.shape[0]):
            g = O.OGame(board=chess.Board(B.fen_from_record(rec[i])))
            for m in (move_lists[i] if move_lists else []):
                g.move(B.move_to_uci(m))
            self.games[first + i] = g
            self.active[first + i] = True

    def games_restart(self, lanes, start_record=None):
        self._count("games_restart")
        for g in np.asarray(lanes, dtype=np.int64):
            self.games[g] = O.OGame()
            self.active[g] = True

    def games_get(self, first=0, n=None):
        self._count("games_get")
        n = self.max_games - first if n is None else n
        rec = np.zeros((n, 9), dtype=np.uint64)
        plies = np.zeros(n, dtype=np.int32)
        res = np.zeros(n, dtype=np.int8)
        for i in range(n):
            g = self.games[first + i]
            rec[i] = B.record_from_fen(g.board.fen())
            plies[i] = len(g.board.move_stack)
            r = g.get_result()
            res[i] = B.RESULT_NONE if r is None else r
        return rec, plies, res

    def games_set_active(self, active, first=0):
        self._count("games_set_active")
        a = np.asarray(active, dtype=bool)
        self.active[first:first + len(a)] = a

    def games_moves(self, lanes, cap=2048):
        self._count("games_moves")
        return [np.array([B.uci_to_move(m.uci()) for m in self.games[g].board.move_stack], dtype=np.uint16)
                for g in np.asarray(lanes, dtype=np.int64)]

    def game_moves(self, game):
        return self.games_moves([game])[0]

    def games_legal(self, first=0, n=None):
        n = self.max_games - first if n is None else n
        legal = np.zeros((n, B.MAX_MOVES), dtype=np.uint16)
        cnt = np.zeros(n, dtype=np.int32)
        for i in range(n):
            ms = self.games[first + i].get_legal_moves()
            cnt[i] = len(ms)
            legal[i, :len(ms)] = [B.uci_to_move(m) for m in ms]
        return legal, cnt

    def games_play(self, moves):
        self._count("games_play")
        acc = np.zeros(self.max_games, dtype=bool)
        for g, m in enumerate(np.asarray(moves, dtype=np.uint16)):
            if self.active[g] and m != B.MOVE_NONE:
                acc[g] = self.games[g].move(B.move_to_uci(m))
        return acc

    def policy_move(self, mask=None):
        self._count("policy_move")
        picks = np.full(self.max_games, B.MOVE_NONE, dtype=np.uint16)
        for g in range(self.max_games):
            if self._running(g) and (mask is None or mask[g]):
                mv = self.agent.best_move(self.games[g], real_game=True)
                picks[g] = B.uci_to_move(mv)
                self.games[g].move(mv)
        return picks

    # ---- search ---------------------------------------------------------------------------------------------
    def mcts_begin_move(self):
        self.trees = [O.OSelfPlayTree(self.games[g]) if self._running(g) else None for g in range(self.max_games)]

    def mcts_simulate(self, n_sims, inflight=1):
        for t in self.trees:
            if t is not None:
                for _ in range(n_sims):
                    t.explore_tree(self.agent)
                    self.n_sims += 1

    def root_stats(self, want=("visits",)):
        G = self.max_games
        out = {"visits": np.zeros((G, B.MAX_MOVES), dtype=np.int32), "n_children": np.zeros(G, dtype=np.int32),
               "root_visits": np.zeros(G, dtype=np.int32), "root_values": np.zeros(G, dtype=np.float64)}
        for g, t in enumerate(self.trees):
            if t is not None:
                kids = t.root.children
                out["n_children"][g] = len(kids)
                out["visits"][g, :len(kids)] = [c.visits for c in kids]
                out["root_visits"][g] = t.root.visits
        return out

    def commit(self, picks, apply=True):
        out = np.full((self.max_games, 2), B.MOVE_NONE, dtype=np.uint16)
        for g, t in enumerate(self.trees):
            k = int(picks[g])
            if t is None or k < 0 or k >= len(t.root.children):
                continue
            stack = t.root.children[k].state.board.move_stack
            if len(stack) >= 2:
                pair = (str(stack[-2]), str(stack[-1]))
                out[g] = [B.uci_to_move(pair[0]), B.uci_to_move(pair[1])]
                if apply:
                    self.games[g].move(pair[0])          # Game.move rejects an illegal first move (selfplay.py:77-78)
                    self.games[g].move(pair[1])
        return out

"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path
(`chessrl_b200/`).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

CPU restatement of the ChessRL self-play hot path, one function
per reference symbol, each citing the reference file:line it follows.  It runs
on the python-chess-0.28.3 restatement in oracle/pychess_compat/chess.

Parity status: the reference has no tests and python-chess / TensorFlow are
not installable here, so python-chess semantics are pinned by public perft
tables and the README move order (tests/test_oracle_chess.py), and everything
ABOVE python-chess (planes, labels, tree) is pinned by golden vectors generated
from the reference's own unmodified game.py / netencoder.py / mctree.py running
on the same shim (tests/golden/make_golden.py; oracle/ref_on_shims.py).
"""

from __future__ import annotations

import hashlib
import os
import sys
from datetime import datetime

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_COMPAT = os.path.join(_HERE, "pychess_compat")
if _COMPAT not in sys.path:
    sys.path.insert(0, _COMPAT)

import chess  # noqa: E402  (the restatement, not the PyPI package)

NULL_MOVE = "00000"          # game.py:13 (five zeros: simply "never legal")
N_LABELS = 1968
N_PLANES = 127
HISTORY_T = 8                # netencoder.py:47
PUCT_C = 10                  # mctree.py:79
VIRTUAL_LOSS = 1             # mctree.py:12
UCI_LABELS_SHA256 = "e67a413cdbce60252cbcf4714d6e4b88549ecaee5e5a613ee4cdb8d5098d8f7b"


# --------------------------------------------------------------------------
# game.py
# --------------------------------------------------------------------------

class OGame:
    """game.Game restated (game.py:11-112), minus plot_board."""

    NULL_MOVE = NULL_MOVE
    WHITE = True
    BLACK = False

    def __init__(self, board=None, player_color=True, date=None):
        self.board = chess.Board() if board is None else board           # game.py:17-21
        self.player_color = player_color
        self.date = date or datetime.now().strftime("%d/%m/%Y %H:%M:%S")  # game.py:24-26

    def get_legal_moves(self):                                            # game.py:43-57
        return [m.uci() for m in self.board.generate_legal_moves()]

    def move(self, uci):                                                  # game.py:28-41
        if uci not in self.get_legal_moves():
            return False
        self.board.push(chess.Move.from_uci(uci))
        return True

    def get_result(self):                                                 # game.py:92-109
        if self.board.can_claim_fifty_moves():
            return 0
        if self.board.is_game_over():
            r = self.board.result()
            return 1 if r == "1-0" else (-1 if r == "0-1" else 0)
        return None

    def get_copy(self):                                                   # game.py:79-80
        return OGame(board=self.board.copy())

    def get_history(self):                                                # game.py:59-66
        return {"moves": [m.uci() for m in self.board.move_stack], "result": self.get_result(),
                "player_color": self.player_color, "date": self.date}

    def get_fen(self):                                                    # game.py:68-69
        return self.board.board_fen()

    @property
    def turn(self):                                                       # game.py:74-77
        return self.board.turn

    def __len__(self):                                                    # game.py:111-112
        return len(self.board.move_stack)


# --------------------------------------------------------------------------
# netencoder.py
# --------------------------------------------------------------------------

def _colour_block(board, colour):
    """netencoder._get_pieces_one_hot (netencoder.py:13-30): [empty, P, N, B, R, Q, K] for one colour,
    row 0 = rank 8, column 0 = file a."""
    block = np.zeros((8, 8, 7))
    for pt in range(1, 7):
        bb = board.pieces_mask(pt, colour)
        while bb:
            low = bb & -bb
            sq = low.bit_length() - 1
            block[7 - (sq >> 3), sq & 7, pt] = 1.0
            bb ^= low
    block[:, :, 0] = 1.0 - block[:, :, 1:].sum(axis=-1)
    return block


def _position_planes(board):
    """netencoder._get_current_game_state (netencoder.py:33-44): black block first, then white."""
    return np.concatenate((_colour_block(board, False), _colour_block(board, True)), axis=-1)


def planes(game, flipped=False):
    """netencoder.get_game_state (netencoder.py:72-91) -> float64[8,8,127]."""
    out = np.zeros((8, 8, N_PLANES))
    out[:, :, 0:14] = _position_planes(game.board)
    walker = game.board.copy()                                            # netencoder.py:58
    for i in range(HISTORY_T):                                            # netencoder.py:61-67
        if not walker.move_stack:
            break
        walker.pop()
        out[:, :, 14 * (i + 1):14 * (i + 2)] = _position_planes(walker)
    out[:, :, 126] = 1.0 if game.turn else 0.0                            # netencoder.py:86
    if flipped:                                                           # netencoder.py:89-90
        out = out[::-1, ::-1, :].copy()
    return out


def uci_labels():
    """netencoder.get_uci_labels (netencoder.py:94-134) -> 1968 labels."""
    files, ranks = "abcdefgh", "12345678"
    knight = ((-2, -1), (-1, -2), (-2, 1), (1, -2), (2, -1), (-1, 2), (2, 1), (1, 2))
    labels = []
    for f in range(8):
        for r in range(8):
            dests = [(t, r) for t in range(8)]
            dests += [(f, t) for t in range(8)]
            dests += [(f + t, r + t) for t in range(-7, 8)]
            dests += [(f + t, r - t) for t in range(-7, 8)]
            dests += [(f + a, r + b) for a, b in knight]
            for f2, r2 in dests:
                if (f2, r2) != (f, r) and 0 <= f2 < 8 and 0 <= r2 < 8:
                    labels.append(files[f] + ranks[r] + files[f2] + ranks[r2])
    for f in range(8):
        for p in "qrbn":
            for df in (0, -1, 1):
                if 0 <= f + df < 8:
                    labels.append(files[f] + "2" + files[f + df] + "1" + p)
                    labels.append(files[f] + "7" + files[f + df] + "8" + p)
    return labels


_LABELS = None
_LABEL_INDEX = None


def label_index():
    global _LABELS, _LABEL_INDEX
    if _LABEL_INDEX is None:
        _LABELS = uci_labels()
        _LABEL_INDEX = {u: i for i, u in enumerate(_LABELS)}
    return _LABEL_INDEX


def labels_sha256():
    return hashlib.sha256("\n".join(uci_labels()).encode()).hexdigest()


# --------------------------------------------------------------------------
# agentdistributed.py (policy-only move, masked gather) with an injected evaluator
# --------------------------------------------------------------------------

class OAgent:
    """AgentDistributed restated (agentdistributed.py:39-111) with the socket replaced by a callable
    `evaluate(game) -> (policy float32[1968], value)`; counts evaluations."""

    def __init__(self, evaluate, color=True):
        self.evaluate = evaluate
        self.color = color
        self.n_evals = 0

    def _eval(self, game):
        self.n_evals += 1
        return self.evaluate(game)

    def predict_policy(self, game, mask_legal_moves=True):                # agentdistributed.py:75-83
        policy = self._eval(game)[0]
        if mask_legal_moves:
            idx = label_index()
            policy = [policy[idx[m]] for m in game.get_legal_moves()]
        return policy

    def predict_outcome(self, game):                                      # agentdistributed.py:70-73
        return float(self._eval(game)[1])                                 # predict_worker.py:111 float(v)

    def policy_move(self, game):                                          # agentdistributed.py:56-58
        policy = self.predict_policy(game)
        return game.get_legal_moves()[int(np.argmax(policy))]

    def best_move(self, game, real_game=False, max_iters=900, ai_move=True, noise=True):
        if real_game:
            return self.policy_move(game)
        if game.get_result() is None:                                     # agentdistributed.py:60-66
            return OSelfPlayTree(game).search_move(self, max_iters=max_iters, ai_move=ai_move, noise=noise)
        return NULL_MOVE


# --------------------------------------------------------------------------
# mctree.py.  threads=1: one in-flight simulation (deterministic in the reference).
# threads=K>1: the reference is schedule-dependent; the oracle runs the WAVE schedule --
# up to K selects one after the other (each adds its virtual loss), then their simulates, then their backprops
# in the same order -- which is one legal interleaving of ThreadPoolExecutor(max_workers=K) (mctree.py:173-176).
# A select that would enter a node created earlier in the same wave by an expansion that needed the opponent's
# reply is deferred to the next wave (legal too: "that worker had not started yet"); the lockstep engine needs
# this because the reply comes out of the wave's own evaluation batch.
# --------------------------------------------------------------------------

class ONode:
    """mctree.Node (mctree.py:15-95)."""

    def __init__(self, state, parent=None):
        self.state = state
        self.children = []
        self.unexpanded_actions = state.get_legal_moves()                 # mctree.py:31
        self.parent = parent
        self.value = 0
        self.visits = 0
        self.prior = 1
        self.vloss = 0
        self.pending = False       # wave schedule only: created in the current wave with an opponent reply
        self._result_known = False
        self._result = None

    @property
    def is_terminal_state(self):                                          # mctree.py:47-49 (pure function of state)
        if not self._result_known:
            self._result = self.state.get_result()
            self._result_known = True
        return self._result is not None

    def get_value(self):                                                  # mctree.py:71-87
        if self.parent is None:
            score = 99999999999
        else:
            n_sub = np.sum([c.visits for c in self.children])
            score = (self.value / (1 + self.visits)) + \
                PUCT_C * self.prior * (np.sqrt(n_sub) / (1 + self.visits))
        return score - self.vloss

    def get_best_child(self):                                             # mctree.py:89-95 (first maximum)
        return self.children[int(np.argmax([c.get_value() for c in self.children]))]


class OSelfPlayTree:
    """mctree.SelfPlayTree (mctree.py:148-322) restated: threads=1 (deterministic) or the wave schedule."""

    def __init__(self, root, threads=1):
        self.root = root if isinstance(root, ONode) else ONode(root.get_copy())  # mctree.py:106-109
        self.root.visits = 1                                                      # mctree.py:111
        self.num_threads = threads                                                # mctree.py:157
        self.n_waves = 0

    def search_move(self, agent, max_iters=200, noise=True, ai_move=False):       # mctree.py:159-198
        if self.num_threads <= 1:
            for _ in range(max_iters):
                self.explore_tree(agent)
        else:
            left = max_iters
            while left > 0:
                left -= self.explore_wave(agent, min(self.num_threads, left))
        pick = int(np.argmax(self.compute_policy(self.root, noise=noise)))
        stack = self.root.children[pick].state.board.move_stack
        ours = str(stack[-2]) if len(stack) >= 2 else NULL_MOVE
        reply = str(stack[-1]) if len(stack) >= 2 else NULL_MOVE                  # IndexError -> both stay null
        return (ours, reply) if ai_move else ours

    def explore_tree(self, agent):                                                # mctree.py:200-214
        leaf = self.select(agent)
        v = self.simulate(leaf, agent)
        self.backprop(leaf, v)

    def explore_wave(self, agent, k):
        """Up to k explore_tree calls interleaved as select*, simulate*, backprop* (see the header comment)."""
        leaves = []
        for j in range(k):
            if j and self._next_select_meets_pending():
                break
            leaves.append(self.select(agent))
        values = [self.simulate(leaf, agent) for leaf in leaves]
        for leaf, v in zip(leaves, values):
            leaf.pending = False
            self.backprop(leaf, v)
        self.n_waves += 1
        return len(leaves)

    def _next_select_meets_pending(self):
        node = self.root                                                          # a dry run of select: pure
        while not node.is_terminal_state:
            if node.unexpanded_actions:
                return False
            node = node.get_best_child()
            if node.pending:
                return True
        return False

    def select(self, agent):                                                      # mctree.py:216-229
        node = self.root
        while not node.is_terminal_state:
            if node.unexpanded_actions:
                node = self.expand(node, agent)
                break
            node = node.get_best_child()
        node.vloss += VIRTUAL_LOSS
        return node

    def expand(self, node, agent):                                                # mctree.py:231-257
        state = node.state.get_copy()
        state.move(node.unexpanded_actions.pop())                                 # last legal move first
        replied = state.get_result() is None
        if replied:
            state.move(agent.policy_move(state))                                  # opponent = policy argmax
        child = ONode(state, parent=node)
        child.pending = replied and self.num_threads > 1
        node.children.append(child)
        if not node.unexpanded_actions:                                           # mctree.py:254-255, 298-303
            for p, c in zip(agent.predict_policy(node.state), reversed(node.children)):
                c.prior = p
        return child

    def simulate(self, node, agent):                                              # mctree.py:259-276
        r = node.state.get_result()
        return agent.predict_outcome(node.state) if r is None else r

    def backprop(self, leaf, value):                                              # mctree.py:278-296
        leaf.vloss -= VIRTUAL_LOSS
        n = leaf
        while n is not None:
            n.visits += 1
            n.value += value
            n = n.parent

    def compute_policy(self, node, noise=True):                                   # mctree.py:305-322
        n = len(node.state.board.move_stack)
        tau = 1 if n < 30 else n / (1 + np.power(n, 1.3))
        pi = np.array([np.power(c.visits, 1 / tau) for c in node.children]) / np.power(node.visits, 1 / tau)
        if noise:
            pi = (1 - 0.25) * pi + np.random.dirichlet([0.03] * len(node.children))
        return pi


# --------------------------------------------------------------------------
# selfplay.py game loop
# --------------------------------------------------------------------------

def play_game(agent, max_iters=900, noise=True, player_color=None, max_agent_moves=None, rng=None):
    """selfplay.play_game (selfplay.py:59-84).  `player_color=None` draws it like the reference."""
    import random
    if player_color is None:
        player_color = (rng or random).random() >= 0.5
    game = OGame(player_color=player_color)
    agent.color = player_color
    if player_color is False:
        game.move(agent.best_move(game, real_game=True))
    n = 0
    while game.get_result() is None:
        bm, am = agent.best_move(game, real_game=False, ai_move=True, max_iters=max_iters, noise=noise)
        game.move(bm)
        game.move(am)
        n += 1
        if max_agent_moves is not None and n >= max_agent_moves:
            break
    return game


# --------------------------------------------------------------------------
# deterministic evaluator shared by the oracle and the CUDA engine (SURVEY.md 8c:
# "injected deterministic evaluator ... identical on CPU and GPU")
# --------------------------------------------------------------------------

_M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def board_words(board):
    """The 9 words of the device board record (include/chessrl_b200.h crl board layout)."""
    cr = board.clean_castling_rights()
    c4 = (1 if cr & chess.BB_H1 else 0) | (2 if cr & chess.BB_A1 else 0) | \
         (4 if cr & chess.BB_H8 else 0) | (8 if cr & chess.BB_A8 else 0)
    ep = 0 if board.ep_square is None else board.ep_square + 1
    meta = (1 if board.turn else 0) | (c4 << 1) | (ep << 5)
    return [board.pawns, board.knights, board.bishops, board.rooks, board.queens, board.kings,
            board.occupied_co[True], board.occupied_co[False], meta]


def position_hash(board, seed=0):
    h = seed & _M64
    for w in board_words(board):
        h = splitmix64(h ^ w)
    return h


def hash_evaluator(seed=0, policy_bits=24):
    """evaluate(game) -> (float32[1968], float32 value); exact in fp32 so CPU and GPU agree bitwise.
    `policy_bits` < 24 quantises the policy to force ties (first-maximum semantics are then exercised)."""
    idx = np.arange(N_LABELS, dtype=np.uint64)
    mul = np.uint64(0xD6E8FEB86659FD93)

    def _sm(x):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))

    def evaluate(game):
        h = position_hash(game.board, seed)
        with np.errstate(over="ignore"):
            r = _sm(np.uint64(h) + idx * mul)
        q = (r >> np.uint64(64 - policy_bits)).astype(np.float32)
        policy = q * np.float32(2.0 ** -policy_bits)
        v = np.float32(splitmix64(h ^ 0xA5A5A5A5A5A5A5A5) >> 40) * np.float32(2.0 ** -23) - np.float32(1.0)
        return policy, np.float32(v)

    return evaluate

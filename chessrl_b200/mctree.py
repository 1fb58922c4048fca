"""mctree: drop-in for the reference's Monte Carlo tree classes (mctree.py:12-322).

The tree of one game lives in the engine's flat node / edge pools on the GPU; Node objects here are read-only
VIEWS built from crl_mcts_node_dump_host.  SelfPlayTree.search_move runs the reference's algorithm -- select,
expand (our move + the opponent's policy-argmax reply), simulate (network value or game result), backprop -- for
`max_iters` simulations, then picks the move on the host exactly as mctree.py:178-198 / 305-322 do (temperature,
optional Dirichlet noise from numpy's global RNG, first argmax).
`threads` = simulations in flight: 1 is the reference's deterministic order; with K > 1 the reference is
schedule-dependent and the engine runs one of its legal schedules deterministically -- waves of up to
K selects (each adding its virtual loss), one evaluation batch, K backprops in the same order
(include/chessrl_b200.h crl_mcts_simulate).
"""

from __future__ import annotations

import numpy as np

from . import boards as B
from . import runtime
from .game import Game
from .lockstep import compute_policy as _compute_policy

VIRTUAL_LOSS = 1


class Node(object):
    """View of one tree node: state, children (creation order), unexpanded_actions, parent, value, visits,
    prior, vloss (mctree.py:15-95)."""

    def __init__(self, state, parent=None):
        self.state = state
        self.children = []
        self.unexpanded_actions = state.get_legal_moves()
        self.parent = parent
        self.value = 0
        self.visits = 0
        self.prior = 1
        self.vloss = 0

    @property
    def is_leaf(self):
        return len(self.children) == 0

    @property
    def is_fully_expanded(self):
        return len(self.unexpanded_actions) == 0

    @property
    def is_terminal_state(self):
        return self.state.get_result() is not None

    @property
    def is_root(self):
        return self.parent is None

    def get_value(self):
        """Q + U as the engine's select kernel computes it (mctree.py:71-87)."""
        if self.is_root:
            return 99999999999 - self.vloss
        n_sub = np.sum([c.visits for c in self.children])
        q = self.value / (1 + self.visits)
        u = 10 * self.prior * (np.sqrt(n_sub) / (1 + self.visits))
        return q + u - self.vloss

    def get_best_child(self):
        return self.children[int(np.argmax([c.get_value() for c in self.children]))]


class Tree(object):
    """Base tree: root = Node over a copy of the game, root.visits = 1 (mctree.py:98-146)."""

    def __init__(self, root):
        self.root = root if type(root) is Node else Node(root.get_copy())
        self.root.visits = 1

    def search_move(self, agent, max_iters=200, verbose=False, noise=True, ai_move=False):
        pass


class SelfPlayTree(Tree):

    def __init__(self, root, threads=6):
        super().__init__(root)
        self.num_threads = threads
        self._engine = None

    def search_move(self, agent, max_iters=200, verbose=False, noise=True, ai_move=False):
        game = self.root.state
        k = max(1, min(int(self.num_threads), 64))
        eng = runtime.scalar_engine(min_nodes=max_iters + 1, min_inflight=k)
        self._engine = eng
        agent._bind_evaluator(eng)
        eng.games_set(game._start[None, :], [[B.uci_to_move(m) for m in game._moves]])
        eng.mcts_begin_move()
        eng.mcts_simulate(max_iters, inflight=k)
        # the pick and the commit come BEFORE anything else touches the engine's game lane: building Node views
        # replays positions through Game.move on the same one-lane engine, which overwrites lane 0
        st = eng.root_stats(want=("visits",))
        n_kids = int(st["n_children"][0])
        moves = (Game.NULL_MOVE, Game.NULL_MOVE)
        if n_kids:
            policy = _compute_policy(st["visits"][0, :n_kids], int(st["root_visits"][0]), len(game._moves), noise)
            picks = np.full(eng.max_games, -1, dtype=np.int32)
            picks[0] = int(np.argmax(policy))
            out = eng.commit(picks, apply=False)
            moves = (B.move_to_uci(out[0, 0]), B.move_to_uci(out[0, 1]))
        self._attach_views(eng.node_dump(0))
        return moves if ai_move else moves[0]

    def compute_policy(self, node, noise=True):
        """Visit-count policy of `node` with temperature and optional Dirichlet noise (mctree.py:305-322)."""
        n_plies = len(node.state.board.move_stack)
        return _compute_policy([c.visits for c in node.children], node.visits, n_plies, noise)

    # ---- views --------------------------------------------------------------------------------------------
    def _attach_views(self, dump):
        """Node views from one flat dump of the device tree.  Statistics are filled in at once; a view's `state`
        (a Game, i.e. a replay on the device) and `unexpanded_actions` are built on first access, so a search costs
        no per-node replays unless a caller walks the tree."""
        views = []
        for n in dump:
            if n.parent < 0:
                v = self.root
                v.children = []
                legal = v.state.get_legal_moves()
                v.unexpanded_actions = legal[:len(legal) - n.n_children]
            else:
                parent = views[n.parent]
                v = _NodeView(parent, n.move, n.reply, n.n_children)
                parent.children.append(v)
            v.visits = int(n.visits)
            v.value = float(n.value)
            v.prior = np.float32(n.prior) if n.parent >= 0 and n.prior != 1.0 else 1
            views.append(v)


class _NodeView(Node):
    """A non-root Node whose Game is replayed only when somebody asks for it."""

    def __init__(self, parent, move, reply, n_children):
        self.children = []
        self.parent = parent
        self.value = 0
        self.visits = 0
        self.prior = 1
        self.vloss = 0
        self._line = (int(move), int(reply), int(n_children))
        self._state = None
        self._unexpanded = None

    @property
    def state(self):
        if self._state is None:
            move, reply, _ = self._line
            st = self.parent.state.get_copy()
            st.move(B.move_to_uci(move))
            if reply != B.MOVE_NONE:
                st.move(B.move_to_uci(reply))
            self._state = st
        return self._state

    @property
    def unexpanded_actions(self):
        if self._unexpanded is None:
            legal = self.state.get_legal_moves()
            self._unexpanded = legal[:len(legal) - self._line[2]]
        return self._unexpanded

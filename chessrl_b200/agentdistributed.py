"""AgentDistributed: drop-in for the reference's socket-based agent (agentdistributed.py:13-111).

Same surface -- AgentDistributed(color, endpoint=None, num_threads=6), best_move, predict_outcome, predict_policy,
predict, get_copy, connect, disconnect -- but a prediction is a call into the CUDA engine instead of a pickled
Game sent to a PredictWorker over TCP.  `endpoint` selects the weights a PredictWorker registered for it
(predict_worker.py); connect / disconnect are no-ops.
"""

from __future__ import annotations

import numpy as np

from . import mctree
from . import netencoder
from . import runtime
from ._lib import EVAL_HASH, EVAL_NET
from .player import Player

_ENDPOINT_MODELS = {}       # endpoint -> ChessModel registered by PredictWorker


def register_endpoint(endpoint, model):
    _ENDPOINT_MODELS[endpoint] = model


class AgentDistributed(Player):

    def __init__(self, color, endpoint=None, num_threads=6, model=None):
        super().__init__(color)
        self.move_encodings = netencoder.get_uci_labels()
        self.uci_dict = {u: i for i, u in enumerate(self.move_encodings)}
        self.conn = None
        self.pool_conns = None
        self.address = endpoint
        self.num_threads = num_threads
        self._model = model
        self._hash_eval = None          # (seed, policy_bits): deterministic test evaluator instead of the network

    # ---- evaluator selection ------------------------------------------------------------------------------
    def use_hash_evaluator(self, seed, policy_bits=24):
        """Tests only: replaces the network by the deterministic position-hash evaluator."""
        self._hash_eval = (int(seed), int(policy_bits))

    def _model_obj(self):
        if self.address is not None and self.address in _ENDPOINT_MODELS:
            # re-resolved on every use: PredictWorker.reload_model may have registered a new model for the endpoint
            self._model = _ENDPOINT_MODELS[self.address]
        if self._model is None:
            from .model import ChessModel
            self._model = ChessModel()
        return self._model

    def _bind_evaluator(self, eng):
        if self._hash_eval is not None:
            eng.set_evaluator(EVAL_HASH, *self._hash_eval)
        else:
            runtime.ensure_weights(eng, self._model_obj())
            eng.set_evaluator(EVAL_NET)

    # ---- reference surface --------------------------------------------------------------------------------
    def best_move(self, game, real_game=False, max_iters=900, ai_move=True, verbose=False):
        best_move = '00000'
        if real_game:
            policy = self.predict_policy(game)
            best_move = game.get_legal_moves()[int(np.argmax(policy))]
        elif game.get_result() is None:
            tree = mctree.SelfPlayTree(game, threads=self.num_threads)
            best_move = tree.search_move(self, max_iters=max_iters, verbose=verbose, ai_move=ai_move)
        return best_move

    def predict_outcome(self, game) -> float:
        return self.predict(game)[1]

    def predict_policy(self, game, mask_legal_moves=True):
        policy = self.predict(game)[0]
        if mask_legal_moves:
            policy = [policy[self.uci_dict[m]] for m in game.get_legal_moves()]
        return policy

    def predict(self, game):
        """(policy float32[1968], value float) of the position, like PredictWorker's reply (predict_worker.py:110-111)."""
        import torch
        eng = runtime.scalar_engine()
        self._bind_evaluator(eng)
        recs = game.history_records()
        boards = eng.boards_to_device(np.asarray(recs[0], dtype=np.uint64)[None, :])
        if self._hash_eval is not None:
            p, v = eng.hash_eval(boards, *self._hash_eval)
        else:
            hist = np.zeros((8, 8, 1), dtype=np.uint64)
            for i, r in enumerate(recs[1:9]):
                hist[i, :, 0] = r[:8]
            planes = eng.encode(boards, torch.from_numpy(hist.view(np.int64)).to(eng.device),
                                torch.tensor([len(recs) - 1], dtype=torch.uint8, device=eng.device))
            p, v = eng.net_forward(planes)
        return p[0].cpu().numpy(), float(v[0].item())

    def get_copy(self):
        c = AgentDistributed(self.color, endpoint=self.address, num_threads=self.num_threads, model=self._model)
        c._hash_eval = self._hash_eval
        return c

    def connect(self):
        self.conn = True

    def disconnect(self):
        self.conn = None

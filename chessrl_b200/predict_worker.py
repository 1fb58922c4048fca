"""PredictWorker: kept as a thin shim of the reference's NN server (predict_worker.py:13-128).  In the reference it
batches pickled Game objects arriving over TCP into one model.predict; here evaluation batches are formed on the GPU
by the lockstep engine, so the worker only owns the model and registers it for its endpoint."""

from __future__ import annotations

from . import agentdistributed
from .model import ChessModel


class PredictWorker():

    def __init__(self, model_path=None, endpoint=('localhost', 9999)):
        self.model = ChessModel(weights=model_path if model_path else None)
        self.address = endpoint
        agentdistributed.register_endpoint(endpoint, self.model)

    def start(self):
        agentdistributed.register_endpoint(self.address, self.model)

    def stop(self):
        pass

    def reload_model(self, model_path):
        self.model = ChessModel(weights=model_path)
        agentdistributed.register_endpoint(self.address, self.model)

"""Small lockstep self-play workload for compute-sanitizer: evaluation reuse on, the fused backup / select head, the
statistics-only softmax, refill and both colours; hash evaluator and the real network."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

from chessrl_b200 import model, selfplay  # noqa: E402

np.random.seed(1)
stats = {}
data = selfplay.play_games_lockstep(None, 12, sims=20, lanes=6, noise=True, seed=2, max_moves=8, evaluator=("hash", 3, 24),
                                    stats=stats, reuse=True)
print("hash evaluator:", len(data), "games", stats["simulations"], "simulations", stats["evaluations"], "evaluations",
      stats["reused_evaluations"], "reused", flush=True)
m = model.ChessModel()
stats = {}
data = selfplay.play_games_lockstep(m, 6, sims=10, lanes=4, noise=False, seed=2, max_moves=4, stats=stats, reuse=True)
print("network:", len(data), "games", stats["simulations"], "simulations", stats["evaluations"], "evaluations",
      stats["reused_evaluations"], "reused", flush=True)
assert stats["reused_evaluations"] > 0
print("reuse sanitize workload ok")

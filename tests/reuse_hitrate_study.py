"""Study (not a test; run by hand in the build container): how many network evaluations of a move search were already
run by the previous move's search of the same game, in the reference's own schedule.  Plays one game with the oracle's
restated SelfPlayTree (threads=1, Dirichlet noise as in the reference) on the torch-CPU fp32 network and counts, per move,
the evaluations whose position -- identified by its full move stack, i.e. including the history the planes encode -- was
evaluated during the previous search.  Measured here (random-init net, 200 simulations, 30 moves): per-move fractions
are bimodal -- ~0.8 when the game follows the most-visited child, ~0 when the (unscaled, nearly one-hot) Dirichlet
noise picks a barely visited one -- mean 0.245.  This is what chessrl_b200's evaluation reuse exploits.

    python tests/reuse_hitrate_study.py init 200 30        # init | lively, simulations per move, agent moves
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import chessrl_oracle as O  # noqa: E402
import model_torch as M  # noqa: E402
import netpacks  # noqa: E402
from chessrl_b200 import model  # noqa: E402


def main(which="init", sims=200, n_moves=30):
    torch.set_num_threads(os.cpu_count() or 1)
    pack = model.random_pack(0) if which == "init" else netpacks.lively_pack()
    state = {"prev": set(), "cur": set(), "evals": 0, "hits": 0}
    cache = {}

    def evaluate(game):
        key = tuple(str(m) for m in game.board.move_stack)
        state["evals"] += 1
        state["hits"] += key in state["prev"]
        state["cur"].add(key)
        if key not in cache:
            x = np.zeros((1, 8, 8, 128), np.float32)
            x[0, ..., :127] = O.planes(game)
            with torch.no_grad():
                p, v = M.forward(pack, torch.from_numpy(x))
            cache[key] = (p[0].numpy(), np.float32(v.reshape(-1)[0].item()))
        return cache[key]

    agent = O.OAgent(evaluate)
    game = O.OGame(player_color=True)
    np.random.seed(0)
    t0, fracs = time.time(), []
    for mv in range(n_moves):
        if game.get_result() is not None:
            break
        state.update(prev=state["cur"], cur=set(), evals=0, hits=0)
        tree = O.OSelfPlayTree(game)
        bm, am = tree.search_move(agent, max_iters=sims, ai_move=True, noise=True)
        top = sorted((c.visits for c in tree.root.children), reverse=True)[:3]
        fracs.append(state["hits"] / max(1, state["evals"]))
        print("move %2d %s %s: %d evaluations, %d seen by the previous search (%.2f); top visits %s; %.0f s"
              % (mv, bm, am, state["evals"], state["hits"], fracs[-1], top, time.time() - t0), flush=True)
        game.move(bm)
        game.move(am)
    print("mean fraction %.3f over %d moves" % (float(np.mean(fracs)), len(fracs)))


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0] if a else "init", int(a[1]) if len(a) > 1 else 200, int(a[2]) if len(a) > 2 else 30)

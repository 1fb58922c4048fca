"""Deep perft: how large should the breadth-first frontier be before the per-lane depth-first walk?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chessrl_b200 import boards as B
from chessrl_b200.engine import Engine

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
e = Engine(max_games=1, max_nodes=8)
for name, fen, depth, want in (("start_d7", B.STARTING_FEN, 7, 3195901860), ("kiwipete_d6", KIWI, 6, 8031647685)):
    for min_boards in (65536, 1 << 20, 1 << 24, 1 << 27):
        for bulk in (True, False):
            best = None
            for rep in range(2):
                torch.cuda.synchronize()
                a, m, b = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record()
                f = e.boards_to_device(B.record_from_fen(fen)[None, :])
                d = 0
                while f.shape[1] < min_boards and d < depth - 1:
                    f, _ = e.expand_frontier(f)
                    d += 1
                m.record()
                nodes = e.perft(f, depth - d, bulk=bulk)
                b.record()
                torch.cuda.synchronize()
                assert int(nodes.sum().item()) == want
                t = (a.elapsed_time(b), m.elapsed_time(b))
                best = t if best is None or t[0] < best[0] else best
            print("%s frontier>=%d: %d lanes x perft(%d) %s: total %.2f ms (walk %.2f ms) -> %.1f G nodes/s" %
                  (name, min_boards, f.shape[1], depth - d, "bulk" if bulk else "every leaf made", best[0], best[1], want / best[0] / 1e6))
            del f, nodes
e.close()

"""crl_perft_root_host: how far should the device-side breadth-first expansion go before the per-lane walk?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chessrl_b200 import boards as B
from chessrl_b200.engine import Engine

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
deep = "--deep" in sys.argv
e = Engine(max_games=1, max_nodes=8)
cases = [(B.STARTING_FEN, 5, 4865609), (KIWI, 5, 193690690)]
if deep:
    cases += [(B.STARTING_FEN, 7, 3195901860), (KIWI, 6, 8031647685)]
for fen, depth, want in cases:
    for mf in (65536, 1 << 20) + ((1 << 26,) if depth > 5 else ()):
        for bulk in (True, False):
            best = None
            for r in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                t, lanes, plies = e.perft_root(B.record_from_fen(fen), depth, bulk=bulk, min_frontier=mf)
                b.record()
                torch.cuda.synchronize()
                assert t == want, (t, want)
                if r:
                    best = min(best or 1e9, a.elapsed_time(b))
            print("%s d%d min_frontier %9d %-6s lanes %9d bfs plies %d: %8.3f ms %7.1f G nodes/s" %
                  (fen[:8], depth, mf, "bulk" if bulk else "nobulk", lanes, plies, best, want / best / 1e6))
e.close()

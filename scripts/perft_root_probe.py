import sys, torch
sys.path.insert(0, "/root/repo")
from chessrl_b200 import boards as B
from chessrl_b200.engine import Engine
KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
e = Engine(max_games=1, max_nodes=8)
for fen, want in ((B.STARTING_FEN, 4865609), (KIWI, 193690690)):
    for mf in (65536, 1 << 18, 1 << 20, 1 << 22):
        for bulk in (True, False):
            best = None
            for r in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                t, lanes, plies = e.perft_root(B.record_from_fen(fen), 5, bulk=bulk, min_frontier=mf)
                b.record(); torch.cuda.synchronize()
                assert t == want
                if r: best = min(best or 1e9, a.elapsed_time(b))
            print(fen[:8], mf, "bulk" if bulk else "nobulk", lanes, plies, "%.3f ms %.1f G/s" % (best, want / best / 1e6))

"""Rule-kernel probe: movegen / make / perft over 65,536 lockstep boards (BASELINE config 2), for ncu and timing."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chessrl_b200 import boards as B
from chessrl_b200.engine import Engine

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
e = Engine(max_games=1, max_nodes=8)
timing = "--time" in sys.argv


def frontier(fen, min_boards):
    f = e.boards_to_device(B.record_from_fen(fen)[None, :])
    d = 0
    while f.shape[1] < min_boards:
        f, _ = e.expand_frontier(f)
        d += 1
    return f, d


def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best, out


for name, fen, want5 in (("start", B.STARTING_FEN, 4865609), ("kiwipete", KIWI, 193690690)):
    f, d = frontier(fen, 65536)
    n = f.shape[1]
    ms_mg, (mv, cn, fl) = timed(lambda: e.movegen(f))
    tot_moves = int(cn.sum())
    ms_b, nodes = timed(lambda: e.perft(f, 5 - d, True))
    ms_nb, nodes2 = timed(lambda: e.perft(f, 5 - d, False))
    assert int(nodes.sum()) == want5 and int(nodes2.sum()) == want5
    pick = mv[:, 0].contiguous()
    g = f.clone()
    ms_mk, _ = timed(lambda: e.make_moves(g, pick), reps=1)
    if timing:
        bytes_mg = n * 72 + tot_moves * 2 + n * 5
        print("%s: %d lanes at depth %d" % (name, n, d))
        print("  movegen  %.3f ms  %.1f M boards/s  %.1f M moves/s  algorithmic %.1f GB/s" %
              (ms_mg, n / ms_mg / 1e3, tot_moves / ms_mg / 1e3, bytes_mg / ms_mg / 1e6))
        print("  make     %.3f ms  %.1f M moves/s  algorithmic %.1f GB/s" % (ms_mk, n / ms_mk / 1e3, n * 146 / ms_mk / 1e6))
        print("  perft(%d) per lane, leaf bulk counting: %.3f ms -> %.2f G nodes/s" % (5 - d, ms_b, want5 / ms_b / 1e6))
        print("  perft(%d) per lane, every leaf made    : %.3f ms -> %.2f G nodes/s" % (5 - d, ms_nb, want5 / ms_nb / 1e6))
# replicated variant: 65,536 copies of each root run perft(3)
for name, fen, want3 in (("start", B.STARTING_FEN, 8902), ("kiwipete", KIWI, 97862)):
    t = e.boards_to_device(np.tile(B.record_from_fen(fen), (65536, 1)))
    ms_b, nodes = timed(lambda: e.perft(t, 3, True), reps=3)
    ms_nb, nodes2 = timed(lambda: e.perft(t, 3, False), reps=3)
    assert bool((nodes == want3).all()) and bool((nodes2 == want3).all())
    if timing:
        print("%s x65536 perft(3): bulk %.3f ms -> %.2f G nodes/s ; no bulk %.3f ms -> %.2f G nodes/s" %
              (name, ms_b, 65536 * want3 / ms_b / 1e6, ms_nb, 65536 * want3 / ms_nb / 1e6))
e.close()

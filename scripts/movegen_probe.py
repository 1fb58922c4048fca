"""k_movegen / k_make alone on ~1.08 M midgame boards (Kiwipete's depth-3 frontier x 11), L2 flushed; plus the rule tests' perft."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chessrl_b200 import boards as B
from chessrl_b200.engine import Engine

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
e = Engine(max_games=1, max_nodes=8)
fr = e.boards_to_device(B.record_from_fen(KIWI)[None, :])
for _ in range(3):
    fr, _ = e.expand_frontier(fr)
boards = fr.repeat(1, 11).contiguous()
n = boards.shape[1]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    mv, cn, fl = e.movegen(boards)
tot = 0.0
for _ in range(10):
    flush.fill_(1); flush.max()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); mv, cn, fl = e.movegen(boards); b.record(); torch.cuda.synchronize()
    tot += a.elapsed_time(b)
ms = tot / 10
print("movegen: %d boards, %.1f moves avg, %.1f us -> %.2f G boards/s, %.1f G moves/s" % (n, cn.float().mean().item(), ms * 1e3, n / ms / 1e6, cn.sum().item() / ms / 1e6))
e.close()

"""Timing probe: does k_trunk4 sustain its rate when every K step is issued as two M=256 x N=128 MMAs (filter halves)
instead of one N=256 MMA?  (CRL_T4_NSPLIT_PROBE=1: the accumulator columns come out permuted, results are wrong -- timing
only.)  This is the precondition of sharing one weight stage between the two tiles of a CTA pair's group."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chessrl_b200 import model
from chessrl_b200.engine import Engine

pack = model.random_pack(0)
torch.manual_seed(0)
planes = (torch.rand(4096, 8, 8, 128, device="cuda") < 0.15).to(torch.bfloat16)
planes[..., 127] = 0
for rnd in range(3):
    for probe in (0, 1):
        os.environ["CRL_T4_NSPLIT_PROBE"] = str(probe)
        e = Engine(max_games=4096, max_nodes=4)
        e.load_weights(pack)
        for _ in range(10):
            e.net_forward(planes)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(300):
            e.net_forward(planes)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 300
        print("%s: %.4f ms per 4096-position evaluation -> %.0f TFLOP/s" %
              ("two N=128 MMAs per K step" if probe else "one N=256 MMA per K step ", ms, 4096 * 1548038656 / ms / 1e9), flush=True)
        e.close()
os.environ["CRL_T4_NSPLIT_PROBE"] = "0"

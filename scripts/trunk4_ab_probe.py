"""A/B of library builds on ONE box: time 4,096-position evaluations back to back for the package found under argv[1]."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(sys.argv[1]))
from chessrl_b200 import model
from chessrl_b200.engine import Engine

pack = model.random_pack(0)
torch.manual_seed(0)
for n in (4096, 296, 4):
    planes = (torch.rand(n, 8, 8, 128, device="cuda") < 0.15).to(torch.bfloat16)
    planes[..., 127] = 0
    e = Engine(max_games=n, max_nodes=4)
    e.load_weights(pack)
    for _ in range(10):
        e.net_forward(planes)
    torch.cuda.synchronize()
    reps = 300 if n == 4096 else 1000
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        e.net_forward(planes)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%s n=%d: %.4f ms per evaluation -> %.0f TFLOP/s" % (sys.argv[1], n, ms, n * 1548038656 / ms / 1e9), flush=True)
    e.close()

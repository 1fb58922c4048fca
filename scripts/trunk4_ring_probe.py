"""k_trunk4 ring sizes (padded-image slots / weight stages): time 4,096-position evaluations back to back."""
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
names = {0: "3/7", 1: "3/8", 2: "2/9", 3: "2/10"}
ref = None
for rnd in range(2):
    for ring in (0, 1, 2, 3):
        os.environ["CRL_T4_RING"] = str(ring)
        e = Engine(max_games=4096, max_nodes=4)
        e.load_weights(pack)
        for _ in range(10):
            p, v = e.net_forward(planes)
        torch.cuda.synchronize()
        if ref is None:
            ref = p.clone()
        assert torch.equal(p, ref), "ring size changed the result"
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            e.net_forward(planes)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 200
        print("ring %s: %.4f ms per 4096-position evaluation -> %.0f TFLOP/s" % (names[ring], ms, 4096 * 1548038656 / ms / 1e9))
        e.close()

"""k_trunk4 (padded-image A operand reused across the nine taps) against k_trunk (v3): outputs and timing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import model_torch
from chessrl_b200 import model
from chessrl_b200.engine import Engine

pack = model.random_pack(0)
torch.manual_seed(0)
res = {}
for n in (1, 3, 8, 37, 600, 4096):
    planes = (torch.rand(n, 8, 8, 128, device="cuda") < 0.15).to(torch.bfloat16)
    planes[..., 127] = 0
    out = {}
    for v3 in ("1", "0"):
        os.environ["CRL_TRUNK_V3"] = v3
        e = Engine(max_games=max(n, 2), max_nodes=4)
        e.load_weights(pack)
        p, v = e.net_forward(planes)
        torch.cuda.synchronize()
        out[v3] = (p.clone(), v.clone())
        if n == 4096:
            for _ in range(5):
                e.net_forward(planes)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(60):
                e.net_forward(planes)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 60
            print("v%s: %.3f ms per 4096-position evaluation -> %.0f TFLOP/s" % ("3" if v3 == "1" else "4", ms, 4096 * 1548038656 / ms / 1e9))
        e.close()
    dp = (out["1"][0] - out["0"][0]).abs().max().item()
    dv = (out["1"][1] - out["0"][1]).abs().max().item()
    print("n=%d  v4 vs v3: max|dpolicy| %.3e  max|dvalue| %.3e  policy sums %.5f" % (n, dp, dv, out["0"][0].sum(1).mean().item()))
    if n <= 600:
        rp, rv = model_torch.forward(pack, planes[..., :127].float().cpu().numpy(), device="cuda")
        for name, key in (("v3", "1"), ("v4", "0")):
            ep = (out[key][0].cpu() - rp.cpu()).abs().max().item()
            ev = (out[key][1].cpu() - rv.cpu().reshape(-1)).abs()
            print("      %s vs torch fp32: max|dpolicy| %.3e  max|dvalue| %.3e  mean|dvalue| %.3e" % (name, ep, ev.max().item(), ev.mean().item()))
    assert dp < 2e-3 and dv < 4e-2, "v4 differs from v3"
print("trunk4 ok")

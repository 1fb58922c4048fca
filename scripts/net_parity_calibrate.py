#!/usr/bin/env python
"""Prints the numbers the network parity tests bound (tests/test_gpu_net.py): for the lively test pack, the emulated
bf16 graph's own distance from fp32 (E) and the tower kernel's distance from the emulated graph (d), per tap.
Usage on a GPU box:  python scripts/net_parity_calibrate.py > gpurun_out/net_parity_calibrate.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import netpacks  # noqa: E402
import test_gpu_net as T  # noqa: E402
from chessrl_b200.engine import Engine  # noqa: E402

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
e = Engine(max_games=4096, max_nodes=2)
pack = netpacks.lively_pack()
e.load_weights(pack)
for n in (1, 5, 37, 592, 593, 4096):
    x = T._planes_for(n)
    xt = torch.from_numpy(x).to(e.device).to(torch.bfloat16)
    got = e.debug_tower(xt, layer=20)
    f32 = T._reference_taps(pack, x, False)
    emu = T._reference_taps(pack, x, True)
    row = []
    for name, k, r32, rem in (("act20", got["act"], f32["act"][20], emu["act"][20]), ("pf", got["pf"], f32["pf"], emu["pf"]),
                              ("vf", got["vf"], f32["vf"], emu["vf"]), ("logits", got["logits"], f32["logits"], emu["logits"])):
        row.append("%s E=%.3e d=%.3e Em=%.3e dm=%.3e" % (name, T._err(rem, r32), T._err(k, rem), T._err_mean(rem, r32),
                                                         T._err_mean(k, rem)))
    Ep = (emu["policy"] - f32["policy"]).abs().max().item()
    Ev = (emu["value"] - f32["value"]).abs().max().item()
    dp = (got["policy"] - f32["policy"]).abs().max().item()
    dv = (got["value"] - f32["value"]).abs().max().item()
    v = f32["value"]
    print("n=%d | %s | policy Ep=%.3e dp=%.3e value Ev=%.3e dv=%.3e | value range %.3f..%.3f logit spread %.2f" %
          (n, " | ".join(row), Ep, dp, Ev, dv, v.min().item(), v.max().item(),
           (f32["logits"].max() - f32["logits"].min()).item()))
x = netpacks.planes_of(netpacks.midgame_games(19, seed=77))
xt = torch.from_numpy(x).to(e.device).to(torch.bfloat16)
f32 = T._reference_taps(pack, x, False)
emu = T._reference_taps(pack, x, True)
for layer in range(21):
    got = e.debug_tower(xt, layer=layer)["act"]
    print("layer %2d E=%.3e d=%.3e Em=%.3e dm=%.3e" % (layer, T._err(emu["act"][layer], f32["act"][layer]),
                                                      T._err(got, emu["act"][layer]),
                                                      T._err_mean(emu["act"][layer], f32["act"][layer]),
                                                      T._err_mean(got, emu["act"][layer])))
e.close()

"""Probe (GPU): does running the lanes as TWO engines on two CUDA streams hide the non-tower kernels of a
simulation (select / encode / policy GEMM / softmax / reply / finalize = 3.7 % of a step, a strict dependency chain
with the tower inside one engine) behind the other engine's tower?  No kernel changes: each engine has its own
pools, network workspaces and CUDA graph; the towers serialise on shared memory, the small kernels can co-reside.
Prints simulations/s for one engine with all lanes and for several two-engine splits."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chessrl_b200 import model  # noqa: E402
from chessrl_b200._lib import EVAL_NET  # noqa: E402
from chessrl_b200.engine import Engine  # noqa: E402

S = int(os.environ.get("SIMS", "200"))
CHUNK = int(os.environ.get("CHUNK", "10"))
STEPS = int(os.environ.get("STEPS", "3"))
pack = model.random_pack(seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def make(lanes, stream, seed):
    with torch.cuda.stream(stream):
        e = Engine(max_games=lanes, max_nodes=S + 1, avg_moves=64)
        e.load_weights(pack)
        e.set_evaluator(EVAL_NET)
        start, ml = bench.synthetic_games(e, lanes, seed=seed)
        e.games_set(start, e.pack_move_lists(ml))
    stream.synchronize()
    return e


def run(engines, streams):
    def step():
        for e, s in zip(engines, streams):
            with torch.cuda.stream(s):
                e.mcts_begin_move()
        for _ in range(S // CHUNK):
            for e, s in zip(engines, streams):
                with torch.cuda.stream(s):
                    e.mcts_simulate(CHUNK)
    times = []
    for it in range(STEPS + 2):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream())
        for s in streams:
            s.wait_event(a)
        step()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        b.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        if it >= 2:
            times.append(a.elapsed_time(b))
    lanes = sum(e.max_games for e in engines)
    ms = sum(times) / len(times)
    return {"lanes": [e.max_games for e in engines], "ms_per_step": round(ms, 2), "sims_per_s": round(lanes * S / ms * 1e3)}


out = []
configs = [((4096,), (0,)), ((2368, 1728), (0, 0)), ((2048, 2048), (0, 0)), ((2368, 1728), (-1, 0)), ((1184, 1184, 1728), (0, 0, 0))]
for lanes, prios in configs:
    streams = [torch.cuda.Stream(priority=p) for p in prios]
    engines = [make(n, s, 1234 + i) for i, (n, s) in enumerate(zip(lanes, streams))]
    r = run(engines, streams)
    r["priorities"] = list(prios)
    print(json.dumps(r), flush=True)
    out.append(r)
    for e in engines:
        e.close()
    del engines
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "two_engine_probe.json"), "w"), indent=1)

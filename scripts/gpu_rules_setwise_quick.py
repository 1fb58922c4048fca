"""One short process for a box with seconds of budget: GPU rule parity tests, then perft / movegen timings of the
set-wise rule core (compare with profiles/r01_perft_pair_probe.json 'pair=0' and the bench's kernels.movegen)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
t0 = time.time()
import pytest  # noqa: E402

rc = pytest.main(["-q", "-x", "-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_rules.py"), "-p", "no:cacheprovider"])
print("== rules tests rc", int(rc), "at %.1f s" % (time.time() - t0), flush=True)
import torch  # noqa: E402
from chessrl_b200 import boards as B  # noqa: E402
from chessrl_b200.engine import Engine, _ptr  # noqa: E402
from chessrl_b200._lib import check  # noqa: E402

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
out = {"rules_tests_rc": int(rc)}
e = Engine(max_games=1, max_nodes=8)
for name, fen, depth, want, mf in (("start d5", B.STARTING_FEN, 5, 4865609, 1 << 20), ("kiwipete d5", KIWI, 5, 193690690, 1 << 20),
                                   ("start d7", B.STARTING_FEN, 7, 3195901860, 1 << 26), ("kiwipete d6", KIWI, 6, 8031647685, 1 << 26)):
    rec = B.record_from_fen(fen)
    for bulk in (True, False):
        best = None
        for r in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            t, lanes, plies = e.perft_root(rec, depth, bulk=bulk, min_frontier=mf)
            b.record()
            torch.cuda.synchronize()
            assert t == want, (name, t, want)
            if r:
                best = min(best or 1e9, a.elapsed_time(b))
        out["%s %s" % (name, "bulk" if bulk else "every leaf made")] = {"ms": round(best, 4), "G_nodes_per_s": round(want / best / 1e6, 1)}
        print(name, bulk, out["%s %s" % (name, "bulk" if bulk else "every leaf made")], flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rules_setwise_quick.json"), "w"), indent=1)
# movegen alone on ~1.08 M midgame boards (the bench's kernels.movegen workload)
fr = e.boards_to_device(B.record_from_fen(KIWI)[None, :])
for _ in range(3):
    fr, _ = e.expand_frontier(fr)
boards = fr.repeat(1, 11).contiguous()
n = boards.shape[1]
moves = torch.empty((n, B.MAX_MOVES), dtype=torch.int16, device=e.device)
counts = torch.empty((n,), dtype=torch.int32, device=e.device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
best = 1e9
for r in range(6):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(e.lib.crl_movegen(e.h, _ptr(boards), n, _ptr(moves), _ptr(counts), None))
    b.record()
    torch.cuda.synchronize()
    if r:
        best = min(best, a.elapsed_time(b))
out["movegen"] = {"boards": n, "us": round(best * 1e3, 1), "G_boards_per_s": round(n / best / 1e6, 2)}
print("movegen", out["movegen"], "total %.1f s" % (time.time() - t0), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rules_setwise_quick.json"), "w"), indent=1)
e.close()

"""Probe (GPU): crl_perft_root_host with the last two plies as one pass (k_perft_pair, 80- and 96-register builds)
against "expand the last-but-one ply into HBM, then walk it" (CRL_PERFT_PAIR=0).  Best of 3 after a sizing call."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chessrl_b200 import boards as B  # noqa: E402
from chessrl_b200.engine import Engine  # noqa: E402

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
CASES = [("start d5", B.STARTING_FEN, 5, 4865609, 1 << 20), ("kiwipete d5", KIWI, 5, 193690690, 1 << 20),
         ("start d6", B.STARTING_FEN, 6, 119060324, 1 << 20), ("start d7", B.STARTING_FEN, 7, 3195901860, 1 << 26),
         ("kiwipete d6", KIWI, 6, 8031647685, 1 << 26)]
out = {}
for mode in ("0", "5", "6"):
    os.environ["CRL_PERFT_PAIR"] = mode
    e = Engine(max_games=1, max_nodes=8)
    for name, fen, depth, want, mf in CASES:
        rec = B.record_from_fen(fen)
        for bulk in (True, False):
            best = None
            for r in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                t, lanes, plies = e.perft_root(rec, depth, bulk=bulk, min_frontier=mf)
                b.record()
                torch.cuda.synchronize()
                assert t == want, (mode, name, t, want)
                if r:
                    best = min(best or 1e9, a.elapsed_time(b))
            key = "%s %s" % (name, "bulk" if bulk else "every leaf made")
            out.setdefault(key, {})["pair=" + mode] = {"ms": round(best, 4), "G_nodes_per_s": round(want / best / 1e6, 1),
                                                      "lanes": lanes, "bfs_plies": plies}
            print(mode, key, out[key]["pair=" + mode], flush=True)
    e.close()
    del e
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "perft_pair_probe.json"), "w"), indent=1)

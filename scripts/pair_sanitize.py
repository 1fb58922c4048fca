"""Small crl_perft_root_host workload for compute-sanitizer: default path and the optional two-ply pass."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chessrl_b200 import boards as B  # noqa: E402
from chessrl_b200.engine import Engine  # noqa: E402

KIWI = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"
for mode in ("0", "5", "6"):
    os.environ["CRL_PERFT_PAIR"] = mode
    e = Engine(max_games=1, max_nodes=8)
    for fen, depth, want in ((B.STARTING_FEN, 4, 197281), (KIWI, 3, 97862), (KIWI, 4, 4085603)):
        for bulk in (True, False):
            for mf in (1, 300, 1 << 16):
                if mf == 1 and want > 200000:
                    continue
                t, lanes, plies = e.perft_root(B.record_from_fen(fen), depth, bulk=bulk, min_frontier=mf)
                assert t == want, (mode, fen, depth, t, want)
    e.close()
    print("pair mode", mode, "ok", flush=True)

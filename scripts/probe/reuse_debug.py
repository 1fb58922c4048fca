"""Debug probe: the most-visited-child reuse scenario of tests/test_gpu_reuse.py, off twice and on, first difference printed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from chessrl_b200 import boards as B
from chessrl_b200._lib import EVAL_HASH
from chessrl_b200.engine import Engine
from chessrl_b200.lockstep import LockstepSelfPlay

WANT = ("visits", "values", "priors", "moves", "replies", "results")
foreign = int(os.environ.get("FOREIGN", "1"))
bits = int(os.environ.get("BITS", "11"))
e = Engine(max_games=32, max_nodes=101, avg_moves=218)
e.set_evaluator(EVAL_HASH, 3, bits)


def run(reuse):
    rng = np.random.default_rng(9)
    sp = LockstepSelfPlay(e, n_games=32, sims=100, noise=False, reuse=reuse)
    sp.start()
    trace = []
    for mv in range(14):
        if not sp.running().any():
            break
        c0 = e.counters()
        e.mcts_begin_move()
        e.mcts_simulate(100, 1)
        st = e.root_stats(want=WANT)
        c1 = e.counters()
        st["ev"] = (c1["evaluations"] - c0["evaluations"], c1["reused_evaluations"] - c0["reused_evaluations"])
        st["plies"] = sp._plies.copy()
        trace.append({k: (st[k].copy() if hasattr(st[k], "copy") else st[k]) for k in st})
        live = sp.running()
        p = np.where(live & (st["n_children"][:32] > 0), np.argmax(st["visits"][:32], axis=1), -1).astype(np.int32)
        trace[-1]["picks"] = p.copy()
        trace[-1]["nodes30"] = [(n.parent, n.slot, n.n_legal, n.n_children, n.result, n.move, n.reply, n.visits) for n in e.node_dump(30)[:12]]
        trace[-1]["committed"] = e.commit(p, apply=True).copy()
        sp._read_status()
        trace[-1]["plies_after"] = sp._plies.copy()
        st2 = e.root_stats(want=WANT)
        same_after = all(np.array_equal(st2[k], st[k]) for k in WANT + ("n_children",))
        trace[-1]["nodes30_after"] = [(n.parent, n.slot, n.n_legal, n.n_children, n.result, n.move, n.reply, n.visits) for n in e.node_dump(30)[:12]]
        if not same_after or trace[-1]["nodes30_after"] != trace[-1]["nodes30"]:
            print("reuse", reuse, "move", mv, "root stats / node dump CHANGED across commit", same_after)
        print("reuse", reuse, "mv", mv, "lane30 plies", int(st["plies"][30]), "->", int(sp._plies[30]), "n_children", int(st["n_children"][30]),
              "pick", int(p[30]), "committed", trace[-1]["committed"][30].tolist(), "result", int(sp._results[30]))
        trace[-1]["boards_after"] = e.games_get(0, 32)[0].copy()
        if foreign and mv % 5 == 4:
            for _ in range(2):
                legal, cnt = e.games_legal()
                mvz = np.full(32, 0xFFFF, dtype=np.uint16)
                for g in range(0, 32, 2):
                    if cnt[g] > 0 and sp._results[g] == B.RESULT_NONE:
                        mvz[g] = legal[g, int(rng.integers(cnt[g]))]
                e.games_play(mvz)
                sp._read_status()
    return trace


def diff(a, b, name):
    for mv, (x, y) in enumerate(zip(a, b)):
        for k in ("picks", "committed", "plies_after", "boards_after"):
            if not np.array_equal(x[k], y[k]):
                bad = [g for g in range(32) if not np.array_equal(x[k][g], y[k][g])]
                print(name, "AFTER move", mv, k, "differs in lanes", bad, "A", x[k][bad[0]].tolist(), "B", y[k][bad[0]].tolist())
        if x["nodes30"] != y["nodes30"]:
            print(name, "move", mv, "node dump of lane 30 differs:\n  A", x["nodes30"], "\n  B", y["nodes30"])
        for k in WANT + ("n_children", "root_visits", "root_values"):
            if not np.array_equal(x[k], y[k]):
                bad = [g for g in range(32) if not np.array_equal(x[k][g], y[k][g])]
                g = bad[0]
                n = int(x["n_children"][g])
                print(name, "first difference: move", mv, "key", k, "lanes", bad, "ev", x["ev"], y["ev"], "plies", x["plies"][g])
                for kk in ("visits", "values", "priors", "moves", "replies", "results"):
                    print("  ", kk, "A", x[kk][g][:n].tolist())
                    print("  ", kk, "B", y[kk][g][:n].tolist())
                print("   root", x["root_visits"][g], y["root_visits"][g], x["root_values"][g], y["root_values"][g])
                return
    print(name, "identical over", len(a), "moves; evaluations per move", [t["ev"] for t in b])


off1 = run(False)

on1 = run(True)
diff(off1, on1, "off vs on")


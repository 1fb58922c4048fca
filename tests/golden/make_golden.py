#!/usr/bin/env python
"""Generates tests/golden/*.json|npz by running the reference's UNMODIFIED game.py / netencoder.py /
mctree.py / agentdistributed.py / dataset.py (imported from /root/reference where they lie, see
oracle/ref_on_shims.py) on the python-chess restatement.  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests read the committed fixtures instead.
Floats are stored as float.hex() strings so the fixtures are bit-exact.
"""

import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import chessrl_oracle as O  # noqa: E402
import ref_on_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

FENS = {
    "start": "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "kiwipete": "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
    "pos3": "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
    "pos4": "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
    "pos5": "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
    "pos6": "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10",
    # mates / stalemates one ply away, to exercise terminal children and the (prev ply, our move) quirk
    "mate_in_1": "6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1",
    "kq_vs_k": "7k/8/5KQ1/8/8/8/8/8 w - - 0 1",
    "black_to_move_mates": "6k1/5ppp/8/8/8/1r6/r4PPP/6K1 b - - 0 1",
    "fifty_near": "8/8/4k3/8/8/3K4/R7/8 w - - 97 80",
    "ep_pin": "8/8/8/K2pP2r/8/8/8/4k3 w - d6 0 2",
}


def fhex(x):
    return float(x).hex()


def random_game_moves(ref, seed, plies):
    rng = random.Random(seed)
    g = ref.game.Game()
    moves = []
    for _ in range(plies):
        if g.get_result() is not None:
            break
        legal = g.get_legal_moves()
        m = legal[rng.randrange(len(legal))]
        g.move(m)
        moves.append(m)
    return moves


def gen_labels(ref):
    labels = ref.netencoder.get_uci_labels()
    import hashlib
    with open(os.path.join(OUT, "uci_labels.json"), "w") as f:
        json.dump({"source": "netencoder.get_uci_labels (netencoder.py:94-134), reference code executed",
                   "sha256": hashlib.sha256("\n".join(labels).encode()).hexdigest(),
                   "labels": labels}, f)


def gen_planes(ref):
    """netencoder.get_game_state on the reference's own Game objects."""
    cases = []
    packed = []
    specs = [("start", [], False), ("e2e4", ["e2e4"], False)]
    for seed, plies in ((1, 5), (2, 12), (3, 30), (4, 61), (5, 100), (6, 7), (7, 8), (8, 9)):
        specs.append(("random%d" % seed, None, False))
        specs[-1] = ("random%d_%d" % (seed, plies), random_game_moves(ref, seed, plies), False)
    specs.append(("random3_flipped", random_game_moves(ref, 3, 30), True))
    castle = ["e2e4", "e7e5", "g1f3", "b8c6", "f1c4", "f8c5", "e1g1", "g8f6", "d2d4", "e5d4", "e4e5", "d7d5", "e5d6"]
    specs.append(("castle_ep", castle, False))
    for name, moves, flipped in specs:
        g = ref.game.Game()
        for m in moves:
            assert g.move(m), (name, m)
        p = ref.netencoder.get_game_state(g, flipped=flipped)
        assert p.shape == (8, 8, 127) and set(np.unique(p)) <= {0.0, 1.0}
        cases.append({"name": name, "moves": moves, "flipped": flipped, "sum": int(p.sum())})
        packed.append(np.packbits(p.astype(np.uint8).reshape(-1)))
    # a position built from a FEN has an empty stack -> all history planes zero
    for name in ("kiwipete", "pos4"):
        g = ref.game.Game(board=ref.chess.Board(FENS[name]))
        p = ref.netencoder.get_game_state(g)
        cases.append({"name": name, "fen": FENS[name], "moves": [], "flipped": False, "sum": int(p.sum())})
        packed.append(np.packbits(p.astype(np.uint8).reshape(-1)))
    np.savez_compressed(os.path.join(OUT, "planes.npz"), packed=np.stack(packed))
    with open(os.path.join(OUT, "planes.json"), "w") as f:
        json.dump({"source": "netencoder.get_game_state (netencoder.py:72-91), reference code executed; "
                             "planes.npz['packed'][i] = np.packbits(planes.reshape(-1)) of case i",
                   "cases": cases}, f)


def dump_tree(tree, agent_evals, returned):
    root = tree.root
    kids = []
    for c in root.children:
        ms = [str(m) for m in c.state.board.move_stack]
        n_root = len(root.state.board.move_stack)
        kids.append({"line": ms[n_root:], "visits": int(c.visits), "value": fhex(c.value),
                     "prior": fhex(c.prior), "score": fhex(c.get_value()),
                     "result": c.state.get_result(), "n_children": len(c.children),
                     "grandchild_visits": [int(gc.visits) for gc in c.children]})
    return {"root_visits": int(root.visits), "root_value": fhex(root.value), "n_evals": agent_evals,
            "returned": list(returned), "children": kids,
            "policy_no_noise": [fhex(x) for x in tree.compute_policy(root, noise=False)]}


def gen_mcts(ref):
    """mctree.SelfPlayTree(threads=1) + AgentDistributed logic + game.Game, all reference code, fed by the
    deterministic hash evaluator (oracle/chessrl_oracle.py hash_evaluator)."""
    cases = []
    roots = [("start", None, []), ("kiwipete", FENS["kiwipete"], []), ("pos3", FENS["pos3"], []),
             ("pos4", FENS["pos4"], []), ("pos5", FENS["pos5"], []),
             ("mate_in_1", FENS["mate_in_1"], []), ("kq_vs_k", FENS["kq_vs_k"], []),
             ("black_to_move_mates", FENS["black_to_move_mates"], []),
             ("fifty_near", FENS["fifty_near"], []), ("ep_pin", FENS["ep_pin"], []),
             ("random11_24", None, random_game_moves(ref, 11, 24)),
             ("random12_47", None, random_game_moves(ref, 12, 47)),
             ("random13_80", None, random_game_moves(ref, 13, 80))]
    for name, fen, moves in roots:
        for seed, bits in ((1, 24), (2, 3)):
            for sims in (1, 30, 120):
                if bits == 3 and sims == 1:
                    continue
                board = ref.chess.Board(fen) if fen else ref.chess.Board()
                g = ref.game.Game(board=board)
                for m in moves:
                    assert g.move(m)
                if g.get_result() is not None:
                    continue
                ev = O.hash_evaluator(seed=seed, policy_bits=bits)
                agent = ref_on_shims.make_ref_agent(ref, ev)
                type(agent).n_evals = 0
                tree = ref.mctree.SelfPlayTree(g, threads=1)
                ret = tree.search_move(agent, max_iters=sims, noise=False, ai_move=True)
                rec = dump_tree(tree, type(agent).n_evals, ret)
                rec.update({"name": name, "fen": fen, "moves": moves, "eval_seed": seed,
                            "policy_bits": bits, "sims": sims})
                cases.append(rec)
    with open(os.path.join(OUT, "mcts_chess.json"), "w") as f:
        json.dump({"source": "mctree.SelfPlayTree.search_move(threads=1, noise=False, ai_move=True) "
                             "(mctree.py:159-198), reference code executed on game.py + python-chess restatement; "
                             "evaluator = chessrl_oracle.hash_evaluator(seed, policy_bits)",
                   "cases": cases}, f)


def ref_wave_search(tree, agent, sims, K):
    """Drives the reference's own SelfPlayTree.select / simulate / backprop (mctree.py:216-296) in the WAVE
    schedule of threads=K (see oracle/chessrl_oracle.py): per wave up to K selects, then their simulates, then
    their backprops in the same order; a select that would enter a node created in the same wave by an expansion
    with an opponent reply is deferred to the next wave.  Returns the number of waves."""
    def meets_pending(pending):
        node = tree.root
        while not node.is_terminal_state:                 # dry run of select: these reference calls are pure
            if not node.is_fully_expanded:
                return False
            node = node.get_best_child()
            if id(node) in pending:
                return True
        return False

    left, waves = sims, 0
    while left > 0:
        leaves, pending = [], set()
        for j in range(min(K, left)):
            if j and meets_pending(pending):
                break
            leaf = tree.select(tree.root, agent)
            new = leaf.visits == 0 and all(leaf is not x for x in leaves)
            if new and len(leaf.state.board.move_stack) == len(leaf.parent.state.board.move_stack) + 2:
                pending.add(id(leaf))
            leaves.append(leaf)
        values = [tree.simulate(leaf, agent) for leaf in leaves]
        for leaf, v in zip(leaves, values):
            tree.backprop(leaf, v, remove_vloss=True)
        left -= len(leaves)
        waves += 1
    return waves


def gen_mcts_wave(ref):
    """threads=K > 1: the reference's own select / simulate / backprop driven in the wave schedule."""
    cases = []
    roots = [("start", None, []), ("kiwipete", FENS["kiwipete"], []), ("pos3", FENS["pos3"], []),
             ("pos4", FENS["pos4"], []), ("mate_in_1", FENS["mate_in_1"], []), ("kq_vs_k", FENS["kq_vs_k"], []),
             ("black_to_move_mates", FENS["black_to_move_mates"], []), ("fifty_near", FENS["fifty_near"], []),
             ("random11_24", None, random_game_moves(ref, 11, 24)),
             ("random13_80", None, random_game_moves(ref, 13, 80)),
             # narrow trees (1-3 legal moves per node): nodes fill up inside one wave, so waves get cut
             ("narrow_one_move", "k7/8/8/8/8/8/1r6/K7 w - - 0 1", []),
             ("narrow_two_moves", "8/8/8/8/8/2k5/7p/K7 w - - 0 1", []),
             ("narrow_pawn_ending", "8/8/8/p7/P7/8/8/K6k w - - 0 1", []),
             ("narrow_check_evasion", "4k3/8/8/8/8/8/4r3/4K3 w - - 0 1", [])]
    for name, fen, moves in roots:
        for seed, bits in ((1, 24), (2, 3)):
            plan = ((2, 40), (6, 120), (8, 100), (16, 150))
            if name.startswith("narrow"):
                plan = ((3, 60), (6, 90), (32, 128), (64, 128))
            for K, sims in plan:
                board = ref.chess.Board(fen) if fen else ref.chess.Board()
                g = ref.game.Game(board=board)
                for m in moves:
                    assert g.move(m)
                if g.get_result() is not None:
                    continue
                ev = O.hash_evaluator(seed=seed, policy_bits=bits)
                agent = ref_on_shims.make_ref_agent(ref, ev)
                type(agent).n_evals = 0
                tree = ref.mctree.SelfPlayTree(g, threads=K)
                waves = ref_wave_search(tree, agent, sims, K)
                assert all(c.vloss == 0 for c in tree.root.children)
                pick = int(np.argmax(tree.compute_policy(tree.root, noise=False)))
                st = tree.root.children[pick].state.board.move_stack
                ret = (str(st[-2]), str(st[-1])) if len(st) >= 2 else ("00000", "00000")
                rec = dump_tree(tree, type(agent).n_evals, ret)
                rec.update({"name": name, "fen": fen, "moves": moves, "eval_seed": seed, "policy_bits": bits,
                            "sims": sims, "threads": K, "waves": waves})
                cases.append(rec)
    with open(os.path.join(OUT, "mcts_wave.json"), "w") as f:
        json.dump({"source": "mctree.SelfPlayTree(threads=K).select / simulate / backprop (mctree.py:216-296), the "
                             "reference's own methods driven in the wave schedule (make_golden.ref_wave_search); "
                             "evaluator = chessrl_oracle.hash_evaluator(seed, policy_bits)",
                   "cases": cases}, f)


def gen_toy(ref):
    """the reference mctree.py on a toy game (no chess rules involved)."""

    class ToyBoard:
        def __init__(self, stack):
            self.move_stack = list(stack)

    class ToyGame:
        NULL_MOVE = "00000"

        def __init__(self, moves=(), maxply=6):
            self.board = ToyBoard(moves)
            self.maxply = maxply

        def get_legal_moves(self):
            return ["0", "1", "2"]

        def move(self, m):
            if m not in ("0", "1", "2"):
                return False
            self.board.move_stack.append(m)
            return True

        def get_result(self):
            if len(self.board.move_stack) >= self.maxply:
                return sum(int(m) for m in self.board.move_stack) % 3 - 1
            return None

        def get_copy(self):
            return ToyGame(self.board.move_stack, self.maxply)

    def h_of(game):
        h = 7
        for m in game.board.move_stack:
            h = (h * 31 + int(m) + 1) % 1000003
        return h

    class ToyAgent:
        def __init__(self):
            self.n = 0

        def get_copy(self):
            return self

        def connect(self):
            pass

        def disconnect(self):
            pass

        def predict_policy(self, game, mask_legal_moves=True):
            self.n += 1
            h = h_of(game)
            return [np.float32((h * 7 + i * 13) % 100 + 1) / np.float32(1000) for i in range(3)]

        def predict_outcome(self, game):
            self.n += 1
            h = h_of(game)
            return float(np.float32(h % 2001 - 1000) / np.float32(1000))

        def best_move(self, game, real_game=True):
            return game.get_legal_moves()[int(np.argmax(self.predict_policy(game)))]

    out = []
    for root_moves, maxply, sims_list in (((), 6, (1, 3, 4, 10, 50, 200)), (("0", "1", "2", "0"), 5, (5,))):
        for sims in sims_list:
            a = ToyAgent()
            tree = ref.mctree.SelfPlayTree(ToyGame(root_moves, maxply), threads=1)
            ret = tree.search_move(a, max_iters=sims, noise=False, ai_move=True)
            r = tree.root
            out.append({"root_moves": list(root_moves), "maxply": maxply, "sims": sims, "n_evals": a.n,
                        "returned": list(ret), "root_visits": r.visits, "root_value": fhex(r.value),
                        "visits": [c.visits for c in r.children],
                        "values": [fhex(c.value) for c in r.children],
                        "priors": [fhex(c.prior) for c in r.children],
                        "scores": [fhex(c.get_value()) for c in r.children],
                        "policy": [fhex(x) for x in tree.compute_policy(r, noise=False)]})
    with open(os.path.join(OUT, "mcts_toy.json"), "w") as f:
        json.dump({"source": "SURVEY.md KAT-5/KAT-5b: unmodified mctree.py on a toy game", "cases": out}, f)


def gen_selfplay(ref):
    """A short self-play run with the reference's search code: the loop of selfplay.play_game
    (selfplay.py:59-84) with max_iters lowered from the hard-coded 900 and numpy's legacy RNG seeded before
    every search so the Dirichlet draw (mctree.py:318-321) is reproducible."""
    runs = []
    for player_color, sims, n_moves in ((True, 40, 6), (False, 25, 5)):
        ev = O.hash_evaluator(seed=5)
        agent = ref_on_shims.make_ref_agent(ref, ev)
        g = ref.game.Game(player_color=player_color, date="01/01/2020 00:00:00")
        agent.color = player_color
        if player_color is False:
            g.move(agent.best_move(g, real_game=True))
        picks = []
        for k in range(n_moves):
            if g.get_result() is not None:
                break
            np.random.seed(1000 + k)
            bm, am = agent.best_move(g, real_game=False, ai_move=True, max_iters=sims)
            picks.append([bm, am])
            g.move(bm)
            g.move(am)
        ds = ref.dataset.DatasetGame([g])
        runs.append({"player_color": player_color, "sims": sims, "eval_seed": 5, "noise_seed_base": 1000,
                     "picks": picks, "history": g.get_history(), "dataset_str": str(ds)})
    with open(os.path.join(OUT, "selfplay.json"), "w") as f:
        json.dump({"source": "selfplay.play_game loop (selfplay.py:59-84) over AgentDistributed.best_move + "
                             "mctree.SelfPlayTree, reference code executed; noise=True, np.random.seed(1000+k)",
                   "runs": runs}, f)


def gen_rules(ref):
    """Move lists / results along games, through the reference's Game wrapper (game.py:43-57, 92-109).
    These pin Game.get_legal_moves order and Game.get_result on top of the python-chess restatement."""
    recs = []
    for name, fen in FENS.items():
        g = ref.game.Game(board=ref.chess.Board(fen))
        recs.append({"name": name, "fen": fen, "moves": [], "legal": g.get_legal_moves(), "result": g.get_result()})
    for seed in range(20, 32):
        moves = random_game_moves(ref, seed, 400)
        g = ref.game.Game()
        trace = []
        for i, m in enumerate(moves):
            g.move(m)
            if i % 7 == 0 or i >= len(moves) - 3:
                trace.append({"ply": i + 1, "legal": g.get_legal_moves(), "result": g.get_result(),
                              "fen": g.board.fen()})
        recs.append({"name": "random%d" % seed, "fen": None, "moves": moves, "final_result": g.get_result(),
                     "trace": trace})
    with open(os.path.join(OUT, "rules.json"), "w") as f:
        json.dump({"source": "game.Game.get_legal_moves / get_result (game.py:43-57, 92-109), reference code "
                             "executed on the python-chess restatement", "cases": recs}, f)


def gen_training_batches(ref):
    """the training input pipeline.  The reference's own DatasetGame.augment_game (dataset.py:21-43) and
    netencoder.DataGameSequence.__getitem__ (netencoder.py:159-181) run on three stored games; the flip draw
    (`np.random.rand() < random_flips`, one per game) is steered by random_flips = 0 / 1 and by seeding numpy."""
    fools = ["f2f3", "e7e5", "g2g4", "d8h4"]                                         # 0-1
    scholars = ["e2e4", "e7e5", "d1h5", "b8c6", "f1c4", "g8f6", "h5f7"]             # 1-0
    games_moves = [fools, scholars, random_game_moves(ref, 41, 30)]                  # the third one is unfinished
    colors = [True, False, True]
    games = []
    for moves, col in zip(games_moves, colors):
        g = ref.game.Game(player_color=col, date="01/01/2020 00:00:00")
        for m in moves:
            assert g.move(m)
        games.append(g)
    ds = ref.dataset.DatasetGame(games)
    aug = []
    for g in games:
        samples = ds.augment_game(g)
        aug.append([{"plies": len(s["game"].board.move_stack), "next_move": s["next_move"], "result": s["result"],
                     "player_color": s["game"].player_color, "fen": s["game"].board.fen()} for s in samples])
    base = ref.netencoder.DataGameSequence(ds, batch_size=3, random_flips=0)
    assert len(base) == 1
    x0, (p0, v0) = base[0]
    cases, packed = [], []
    for name, rf, seed in (("no_flips", 0, None), ("all_flipped", 1.0, None), ("seeded_half_a", 0.5, 1),
                           ("seeded_half_b", 0.5, 6), ("training_rate_0.1", 0.1, 7)):
        seq = ref.netencoder.DataGameSequence(ds, batch_size=3, random_flips=rf)
        if seed is not None:
            np.random.seed(seed)
        x, (pol, val) = seq[0]
        after = float(np.random.rand()) if seed is not None else None     # where the global stream stands afterwards
        assert x.shape == (sum(len(m) for m in games_moves), 8, 8, 127) and pol.shape == (x.shape[0], 1968)
        assert (pol == p0).all()                                          # the policy target is never flipped
        flips, o = [], 0
        for m in games_moves:
            same = bool((x[o:o + len(m)] == x0[o:o + len(m)]).all())
            rot = bool((x[o:o + len(m)] == x0[o:o + len(m), ::-1, ::-1]).all())
            assert same != rot
            flips.append(rot)
            o += len(m)
        cases.append({"name": name, "random_flips": rf, "seed": seed, "flips": flips, "next_rand": after,
                      "policy_index": [int(i) for i in pol.argmax(1)], "values": [None if v is None else int(v) for v in val]})
        packed.append(np.packbits(x.astype(np.uint8).reshape(-1)))
    np.savez_compressed(os.path.join(OUT, "training_batches.npz"), packed=np.stack(packed))
    with open(os.path.join(OUT, "training_batches.json"), "w") as f:
        json.dump({"source": "dataset.DatasetGame.augment_game (dataset.py:21-43) + netencoder.DataGameSequence.__getitem__ "
                             "(netencoder.py:159-181), reference code executed; training_batches.npz['packed'][i] = "
                             "np.packbits(x.reshape(-1)) of case i, x float64 [N,8,8,127]",
                   "games": [{"moves": m, "player_color": c, "result": g.get_result()}
                             for m, c, g in zip(games_moves, colors, games)],
                   "augment": aug, "cases": cases}, f)


def main():
    ref = ref_on_shims.load_reference()
    if ref is None:
        sys.exit("reference tree not mounted; fixtures can only be regenerated in the build container")
    gen_labels(ref)
    gen_planes(ref)
    gen_toy(ref)
    gen_rules(ref)
    gen_mcts(ref)
    gen_mcts_wave(ref)
    gen_selfplay(ref)
    gen_training_batches(ref)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()

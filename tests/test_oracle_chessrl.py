"""Pins the in-repo restatement of the ChessRL layer (oracle/chessrl_oracle.py) against golden vectors that
were produced by the reference's OWN unmodified netencoder.py / mctree.py / agentdistributed.py / game.py
(tests/golden/make_golden.py), and -- where /root/reference is mounted -- against that code live."""
import json
import os

import numpy as np
import pytest

import chessrl_oracle as O
import ref_on_shims

chess = O.chess


def _game(fen, moves):
    g = O.OGame(board=chess.Board(fen) if fen else None)
    for m in moves:
        assert g.move(m)
    return g


def test_labels_golden(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "uci_labels.json")))
    labels = O.uci_labels()
    assert labels == gold["labels"] and len(set(labels)) == 1968
    assert O.labels_sha256() == gold["sha256"] == O.UCI_LABELS_SHA256
    idx = O.label_index()
    kat3 = {"e2e4": 930, "g1f3": 1402, "e1g1": 901, "e1c1": 898, "e8g8": 1120, "e8c8": 1117, "e7e8": 1099,
            "e7e8q": 1881, "a7a8n": 1805, "h2g1r": 1958, "a1a2": 7, "h8h7": 1782}
    assert {k: idx[k] for k in kat3} == kat3
    assert labels[0:3] == ["a1b1", "a1c1", "a1d1"] and labels[1791] == "h8g6" and labels[-1] == "h7g8n"
    assert labels[1792:1796] == ["a2a1q", "a7a8q", "a2b1q", "a7b8q"]


def test_planes_kat4_hand_derived():
    p = O.planes(O.OGame())
    assert p.shape == (8, 8, 127) and p.sum() == 192
    assert p[2:8, :, 0].all() and not p[0:2, :, 0].any()          # no black piece on ranks 6..1
    assert p[1, :, 1].all() and p[:, :, 1].sum() == 8               # black pawns on row 1
    assert p[0, 0, 4] == 1 and p[0, 7, 4] == 1 and p[0, 4, 6] == 1  # black rooks, king
    assert p[0:6, :, 7].all() and p[6, :, 8].all() and p[7, 4, 13] == 1
    assert not p[:, :, 14:126].any() and p[:, :, 126].all()
    g = O.OGame()
    g.move("e2e4")
    q = O.planes(g)
    assert not q[:, :, 126].any() and q[4, 4, 8] == 1 and q[6, 4, 8] == 0
    assert (q[:, :, 14:28] == p[:, :, 0:14]).all() and not q[:, :, 28:126].any()


def test_planes_golden(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "planes.json")))["cases"]
    packed = np.load(os.path.join(golden_dir, "planes.npz"))["packed"]
    for c, bits in zip(meta, packed):
        g = _game(c.get("fen"), c["moves"])
        p = O.planes(g, flipped=c["flipped"])
        want = np.unpackbits(bits)[:8 * 8 * 127].reshape(8, 8, 127)
        assert p.sum() == c["sum"]
        assert (p == want).all(), c["name"]


def test_mcts_toy_golden_kat5(golden_dir):
    """SURVEY.md KAT-5 / KAT-5b, restated tree on the toy game."""

    class ToyBoard:
        def __init__(self, s):
            self.move_stack = list(s)

    class ToyGame:
        def __init__(self, moves=(), maxply=6):
            self.board = ToyBoard(moves)
            self.maxply = maxply

        def get_legal_moves(self):
            return ["0", "1", "2"]

        def move(self, m):
            self.board.move_stack.append(m)
            return True

        def get_result(self):
            s = self.board.move_stack
            return sum(int(m) for m in s) % 3 - 1 if len(s) >= self.maxply else None

        def get_copy(self):
            return ToyGame(self.board.move_stack, self.maxply)

    def h_of(g):
        h = 7
        for m in g.board.move_stack:
            h = (h * 31 + int(m) + 1) % 1000003
        return h

    class ToyAgent:
        n_evals = 0

        def predict_policy(self, g, mask_legal_moves=True):
            self.n_evals += 1
            return [np.float32((h_of(g) * 7 + i * 13) % 100 + 1) / np.float32(1000) for i in range(3)]

        def predict_outcome(self, g):
            self.n_evals += 1
            return float(np.float32(h_of(g) % 2001 - 1000) / np.float32(1000))

        def policy_move(self, g):
            return g.get_legal_moves()[int(np.argmax(self.predict_policy(g)))]

    for c in json.load(open(os.path.join(golden_dir, "mcts_toy.json")))["cases"]:
        a = ToyAgent()
        t = O.OSelfPlayTree(ToyGame(c["root_moves"], c["maxply"]))
        ret = t.search_move(a, max_iters=c["sims"], noise=False, ai_move=True)
        r = t.root
        assert list(ret) == c["returned"] and a.n_evals == c["n_evals"]
        assert [k.visits for k in r.children] == c["visits"] and r.visits == c["root_visits"]
        assert float(r.value) == float.fromhex(c["root_value"])
        assert [float(k.value) for k in r.children] == [float.fromhex(x) for x in c["values"]]
        assert [float(k.get_value()) for k in r.children] == [float.fromhex(x) for x in c["scores"]]
        assert [float(x) for x in t.compute_policy(r, noise=False)] == [float.fromhex(x) for x in c["policy"]]


def test_mcts_chess_golden(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "mcts_chess.json")))["cases"]
    for c in cases:
        if c["sims"] > 30:
            continue                                   # the 120-sim cases run in the GPU parity test
        ag = O.OAgent(O.hash_evaluator(c["eval_seed"], c["policy_bits"]))
        t = O.OSelfPlayTree(_game(c["fen"], c["moves"]))
        ret = t.search_move(ag, max_iters=c["sims"], noise=False, ai_move=True)
        assert list(ret) == c["returned"], c["name"]
        assert t.root.visits == c["root_visits"] and float(t.root.value) == float.fromhex(c["root_value"])
        for k, kid in zip(t.root.children, c["children"]):
            assert k.visits == kid["visits"] and float(k.value) == float.fromhex(kid["value"])
            assert k.state.get_result() == kid["result"]
            assert [g.visits for g in k.children] == kid["grandchild_visits"]
            assert float(k.get_value()) == float.fromhex(kid["score"])
        assert len(t.root.children) == len(c["children"])


def test_mcts_wave_golden(golden_dir):
    """threads=K > 1: the oracle's wave schedule against the reference's own select / simulate / backprop driven
    in that schedule (tests/golden/make_golden.py ref_wave_search)."""
    cases = json.load(open(os.path.join(golden_dir, "mcts_wave.json")))["cases"]
    ran = 0
    for c in cases:
        if c["sims"] > 100 or c["policy_bits"] != 24:
            continue                                   # the rest run in the hostsim / GPU parity tests
        ag = O.OAgent(O.hash_evaluator(c["eval_seed"], c["policy_bits"]))
        t = O.OSelfPlayTree(_game(c["fen"], c["moves"]), threads=c["threads"])
        ret = t.search_move(ag, max_iters=c["sims"], noise=False, ai_move=True)
        assert list(ret) == c["returned"], c["name"]
        assert t.n_waves == c["waves"]
        assert t.root.visits == c["root_visits"] and float(t.root.value) == float.fromhex(c["root_value"])
        assert len(t.root.children) == len(c["children"])
        for k, kid in zip(t.root.children, c["children"]):
            assert k.visits == kid["visits"] and float(k.value) == float.fromhex(kid["value"]) and k.vloss == 0
            assert [g.visits for g in k.children] == kid["grandchild_visits"]
            assert float(k.get_value()) == float.fromhex(kid["score"])
        ran += 1
    assert ran >= 10


def test_wave_schedule_threads1_is_the_plain_schedule():
    g = _game(None, ["e2e4", "e7e5"])
    a = O.OSelfPlayTree(g, threads=1)
    a.search_move(O.OAgent(O.hash_evaluator(3, 24)), max_iters=40, noise=False)
    b = O.OSelfPlayTree(g)
    b.search_move(O.OAgent(O.hash_evaluator(3, 24)), max_iters=40, noise=False)
    assert [c.visits for c in a.root.children] == [c.visits for c in b.root.children]


def test_selfplay_golden(golden_dir):
    for run in json.load(open(os.path.join(golden_dir, "selfplay.json")))["runs"]:
        ag = O.OAgent(O.hash_evaluator(run["eval_seed"]))
        g = O.OGame(player_color=run["player_color"])
        if not run["player_color"]:
            g.move(ag.best_move(g, real_game=True))
        for k, (bm, am) in enumerate(run["picks"]):
            np.random.seed(run["noise_seed_base"] + k)
            got = ag.best_move(g, real_game=False, ai_move=True, max_iters=run["sims"])
            assert list(got) == [bm, am]
            g.move(bm)
            g.move(am)
        assert g.get_history()["moves"] == run["history"]["moves"]


def test_dirichlet_stream_kat7():
    np.random.seed(0)
    d = np.random.dirichlet([0.03] * 3)
    assert np.allclose(d, [4.20658792e-02, 9.57926590e-01, 7.53080427e-06], rtol=1e-6)


@pytest.mark.skipif(ref_on_shims.reference_dir() is None, reason="reference tree not mounted (GPU box)")
def test_restatement_matches_reference_code_live():
    ref = ref_on_shims.load_reference()
    assert ref.netencoder.get_uci_labels() == O.uci_labels()
    import random
    rng = random.Random(99)
    rg, og = ref.game.Game(), O.OGame()
    for ply in range(60):
        if rg.get_result() is not None:
            break
        assert rg.get_legal_moves() == og.get_legal_moves() and rg.get_result() == og.get_result()
        m = rng.choice(rg.get_legal_moves())
        rg.move(m)
        og.move(m)
        if ply % 5 == 0:
            assert (ref.netencoder.get_game_state(rg) == O.planes(og)).all()
    ev = O.hash_evaluator(seed=3)
    ra = ref_on_shims.make_ref_agent(ref, ev)
    rt = ref.mctree.SelfPlayTree(rg, threads=1)
    r1 = rt.search_move(ra, max_iters=40, noise=False, ai_move=True)
    ot = O.OSelfPlayTree(og)
    r2 = ot.search_move(O.OAgent(ev), max_iters=40, noise=False, ai_move=True)
    assert tuple(r1) == tuple(r2)
    assert [c.visits for c in rt.root.children] == [c.visits for c in ot.root.children]
    assert [float(c.value) for c in rt.root.children] == [float(c.value) for c in ot.root.children]

"""SURVEY 8(f)-2, CPU half: the oracle's netencoder restatement against the training batches the reference's own
DataGameSequence / augment_game produced (tests/golden/training_batches.*, written by make_golden.gen_training_batches).
The GPU half (tests/test_gpu_training.py) checks the product path -- training.encode_games on the CUDA encode kernel --
against the same fixture."""
import json
import os

import numpy as np

import chessrl_oracle as O
from conftest import GOLDEN


def load_training_golden():
    with open(os.path.join(GOLDEN, "training_batches.json")) as f:
        meta = json.load(f)
    packed = np.load(os.path.join(GOLDEN, "training_batches.npz"))["packed"]
    n = sum(len(g["moves"]) for g in meta["games"])
    xs = [np.unpackbits(p)[:n * 8 * 8 * 127].reshape(n, 8, 8, 127) for p in packed]
    return meta, xs


def test_oracle_reproduces_reference_training_batches():
    meta, xs = load_training_golden()
    labels = O.label_index()
    assert [c["flips"] for c in meta["cases"][:2]] == [[False] * 3, [True] * 3]
    assert any(len(set(c["flips"])) == 2 for c in meta["cases"])            # mixed flips are covered
    for case, x in zip(meta["cases"], xs):
        row = 0
        for gi, g in enumerate(meta["games"]):
            og = O.OGame()
            for ply, m in enumerate(g["moves"]):
                assert (O.planes(og, flipped=case["flips"][gi]) == x[row]).all(), (case["name"], gi, ply)
                assert labels[m] == case["policy_index"][row]
                assert case["values"][row] == g["result"]
                aug = meta["augment"][gi][ply]
                assert aug["plies"] == ply and aug["next_move"] == m and aug["fen"] == og.board.fen()
                og.move(m)
                row += 1
        assert row == x.shape[0]

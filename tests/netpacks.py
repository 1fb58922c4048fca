"""Weight packs and positions for the network parity tests.

Random-init packs are easily DEGENERATE test vectors: with Keras-default initialisation the value head's single ReLU
channel is often dead (value exactly 0) or the tanh saturated (value exactly 1), and every policy entry sits within
1e-3 of 1/1968 -- a tower that outputs nothing would pass a tolerance test on such outputs.  `lively_pack` perturbs
biases / BatchNorm statistics as a trained network would have them and sharpens the policy head;
`assert_lively` is called IN the tests on the fp32 reference outputs so a pack can never silently go flat."""
import random

import numpy as np

import chessrl_oracle as O
from chessrl_b200 import model

LIVELY_SEED = 21


def lively_pack(seed=LIVELY_SEED):
    pack = model.random_pack(seed, perturb_bn=True)
    pack[128] = pack[128] * 3.0                       # policy dense kernel: logits spread over ~4.5 instead of ~1.5
    return pack


def assert_lively(policy, value):
    """policy [n,1968], value [n] from the fp32 reference (torch tensors or arrays)."""
    p = np.asarray(policy.detach().cpu() if hasattr(policy, "detach") else policy, dtype=np.float64)
    v = np.asarray(value.detach().cpu() if hasattr(value, "detach") else value, dtype=np.float64).reshape(-1)
    assert np.all(p.max(1) / p.min(1) > 10), "flat policy: max/min %.2f" % (p.max(1) / p.min(1)).min()
    assert np.all(np.abs(v) < 0.999), "saturated value head"
    if len(v) >= 8:
        assert v.max() - v.min() > 0.1, "dead value head: spread %.3g" % (v.max() - v.min())


def midgame_games(n, seed=0, max_plies=40):
    rng = random.Random(seed)
    games = []
    for _ in range(n):
        g = O.OGame()
        for _ in range(rng.randrange(0, max_plies)):
            ms = g.get_legal_moves()
            if not ms or g.get_result() is not None:
                break
            g.move(rng.choice(ms))
        games.append(g)
    return games


def planes_of(games):
    """float32 [n,8,8,128] (channel 127 = zero pad) from the oracle's netencoder restatement."""
    x = np.zeros((len(games), 8, 8, 128), dtype=np.float32)
    x[..., :127] = np.stack([O.planes(g) for g in games])
    return x


def synthetic_planes(n, seed, density=0.2):
    """Random 0/1 planes for large batches (real positions cost a Python movegen each)."""
    rng = np.random.default_rng(seed)
    x = (rng.random((n, 8, 8, 128)) < density).astype(np.float32)
    x[..., 127] = 0
    return x

"""Host-side move policy (chessrl_b200.lockstep.compute_policy) against the reference formula
(mctree.SelfPlayTree.compute_policy, mctree.py:305-322), bit for bit, including the Dirichlet stream."""
import numpy as np

from chessrl_b200.lockstep import compute_policy


def _reference(child_visits, root_visits, n_plies, noise):
    tau = 1
    if n_plies >= 30:
        tau = n_plies / (1 + np.power(n_plies, 1.3))
    policy = np.array([np.power(int(v), 1 / tau) for v in child_visits]) / np.power(int(root_visits), 1 / tau)
    if noise:
        policy = (1 - 0.25) * policy + np.random.dirichlet([0.03] * len(child_visits))
    return policy


def test_compute_policy_bit_identical():
    rng = np.random.default_rng(0)
    for t in range(1500):
        k = int(rng.integers(1, 40))
        cv = rng.integers(0, 300, k).astype(np.int32)
        rv = int(cv.sum() + 1)
        n = int(rng.integers(0, 200))
        for noise in (False, True):
            np.random.seed(t)
            a = compute_policy(cv, rv, n, noise)
            np.random.seed(t)
            b = _reference(cv, rv, n, noise)
            assert a.dtype == b.dtype and a.tobytes() == b.tobytes(), (t, n, noise)


def test_compute_policy_edge_cases():
    assert compute_policy([5], 6, 0, False).tolist() == [5 / 6]
    assert compute_policy([0, 0, 1], 2, 29, False).tolist() == [0.0, 0.0, 0.5]
    p = compute_policy([0, 3], 4, 30, False)          # first ply with tau != 1
    tau = 30 / (1 + np.power(30, 1.3))
    assert p.tolist() == [0.0, float(np.power(3, 1 / tau) / np.power(4, 1 / tau))]


def test_pick_moves_batch_is_the_per_game_sequence():
    """The vectorised lockstep pick (one standard_gamma call for all lanes) = compute_policy + argmax lane by lane,
    including where numpy's legacy global RNG stands afterwards."""
    from chessrl_b200.lockstep import pick_moves
    rng = np.random.default_rng(1)
    for t in range(40):
        G = int(rng.integers(1, 300))
        kk = rng.integers(0, 50, G)
        vis = np.zeros((G, 256), dtype=np.int32)
        for g in range(G):
            vis[g, :kk[g]] = rng.integers(0, 200, kk[g])
        rv = (vis.sum(1) + 1).astype(np.int32)
        plies = rng.integers(0, 120, G).astype(np.int32)
        live = rng.random(G) < 0.8
        for noise in (False, True):
            np.random.seed(t)
            want = np.full(G, -1, dtype=np.int32)
            for g in range(G):
                if live[g] and kk[g]:
                    want[g] = int(np.argmax(compute_policy(vis[g, :kk[g]], rv[g], int(plies[g]), noise)))
            after_want = np.random.random_sample()
            np.random.seed(t)
            got = pick_moves(vis, kk, rv, plies, live, noise)
            after_got = np.random.random_sample()
            assert (got == want).all(), (t, noise)
            assert after_got == after_want
    assert (pick_moves(np.zeros((3, 256), dtype=np.int32), [0, 0, 0], [1, 1, 1], [0, 0, 0], [True] * 3) == -1).all()

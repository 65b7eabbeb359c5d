"""Neighbour-table construction (host logic)."""

import numpy as np

from depthdensifier_b200.neighbours import all_views_table, camera_centers, covisibility_table, default_vote_threshold, nearest_views_table


def _ring_poses(n, radius=4.0):
    poses = np.zeros((n, 3, 4))
    for i in range(n):
        a = 2 * np.pi * i / n
        c = np.array([radius * np.cos(a), radius * np.sin(a), 1.0])
        z = -c / np.linalg.norm(c)
        x = np.cross([0, 0, 1.0], z)
        x /= np.linalg.norm(x)
        R = np.stack([x, np.cross(z, x), z])
        poses[i, :, :3], poses[i, :, 3] = R, -R @ c
    return poses


def test_nearest_and_all_views():
    poses = _ring_poses(12)
    assert np.allclose(np.linalg.norm(camera_centers(poses)[:, :2], axis=1), 4.0)
    nbr = nearest_views_table(poses, 4)
    for s in range(12):
        assert sorted(nbr[s]) == sorted([(s - 2) % 12, (s - 1) % 12, (s + 1) % 12, (s + 2) % 12]) and s not in nbr[s]
    assert np.array_equal(nearest_views_table(poses[:3], 4)[0], [1, 2, -1, -1])  # padded when V-1 < K
    assert np.array_equal(all_views_table(3), [[0, 1, 2]] * 3)
    assert [default_vote_threshold(k) for k in (1, 4, 8, 9)] == [1, 2, 4, 5]


def test_covisibility_table():
    # five views; 0/1/2 see one surface, 3/4 another; view 2 also sees a little of the second one
    obs = [np.arange(0, 100), np.arange(20, 120), np.r_[np.arange(60, 160), 500, 501], np.arange(500, 560), np.arange(530, 600)]
    nbr = covisibility_table(obs, 2)
    assert list(nbr[0]) == [1, 2] and list(nbr[1]) == [0, 2] and list(nbr[3]) == [4, 2] and nbr[4][0] == 3
    assert all(s not in nbr[s] for s in range(5))
    # duplicates inside a view count once; ties fall back to camera distance when poses are given
    poses = _ring_poses(5)
    nbr2 = covisibility_table([np.r_[o, o] for o in obs], 2, poses)
    assert np.array_equal(nbr2[:4], nbr[:4])
    none = covisibility_table([np.zeros(0, np.int64)] * 4, 2, _ring_poses(4))  # nothing shared -> nearest views
    assert np.array_equal(np.sort(none, 1), np.sort(nearest_views_table(_ring_poses(4), 2), 1))
    assert covisibility_table(obs[:2], 3).tolist() == [[1, -1, -1], [0, -1, -1]]


def test_covisibility_counts_match_brute_force():
    """The sparse incidence product counts, for every pair of views, the 3D points both observe - against the
    straightforward set intersection on a random model (with duplicate observations inside a view)."""
    rng = np.random.default_rng(3)
    V, P = 23, 900
    obs = [rng.choice(P, size=rng.integers(0, 200), replace=True).astype(np.int64) for _ in range(V)]
    k = 5
    nbr = covisibility_table(obs, k)
    sets = [set(o.tolist()) for o in obs]
    for s in range(V):
        shared = np.array([len(sets[s] & sets[t]) if t != s else -1 for t in range(V)])
        order = sorted((t for t in range(V) if t != s), key=lambda t: (-shared[t], t))[:k]
        assert list(nbr[s]) == order, (s, list(nbr[s]), order)

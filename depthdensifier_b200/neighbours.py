"""Neighbour-view tables.

The reference has no neighbour selection: every point is tested against every view
(/root/reference/scripts/test.py:275).  ``all_views_table`` reproduces that (K = V, own view
included, image order) and is the parity configuration; ``nearest_views_table`` is the
north-star default for K < V.
"""

from __future__ import annotations

import math

import numpy as np


def camera_centers(cam_from_world: np.ndarray) -> np.ndarray:
    """c = -R^T t for [V,3,4] float64 poses."""
    cam_from_world = np.asarray(cam_from_world, dtype=np.float64)
    R = cam_from_world[:, :, :3]
    t = cam_from_world[:, :, 3]
    return -np.einsum("vji,vj->vi", R, t)


def all_views_table(n_views: int) -> np.ndarray:
    """nbr[s] = [0..V-1]: the reference's all-views semantics."""
    return np.tile(np.arange(n_views, dtype=np.int32), (n_views, 1))


def nearest_views_table(cam_from_world: np.ndarray, k: int) -> np.ndarray:
    """K nearest *other* views by Euclidean distance of camera centres (float64), ties -> lower
    index.  Rows are padded with -1 when V-1 < K."""
    c = camera_centers(cam_from_world)
    V = c.shape[0]
    nbr = np.full((V, k), -1, dtype=np.int32)
    for s in range(V):
        d = np.linalg.norm(c - c[s], axis=1)
        d[s] = np.inf
        order = np.argsort(d, kind="stable")[: min(k, V - 1)]
        nbr[s, : len(order)] = order
    return nbr


def covisibility_table(observed_ids: list[np.ndarray], k: int, cam_from_world: np.ndarray | None = None) -> np.ndarray:
    """K other views sharing the most sparse 3D points with each view (``observed_ids[v]`` = the point3D ids
    view v observes - what the reference collects per image at scripts/test.py:135).  Ties, and views that share
    nothing, fall back to camera-centre distance when poses are given, else to the lower index.  Rows are padded
    with -1 when V-1 < K.  A COLMAP model already knows which views see the same surface; camera distance does
    not (two cameras back to back are close and share nothing)."""
    V = len(observed_ids)
    nbr = np.full((V, k), -1, dtype=np.int32)
    if V == 0:
        return nbr
    # incidence matrix A [V, points] (1 where the view observes the point); shared = A A^T counts, for every pair
    # of views, the tracks that contain both - one sparse product instead of a Python loop over the tracks
    import scipy.sparse as sp

    uniq = [np.unique(np.asarray(o, dtype=np.int64)) for o in observed_ids]
    all_ids = np.concatenate(uniq)
    shared = np.zeros((V, V), dtype=np.int64)
    if len(all_ids):
        owner = np.concatenate([np.full(len(u), v, dtype=np.int64) for v, u in enumerate(uniq)])
        _, col = np.unique(all_ids, return_inverse=True)
        A = sp.csr_matrix((np.ones(len(col), dtype=np.int64), (owner, col)), shape=(V, int(col.max()) + 1))
        shared = np.asarray((A @ A.T).todense(), dtype=np.int64)
    dist = None
    if cam_from_world is not None:
        c = camera_centers(cam_from_world)
        dist = np.linalg.norm(c[:, None] - c[None], axis=2)
    for s in range(V):
        score = shared[s].astype(np.float64)
        score[s] = -1.0
        tie = dist[s] if dist is not None else np.arange(V, dtype=np.float64)
        order = np.lexsort((tie, -score))  # most shared points first, then nearest / lowest index
        order = order[order != s][: min(k, V - 1)]
        nbr[s, : len(order)] = order
    return nbr


def default_vote_threshold(k: int) -> int:
    """ceil(K/2): the reference default of 5 (scripts/test.py:43) is meaningless for K <= 4."""
    return int(math.ceil(k / 2))

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


def default_vote_threshold(k: int) -> int:
    """ceil(K/2): the reference default of 5 (scripts/test.py:43) is meaningless for K <= 4."""
    return int(math.ceil(k / 2))

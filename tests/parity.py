"""Tie-aware comparison of CUDA consistency votes with the float64 oracle.

The reference evaluates stage 3 in float64 (scripts/test.py:58-76, 273-330); the CUDA kernel works in
float32 with float64-precomputed transforms.  The two can only disagree on a (point, view) pair whose
float64 quantities sit within a small band of a decision boundary.  The bands are STATED here:

  EPS_PX   pixel band around image borders and around integer pixel coordinates (truncation)
  EPS_DOT  band around the grazing threshold 0.087
  EPS_REL  relative band around z = depth_threshold * D
  EPS_Z    band around z = 0

A truncation-band pair only counts as a tie if one of the candidate pixels it could fall into
changes the decision.  Everything outside the bands must match bit for bit."""

from __future__ import annotations

import numpy as np

from oracle import restatement as R

EPS_PX = 4e-3
EPS_DOT = 2e-5
EPS_REL = 2e-5
EPS_Z = 1e-5


def pair_ties(points, normals, src_view, t, refined_t, pose_t, intr_t, depth_threshold=0.7, grazing=0.087):
    """Returns (vote[N] bool, tie[N] bool) of all points against target view t (nearest mode)."""
    h, w = refined_t.shape
    vote, det = R.votes_against_view(points, normals, refined_t, pose_t, intr_t, depth_threshold=depth_threshold,
                                     grazing=grazing, return_detail=True)
    u, v, z, dot = det["u"], det["v"], det["z"], det["dot"]
    own = src_view == t
    tie = np.zeros(len(points), dtype=bool)
    near_border = (np.abs(u) < EPS_PX) | (np.abs(u - w) < EPS_PX) | (np.abs(v) < EPS_PX) | (np.abs(v - h) < EPS_PX)
    tie |= near_border & ~own
    tie |= np.abs(z) < EPS_Z
    tie |= np.abs(dot - grazing) < EPS_DOT
    # decision with the looked-up depth D
    thr32 = np.float32(depth_threshold)

    def decide(ui, vi):
        ok = (ui >= 0) & (ui < w) & (vi >= 0) & (vi < h)
        D = np.zeros(len(points), dtype=np.float32)
        D[ok] = refined_t[vi[ok], ui[ok]]
        dec = ok & (D > 0) & (z < (thr32 * D))
        near_thr = ok & (D > 0) & (np.abs(z - thr32 * D.astype(np.float64)) < EPS_REL * D)
        return dec, near_thr

    cand = det["inb"] & ~own
    ui = np.where(cand, u, 0).astype(int)
    vi = np.where(cand, v, 0).astype(int)
    base, near = decide(ui, vi)
    tie |= cand & near
    fu = u - np.floor(u)
    fv = v - np.floor(v)
    for du, dv, sel in (
        (-1, 0, fu < EPS_PX), (1, 0, fu > 1 - EPS_PX), (0, -1, fv < EPS_PX), (0, 1, fv > 1 - EPS_PX),
        (-1, -1, (fu < EPS_PX) & (fv < EPS_PX)), (1, 1, (fu > 1 - EPS_PX) & (fv > 1 - EPS_PX)),
        (-1, 1, (fu < EPS_PX) & (fv > 1 - EPS_PX)), (1, -1, (fu > 1 - EPS_PX) & (fv < EPS_PX)),
    ):
        sel = sel & cand
        if sel.any():
            alt, near_alt = decide(ui + du, vi + dv)
            tie |= sel & ((alt != base) | near_alt)
    # own view: u = x*(1 - 1e-8/z) sits a deterministic ~1e-9 px below the integer x, so the lookup
    # pixel is (x-1, y-1) except in column/row 0 where float64 round-off decides (tie); otherwise
    # only the threshold band applies
    if own.any():
        D = det["D"]
        tie |= own & det["inb"] & (D > 0) & (np.abs(z - thr32 * D.astype(np.float64)) < EPS_REL * D)
        tie |= own & ((np.abs(u) < 0.5) | (np.abs(v) < 0.5))
    return vote, tie


def votes_with_ties(points, normals, src_view, refined_all, poses, intr, nbr, active=None, **kw):
    """Oracle votes[N] and the number of tie pairs per point, honouring the neighbour table."""
    V = refined_all.shape[0]
    votes = np.zeros(len(points), dtype=np.int64)
    nties = np.zeros(len(points), dtype=np.int64)
    member = np.zeros((V, V), dtype=bool)
    for s in range(V):
        for t in nbr[s]:
            if t >= 0:
                member[s, t] = True
    for t in range(V):
        if active is not None and t not in active:
            continue
        sel = np.where(member[src_view, t])[0]
        if len(sel) == 0:
            continue
        vt, tt = pair_ties(points[sel], normals[sel], src_view[sel], t, refined_all[t], poses[t], intr[t], **kw)
        votes[sel] += vt
        nties[sel] += tt
    return votes, nties


def assert_votes_match(gpu_votes, ref_votes, nties, max_tie_fraction=0.02):
    gpu_votes = np.asarray(gpu_votes, dtype=np.int64)
    clean = nties == 0
    bad = clean & (gpu_votes != ref_votes)
    assert not bad.any(), f"{bad.sum()} of {clean.sum()} tie-free points differ; first {np.where(bad)[0][:5]}"
    diff = np.abs(gpu_votes - ref_votes)
    assert (diff <= nties).all(), "vote difference exceeds the number of tie pairs"
    frac = 1.0 - clean.mean() if len(clean) else 0.0
    assert frac <= max_tie_fraction, f"tie fraction {frac:.4f} too large for a meaningful test"
    return frac

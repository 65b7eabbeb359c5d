"""Property tests (hypothesis) of the host-side logic: COLMAP model round trips and the halo-exchange plan."""

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from depthdensifier_b200.colmap_io import CAMERA_MODELS, Camera, Image, Reconstruction, Track
from depthdensifier_b200.distributed import make_halo_plan, shard_bounds

names = st.text(alphabet=st.characters(blacklist_characters="\x00\n\r", blacklist_categories=("Cs",)), min_size=1, max_size=24).map(
    lambda s: s.strip() or "x").filter(lambda s: s == " ".join(s.split()))  # images.txt separates fields by single blanks
finite = st.floats(allow_nan=False, allow_infinity=False, width=64, min_value=-1e6, max_value=1e6)


@st.composite
def models(draw):
    rec = Reconstruction()
    n_cam = draw(st.integers(1, 3))
    for c in range(n_cam):
        mid = draw(st.sampled_from(sorted(CAMERA_MODELS)))
        rec.add_camera(Camera(c + 1, mid, draw(st.integers(1, 5000)), draw(st.integers(1, 5000)),
                              [draw(finite) for _ in range(CAMERA_MODELS[mid][1])]))
    n_pts = draw(st.integers(0, 6))
    for _ in range(n_pts):
        tl = draw(st.integers(0, 3))
        rec.add_point3D([draw(finite) for _ in range(3)], Track(draw(st.lists(st.integers(1, 9), min_size=tl, max_size=tl)),
                                                                 draw(st.lists(st.integers(0, 99), min_size=tl, max_size=tl))),
                        [draw(st.integers(0, 255)) for _ in range(3)])
    for i in range(draw(st.integers(0, 4))):
        n2d = draw(st.integers(0, 5))
        q = np.array([draw(finite) for _ in range(4)])
        if not np.any(q):
            q[0] = 1.0
        rec.add_image(Image(i + 1, q, [draw(finite) for _ in range(3)], draw(st.integers(1, n_cam)), draw(names),
                            [[draw(finite), draw(finite)] for _ in range(n2d)],
                            [draw(st.integers(-1, max(n_pts, 1))) for _ in range(n2d)]))
    if draw(st.booleans()):
        k = draw(st.integers(1, 5))
        rec.add_points3D(np.array([[draw(finite) for _ in range(3)] for _ in range(k)]),
                         np.array([[draw(st.integers(0, 255)) for _ in range(3)] for _ in range(k)], np.uint8))
    return rec


def _same(a, b):
    """b is a re-read of a: identical, except that a short bulk-appended block comes back as ordinary track-less
    points (a long one is bulk-loaded again)."""
    assert list(a.cameras) == list(b.cameras) and list(a.images) == list(b.images)
    assert list(b.points3D)[: len(a.points3D)] == list(a.points3D)
    ax, ac = a.dense_points()
    extra = list(b.points3D)[len(a.points3D):]
    if extra:
        assert b.num_dense_points() == 0 and len(extra) == len(ax)
        assert np.array_equal(np.stack([b.points3D[k].xyz for k in extra]), ax)
        assert np.array_equal(np.stack([b.points3D[k].color for k in extra]), ac)
        assert all(b.points3D[k].track.length() == 0 for k in extra)
    for k in a.cameras:
        assert (a.cameras[k].model_name, a.cameras[k].width, a.cameras[k].height) == (b.cameras[k].model_name, b.cameras[k].width, b.cameras[k].height)
        assert np.array_equal(a.cameras[k].params, b.cameras[k].params)
    for k in a.images:
        x, y = a.images[k], b.images[k]
        assert (x.name, x.camera_id) == (y.name, y.camera_id) and np.array_equal(x.qvec, y.qvec) and np.array_equal(x.tvec, y.tvec)
        assert np.array_equal(x.xys, y.xys) and np.array_equal(x.point3D_ids, y.point3D_ids)
    for k in a.points3D:
        x, y = a.points3D[k], b.points3D[k]
        assert np.array_equal(x.xyz, y.xyz) and np.array_equal(x.color, y.color)
        assert np.array_equal(x.track.image_ids, y.track.image_ids) and np.array_equal(x.track.point2D_idxs, y.track.point2D_idxs)
    return ax, ac


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(rec=models())
def test_colmap_model_round_trips(tmp_path_factory, rec):
    d = tmp_path_factory.mktemp("m")
    rec.write_binary(d / "bin")
    back = Reconstruction(d / "bin")
    ax, ac = _same(rec, back)
    # bulk-appended points come back either as the bulk block (long track-less tail) or as ordinary points
    assert back.num_points3D() == rec.num_points3D()
    rec.write_text(d / "txt")
    again = Reconstruction(d / "txt")
    assert again.num_points3D() == rec.num_points3D() and again.num_images() == rec.num_images()
    for k in rec.images:
        assert again.images[k].name == rec.images[k].name and np.array_equal(again.images[k].qvec, rec.images[k].qvec)
    for k in rec.points3D:
        assert np.array_equal(again.points3D[k].xyz, rec.points3D[k].xyz)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(V=st.integers(1, 40), world=st.integers(1, 6), K=st.integers(1, 6), seed=st.integers(0, 10_000))
def test_halo_plan_is_consistent(V, world, K, seed):
    rng = np.random.default_rng(seed)
    nbr = rng.integers(-1, V, size=(V, K)).astype(np.int32)
    bounds = shard_bounds(V, world)
    assert bounds[0][0] == 0 and bounds[-1][1] == V and all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    plans = [make_halo_plan(nbr, bounds, r) for r in range(world)]
    for r, pl in enumerate(plans):
        lo, hi = bounds[r]
        n_local = hi - lo
        assert list(pl.slots[:n_local]) == list(range(lo, hi))
        halo = list(pl.slots[n_local:])
        assert len(set(halo)) == len(halo) and all(not (lo <= v < hi) for v in halo)
        # every referenced neighbour has a slot holding exactly that view
        for i in range(n_local):
            for k in range(K):
                t = nbr[lo + i, k]
                assert (pl.nbr_slots[i, k] == -1) if t < 0 else (pl.slots[pl.nbr_slots[i, k]] == t)
        # what r expects from q is exactly what q plans to send to r, in the same order
        off = n_local
        for q in range(world):
            cnt = pl.recv_counts[q]
            sent = [bounds[q][0] + int(j) for j in plans[q].send_views[r]] if q != r else []
            assert halo[off - n_local: off - n_local + cnt] == sent
            off += cnt

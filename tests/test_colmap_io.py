"""COLMAP model reader / writer (depthdensifier_b200/colmap_io.py): round trips through both on-disk formats,
the rigid-transform / camera surface the reference uses through pycolmap, and the bulk point append."""

import struct

import numpy as np
import pytest

from depthdensifier_b200.colmap_io import (Camera, Image, Reconstruction, Rigid3d, Track, quat_to_rotmat, rotmat_to_quat)


def _random_rotation(rng):
    q = rng.normal(size=4)
    return quat_to_rotmat(q / np.linalg.norm(q))


def _toy_model(rng, n_images=5, n_points=40):
    rec = Reconstruction()
    rec.add_camera(Camera(1, "PINHOLE", 640, 480, [500.0, 510.0, 320.0, 240.0]))
    rec.add_camera(Camera(7, "SIMPLE_RADIAL", 800, 600, [700.0, 400.0, 300.0, 0.01]))
    pts = rng.normal(size=(n_points, 3))
    tracks = {pid: ([], []) for pid in range(1, n_points + 1)}
    for iid in range(1, n_images + 1):
        R, t = _random_rotation(rng), rng.normal(size=3)
        n2d = int(rng.integers(0, 30))
        xys = rng.uniform(0, 600, size=(n2d, 2))
        ids = rng.integers(1, n_points + 1, size=n2d).astype(np.int64)
        ids[rng.random(n2d) < 0.3] = -1
        for j, pid in enumerate(ids):
            if pid != -1:
                tracks[int(pid)][0].append(iid)
                tracks[int(pid)][1].append(j)
        rec.add_image(Image(iid, rotmat_to_quat(R), t, 1 if iid % 2 else 7, f"dir with space/img_{iid:03d}.jpg", xys, ids))
    for pid in range(1, n_points + 1):
        got = rec.add_point3D(pts[pid - 1], Track(*tracks[pid]), rng.integers(0, 256, 3))
        assert got == pid
    return rec


def _assert_same(a: Reconstruction, b: Reconstruction, exact=True):
    eq = np.array_equal if exact else (lambda x, y: np.allclose(x, y, rtol=1e-15, atol=0))
    assert list(a.cameras) == list(b.cameras) and list(a.images) == list(b.images) and list(a.points3D) == list(b.points3D)
    for k in a.cameras:
        ca, cb = a.cameras[k], b.cameras[k]
        assert (ca.model_name, ca.width, ca.height) == (cb.model_name, cb.width, cb.height) and eq(ca.params, cb.params)
    for k in a.images:
        ia, ib = a.images[k], b.images[k]
        assert (ia.name, ia.camera_id) == (ib.name, ib.camera_id)
        assert eq(ia.qvec, ib.qvec) and eq(ia.tvec, ib.tvec) and eq(ia.xys, ib.xys) and np.array_equal(ia.point3D_ids, ib.point3D_ids)
    for k in a.points3D:
        pa, pb = a.points3D[k], b.points3D[k]
        assert eq(pa.xyz, pb.xyz) and np.array_equal(pa.color, pb.color) and pa.error == pb.error
        assert np.array_equal(pa.track.image_ids, pb.track.image_ids) and np.array_equal(pa.track.point2D_idxs, pb.track.point2D_idxs)


def test_binary_and_text_round_trip(tmp_path):
    rec = _toy_model(np.random.default_rng(0))
    rec.write_binary(tmp_path / "bin")
    back = Reconstruction(tmp_path / "bin")
    _assert_same(rec, back)
    back.write_text(tmp_path / "txt")
    again = Reconstruction(tmp_path / "txt")  # repr(float) round-trips float64 exactly
    _assert_same(rec, again)
    assert back.num_reg_images() == 5 and back.num_points3D() == 40
    # byte-level check of the documented layout: first camera record
    buf = (tmp_path / "bin" / "cameras.bin").read_bytes()
    assert struct.unpack_from("<Q", buf, 0)[0] == 2
    assert struct.unpack_from("<iiQQ", buf, 8) == (1, 1, 640, 480)
    assert struct.unpack_from("<4d", buf, 32) == (500.0, 510.0, 320.0, 240.0)


def test_empty_and_missing(tmp_path):
    Reconstruction().write_binary(tmp_path / "empty")
    e = Reconstruction(tmp_path / "empty")
    assert e.num_cameras() == e.num_images() == e.num_points3D() == 0
    with pytest.raises(FileNotFoundError):
        Reconstruction(tmp_path / "nothing_here")


def test_bulk_append_matches_per_point_append(tmp_path):
    rng = np.random.default_rng(1)
    a, b = _toy_model(np.random.default_rng(2)), _toy_model(np.random.default_rng(2))
    xyz = rng.normal(size=(5000, 3))
    rgb = rng.integers(0, 256, (5000, 3)).astype(np.uint8)
    first, last = a.add_points3D(xyz, rgb)
    assert (first, last) == (41, 5040)
    for p, c in zip(xyz, rgb):  # the reference's loop, scripts/test.py:355-358
        b.add_point3D(xyz=p, track=Track(), color=c)
    a.write_binary(tmp_path / "bulk")
    b.write_binary(tmp_path / "loop")
    assert (tmp_path / "bulk" / "points3D.bin").read_bytes() == (tmp_path / "loop" / "points3D.bin").read_bytes()
    back = Reconstruction(tmp_path / "bulk")  # the track-less tail is bulk-loaded again
    assert back.num_points3D() == 5040 and len(back.points3D) == 40
    dx, dc = back.dense_points()
    assert np.array_equal(dx, xyz) and np.array_equal(dc, rgb)
    assert back.add_point3D([0, 0, 0]) == 5041


def test_rigid3d_and_camera_surface():
    rng = np.random.default_rng(3)
    R, t = _random_rotation(rng), rng.normal(size=3)
    im = Image(1, rotmat_to_quat(R), t, 1, "a.png")
    T = im.cam_from_world()
    assert np.allclose(T.matrix(), np.hstack([R, t[:, None]]), atol=1e-15)
    p = rng.normal(size=(10, 3))
    assert np.allclose(T * p, p @ R.T + t, atol=1e-14)
    assert np.allclose(T.inverse() * (T * p), p, atol=1e-13)
    assert np.allclose(im.projection_center(), -R.T @ t, atol=1e-14)
    assert np.allclose((T * T.inverse()).matrix(), np.eye(4)[:3], atol=1e-14)
    for _ in range(50):  # quaternion <-> matrix in every pivot branch
        Rr = _random_rotation(rng)
        assert np.allclose(quat_to_rotmat(rotmat_to_quat(Rr)), Rr, atol=1e-14)
    cam = Camera(1, "PINHOLE", 1000, 800, [900.0, 880.0, 500.0, 400.0])
    assert np.array_equal(cam.calibration_matrix(), [[900, 0, 500], [0, 880, 400], [0, 0, 1]])
    cam.rescale(new_width=500, new_height=200)
    assert (cam.width, cam.height) == (500, 200) and np.allclose(cam.params, [450.0, 220.0, 250.0, 100.0])
    cam.rescale(new_width=500, new_height=200)  # idempotent at the same size (the reference rescales per image)
    assert np.allclose(cam.params, [450.0, 220.0, 250.0, 100.0])
    sp = Camera(2, "SIMPLE_PINHOLE", 100, 100, [80.0, 50.0, 50.0])
    assert np.array_equal(sp.calibration_matrix(), [[80, 0, 50], [0, 80, 50], [0, 0, 1]])
    pts2d = im.points2D
    assert len(pts2d) == 0
    im2 = Image(2, [1, 0, 0, 0], [0, 0, 0], 1, "b.png", [[1.0, 2.0], [3.0, 4.0]], [5, -1])
    assert [q.has_point3D() for q in im2.points2D] == [True, False] and im2.points2D[0].point3D_id == 5
    assert np.array_equal(im2.observed_point3D_ids(), [5])

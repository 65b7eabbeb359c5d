"""Byte-level pin of the COLMAP binary model format (colmap_io.py), independent of the module's own writer.

The three files are assembled here field by field from the layout COLMAP documents for its binary models
(https://colmap.github.io/format.html#binary-file-format, src/colmap/scene/reconstruction_io.cc): little endian,
  cameras.bin   u64 n | per camera: i32 camera_id, i32 model_id, u64 width, u64 height, f64 params[model]
                (model 1 = PINHOLE fx fy cx cy; 0 = SIMPLE_PINHOLE f cx cy; 2 = SIMPLE_RADIAL f cx cy k)
  images.bin    u64 n | per image: u32 image_id, f64 qw qx qy qz, f64 tx ty tz, u32 camera_id, name + NUL,
                u64 n_points2D, per point: f64 x, f64 y, u64 point3D_id (all ones = none)
  points3D.bin  u64 n | per point: u64 id, f64 x y z, u8 r g b, f64 error, u64 track_length,
                per element: u32 image_id, u32 point2D_idx
The reader must parse these bytes into the values below, and the writer must reproduce them byte for byte."""

import struct

import numpy as np

from depthdensifier_b200.colmap_io import Reconstruction

NONE = 0xFFFFFFFFFFFFFFFF


def u64(v):
    return int(v).to_bytes(8, "little")


def u32(v):
    return int(v).to_bytes(4, "little")


def f64(*vs):
    return b"".join(struct.pack("<d", float(v)) for v in vs)


CAMERAS = (u64(2)
           + u32(1) + u32(1) + u64(640) + u64(480) + f64(500.5, 501.25, 320.0, 240.0)          # PINHOLE
           + u32(7) + u32(2) + u64(1297) + u64(840) + f64(1037.6, 648.5, 420.0, -0.0125))        # SIMPLE_RADIAL
IMAGES = (u64(2)
          + u32(3) + f64(1.0, 0.0, 0.0, 0.0) + f64(0.5, -0.25, 2.0) + u32(1) + b"frame_0003.jpg\x00"
          + u64(3) + f64(10.5, 20.25) + u64(11) + f64(300.0, 100.75) + u64(NONE) + f64(639.0, 479.0) + u64(12)
          + u32(9) + f64(0.5, 0.5, -0.5, 0.5) + f64(-1.0, 0.0, 4.5) + u32(7) + b"sub/dir/b.png\x00"
          + u64(1) + f64(1.0, 2.0) + u64(11))
POINTS = (u64(2)
          + u64(11) + f64(0.1, -0.2, 3.0) + bytes([255, 128, 0]) + f64(0.75) + u64(2) + u32(3) + u32(0) + u32(9) + u32(0)
          + u64(12) + f64(-4.0, 5.5, 6.25) + bytes([1, 2, 3]) + f64(1.5) + u64(1) + u32(3) + u32(2))


def test_reader_parses_documented_bytes(tmp_path):
    for name, blob in (("cameras.bin", CAMERAS), ("images.bin", IMAGES), ("points3D.bin", POINTS)):
        (tmp_path / name).write_bytes(blob)
    rec = Reconstruction(tmp_path)
    assert sorted(rec.cameras) == [1, 7] and sorted(rec.images) == [3, 9] and sorted(rec.points3D) == [11, 12]
    c1, c7 = rec.cameras[1], rec.cameras[7]
    assert (c1.model_name, c1.width, c1.height) == ("PINHOLE", 640, 480) and list(c1.params) == [500.5, 501.25, 320.0, 240.0]
    assert (c7.model_name, c7.width, c7.height) == ("SIMPLE_RADIAL", 1297, 840) and list(c7.params) == [1037.6, 648.5, 420.0, -0.0125]
    assert np.array_equal(c1.calibration_matrix(), [[500.5, 0, 320.0], [0, 501.25, 240.0], [0, 0, 1]])
    im3, im9 = rec.images[3], rec.images[9]
    assert (im3.name, im3.camera_id, im9.name, im9.camera_id) == ("frame_0003.jpg", 1, "sub/dir/b.png", 7)
    assert np.array_equal(im3.cam_from_world().matrix(), [[1, 0, 0, 0.5], [0, 1, 0, -0.25], [0, 0, 1, 2.0]])
    # q = (0.5, 0.5, -0.5, 0.5): R from the Hamilton convention COLMAP uses
    assert np.allclose(im9.cam_from_world().matrix()[:, :3], [[0, -1, 0], [0, 0, -1], [1, 0, 0]], atol=1e-15)
    p2 = im3.points2D
    assert len(p2) == 3 and [p.has_point3D() for p in p2] == [True, False, True]
    assert [int(p.point3D_id) for p in p2 if p.has_point3D()] == [11, 12] and list(p2[1].xy) == [300.0, 100.75]
    assert np.array_equal(rec.sparse_xyz_of_image(im3), [[0.1, -0.2, 3.0], [-4.0, 5.5, 6.25]])
    q = rec.points3D[11]
    assert list(q.xyz) == [0.1, -0.2, 3.0] and [int(c) for c in q.color] == [255, 128, 0] and q.error == 0.75
    assert list(q.track.image_ids) == [3, 9] and list(q.track.point2D_idxs) == [0, 0]


def test_writer_reproduces_documented_bytes(tmp_path):
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    for name, blob in (("cameras.bin", CAMERAS), ("images.bin", IMAGES), ("points3D.bin", POINTS)):
        (src / name).write_bytes(blob)
    rec = Reconstruction(src)
    dst.mkdir()
    rec.write_binary(dst)
    for name, blob in (("cameras.bin", CAMERAS), ("images.bin", IMAGES), ("points3D.bin", POINTS)):
        assert (dst / name).read_bytes() == blob, name
    # dense points appended in bulk follow the documented point record (empty track), ids continue after the largest
    rec.add_points3D(np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]), np.array([[9, 8, 7], [6, 5, 4]], np.uint8))
    rec.write_binary(dst)
    blob = (dst / "points3D.bin").read_bytes()
    assert blob[:8] == u64(4) and blob[8:8 + len(POINTS) - 8] == POINTS[8:]
    tail = blob[len(POINTS):]
    assert tail == (u64(13) + f64(1.0, 2.0, 3.0) + bytes([9, 8, 7]) + f64(-1.0) + u64(0)
                    + u64(14) + f64(4.0, 5.0, 6.0) + bytes([6, 5, 4]) + f64(-1.0) + u64(0))

"""COLMAP sparse-model reader / writer (cameras, images, points3D as .bin or .txt).

The reference loads and saves its scenes through ``pycolmap.Reconstruction`` (scripts/test.py:111, :363;
src/depthdensifier/utils.py:46-51) and touches the objects only through the attributes listed in
SURVEY.md §8(b).  pycolmap is a compiled third-party binding that is not part of this repo; this module
implements the same surface directly on COLMAP's published on-disk formats so the densification driver
(``pipeline.main``) reads COLMAP models "as today":

  rec = Reconstruction(path)        rec.images / rec.cameras / rec.points3D (dict-like, insertion ordered)
  image.cam_from_world()            -> Rigid3d with .matrix() [3,4], .inverse(), ``rigid * points[N,3]``
  image.projection_center(), .has_pose, .name, .image_id, .camera_id, .points2D[i].point3D_id / .has_point3D()
  camera.params, .calibration_matrix(), .rescale(new_width=, new_height=)
  rec.add_point3D(xyz=, track=, color=), rec.write_binary(dir), rec.write_text(dir)

plus the bulk form the reference lacks: ``rec.add_points3D(xyz[N,3], colors[N,3])`` appends a whole fused
cloud without a Python loop (the reference spends seconds in ``add_point3D`` per million points,
scripts/test.py:355-358) and the writers stream it with numpy structured arrays.

Formats (COLMAP documentation, "Output format"): little endian.
  cameras.bin   u64 n; n x { i32 camera_id, i32 model_id, u64 width, u64 height, f64 params[model] }
  images.bin    u64 n; n x { i32 image_id, f64 qw qx qy qz, f64 tx ty tz, i32 camera_id, char name[] NUL,
                             u64 n2d, n2d x { f64 x, f64 y, i64 point3D_id (-1 = none) } }
  points3D.bin  u64 n; n x { u64 id, f64 x y z, u8 r g b, f64 error, u64 track_len,
                             track_len x { i32 image_id, i32 point2D_idx } }
COLMAP >= 3.12 additionally writes rigs/frames files; models without them load as trivial rigs there, and
this module neither needs nor writes them.
"""

from __future__ import annotations

import os
import struct
from pathlib import Path

import numpy as np

# model_id -> (name, number of parameters)   (COLMAP src/colmap/sensor/models.h)
CAMERA_MODELS = {
    0: ("SIMPLE_PINHOLE", 3),
    1: ("PINHOLE", 4),
    2: ("SIMPLE_RADIAL", 4),
    3: ("RADIAL", 5),
    4: ("OPENCV", 8),
    5: ("OPENCV_FISHEYE", 8),
    6: ("FULL_OPENCV", 12),
    7: ("FOV", 5),
    8: ("SIMPLE_RADIAL_FISHEYE", 4),
    9: ("RADIAL_FISHEYE", 5),
    10: ("THIN_PRISM_FISHEYE", 12),
}
CAMERA_MODEL_IDS = {name: mid for mid, (name, _) in CAMERA_MODELS.items()}
# models whose first parameters are (f, cx, cy) rather than (fx, fy, cx, cy)
_SINGLE_FOCAL = {"SIMPLE_PINHOLE", "SIMPLE_RADIAL", "RADIAL", "SIMPLE_RADIAL_FISHEYE", "RADIAL_FISHEYE"}

INVALID_POINT3D_ID = -1


def quat_to_rotmat(q) -> np.ndarray:
    """(qw, qx, qy, qz) -> 3x3 rotation matrix (the quaternion is normalised first)."""
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def rotmat_to_quat(R) -> np.ndarray:
    """3x3 rotation matrix -> (qw, qx, qy, qz) with qw >= 0 (largest-pivot branch for stability)."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        v = [0.0, 0.0, 0.0]
        v[i] = 0.25 * s
        v[j] = (R[j, i] + R[i, j]) / s
        v[k] = (R[k, i] + R[i, k]) / s
        q = [(R[k, j] - R[j, k]) / s, *v]
    q = np.array(q)
    return -q if q[0] < 0 else q


class Rigid3d:
    """Rigid transform x -> R x + t (what ``pycolmap.Rigid3d`` is to the reference: scripts/test.py:63, :233)."""

    __slots__ = ("R", "t")

    def __init__(self, rotation=None, translation=None):
        self.R = np.eye(3) if rotation is None else np.asarray(rotation, dtype=np.float64).reshape(3, 3)
        self.t = np.zeros(3) if translation is None else np.asarray(translation, dtype=np.float64).reshape(3)

    @classmethod
    def from_quat(cls, qvec, tvec):
        return cls(quat_to_rotmat(qvec), tvec)

    def matrix(self) -> np.ndarray:
        return np.hstack([self.R, self.t[:, None]])

    def inverse(self) -> "Rigid3d":
        return Rigid3d(self.R.T, -self.R.T @ self.t)

    def quat(self) -> np.ndarray:
        return rotmat_to_quat(self.R)

    def __mul__(self, other):
        if isinstance(other, Rigid3d):
            return Rigid3d(self.R @ other.R, self.R @ other.t + self.t)
        p = np.asarray(other, dtype=np.float64)
        return p @ self.R.T + self.t


class Camera:
    def __init__(self, camera_id: int, model, width: int, height: int, params):
        self.camera_id = int(camera_id)
        self.model_id = CAMERA_MODEL_IDS[model] if isinstance(model, str) else int(model)
        if self.model_id not in CAMERA_MODELS:
            raise ValueError(f"unknown COLMAP camera model id {self.model_id}")
        self.width, self.height = int(width), int(height)
        self.params = np.asarray(params, dtype=np.float64).copy()
        if len(self.params) != CAMERA_MODELS[self.model_id][1]:
            raise ValueError(f"camera model {self.model_name} takes {CAMERA_MODELS[self.model_id][1]} parameters")

    @property
    def model_name(self) -> str:
        return CAMERA_MODELS[self.model_id][0]

    # pycolmap spells it both ways depending on the version
    model = property(lambda self: self.model_name)

    def _focal_pp(self):
        if self.model_name in _SINGLE_FOCAL:
            f, cx, cy = self.params[:3]
            return f, f, cx, cy
        return tuple(self.params[:4])

    def calibration_matrix(self) -> np.ndarray:
        fx, fy, cx, cy = self._focal_pp()
        return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])

    def rescale(self, new_width: int | None = None, new_height: int | None = None, scale: float | None = None) -> None:
        """COLMAP ``Camera::Rescale``: focal lengths and principal point scale with the image size (in place,
        as the reference relies on at scripts/test.py:172-173)."""
        if scale is not None:
            new_width, new_height = round(self.width * scale), round(self.height * scale)
        sx, sy = new_width / self.width, new_height / self.height
        self.width, self.height = int(new_width), int(new_height)
        if self.model_name in _SINGLE_FOCAL:
            self.params[0] *= (sx + sy) / 2.0
            self.params[1] *= sx
            self.params[2] *= sy
        else:
            self.params[0] *= sx
            self.params[1] *= sy
            self.params[2] *= sx
            self.params[3] *= sy


class Point2D:
    __slots__ = ("xy", "point3D_id")

    def __init__(self, xy, point3D_id: int = INVALID_POINT3D_ID):
        self.xy = xy
        self.point3D_id = int(point3D_id)

    def has_point3D(self) -> bool:
        return self.point3D_id != INVALID_POINT3D_ID


class _Points2D:
    """List-like view over an image's observation arrays (objects are made on access)."""

    def __init__(self, xys, ids):
        self._xys, self._ids = xys, ids

    def __len__(self):
        return len(self._ids)

    def __getitem__(self, i):
        return Point2D(self._xys[i], self._ids[i])

    def __iter__(self):
        for xy, pid in zip(self._xys, self._ids):
            yield Point2D(xy, pid)


class Image:
    def __init__(self, image_id: int, qvec, tvec, camera_id: int, name: str, xys=None, point3D_ids=None, has_pose: bool = True):
        self.image_id, self.camera_id, self.name = int(image_id), int(camera_id), str(name)
        self.qvec = np.asarray(qvec, dtype=np.float64).copy()
        self.tvec = np.asarray(tvec, dtype=np.float64).copy()
        self.xys = np.zeros((0, 2)) if xys is None else np.asarray(xys, dtype=np.float64).reshape(-1, 2)
        self.point3D_ids = (np.zeros(0, np.int64) if point3D_ids is None else np.asarray(point3D_ids, dtype=np.int64).reshape(-1))
        self.has_pose = bool(has_pose)

    def cam_from_world(self) -> Rigid3d:
        return Rigid3d.from_quat(self.qvec, self.tvec)

    def projection_center(self) -> np.ndarray:
        R = quat_to_rotmat(self.qvec)
        return -R.T @ self.tvec

    @property
    def points2D(self) -> _Points2D:
        return _Points2D(self.xys, self.point3D_ids)

    def observed_point3D_ids(self) -> np.ndarray:
        """ids of the 3D points this image observes, in points2D order (vectorised form of the list
        comprehension at scripts/test.py:135)."""
        return self.point3D_ids[self.point3D_ids != INVALID_POINT3D_ID]


class Track:
    def __init__(self, image_ids=None, point2D_idxs=None):
        self.image_ids = np.zeros(0, np.int32) if image_ids is None else np.asarray(image_ids, dtype=np.int32)
        self.point2D_idxs = np.zeros(0, np.int32) if point2D_idxs is None else np.asarray(point2D_idxs, dtype=np.int32)

    def length(self) -> int:
        return len(self.image_ids)


class Point3D:
    __slots__ = ("xyz", "color", "error", "track")

    def __init__(self, xyz, color=(0, 0, 0), error: float = -1.0, track: Track | None = None):
        self.xyz = np.asarray(xyz, dtype=np.float64)
        self.color = np.asarray(color, dtype=np.uint8)
        self.error = float(error)
        self.track = track if track is not None else Track()


_DENSE_DTYPE = np.dtype([("id", "<u8"), ("xyz", "<f8", 3), ("rgb", "u1", 3), ("error", "<f8"), ("track_len", "<u8")])
assert _DENSE_DTYPE.itemsize == 51


class Reconstruction:
    """A COLMAP sparse model.  ``Reconstruction(path)`` reads ``path`` (a directory holding cameras/images/
    points3D as .bin or .txt) like ``pycolmap.Reconstruction(path)`` (scripts/test.py:111)."""

    def __init__(self, path=None):
        self.cameras: dict[int, Camera] = {}
        self.images: dict[int, Image] = {}
        self.points3D: dict[int, Point3D] = {}
        self._dense_xyz: list[np.ndarray] = []  # bulk-appended clouds (track-less points)
        self._dense_rgb: list[np.ndarray] = []
        self._dense_first: list[int] = []  # id of the first point of each bulk block
        self._next_point3D_id = 1
        if path is not None:
            self.read(path)

    # -- counts --------------------------------------------------------------------------------------
    def num_reg_images(self) -> int:
        return sum(1 for im in self.images.values() if im.has_pose)

    def num_images(self) -> int:
        return len(self.images)

    def num_cameras(self) -> int:
        return len(self.cameras)

    def num_dense_points(self) -> int:
        return int(sum(len(a) for a in self._dense_xyz))

    def num_points3D(self) -> int:
        return len(self.points3D) + self.num_dense_points()

    # -- editing -------------------------------------------------------------------------------------
    def add_camera(self, camera: Camera) -> None:
        self.cameras[camera.camera_id] = camera

    def add_image(self, image: Image) -> None:
        self.images[image.image_id] = image

    def add_point3D(self, xyz, track: Track | None = None, color=(0, 0, 0)) -> int:
        pid = self._next_point3D_id
        self._next_point3D_id += 1
        self.points3D[pid] = Point3D(xyz, color, -1.0, track)
        return pid

    def add_points3D(self, xyz, colors) -> tuple[int, int]:
        """Bulk append of track-less points (the fused dense cloud).  Returns the id range [first, last]."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        colors = np.ascontiguousarray(colors, dtype=np.uint8).reshape(-1, 3)
        if len(xyz) != len(colors):
            raise ValueError("xyz and colors differ in length")
        first = self._next_point3D_id
        self._next_point3D_id += len(xyz)
        self._dense_xyz.append(xyz)
        self._dense_rgb.append(colors)
        self._dense_first.append(first)
        return first, self._next_point3D_id - 1

    def dense_points(self) -> tuple[np.ndarray, np.ndarray]:
        if not self._dense_xyz:
            return np.zeros((0, 3)), np.zeros((0, 3), np.uint8)
        return np.concatenate(self._dense_xyz), np.concatenate(self._dense_rgb)

    def sparse_xyz_of_image(self, image: Image) -> np.ndarray:
        """[C,3] float64 world points observed by ``image`` in points2D order (scripts/test.py:135-139)."""
        ids = image.observed_point3D_ids()
        if len(ids) == 0:
            return np.zeros((0, 3))
        return np.stack([self.points3D[int(i)].xyz for i in ids])

    # -- reading -------------------------------------------------------------------------------------
    def read(self, path) -> None:
        path = Path(path)
        if (path / "cameras.bin").exists():
            self.read_binary(path)
        elif (path / "cameras.txt").exists():
            self.read_text(path)
        else:
            raise FileNotFoundError(f"no COLMAP model (cameras.bin / cameras.txt) under {path}")

    def read_binary(self, path) -> None:
        path = Path(path)
        buf = (path / "cameras.bin").read_bytes()
        (n,), off = struct.unpack_from("<Q", buf, 0), 8
        for _ in range(n):
            cid, mid, w, h = struct.unpack_from("<iiQQ", buf, off)
            off += 24
            if mid not in CAMERA_MODELS:
                raise ValueError(f"cameras.bin: unknown camera model id {mid}")
            npar = CAMERA_MODELS[mid][1]
            params = np.frombuffer(buf, "<f8", npar, off)
            off += 8 * npar
            self.cameras[cid] = Camera(cid, mid, w, h, params)

        buf = (path / "images.bin").read_bytes()
        (n,), off = struct.unpack_from("<Q", buf, 0), 8
        obs = np.dtype([("xy", "<f8", 2), ("pid", "<i8")])
        for _ in range(n):
            iid, qw, qx, qy, qz, tx, ty, tz, cid = struct.unpack_from("<idddddddi", buf, off)
            off += 64
            end = buf.index(b"\x00", off)
            name = buf[off:end].decode("utf-8")
            off = end + 1
            (n2d,) = struct.unpack_from("<Q", buf, off)
            off += 8
            rec = np.frombuffer(buf, obs, n2d, off)
            off += 24 * n2d
            self.images[iid] = Image(iid, (qw, qx, qy, qz), (tx, ty, tz), cid, name, rec["xy"].copy(), rec["pid"].copy())

        buf = (path / "points3D.bin").read_bytes()
        (n,), off = struct.unpack_from("<Q", buf, 0), 8
        max_id = 0
        for i in range(n):
            # a previously densified model ends in millions of track-less 51-byte records: bulk-load them
            if (n - i) * 51 == len(buf) - off and n - i > 1024:
                rec = np.frombuffer(buf, _DENSE_DTYPE, n - i, off)
                if not rec["track_len"].any() and np.array_equal(rec["id"], np.arange(rec["id"][0], rec["id"][0] + len(rec), dtype=np.uint64)):
                    self._dense_xyz.append(rec["xyz"].copy())
                    self._dense_rgb.append(rec["rgb"].copy())
                    self._dense_first.append(int(rec["id"][0]))
                    max_id = max(max_id, int(rec["id"][-1]))
                    break
            pid, x, y, z, r, g, b, err = struct.unpack_from("<QdddBBBd", buf, off)
            off += 43
            (tl,) = struct.unpack_from("<Q", buf, off)
            off += 8
            tr = np.frombuffer(buf, "<i4", 2 * tl, off).reshape(-1, 2)
            off += 8 * tl
            self.points3D[pid] = Point3D((x, y, z), (r, g, b), err, Track(tr[:, 0].copy(), tr[:, 1].copy()))
            max_id = max(max_id, pid)
        self._next_point3D_id = max_id + 1

    def read_text(self, path) -> None:
        path = Path(path)

        def lines(name):
            with open(path / name, "r", encoding="utf-8") as f:
                for ln in f:
                    if ln.startswith("#"):
                        continue
                    yield ln.rstrip("\n")

        for ln in lines("cameras.txt"):
            if not ln.strip():
                continue
            el = ln.split()
            self.cameras[int(el[0])] = Camera(int(el[0]), el[1], int(el[2]), int(el[3]), [float(v) for v in el[4:]])
        it = lines("images.txt")
        for ln in it:
            if not ln.strip():
                continue
            el = ln.split()
            iid, q, t, cid, name = int(el[0]), [float(v) for v in el[1:5]], [float(v) for v in el[5:8]], int(el[8]), " ".join(el[9:])
            o = next(it, "").split()  # the observation line may be empty
            xs = np.array(o[0::3], dtype=np.float64)
            ys = np.array(o[1::3], dtype=np.float64)
            ids = np.array(o[2::3], dtype=np.int64)
            self.images[iid] = Image(iid, q, t, cid, name, np.stack([xs, ys], 1) if len(xs) else None, ids)
        max_id = 0
        for ln in lines("points3D.txt"):
            if not ln.strip():
                continue
            el = ln.split()
            pid = int(el[0])
            tr = np.array(el[8:], dtype=np.int64).reshape(-1, 2)
            self.points3D[pid] = Point3D([float(v) for v in el[1:4]], [int(v) for v in el[4:7]], float(el[7]),
                                        Track(tr[:, 0], tr[:, 1]))
            max_id = max(max_id, pid)
        self._next_point3D_id = max_id + 1

    # -- writing -------------------------------------------------------------------------------------
    def write(self, path) -> None:
        self.write_binary(path)

    def write_binary(self, path) -> None:
        path = Path(path)
        os.makedirs(path, exist_ok=True)
        with open(path / "cameras.bin", "wb") as f:
            f.write(struct.pack("<Q", len(self.cameras)))
            for c in self.cameras.values():
                f.write(struct.pack("<iiQQ", c.camera_id, c.model_id, c.width, c.height))
                f.write(np.asarray(c.params, "<f8").tobytes())
        obs = np.dtype([("xy", "<f8", 2), ("pid", "<i8")])
        with open(path / "images.bin", "wb") as f:
            f.write(struct.pack("<Q", len(self.images)))
            for im in self.images.values():
                f.write(struct.pack("<idddddddi", im.image_id, *im.qvec, *im.tvec, im.camera_id))
                f.write(im.name.encode("utf-8") + b"\x00")
                f.write(struct.pack("<Q", len(im.point3D_ids)))
                rec = np.empty(len(im.point3D_ids), obs)
                rec["xy"], rec["pid"] = im.xys, im.point3D_ids
                f.write(rec.tobytes())
        with open(path / "points3D.bin", "wb") as f:
            f.write(struct.pack("<Q", self.num_points3D()))
            for pid, p in self.points3D.items():
                f.write(struct.pack("<QdddBBBd", pid, *p.xyz, *(int(c) for c in p.color), p.error))
                f.write(struct.pack("<Q", p.track.length()))
                tr = np.empty((p.track.length(), 2), "<i4")
                tr[:, 0], tr[:, 1] = p.track.image_ids, p.track.point2D_idxs
                f.write(tr.tobytes())
            # bulk-appended points: fixed 51-byte records (no track), streamed in slabs
            for first, xyz, rgb in zip(self._dense_first, self._dense_xyz, self._dense_rgb):
                for s in range(0, len(xyz), 1 << 20):
                    e = min(s + (1 << 20), len(xyz))
                    rec = np.zeros(e - s, _DENSE_DTYPE)
                    rec["id"] = np.arange(first + s, first + e, dtype=np.uint64)
                    rec["xyz"], rec["rgb"], rec["error"] = xyz[s:e], rgb[s:e], -1.0
                    f.write(rec.tobytes())

    def write_text(self, path) -> None:
        path = Path(path)
        os.makedirs(path, exist_ok=True)
        with open(path / "cameras.txt", "w", encoding="utf-8") as f:
            f.write("# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n")
            f.write(f"# Number of cameras: {len(self.cameras)}\n")
            for c in self.cameras.values():
                f.write(f"{c.camera_id} {c.model_name} {c.width} {c.height} " + " ".join(repr(float(v)) for v in c.params) + "\n")
        with open(path / "images.txt", "w", encoding="utf-8") as f:
            f.write("# Image list with two lines of data per image:\n#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n"
                    "#   POINTS2D[] as (X, Y, POINT3D_ID)\n")
            f.write(f"# Number of images: {len(self.images)}\n")
            for im in self.images.values():
                f.write(f"{im.image_id} " + " ".join(repr(float(v)) for v in (*im.qvec, *im.tvec)) + f" {im.camera_id} {im.name}\n")
                f.write(" ".join(f"{repr(float(x))} {repr(float(y))} {int(p)}" for (x, y), p in zip(im.xys, im.point3D_ids)) + "\n")
        with open(path / "points3D.txt", "w", encoding="utf-8") as f:
            f.write("# 3D point list with one line of data per point:\n"
                    "#   POINT3D_ID, X, Y, Z, R, G, B, ERROR, TRACK[] as (IMAGE_ID, POINT2D_IDX)\n")
            f.write(f"# Number of points: {self.num_points3D()}\n")
            for pid, p in self.points3D.items():
                tr = " ".join(f"{int(a)} {int(b)}" for a, b in zip(p.track.image_ids, p.track.point2D_idxs))
                f.write(f"{pid} " + " ".join(repr(float(v)) for v in p.xyz) + " " + " ".join(str(int(c)) for c in p.color)
                        + f" {repr(p.error)}" + (f" {tr}" if tr else "") + "\n")
            for first, xyz, rgb in zip(self._dense_first, self._dense_xyz, self._dense_rgb):
                for i in range(len(xyz)):
                    f.write(f"{first + i} " + " ".join(repr(float(v)) for v in xyz[i]) + " "
                            + " ".join(str(int(c)) for c in rgb[i]) + " -1.0\n")

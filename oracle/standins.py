"""TEST INFRASTRUCTURE ONLY - stand-in ``pycolmap`` and ``moge.model.v2`` modules.

pycolmap 3.12.5 and MoGe 2.0.0 are pinned by the reference (uv.lock:1280-1281, :865-867) but are
not installable here.  These stand-ins expose exactly the duck-typed surface that the reference's
``scripts/test.py`` touches (SURVEY.md §8b "Callers"), so that its UNMODIFIED ``main()`` runs on a
synthetic scene.  All SE(3)/pinhole arithmetic is closed-form float64, as in COLMAP's Rigid3d.

Nothing in the product package imports this file.
"""

from __future__ import annotations

import sys
import types

import numpy as np
import torch


class Rigid3d:
    def __init__(self, R: np.ndarray, t: np.ndarray):
        self.R = np.asarray(R, dtype=np.float64)
        self.t = np.asarray(t, dtype=np.float64)

    def matrix(self) -> np.ndarray:
        return np.concatenate([self.R, self.t[:, None]], axis=1)

    def inverse(self) -> "Rigid3d":
        return Rigid3d(self.R.T, -self.R.T @ self.t)

    def __mul__(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        return pts @ self.R.T + self.t


class Point2D:
    def __init__(self, point3D_id: int):
        self.point3D_id = point3D_id

    def has_point3D(self) -> bool:
        return self.point3D_id >= 0


class Image:
    def __init__(self, image_id, name, camera_id, R, t, point3D_ids, has_pose=True):
        self.image_id = image_id
        self.name = name
        self.camera_id = camera_id
        self._pose = Rigid3d(R, t)
        self.points2D = [Point2D(int(p)) for p in point3D_ids]
        self.has_pose = has_pose

    def cam_from_world(self) -> Rigid3d:
        return self._pose

    def projection_center(self) -> np.ndarray:
        return -self._pose.R.T @ self._pose.t


class Camera:
    def __init__(self, camera_id, width, height, params):
        self.camera_id = camera_id
        self.width = width
        self.height = height
        self.params = np.asarray(params, dtype=np.float64)  # fx, fy, cx, cy (PINHOLE)

    def calibration_matrix(self) -> np.ndarray:
        fx, fy, cx, cy = self.params
        return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)

    def rescale(self, new_width: int, new_height: int) -> None:
        sx = new_width / self.width
        sy = new_height / self.height
        fx, fy, cx, cy = self.params
        self.params = np.array([fx * sx, fy * sy, cx * sx, cy * sy])
        self.width, self.height = new_width, new_height


class Point3D:
    def __init__(self, xyz, color=None):
        self.xyz = np.asarray(xyz, dtype=np.float64)
        self.color = color


class Track:
    pass


class Reconstruction:
    """Built in memory by the fixture and handed to ``main`` through the ``_REGISTRY`` keyed by
    the ``recon_path`` string (the reference calls ``pycolmap.Reconstruction(path)``)."""

    _REGISTRY: dict[str, "Reconstruction"] = {}

    def __new__(cls, path=None):
        if path is not None and str(path) in cls._REGISTRY:
            return cls._REGISTRY[str(path)]
        return super().__new__(cls)

    def __init__(self, path=None):
        if getattr(self, "_init", False):
            return
        self._init = True
        self.images: dict[int, Image] = {}
        self.cameras: dict[int, Camera] = {}
        self.points3D: dict[int, Point3D] = {}
        self.added_xyz: list = []
        self.added_rgb: list = []
        self.written_to = None

    def num_reg_images(self) -> int:
        return sum(1 for i in self.images.values() if i.has_pose)

    def num_points3D(self) -> int:
        return len(self.points3D) + len(self.added_xyz)

    def add_point3D(self, xyz, track, color):
        self.added_xyz.append(xyz)
        self.added_rgb.append(color)
        return len(self.points3D) + len(self.added_xyz)

    def write_binary(self, path: str) -> None:
        self.written_to = path


class FakeMoGeModel:
    """``infer`` pops pre-computed synthetic outputs in call order (one per image)."""

    queue: list = []

    @classmethod
    def from_pretrained(cls, path):
        return cls()

    def to(self, device):
        return self

    def eval(self):
        return self

    def infer(self, img_tensor):
        depth, normal, mask = FakeMoGeModel.queue.pop(0)
        return {
            "depth": torch.from_numpy(depth)[None],
            "normal": torch.from_numpy(normal)[None],
            "mask": torch.from_numpy(mask)[None],
        }


def install() -> None:
    """Inject the stand-ins into ``sys.modules`` (idempotent)."""
    pc = types.ModuleType("pycolmap")
    for name in ("Rigid3d", "Image", "Camera", "Point3D", "Track", "Reconstruction", "Point2D"):
        setattr(pc, name, globals()[name])
    sys.modules["pycolmap"] = pc
    moge = types.ModuleType("moge")
    moge_model = types.ModuleType("moge.model")
    moge_v2 = types.ModuleType("moge.model.v2")
    moge_v2.MoGeModel = FakeMoGeModel
    moge.model = moge_model
    moge_model.v2 = moge_v2
    sys.modules["moge"] = moge
    sys.modules["moge.model"] = moge_model
    sys.modules["moge.model.v2"] = moge_v2

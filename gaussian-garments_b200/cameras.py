"""Synthetic pinhole cameras that reproduce the reference's matrix conventions.

Restates (device-agnostic, no image payload) what the reference computes in
  scene/cameras.py:53-62            znear/zfar, world_view_transform, projection,
                                    full_proj_transform, camera_center
  utils/graphics_utils.py:38-49     getWorld2View2 (translate=0, scale=1)
  utils/graphics_utils.py:51-75     getProjectionMatrix (off-centre frustum from fx,fy,cx,cy)
  utils/graphics_utils.py:77-81     fov2focal / focal2fov
  scene/dataloader.py:175           R = W2C[:3,:3].T , T = W2C[:3,3]
so that a `PinholeCamera` can be handed to the reference's `render()` facade
(`gaussian_renderer/__init__.py:36-50` reads FoVx, FoVy, image_height,
image_width, world_view_transform, full_proj_transform, camera_center).

Camera frame: +Z forward, +X right, +Y down (OpenCV / COLMAP).
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


def focal2fov(focal: float, pixels: float) -> float:
    # utils/graphics_utils.py:80-81
    return 2.0 * math.atan(pixels / (2.0 * focal))


def fov2focal(fov: float, pixels: float) -> float:
    # utils/graphics_utils.py:77-78
    return pixels / (2.0 * math.tan(fov / 2.0))


def world_to_view(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """4x4 W2C, float32; R is the *transposed* W2C rotation (reference convention)."""
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = R.T
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return Rt.astype(np.float32)


def projection_matrix(znear, zfar, fx, fy, cx, cy, w, h) -> torch.Tensor:
    """Off-centre perspective matrix, same entries as utils/graphics_utils.py:51-75."""
    top = cy / fy * znear
    bottom = -(h - cy) / fy * znear
    right = cx / fx * znear
    left = -(w - cx) / fx * znear
    P = torch.zeros(4, 4, dtype=torch.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class PinholeCamera:
    """Attribute-compatible stand-in for scene/cameras.py:Camera (render()-facing part)."""

    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    fx: float
    fy: float
    cx: float
    cy: float
    R: np.ndarray
    T: np.ndarray
    world_view_transform: torch.Tensor
    projection_matrix: torch.Tensor
    full_proj_transform: torch.Tensor
    camera_center: torch.Tensor
    znear: float = 0.01
    zfar: float = 100.0

    def to(self, device) -> "PinholeCamera":
        return PinholeCamera(
            self.image_width, self.image_height, self.FoVx, self.FoVy, self.fx, self.fy,
            self.cx, self.cy, self.R, self.T,
            self.world_view_transform.to(device), self.projection_matrix.to(device),
            self.full_proj_transform.to(device), self.camera_center.to(device),
            self.znear, self.zfar,
        )

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)  # gaussian_renderer/__init__.py:36

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)  # gaussian_renderer/__init__.py:37


def make_camera(R, T, fx, fy, cx, cy, width, height, znear=0.01, zfar=100.0) -> PinholeCamera:
    R = np.asarray(R, dtype=np.float64)
    T = np.asarray(T, dtype=np.float64)
    w2c = world_to_view(R, T)
    wvt = torch.tensor(w2c).transpose(0, 1).contiguous()            # scene/cameras.py:59
    proj = projection_matrix(znear, zfar, fx, fy, cx, cy, width, height).transpose(0, 1).contiguous()
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()   # scene/cameras.py:61
    center = wvt.inverse()[3, :3].contiguous()                       # scene/cameras.py:62
    return PinholeCamera(
        image_width=int(width), image_height=int(height),
        FoVx=focal2fov(fx, width), FoVy=focal2fov(fy, height),
        fx=float(fx), fy=float(fy), cx=float(cx), cy=float(cy), R=R, T=T,
        world_view_transform=wvt, projection_matrix=proj, full_proj_transform=full,
        camera_center=center, znear=znear, zfar=zfar,
    )


def look_at_camera(eye, target, fovy_deg, width, height, cx_jitter=0.0, cy_jitter=0.0,
                   up=(0.0, 1.0, 0.0)) -> PinholeCamera:
    """Camera at `eye` looking at `target`; world +Y is up, camera +Y is down."""
    eye = np.asarray(eye, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    upv = np.asarray(up, dtype=np.float64)
    right = np.cross(fwd, upv)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    w2c_rot = np.stack([right, down, fwd], axis=0)                   # rows = camera axes in world
    T = -w2c_rot @ eye
    R = w2c_rot.T                                                    # reference stores the transpose
    fy = fov2focal(math.radians(fovy_deg), height)
    fx = fy
    return make_camera(R, T, fx, fy, width / 2.0 + cx_jitter, height / 2.0 + cy_jitter, width, height)


def ring_cameras(n, radius=3.0, height_y=0.6, target=(0.0, 0.6, 0.0), fovy_deg=30.0,
                 width=1920, height=1080, seed=31359, jitter_px=5.0, phase=0.0):
    """ring of cameras around the garment cylinder."""
    rng = np.random.RandomState(seed)
    cams = []
    for i in range(n):
        a = phase + 2.0 * math.pi * i / max(n, 1)
        eye = (radius * math.sin(a), height_y, radius * math.cos(a))
        jx, jy = rng.uniform(-jitter_px, jitter_px, size=2)
        cams.append(look_at_camera(eye, target, fovy_deg, width, height, jx, jy))
    return cams

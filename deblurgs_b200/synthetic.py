"""Seeded synthetic scenes and camera trajectories (the reference ships no data or benchmark
inputs; this is the workload definition of SURVEY.md 8(d) / BASELINE.md 4).

Everything is drawn on the CPU from `torch.Generator().manual_seed(seed)` so that every rank /
every implementation sees bit-identical inputs, then moved to the requested device.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

SH_C0 = 0.28209479177387814

CONFIGS = {
    # name: (P, W, H, F, curve_order)
    "c1": (50_000, 256, 256, 4, 3),
    "c2": (300_000, 600, 400, 16, 9),
    "c3": (1_000_000, 1920, 1080, 16, 9),
    "c4": (3_000_000, 1280, 720, 32, 9),
    "tiny": (2_000, 96, 64, 3, 3),
    "small": (20_000, 200, 136, 5, 3),
}


def get_config(name):
    """(P, W, H, F, curve_order) of a named config, or of an ad-hoc "P=<n>,F=<n>[,W=<n>,H=<n>,C=<n>]" string
    (sweep points of BASELINE.json config 5; defaults are c2's image size and curve order)."""
    if name in CONFIGS:
        return CONFIGS[name]
    kv = dict(item.split("=") for item in name.split(","))
    return (int(kv["P"]), int(kv.get("W", 600)), int(kv.get("H", 400)), int(kv["F"]), int(kv.get("C", 9)))


@dataclass
class Camera:
    """Pinhole intrinsics shared by all sub-frames (scene/motion.py:178 uses original_cam[0])."""
    width: int
    height: int
    fovx: float
    fovy: float
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def tanfovx(self):
        return math.tan(self.fovx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.fovy * 0.5)

    def projection_matrix_t(self):
        """getProjectionMatrix(...).transpose(0,1) of the reference (utils/graphics_utils.py:51-71,
        scene/cameras.py:58), as a CPU fp32 [4,4] tensor."""
        tan_y, tan_x = math.tan(self.fovy / 2), math.tan(self.fovx / 2)
        top, right = tan_y * self.znear, tan_x * self.znear
        bottom, left = -top, -right
        P = torch.zeros(4, 4)
        P[0, 0] = 2.0 * self.znear / (right - left)
        P[1, 1] = 2.0 * self.znear / (top - bottom)
        P[0, 2] = (right + left) / (right - left)
        P[1, 2] = (top + bottom) / (top - bottom)
        P[3, 2] = 1.0
        P[2, 2] = self.zfar / (self.zfar - self.znear)
        P[2, 3] = -(self.zfar * self.znear) / (self.zfar - self.znear)
        return P.transpose(0, 1).contiguous()


def make_camera(width, height, focal_factor=1.2):
    focal = focal_factor * width
    return Camera(width, height, 2 * math.atan(width / (2 * focal)), 2 * math.atan(height / (2 * focal)))


def _se3_exp_np(v):
    """fp64 numpy se3 exponential, same clamped-theta formula as the reference."""
    u, w = np.asarray(v[:3], np.float64), np.asarray(v[3:], np.float64)
    th = math.sqrt(max(float(w @ w), 1e-4))
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], np.float64)
    K2 = K @ K
    R = np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * K2
    V = np.eye(3) + (1 - math.cos(th)) / th ** 2 * K + (th - math.sin(th)) / th ** 3 * K2
    return R, V @ u


BASE_SE3 = (0.3, -0.2, 0.1, 0.12, -0.25, 0.08)   # [u | omega] of the base camera-to-world pose


@dataclass
class Scene:
    means3D: torch.Tensor      # [P,3]
    scales: torch.Tensor       # [P,3]  (activated, > 0)
    rotations: torch.Tensor    # [P,4]  (unit quaternions r,x,y,z)
    opacities: torch.Tensor    # [P,1]
    shs: torch.Tensor          # [P,M,3]
    sh_degree: int

    def to(self, device):
        return Scene(*(t.to(device) if torch.is_tensor(t) else t for t in
                       (self.means3D, self.scales, self.rotations, self.opacities, self.shs)),
                     sh_degree=self.sh_degree)

    def param_bytes(self):
        return sum(t.numel() * 4 for t in (self.means3D, self.scales, self.rotations, self.opacities, self.shs))


def make_scene(P, cam, seed=0, sh_degree=3, sigma_px=2.0, base_se3=BASE_SE3):
    g = torch.Generator().manual_seed(seed)
    z = torch.rand(P, generator=g) * 7.0 + 1.0
    ndc = (torch.rand(P, 2, generator=g) * 2.0 - 1.0) * 1.1
    pv = torch.stack([ndc[:, 0] * cam.tanfovx * z, ndc[:, 1] * cam.tanfovy * z, z], dim=1).double()
    R, T = _se3_exp_np(base_se3)
    # p_view = (p_world - T) @ R   =>   p_world = p_view @ R^T + T
    means = (pv @ torch.from_numpy(R).T + torch.from_numpy(T)).float()
    focal_x = cam.width / (2 * cam.tanfovx)
    log_s = torch.log(sigma_px * z / focal_x)[:, None] + 0.5 * torch.randn(P, 3, generator=g)
    scales = torch.exp(log_s)
    q = torch.randn(P, 4, generator=g)
    rot = q / q.norm(dim=1, keepdim=True)
    opac = torch.rand(P, 1, generator=g) * 0.95 + 0.05
    M = (sh_degree + 1) ** 2
    shs = torch.randn(P, M, 3, generator=g) * 0.1
    shs[:, 0, :] = torch.randn(P, 3, generator=g) / SH_C0 * 0.3
    return Scene(means.contiguous(), scales.contiguous(), rot.contiguous(), opac.contiguous(), shs.contiguous(),
                 sh_degree)


@dataclass
class Trajectory:
    ctrl_trans: torch.Tensor   # [C+1,3] fp32 se(3) translation part u
    ctrl_rot: torch.Tensor     # [C+1,3] fp32 se(3) rotation part omega
    nu: torch.Tensor           # [F] fp32
    curve_order: int

    def to(self, device):
        return Trajectory(self.ctrl_trans.to(device), self.ctrl_rot.to(device), self.nu.to(device), self.curve_order)


def make_trajectory(F, curve_order, seed=1, sigma_trans=0.02, sigma_rot=0.005, base_se3=BASE_SE3):
    g = torch.Generator().manual_seed(seed)
    base = torch.tensor(base_se3, dtype=torch.float32)
    ct = base[None, :3] + sigma_trans * torch.randn(curve_order + 1, 3, generator=g)
    cr = base[None, 3:] + sigma_rot * torch.randn(curve_order + 1, 3, generator=g)
    nu = torch.linspace(0.0, 1.0, F) if F > 1 else torch.zeros(1)
    return Trajectory(ct.contiguous(), cr.contiguous(), nu.contiguous(), curve_order)


def make_target(cam, seed=2):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(3, cam.height, cam.width, generator=g)


def make_background(seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(3, generator=g)


def make_config(name, device="cpu"):
    P, W, H, F, C = CONFIGS[name]
    cam = make_camera(W, H)
    scene = make_scene(P, cam).to(device)
    traj = make_trajectory(F, C).to(device)
    return cam, scene, traj, make_target(cam).to(device), make_background().to(device)

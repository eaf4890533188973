"""Generates tests/golden/quat_golden.pt by running the REFERENCE's own Python for the two places that use
unit quaternions -- the "quarternion_cartesian" curve type (scene/motion.py:191-194, 242-246 with
scene/bezier.py BezierModel and _c2w_to_minicam) and the test-pose model (test.py:39-91, OptimPoseModel) --
on the CPU in the authoring container, where /root/reference is mounted.

The reference takes its two quaternion conversions from the third-party package `roma`
(`roma.unitquat_to_rotmat`, `roma.rotmat_to_unitquat`; environment.yml lists `roma` WITHOUT a version).  `roma`
is not installed in this image and cannot be fetched, so `_RomaShim` below restates the published algorithm of
those two functions (roma/mappings.py: XYZW order; matrix -> quaternion "adapted from SciPy": pick the largest of
the three diagonal entries and the trace, build the quaternion from that row, normalise; no sign
canonicalisation) and checks itself against scipy.spatial.transform.Rotation before anything is generated.
Everything else that runs is the reference's own code at its own call sites; `.cuda()` is made a no-op.

Run:  python tests/golden/make_quat_golden.py      (needs /root/reference; not needed on the GPU box)
"""
import ast
import copy
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("DEBLURGS_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


class _RomaShim(types.ModuleType):
    """roma.unitquat_to_rotmat / roma.rotmat_to_unitquat, restated (batch of quaternions / matrices in the
    leading dimension, as the reference calls them)."""

    @staticmethod
    def unitquat_to_rotmat(quat):
        x, y, z, w = quat[..., 0], quat[..., 1], quat[..., 2], quat[..., 3]
        x2, y2, z2, w2 = x * x, y * y, z * z, w * w
        xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
        rows = [torch.stack([x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)], -1),
                torch.stack([2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)], -1),
                torch.stack([2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2], -1)]
        return torch.stack(rows, -2)

    @staticmethod
    def rotmat_to_unitquat(R):
        matrix = R.reshape(-1, 3, 3)
        n = matrix.shape[0]
        decision = torch.empty((n, 4), dtype=matrix.dtype)
        decision[:, :3] = matrix.diagonal(dim1=1, dim2=2)
        decision[:, 3] = decision[:, :3].sum(dim=1)
        choices = decision.argmax(dim=1)
        quat = torch.empty((n, 4), dtype=matrix.dtype)
        ind = torch.nonzero(choices != 3, as_tuple=True)[0]
        i = choices[ind]
        j = (i + 1) % 3
        k = (j + 1) % 3
        quat[ind, i] = 1 - decision[ind, 3] + 2 * matrix[ind, i, i]
        quat[ind, j] = matrix[ind, j, i] + matrix[ind, i, j]
        quat[ind, k] = matrix[ind, k, i] + matrix[ind, i, k]
        quat[ind, 3] = matrix[ind, k, j] - matrix[ind, j, k]
        ind = torch.nonzero(choices == 3, as_tuple=True)[0]
        quat[ind, 0] = matrix[ind, 2, 1] - matrix[ind, 1, 2]
        quat[ind, 1] = matrix[ind, 0, 2] - matrix[ind, 2, 0]
        quat[ind, 2] = matrix[ind, 1, 0] - matrix[ind, 0, 1]
        quat[ind, 3] = 1 + decision[ind, 3]
        quat = quat / torch.norm(quat, dim=1)[:, None]
        return quat.reshape(R.shape[:-2] + (4,))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed module called: " + self.__name__)


def _check_shim_against_scipy(roma):
    from scipy.spatial.transform import Rotation as Rot
    g = torch.Generator().manual_seed(5)
    q = torch.randn(256, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    R = torch.from_numpy(Rot.from_quat(q.numpy()).as_matrix())
    assert (roma.unitquat_to_rotmat(q) - R).abs().max() < 1e-14
    axes = torch.nn.functional.normalize(torch.randn(64, 3, generator=g, dtype=torch.float64), dim=-1)
    R_pi = torch.from_numpy(Rot.from_rotvec((axes * (math.pi - 1e-3)).numpy()).as_matrix())
    for M in (R, R_pi):
        theirs = torch.from_numpy(Rot.from_matrix(M.numpy()).as_quat())
        assert (roma.rotmat_to_unitquat(M) - theirs).abs().max() < 1e-12, "shim differs from SciPy (sign included)"


def _rotations(n, seed, near_pi=False):
    g = torch.Generator().manual_seed(seed)
    if near_pi:
        from scipy.spatial.transform import Rotation as Rot
        axes = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=-1)
        ang = math.pi - torch.rand(n, 1, generator=g, dtype=torch.float64) * 0.2
        return torch.from_numpy(Rot.from_rotvec((axes * ang).numpy()).as_matrix())
    R = torch.linalg.qr(torch.randn(n, 3, 3, generator=g, dtype=torch.float64))[0]
    return R * torch.sign(torch.det(R))[:, None, None]


def main():
    roma = _RomaShim("roma")
    _check_shim_against_scipy(roma)
    sys.modules["roma"] = roma
    for name in ["open3d", "plyfile", "cv2", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "imageio",
                 "lpipsPyTorch", "PIL", "PIL.Image", "tqdm", "simple_knn", "simple_knn._C",
                 "diff_gaussian_rasterization"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    torch.Tensor.cuda = lambda self, *a, **k: self          # the reference hard-codes .cuda(); run it on the CPU
    nn.Module.cuda = lambda self, *a, **k: self

    from deblurgs_b200 import synthetic
    from scene.bezier import BezierModel
    from scene.motion import CameraMotionModule
    from utils.graphics_utils import getProjectionMatrix

    cases = {}
    # ---- 1. the quaternion curve: _sample_nu_from_alignment -> _sample_c2w_from_nu -> _c2w_to_minicam
    for tag, (F, order, seed) in {"q_c3f6": (6, 3, 1), "q_c9f16": (16, 9, 2), "q_c1f3": (3, 1, 3)}.items():
        g = torch.Generator().manual_seed(100 + seed)
        cam = synthetic.make_camera(600, 400)
        proj_t = cam.projection_matrix_t()
        q0 = roma.rotmat_to_unitquat(_rotations(1, seed).float())[0]
        ctrl_rot = q0[None].repeat(order + 1, 1) + 0.05 * torch.randn(order + 1, 4, generator=g)
        ctrl_trans = torch.tensor([0.3, -0.2, -4.0]) + 0.1 * torch.randn(order + 1, 3, generator=g)

        def bez(ctrl):
            m = BezierModel.__new__(BezierModel)
            nn.Module.__init__(m)
            m.curve_order = order
            m._control_points = nn.Parameter(ctrl[None].clone())
            import scipy.special
            m._bezier_binom_coeff = torch.tensor([scipy.special.binom(order, k) for k in range(order + 1)])
            return m

        cmm = CameraMotionModule.__new__(CameraMotionModule)
        cmm.curve_order, cmm.n_subframes, cmm.curve_type, cmm.curve_random_sample = order, F, "quarternion_cartesian", False
        cmm._trans, cmm._rot = bez(ctrl_trans), bez(ctrl_rot)
        lin = torch.linspace(1 / (F - 1), 1.0 - (1 / (F - 1)), F - 2)
        nu_raw = torch.log(lin / (1 - lin)) + 0.2 * torch.randn(F - 2, generator=g)
        cmm._nu = nn.Parameter(nu_raw[None].clone())

        class RefCam:
            projection_matrix = proj_t
            image_width, image_height, FoVx, FoVy, znear, zfar = cam.width, cam.height, cam.fovx, cam.fovy, 0.01, 100.0

        nu = cmm._sample_nu_from_alignment(0)
        rots, transes = cmm._sample_c2w_from_nu(0, nu)
        cams = cmm._c2w_to_minicam(rots, transes, RefCam)
        view = torch.stack([c.world_view_transform for c in cams])
        proj = torch.stack([c.full_proj_transform for c in cams])
        center = torch.stack([c.camera_center for c in cams])
        wv, wp = torch.randn(F, 4, 4, generator=g), torch.randn(F, 4, 4, generator=g)
        loss = (view * wv).sum() + (proj * wp).sum()
        gt, gr, gn = torch.autograd.grad(loss, [cmm._trans._control_points, cmm._rot._control_points, cmm._nu])
        cases[tag] = dict(F=F, order=order, ctrl_trans=ctrl_trans, ctrl_rot=ctrl_rot, nu_param=nu_raw, nu=nu.detach(),
                          proj_t=proj_t, view=view.detach(), proj=proj.detach(), center=center.detach(),
                          w_view=wv, w_proj=wp, g_ctrl_trans=gt[0], g_ctrl_rot=gr[0], g_nu_param=gn[0],
                          rots=rots.detach(), transes=transes.detach())
        print(tag, "ok")

    # ---- 2. _set_initial_parameters in quaternion mode (scene/motion.py:191-194): control points = quaternion of the
    #         c2w rotation / camera position, repeated, + N(0, 0.001^2) resp. N(0, 0.01^2)
    for tag, near_pi in (("init_generic", False), ("init_near_pi", True)):
        R = _rotations(12, 21, near_pi=near_pi)
        pos = torch.randn(12, 3, generator=torch.Generator().manual_seed(22), dtype=torch.float64)
        cmm = CameraMotionModule.__new__(CameraMotionModule)
        cmm.curve_order, cmm.curve_type = 3, "quarternion_cartesian"
        torch.manual_seed(0)
        cmm._set_initial_parameters(R, pos)
        cases[tag] = dict(rotations=R, translations=pos, quat=roma.rotmat_to_unitquat(R),
                          ctrl_rot=cmm._rot._control_points.detach().clone(),
                          ctrl_trans=cmm._trans._control_points.detach().clone())
        print(tag, "ok")

    # ---- 3. OptimPoseModel (test.py:39-91): the class is executed from the reference's own file (only that class;
    #         importing test.py as a module would pull in the whole training stack)
    src = open(os.path.join(REF, "test.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "OptimPoseModel")
    ns = dict(torch=torch, nn=nn, copy=copy, roma=roma, getProjectionMatrix=getProjectionMatrix, Camera=object,
              print=lambda *a, **k: None)
    exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, "test.py"), "exec"), ns)
    scam = synthetic.make_camera(96, 64)

    class Cam:
        pass
    n = 6
    R = torch.cat([_rotations(3, 31), _rotations(3, 32, near_pi=True)]).float()
    T = torch.randn(n, 3, generator=torch.Generator().manual_seed(33))
    cams = []
    for i in range(n):
        c = Cam()
        c.R, c.T = R[i].numpy(), T[i].numpy()
        c.image_width, c.image_height, c.FoVx, c.FoVy, c.znear, c.zfar = 96, 64, scam.fovx, scam.fovy, 0.01, 100.0
        cams.append(c)
    model = ns["OptimPoseModel"](cams)
    g = torch.Generator().manual_seed(34)
    out = dict(R=R, T=T, fovx=scam.fovx, fovy=scam.fovy, rot_param=model._rot.detach().clone(),
               trans_param=model._trans.detach().clone(), view=[], proj=[], center=[], g_rot=[], g_trans=[],
               w_view=[], w_proj=[])
    for i in range(n):
        cam = model(i)
        wv, wp = torch.randn(4, 4, generator=g), torch.randn(4, 4, generator=g)
        loss = (cam.world_view_transform * wv).sum() + (cam.full_proj_transform * wp).sum()
        gr, gt = torch.autograd.grad(loss, [model._rot, model._trans])
        for k, v in (("view", cam.world_view_transform), ("proj", cam.full_proj_transform),
                     ("center", cam.camera_center), ("g_rot", gr[i]), ("g_trans", gt[i]), ("w_view", wv), ("w_proj", wp)):
            out[k].append(v.detach().clone())
    cases["optim_pose_model"] = {k: (torch.stack(v) if isinstance(v, list) else v) for k, v in out.items()}
    print("optim_pose_model ok")

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quat_golden.pt")
    torch.save(cases, path)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()

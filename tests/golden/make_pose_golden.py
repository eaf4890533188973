"""Generates tests/golden/pose_golden.pt by running the REFERENCE's own Python pose chain
(scene/bezier.py BezierModel.forward, scene/motion.py _sample_nu_from_alignment / _sample_c2w_from_nu /
_c2w_to_minicam, utils/pytorch3d_functions.py se3_exp_map, scene/cameras.py MiniCam) on the CPU in the
authoring container, where /root/reference is mounted. Missing third-party modules that the default
se3 path never calls (roma, open3d, plyfile, ...) are stubbed. Also records autograd gradients of a
fixed linear functional of the matrices w.r.t. control points and the alignment parameter.

Run:  python tests/golden/make_pose_golden.py      (needs /root/reference; not needed on the GPU box)
"""
import os
import sys
import types

import torch

REF = os.environ.get("DEBLURGS_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, REF)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed module called: " + self.__name__)


for name in ["roma", "open3d", "plyfile", "cv2", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "imageio",
             "lpipsPyTorch", "PIL", "PIL.Image", "tqdm", "simple_knn", "simple_knn._C", "diff_gaussian_rasterization"]:
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)

from deblurgs_b200 import synthetic  # noqa: E402


def main():
    from scene.bezier import BezierModel
    from scene.motion import CameraMotionModule
    from scene.cameras import MiniCam  # noqa: F401

    cases = {}
    for tag, (F, order, seed) in {"c3f4": (4, 3, 1), "c9f16": (16, 9, 1), "c9f21": (21, 9, 5), "c1f3": (3, 1, 2),
                                  "small_rot": (5, 3, 9)}.items():
        base = synthetic.BASE_SE3 if tag != "small_rot" else (0.1, 0.2, -0.3, 0.001, -0.002, 0.0015)
        traj = synthetic.make_trajectory(F, order, seed=seed, base_se3=base)
        cam = synthetic.make_camera(600, 400)
        proj_t = cam.projection_matrix_t()

        def bez(ctrl):
            m = BezierModel.__new__(BezierModel)
            torch.nn.Module.__init__(m)
            m.curve_order = order
            m._control_points = torch.nn.Parameter(ctrl[None].clone())
            import scipy.special
            m._bezier_binom_coeff = torch.tensor([scipy.special.binom(order, k) for k in range(order + 1)])
            return m

        cmm = CameraMotionModule.__new__(CameraMotionModule)
        cmm.curve_order, cmm.n_subframes, cmm.curve_type, cmm.curve_random_sample = order, F, "se3", False
        cmm._trans, cmm._rot = bez(traj.ctrl_trans), bez(traj.ctrl_rot)
        f = F
        nu_raw = torch.log(torch.linspace(1 / (f - 1), 1.0 - (1 / (f - 1)), f - 2) /
                           (1 - torch.linspace(1 / (f - 1), 1.0 - (1 / (f - 1)), f - 2)))
        if tag == "c9f21":
            nu_raw = nu_raw + 0.3 * torch.randn(f - 2, generator=torch.Generator().manual_seed(3))
        cmm._nu = torch.nn.Parameter(nu_raw[None].clone())

        class RefCam:
            projection_matrix = proj_t
            image_width, image_height, FoVx, FoVy, znear, zfar = cam.width, cam.height, cam.fovx, cam.fovy, 0.01, 100.0

        nu = cmm._sample_nu_from_alignment(0)
        rots, transes = cmm._sample_c2w_from_nu(0, nu)
        cams = cmm._c2w_to_minicam(rots, transes, RefCam)
        view = torch.stack([c.world_view_transform for c in cams])
        proj = torch.stack([c.full_proj_transform for c in cams])
        center = torch.stack([c.camera_center for c in cams])
        g = torch.Generator().manual_seed(11)
        wv, wp = torch.randn(F, 4, 4, generator=g), torch.randn(F, 4, 4, generator=g)
        loss = (view * wv).sum() + (proj * wp).sum()
        gt, gr, gn = torch.autograd.grad(loss, [cmm._trans._control_points, cmm._rot._control_points, cmm._nu])
        cases[tag] = dict(F=F, order=order, ctrl_trans=traj.ctrl_trans, ctrl_rot=traj.ctrl_rot, nu_param=nu_raw,
                          nu=nu.detach(), proj_t=proj_t, view=view.detach(), proj=proj.detach(),
                          center=center.detach(), w_view=wv, w_proj=wp, g_ctrl_trans=gt[0], g_ctrl_rot=gr[0],
                          g_nu_param=gn[0], rots=rots.detach(), transes=transes.detach())
        print(tag, "ok", view.dtype, rots.dtype)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pose_golden.pt")
    torch.save(cases, out)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()

"""GPU tests (pytest -m gpu) of the SURVEY.md 8f rank-4 consumers of the pose gradients: the quaternion curve
type rendered through the batched path, and test-time pose refinement (test.py:39-91, 131-186 of the reference)
through the single-view drop-in `render` + FusedAdam."""
import pytest
import torch

from tests import parity_utils as pu
from deblurgs_b200 import renderer
from deblurgs_b200.motion import CameraMotionModule, GaussianParams
from deblurgs_b200.pose import rotmat_to_unitquat
from deblurgs_b200.refine import OptimPoseModel, optimize_test_pose

pytestmark = pytest.mark.gpu


class _Cam:
    pass


def _camera_from_view(cam, view_row):
    c = _Cam()
    c.R = view_row[:3, :3].double().cpu().numpy()      # c2w rotation (world_view_transform[:3,:3])
    c.T = view_row[3, :3].double().cpu().numpy()       # w2c translation (world_view_transform[3,:3])
    c.image_width, c.image_height, c.FoVx, c.FoVy, c.znear, c.zfar = cam.width, cam.height, cam.fovx, cam.fovy, 0.01, 100.0
    c.projection_matrix = cam.projection_matrix_t().cuda()
    return c


def test_quaternion_curve_renders_like_the_same_poses_given_explicitly():
    cam, scene, traj, bg, view, proj, campos = pu.make_inputs("tiny")
    gauss = GaussianParams.from_scene(scene)
    F = view.shape[0]
    rc = _camera_from_view(cam, view[0])
    R = view[:, :3, :3].contiguous()                                   # c2w rotations of the sub-frames
    pos = campos.clone()
    # a degree-0 "curve" per sub-frame pose would need F modules; instead build one module whose control points
    # all equal pose 1 (no initial noise) and compare with an explicit single-pose batched render
    m = CameraMotionModule.from_poses([rc], R[1:2], pos[1:2], curve_type="quarternion_cartesian", curve_order=2,
                                      num_subframes=F)
    with torch.no_grad():
        q = rotmat_to_unitquat(R[1:2].double()).float()
        m._rot._control_points.copy_(q[:, None, :].expand_as(m._rot._control_points))
        m._trans._control_points.copy_(pos[1][None, None].expand_as(m._trans._control_points))
    m.link_gaussian(gauss)
    out = m.query(0, "all", background=bg)
    v1 = view[1:2].expand(F, 4, 4).contiguous()
    p1 = proj[1:2].expand(F, 4, 4).contiguous()
    c1 = campos[1:2].expand(F, 3).contiguous()
    ref = renderer.render_blurry(v1, p1, c1, rc, gauss, bg)
    assert (out["blurred"] - ref["blurred"]).abs().max().item() < 2e-3     # quaternion round trip of the pose: ~1e-6 in the matrices
    loss = (out["blurred"] - 0.5).abs().mean()
    loss.backward()
    for p in m.parameters()[:2]:
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0


def test_pose_refinement_moves_a_perturbed_camera_back():
    cam, scene, traj, bg, view, proj, campos = pu.make_inputs("small")
    gauss = GaussianParams.from_scene(scene)
    true_cam = _camera_from_view(cam, view[2])
    model = OptimPoseModel([true_cam])
    with torch.no_grad():
        gt = renderer.render(model(0), gauss, bg)["render"].clamp(0.0, 1.0)
    off = _camera_from_view(cam, view[2])
    off.T = off.T + [0.012, -0.008, 0.01]
    off.original_image = gt
    with torch.no_grad():
        start = (renderer.render(OptimPoseModel([off])(0), gauss, bg)["render"].clamp(0, 1) - gt).abs().mean().item()
    refined, hist = optimize_test_pose([off], gauss, bg, num_iter_per_view=80, seed=0)
    assert abs(hist[0] - start) < 1e-6
    assert hist[-1] < 0.6 * hist[0], (hist[0], hist[-1])
    err0 = float(abs(torch.tensor(off.T - true_cam.T)).max())
    err1 = float((refined[0].world_view_transform[3, :3].cpu().double() - torch.tensor(true_cam.T)).abs().max())
    assert err1 < err0

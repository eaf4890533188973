"""CPU: host-side pose algebra of SURVEY.md 8f rank 4 (`deblurgs_b200/pose.py`, `motion.py`, `refine.py`) against
golden vectors produced by the reference's own se3_log_map / se3_exp_map (tests/golden/se3_log_golden.pt, generator
tests/golden/make_se3log_golden.py) and against the oracle's exponential map.

Tolerances: se3_log_map within 5e-6 (fp32) / 1e-10 (fp64) of the reference's output on the same matrices (the
two differ only in how the clamped acos and the 3x3 solve are evaluated); exp(log(T)) == T to 1e-5 / 1e-9."""
import math
import os

import pytest
import torch

from deblurgs_b200 import pose, synthetic
from deblurgs_b200.motion import CameraMotionModule, GaussianParams
from oracle import pose_torch as pt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "se3_log_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


@pytest.mark.parametrize("tag,tol", [("f32", 5e-6), ("f64", 1e-10), ("tiny_f32", 5e-6), ("tiny_f64", 1e-10),
                                     ("large_f64", 1e-9)])
def test_se3_log_map_matches_reference_golden(gold, tag, tol):
    c = gold[tag]
    ours = pose.se3_log_map(c["transform"])
    assert ours.dtype == c["log_out"].dtype
    assert (ours - c["log_out"]).abs().max().item() <= tol
    # and it inverts the exponential map the pose kernel implements (oracle restatement of se3_exp_map)
    back = pt.se3_exp(ours)
    assert (back - c["transform"]).abs().max().item() <= (1e-5 if c["transform"].dtype == torch.float32 else 1e-9)


def test_se3_log_map_on_the_reference_init_layout(gold):
    c = gold["init_layout"]
    assert (pose.se3_log_map(c["transform"]) - c["log_out"]).abs().max().item() <= 1e-10


def test_se3_log_map_rejects_what_the_reference_rejects():
    with pytest.raises(ValueError, match=r"\(N, 4, 4\)"):
        pose.se3_log_map(torch.eye(4))
    bad = torch.eye(4)[None].clone()
    bad[0, 0, 3] = 0.5
    with pytest.raises(ValueError, match="should be 0"):
        pose.se3_log_map(bad)
    with pytest.raises(ValueError, match="trace outside"):
        pose.so3_log_map(4.0 * torch.eye(3)[None])


def test_quaternion_convention_is_xyzw_and_round_trips():
    a = 0.7
    q = torch.tensor([[math.sin(a / 2), 0.0, 0.0, math.cos(a / 2)]], dtype=torch.float64)   # rotation about x
    R = pose.unitquat_to_rotmat(q)[0]
    expect = torch.tensor([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]],
                          dtype=torch.float64)
    assert (R - expect).abs().max().item() < 1e-15
    g = torch.Generator().manual_seed(1)
    q = torch.randn(256, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    q = torch.where(q[:, 3:4] < 0, -q, q)
    R = pose.unitquat_to_rotmat(q)
    assert (R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max().item() < 1e-14
    assert (torch.det(R) - 1).abs().max().item() < 1e-14
    back = pose.rotmat_to_unitquat(R)
    # the sign is that of the branch taken (largest of m00, m11, m22, trace), like roma / SciPy: q or -q
    sign = torch.sign((back * q).sum(-1, keepdim=True))
    assert (back - sign * q).abs().max().item() < 1e-12
    big = q.abs().argmax(-1)
    assert (back[torch.arange(256), big] > 0).all()
    # every branch of the matrix -> quaternion conversion (largest of the three diagonal entries / the trace)
    for axis in range(3):
        v = torch.zeros(1, 4, dtype=torch.float64)
        v[0, axis], v[0, 3] = math.cos(0.05), math.sin(0.05)      # rotation by almost pi about `axis`
        assert (pose.rotmat_to_unitquat(pose.unitquat_to_rotmat(v)) - v).abs().max().item() < 1e-12


def test_quaternion_conversions_match_scipy_rotation():
    """`roma` (the reference's quaternion library, scene/motion.py:192,245, test.py:66,79) is not installed here and the
    reference does not pin its version; its documented conventions are XYZW order and a matrix -> quaternion routine
    "adapted from SciPy".  scipy.spatial.transform.Rotation is installed and uses the same conventions: pin both
    conversions against it (q and -q are the same rotation)."""
    Rot = pytest.importorskip("scipy.spatial.transform").Rotation
    g = torch.Generator().manual_seed(5)
    q = torch.randn(512, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    R_sp = torch.from_numpy(Rot.from_quat(q.numpy()).as_matrix())
    assert (pose.unitquat_to_rotmat(q) - R_sp).abs().max().item() < 1e-14
    mine = pose.rotmat_to_unitquat(R_sp)
    theirs = torch.from_numpy(Rot.from_matrix(R_sp.numpy()).as_quat())     # sign included (no canonicalisation)
    assert (mine - theirs).abs().max().item() < 1e-12
    # near-pi rotations exercise the three non-trace branches
    axes = torch.nn.functional.normalize(torch.randn(64, 3, generator=g, dtype=torch.float64), dim=-1)
    ang = math.pi - torch.rand(64, 1, generator=g, dtype=torch.float64) * 1e-3
    R_pi = torch.from_numpy(Rot.from_rotvec((axes * ang).numpy()).as_matrix())
    mine = pose.rotmat_to_unitquat(R_pi)
    theirs = torch.from_numpy(Rot.from_matrix(R_pi.numpy()).as_quat())
    assert (mine - theirs).abs().max().item() < 1e-9


class _RefCam:
    def __init__(self, cam):
        self.image_width, self.image_height = cam.width, cam.height
        self.FoVx, self.FoVy, self.znear, self.zfar = cam.fovx, cam.fovy, cam.znear, cam.zfar
        self.projection_matrix = cam.projection_matrix_t()


def _poses(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    R = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))[0]
    R = R * torch.sign(torch.det(R))[:, None, None]
    return R, torch.randn(n, 3, generator=g)


def test_from_poses_se3_starts_at_the_log_of_the_camera_pose():
    R, t = _poses(4)
    cam = _RefCam(synthetic.make_camera(96, 64))
    m = CameraMotionModule.from_poses([cam], R, t, curve_order=3, num_subframes=5)
    assert m.curve_type == "se3" and len(m) == 4
    assert m._trans._control_points.shape == (4, 4, 3) and m._nu.shape == (4, 3)
    c2w = torch.zeros(4, 4, 4)
    c2w[:, :3, :3], c2w[:, 3, :3], c2w[:, 3, 3] = R.transpose(1, 2), t, 1.0
    log = pose.se3_log_map(c2w)
    # every control point = the log of the pose + N(0, 0.001^2) (scene/bezier.py:36-40)
    assert (m._trans._control_points - log[:, None, :3]).abs().max().item() < 0.01
    assert (m._rot._control_points - log[:, None, 3:]).abs().max().item() < 0.01


def test_quaternion_curve_type_reproduces_the_reference_chain_on_cpu():
    """curve_type == 'quarternion_cartesian' (scene/motion.py:191-194, 242-246) is pure torch on tiny tensors:
    compare with a per-sub-frame restatement of _c2w_to_minicam (oracle) and check that gradients reach the
    control points and the alignment parameter."""
    R, t = _poses(3, seed=2)
    cam = _RefCam(synthetic.make_camera(96, 64))
    m = CameraMotionModule.from_poses([cam], R, t, curve_type="quarternion_cartesian", curve_order=3,
                                      num_subframes=6)
    nu = m._sample_nu_from_alignment(1)
    view, proj, center = m.get_trajectory_tensors(1, nu)
    assert view.shape == (6, 4, 4) and proj.shape == (6, 4, 4) and center.shape == (6, 3)
    q = pt.bezier_sample(nu, m._rot._control_points[1])
    q = q / q.norm(dim=1, keepdim=True)
    rots = pose.unitquat_to_rotmat(q)
    ref = pt.c2w_to_view_proj(rots.float(), pt.bezier_sample(nu, m._trans._control_points[1]).float(),
                              cam.projection_matrix)
    for s, (wvt, full, c) in enumerate(ref):
        assert (view[s] - wvt).abs().max().item() < 1e-6
        assert (proj[s] - full).abs().max().item() < 1e-5
        assert (center[s] - c).abs().max().item() < 1e-5
    loss = (view * torch.arange(16.0).view(4, 4)).sum() + proj.sum()
    grads = torch.autograd.grad(loss, [m._rot._control_points, m._trans._control_points, m._nu])
    assert all(torch.isfinite(g).all() for g in grads)
    assert grads[0][1].abs().max() > 0 and grads[1][1].abs().max() > 0 and grads[2][1].abs().max() > 0
    assert grads[0][0].abs().max() == 0       # other images' curves untouched


def test_add_training_setup_appends_curve_groups_like_the_reference():
    R, t = _poses(2)
    cam = _RefCam(synthetic.make_camera(96, 64))
    m = CameraMotionModule.from_poses([cam], R, t, curve_order=2, num_subframes=4)
    P = 5
    g = GaussianParams(torch.zeros(P, 3), torch.zeros(P, 1, 3), torch.zeros(P, 3, 3), torch.zeros(P, 3),
                       torch.ones(P, 4), torch.ones(P, 1), 1)
    g.training_setup()
    lr = {"curve_rot": 1e-3, "curve_trans": 2e-3, "curve_alignment": 0.0}
    m.add_training_setup(g, lr)
    names = [gr["name"] for gr in g.optimizer.param_groups]
    assert names == ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "curve_rot", "curve_trans",
                     "curve_alignment"]
    m.add_training_setup(g, {**lr, "curve_rot": 5e-4})          # again: replaced, not duplicated
    names = [gr["name"] for gr in g.optimizer.param_groups]
    assert names.count("curve_rot") == 1 and g.optimizer.param_groups[6]["lr"] == 5e-4
    assert g.optimizer.param_groups[8]["params"][0] is m._nu


def test_optim_pose_model_builds_the_reference_matrices_on_cpu():
    from deblurgs_b200.refine import OptimPoseModel, get_projection_matrix
    scam = synthetic.make_camera(96, 64)
    R, T = _poses(3, seed=5)

    class Cam:
        pass
    cams = []
    for i in range(3):
        c = Cam()
        c.R, c.T = R[i].numpy(), T[i].numpy()
        c.image_width, c.image_height, c.FoVx, c.FoVy, c.znear, c.zfar = 96, 64, scam.fovx, scam.fovy, 0.01, 100.0
        cams.append(c)
    model = OptimPoseModel(cams, device="cpu")
    assert (get_projection_matrix(0.01, 100.0, scam.fovx, scam.fovy).T - scam.projection_matrix_t()).abs().max() < 1e-6
    for i in range(3):
        cam = model(i)
        # getWorld2View(R, T).T of the reference: upper-left block = R, row 3 = T (utils/graphics_utils.py:28-49)
        assert (cam.world_view_transform[:3, :3] - R[i]).abs().max().item() < 1e-6
        assert (cam.world_view_transform[3, :3] - T[i]).abs().max().item() < 1e-7
        assert (cam.full_proj_transform - cam.world_view_transform @ scam.projection_matrix_t()).abs().max() < 1e-5
        # camera centre = -R T (c2w translation)
        assert (cam.camera_center - (-(R[i] @ T[i]))).abs().max().item() < 1e-5
        g = torch.autograd.grad(cam.full_proj_transform.sum(), [model._rot, model._trans])
        assert g[0][i].abs().max() > 0 and g[1][i].abs().max() > 0

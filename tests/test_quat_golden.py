"""CPU: the two quaternion users of the reference -- the "quarternion_cartesian" curve type (scene/motion.py:191-194,
242-246) and the test-pose model (test.py:39-91) -- against tests/golden/quat_golden.pt, which
tests/golden/make_quat_golden.py produced by running the REFERENCE's own Python at those call sites.  The only
substituted part is the absent third-party `roma` (two conversion functions restated from its published algorithm and
checked against SciPy in the generator, sign convention included)."""
import os

import pytest
import torch

from deblurgs_b200 import pose, synthetic
from deblurgs_b200.motion import CameraMotionModule

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quat_golden.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


class _RefCam:
    def __init__(self, proj_t, cam):
        self.image_width, self.image_height = cam.width, cam.height
        self.FoVx, self.FoVy, self.znear, self.zfar = cam.fovx, cam.fovy, 0.01, 100.0
        self.projection_matrix = proj_t


@pytest.mark.parametrize("tag", ["q_c3f6", "q_c9f16", "q_c1f3"])
def test_quaternion_curve_matches_the_reference_python(golden, tag):
    c = golden[tag]
    F, order = c["F"], c["order"]
    cam = _RefCam(c["proj_t"], synthetic.make_camera(600, 400))
    q0 = c["ctrl_rot"][:1] / c["ctrl_rot"][:1].norm()
    m = CameraMotionModule([cam], (q0, c["ctrl_trans"][:1]), curve_order=order, num_subframes=F,
                           curve_type="quarternion_cartesian")
    with torch.no_grad():
        m._rot._control_points.copy_(c["ctrl_rot"][None])
        m._trans._control_points.copy_(c["ctrl_trans"][None])
        m._nu.copy_(c["nu_param"][None])
    nu = m._sample_nu_from_alignment(0)
    assert torch.equal(nu, c["nu"])
    view, proj, center = m.get_trajectory_tensors(0)
    # fp32 chain of ~20 operations: a few ulp of the O(1..5) entries
    assert (view - c["view"]).abs().max().item() < 2e-6
    assert (proj - c["proj"]).abs().max().item() < 1e-5
    assert (center - c["center"]).abs().max().item() < 2e-6
    loss = (view * c["w_view"]).sum() + (proj * c["w_proj"]).sum()
    gt, gr, gn = torch.autograd.grad(loss, [m._trans._control_points, m._rot._control_points, m._nu])
    for mine, ref in ((gt[0], c["g_ctrl_trans"]), (gr[0], c["g_ctrl_rot"]), (gn[0], c["g_nu_param"])):
        assert (mine - ref).abs().max().item() <= 1e-4 * max(ref.abs().max().item(), 1.0)
    # get_trajectory returns the reference's MiniCam list with the same matrices
    cams = m.get_trajectory(0)
    assert len(cams) == F and torch.equal(cams[1].world_view_transform, view[1])


@pytest.mark.parametrize("tag", ["init_generic", "init_near_pi"])
def test_quaternion_initialisation_matches_the_reference_python(golden, tag):
    """_set_initial_parameters, quaternion branch: control points = roma.rotmat_to_unitquat(c2w rotation) resp. the
    camera position, repeated C+1 times, plus N(0, 0.001^2) resp. N(0, 0.01^2) noise (scene/bezier.py:36-40)."""
    c = golden[tag]
    q = pose.rotmat_to_unitquat(c["rotations"])
    assert (q - c["quat"]).abs().max().item() < 1e-12          # sign included: the same representative as roma's
    cam = _RefCam(synthetic.make_camera(96, 64).projection_matrix_t(), synthetic.make_camera(96, 64))
    m = CameraMotionModule.from_poses([cam], c["rotations"], c["translations"], curve_type="quarternion_cartesian",
                                      curve_order=3, num_subframes=5, generator=torch.Generator().manual_seed(0))
    assert m._rot._control_points.shape == c["ctrl_rot"].shape == (12, 4, 4)
    assert m._trans._control_points.shape == c["ctrl_trans"].shape == (12, 4, 3)
    assert m._rot._control_points.dtype == torch.float32
    # both sides drew their own noise: equal up to 6 sigma of it
    assert (m._rot._control_points - c["quat"][:, None, :].float()).abs().max().item() < 6e-3
    assert (c["ctrl_rot"] - c["quat"][:, None, :].float()).abs().max().item() < 6e-3
    assert (m._trans._control_points - c["translations"][:, None, :].float()).abs().max().item() < 6e-2
    assert (c["ctrl_trans"] - c["translations"][:, None, :].float()).abs().max().item() < 6e-2
    if tag == "init_near_pi":
        assert (c["quat"][:, 3] < 0).any()      # the case canonicalising to w >= 0 would get wrong


def test_optim_pose_model_matches_the_reference_python(golden):
    from deblurgs_b200.refine import OptimPoseModel
    c = golden["optim_pose_model"]
    n = c["R"].shape[0]

    class Cam:
        pass
    cams = []
    for i in range(n):
        cam = Cam()
        cam.R, cam.T = c["R"][i].numpy(), c["T"][i].numpy()
        cam.image_width, cam.image_height, cam.FoVx, cam.FoVy, cam.znear, cam.zfar = 96, 64, c["fovx"], c["fovy"], 0.01, 100.0
        cams.append(cam)
    model = OptimPoseModel(cams, device="cpu")
    assert (model._rot - c["rot_param"]).abs().max().item() < 1e-6
    assert torch.equal(model._trans.detach(), c["trans_param"])
    for i in range(n):
        cam = model(i)
        assert (cam.world_view_transform - c["view"][i]).abs().max().item() < 1e-6
        assert (cam.full_proj_transform - c["proj"][i]).abs().max().item() < 1e-5
        assert (cam.camera_center - c["center"][i]).abs().max().item() < 1e-5
        loss = (cam.world_view_transform * c["w_view"][i]).sum() + (cam.full_proj_transform * c["w_proj"][i]).sum()
        gr, gt = torch.autograd.grad(loss, [model._rot, model._trans])
        assert (gr[i] - c["g_rot"][i]).abs().max().item() <= 1e-4 * max(c["g_rot"][i].abs().max().item(), 1.0)
        assert (gt[i] - c["g_trans"][i]).abs().max().item() <= 1e-4 * max(c["g_trans"][i].abs().max().item(), 1.0)

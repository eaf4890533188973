"""CPU: the oracle's pose restatement (oracle/pose_torch.py) against golden vectors produced by the
reference's own Python code (tests/golden/make_pose_golden.py)."""
import os

import pytest
import torch

from oracle import pose_torch as pt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


@pytest.mark.parametrize("tag", ["c3f4", "c9f16", "c9f21", "c1f3", "small_rot"])
def test_pose_oracle_matches_reference_python(gold, tag):
    c = gold[tag]
    nu_p = c["nu_param"].clone().requires_grad_(True)
    ct = c["ctrl_trans"].clone().requires_grad_(True)
    cr = c["ctrl_rot"].clone().requires_grad_(True)
    nu = pt.sample_nu(nu_p, c["F"])
    assert torch.equal(nu.detach(), c["nu"])
    rots, transes = pt.sample_c2w(ct, cr, nu)
    assert rots.dtype == torch.float64
    assert torch.equal(rots.detach(), c["rots"]) and torch.equal(transes.detach(), c["transes"])
    poses = pt.c2w_to_view_proj(rots, transes, c["proj_t"])
    view = torch.stack([p[0] for p in poses])
    proj = torch.stack([p[1] for p in poses])
    center = torch.stack([p[2] for p in poses])
    # identical torch ops on the same machine: bit-exact
    assert torch.equal(view.detach(), c["view"])
    assert torch.equal(proj.detach(), c["proj"])
    assert torch.equal(center.detach(), c["center"])
    loss = (view * c["w_view"]).sum() + (proj * c["w_proj"]).sum()
    gt, gr, gn = torch.autograd.grad(loss, [ct, cr, nu_p])
    torch.testing.assert_close(gt, c["g_ctrl_trans"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(gr, c["g_ctrl_rot"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(gn, c["g_nu_param"], rtol=1e-6, atol=1e-7)


def test_bernstein_index_is_reversed():
    # control point k=0 is the t=1 end (scene/bezier.py:62)
    ctrl = torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]])
    s = pt.bezier_sample(torch.tensor([0.0, 1.0]), ctrl)
    assert torch.allclose(s[0], ctrl[2].double()) and torch.allclose(s[1], ctrl[0].double())


def test_se3_small_rotation_uses_clamped_angle():
    # |omega| < 0.01 -> theta = 0.01 exactly (utils/pytorch3d_functions.py:230-236)
    v = torch.tensor([[0.1, 0.2, 0.3, 1e-3, 0.0, 0.0]], dtype=torch.float64)
    T = pt.se3_exp(v)
    th = 0.01
    import math
    assert abs(T[0, 2, 1].item() - (-(math.sin(th) / th) * 1e-3)) < 1e-15 or abs(T[0, 1, 2].item() - (-(math.sin(th) / th) * 1e-3)) < 1e-15

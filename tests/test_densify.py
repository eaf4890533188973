"""CPU: densify / prune mechanics of the parameter store (`deblurgs_b200/densify.py`) against the behaviour the
reference defines (scene/gaussian_model.py:247-254, 300-454): which Gaussians are selected, what the new ones look
like, and what happens to the Adam moments, the per-tensor step count and the statistics.  The operations are
device-agnostic torch ops, so they are exercised here with CPU tensors and both optimizers that share torch's
`param_groups` / `state` layout (FusedAdam's host logic, torch.optim.Adam)."""
import pytest
import torch

from deblurgs_b200.densify import build_rotation
from deblurgs_b200.motion import GaussianParams
from deblurgs_b200.params import FusedAdam

NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]
ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
        "scaling": "_scaling", "rotation": "_rotation"}


def _store(P=40, M=4, seed=0, optimizer_cls=None):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    gp = GaussianParams(r(P, 3), r(P, 1, 3), r(P, M - 1, 3), r(P, 3) * 0.5 - 2.0, r(P, 4),
                        torch.rand(P, 1, generator=g), 1)
    opt = gp.training_setup(optimizer_cls=optimizer_cls)
    # give every tensor optimizer state, as after a few training steps
    for grp in opt.param_groups:
        p = grp["params"][0]
        opt.state[p] = {"step": 7 if optimizer_cls is None else torch.tensor(7.0), "exp_avg": r(*p.shape),
                        "exp_avg_sq": r(*p.shape).abs()}
    gp._ensure_stats()
    gp.xyz_gradient_accum += torch.rand(P, 1, generator=g)
    gp.denom += 1.0
    gp.max_radii2D += torch.arange(P).float()
    return gp, opt, g


def _snapshot(gp, opt):
    snap = {}
    for grp in opt.param_groups:
        p = grp["params"][0]
        st = opt.state[p]
        snap[grp["name"]] = (p.detach().clone(), st["exp_avg"].clone(), st["exp_avg_sq"].clone(), float(st["step"]))
    return snap


@pytest.mark.parametrize("optimizer_cls", [None, torch.optim.Adam])
def test_prune_points_compacts_parameters_moments_and_statistics(optimizer_cls):
    gp, opt, g = _store(optimizer_cls=optimizer_cls)
    before = _snapshot(gp, opt)
    opt.add_param_group({"params": [torch.nn.Parameter(torch.zeros(2, 4, 3))], "lr": 1e-3, "name": "curve_rot"})
    curve = opt.param_groups[-1]["params"][0]
    stats = (gp.xyz_gradient_accum.clone(), gp.denom.clone(), gp.max_radii2D.clone())
    mask = torch.zeros(40, dtype=torch.bool)
    mask[[0, 3, 17, 39]] = True
    gp.prune_points(mask)
    keep = ~mask
    for grp in opt.param_groups[:6]:
        name, p = grp["name"], grp["params"][0]
        assert p is getattr(gp, ATTR[name]) and isinstance(p, torch.nn.Parameter) and p.requires_grad
        st = opt.state[p]
        assert torch.equal(p.detach(), before[name][0][keep])
        assert torch.equal(st["exp_avg"], before[name][1][keep]) and torch.equal(st["exp_avg_sq"], before[name][2][keep])
        assert float(st["step"]) == before[name][3]
    assert len(opt.state) == 6                                   # no stale entries of the replaced parameters
    assert opt.param_groups[-1]["params"][0] is curve            # 'curve_*' groups are left alone
    assert torch.equal(gp.xyz_gradient_accum, stats[0][keep]) and torch.equal(gp.denom, stats[1][keep])
    assert torch.equal(gp.max_radii2D, stats[2][keep])
    if optimizer_cls is torch.optim.Adam:                        # the optimizer is still steppable
        for p in gp.parameters():
            p.grad = torch.ones_like(p)
        opt.step()
        assert all(float(opt.state[p]["step"]) == 8.0 for p in gp.parameters())


def test_clone_appends_small_high_gradient_gaussians_with_zero_moments():
    gp, opt, g = _store()
    grads = torch.zeros(40, 1)
    grads[[2, 5, 9]] = 1.0
    with torch.no_grad():
        gp._scaling[5] = 3.0                                     # exp(3) = 20 > percent_dense * extent: too large to clone
    before = _snapshot(gp, opt)
    gp.densify_and_clone(grads, 0.5, scene_extent=100.0)         # percent_dense * extent = 1.0
    sel = torch.tensor([2, 9])
    assert gp._xyz.shape[0] == 42
    for grp in opt.param_groups:
        name, p = grp["name"], grp["params"][0]
        st = opt.state[p]
        assert torch.equal(p.detach()[:40], before[name][0]) and torch.equal(p.detach()[40:], before[name][0][sel])
        assert torch.equal(st["exp_avg"][:40], before[name][1]) and float(st["exp_avg"][40:].abs().max()) == 0.0
        assert torch.equal(st["exp_avg_sq"][:40], before[name][2]) and float(st["exp_avg_sq"][40:].abs().max()) == 0.0
        assert float(st["step"]) == 7.0
    assert gp.xyz_gradient_accum.shape == (42, 1) and float(gp.xyz_gradient_accum.abs().max()) == 0.0
    assert gp.denom.shape == (42, 1) and gp.max_radii2D.shape == (42,) and float(gp.max_radii2D.max()) == 0.0


def test_split_replaces_large_gaussians_by_scaled_samples():
    gp, opt, g = _store()
    with torch.no_grad():
        gp._scaling[[4, 11]] = torch.tensor([[1.0, 0.5, 0.2], [0.7, 1.2, 0.1]])   # exp(.) > 1.0 = threshold
    before = _snapshot(gp, opt)
    grads = torch.zeros(40, 1)
    grads[[4, 11, 20]] = 1.0                                     # 20 is small: not split
    gen = torch.Generator().manual_seed(99)
    gp.densify_and_split(grads, 0.5, scene_extent=100.0, N=2, generator=gen)
    assert gp._xyz.shape[0] == 40 - 2 + 4
    sel = torch.tensor([4, 11])
    keep = torch.ones(40, dtype=torch.bool)
    keep[sel] = False
    # survivors keep their values and moments; the 4 new ones: N copies in (sel, sel) order
    assert torch.equal(gp._xyz.detach()[:38], before["xyz"][0][keep])
    assert torch.equal(opt.state[gp._xyz]["exp_avg"][:38], before["xyz"][1][keep])
    assert float(opt.state[gp._xyz]["exp_avg"][38:].abs().max()) == 0.0
    old_scale = torch.exp(before["scaling"][0][sel]).repeat(2, 1)
    assert torch.allclose(torch.exp(gp._scaling.detach()[38:]), old_scale / 1.6, rtol=1e-6)
    assert torch.equal(gp._rotation.detach()[38:], before["rotation"][0][sel].repeat(2, 1))
    assert torch.equal(gp._features_rest.detach()[38:], before["f_rest"][0][sel].repeat(2, 1, 1))
    assert torch.equal(gp._opacity.detach()[38:], before["opacity"][0][sel].repeat(2, 1))
    # positions: old mean + R(q) * N(0, diag(scale^2)) with the same random stream
    gen2 = torch.Generator().manual_seed(99)
    samples = torch.normal(mean=torch.zeros(4, 3), std=old_scale, generator=gen2)
    R = build_rotation(before["rotation"][0][sel]).repeat(2, 1, 1)
    expect = torch.bmm(R, samples[:, :, None])[:, :, 0] + before["xyz"][0][sel].repeat(2, 1)
    assert torch.allclose(gp._xyz.detach()[38:], expect, atol=1e-6)
    q = before["rotation"][0][sel]
    Rm = build_rotation(q)
    assert torch.allclose(Rm @ Rm.transpose(1, 2), torch.eye(3).expand(2, 3, 3), atol=1e-6)


def test_densify_and_prune_round_and_opacity_reset():
    gp, opt, g = _store(P=60, seed=3)
    with torch.no_grad():
        gp._opacity[:5] = 0.001                                  # below min_opacity = 0.005: pruned at the end
        gp._opacity[5:] = gp._opacity[5:].clamp(0.05, 1.0)
        gp._scaling[10:14] = 1.0                                 # large -> split candidates
        gp.xyz_gradient_accum.zero_()
        gp.xyz_gradient_accum[8:16] = 1.0                        # 8, 9, 14, 15 small -> cloned; 10..13 -> split
        gp.denom.fill_(2.0)
        gp.denom[30] = 0.0                                       # 0 / 0 -> nan -> treated as 0
    gp.densify_and_prune(0.25, 100.0, generator=torch.Generator().manual_seed(1))
    # 60 + 4 clones = 64; split: + 8 new - 4 originals = 68; minus the 5 transparent ones = 63
    assert gp._xyz.shape[0] == 63
    for grp in opt.param_groups:
        p = grp["params"][0]
        assert p.shape[0] == 63 and opt.state[p]["exp_avg"].shape == p.shape and opt.state[p]["exp_avg_sq"].shape == p.shape
    assert gp.xyz_gradient_accum.shape == (63, 1) and gp.denom.shape == (63, 1) and gp.max_radii2D.shape == (63,)
    assert float(gp.get_opacity.detach().min()) >= 0.005
    # reset_opacity: capped at 0.1, moments of the opacity tensor zeroed, step count kept
    gp.reset_opacity()
    assert float(gp._opacity.detach().max()) <= 0.1 + 1e-7 and float(gp._opacity.detach().min()) >= 0.005
    st = opt.state[gp._opacity]
    assert float(st["exp_avg"].abs().max()) == 0.0 and float(st["exp_avg_sq"].abs().max()) == 0.0 and st["step"] == 7
    assert opt.param_groups[3]["params"][0] is gp._opacity

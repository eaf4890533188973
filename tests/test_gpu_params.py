"""GPU parity tests (pytest -m gpu) of the parameter-store kernels: the fused
activations against the reference's torch getters (scene/gaussian_model.py:114-137) and their autograd,
and the fused Adam step against torch.optim.Adam as the reference configures it
(scene/gaussian_model.py:181-190, train.py:204-208).

Bars: exp / clamp / cat outputs bit-exact; normalize and every gradient within 2 fp32 ulp-level relative
error (1e-6); Adam parameters after 25 steps within 1e-6 of torch's (same operation order, one rounding
per operation)."""
import math

import pytest
import torch

from deblurgs_b200 import _lib
from deblurgs_b200.params import FusedAdam, activate_gaussians

pytestmark = pytest.mark.gpu


def _raw(P, M, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    t = dict(dc=r(P, 1, 3), rest=r(P, M - 1, 3) * 0.1, scaling=r(P, 3) - 3.0, rotation=r(P, 4),
             opacity=torch.rand(P, 1, generator=g) * 1.4 - 0.2)   # some outside [0,1]: clamp gradient mask
    k = min(3, P)
    t["opacity"][:k, 0] = torch.tensor([0.0, 1.0, 0.5])[:k]       # exactly on the clamp bounds
    return {k: v.cuda().requires_grad_(True) for k, v in t.items()}


def _torch_getters(t, lb, isotropic):
    sc = t["scaling"][:, :1].expand(-1, 3) if isotropic else t["scaling"]
    return (torch.cat((t["dc"], t["rest"]), dim=1), torch.exp(sc) + lb,
            torch.nn.functional.normalize(t["rotation"]), t["opacity"].clamp(0.0, 1.0))


@pytest.mark.parametrize("P,M,lb,isotropic", [(1000, 16, 0.0, False), (4097, 4, 0.01, False), (333, 1, 0.0, False),
                                              (2048, 16, 0.0, True), (1, 9, 0.0, False)])
def test_activations_match_reference_getters(P, M, lb, isotropic):
    t = _raw(P, M)
    ours = activate_gaussians(t["dc"], t["rest"], t["scaling"], t["rotation"], t["opacity"], lb, isotropic)
    ref = _torch_getters(t, lb, isotropic)
    assert torch.equal(ours[0], ref[0])          # cat
    assert torch.equal(ours[1], ref[1])          # exp + lb (same expf)
    assert torch.equal(ours[3], ref[3])          # clamp
    assert (ours[2] - ref[2]).abs().max().item() <= 2e-7   # normalize: summation order may differ by an ulp
    # backward: random upstream gradients through both graphs
    g = torch.Generator(device="cpu").manual_seed(5)
    ups = [torch.randn(o.shape, generator=g).cuda() for o in ref]
    names = ["dc", "rest", "scaling", "rotation", "opacity"]
    gr_ref = torch.autograd.grad(ref, [t[n] for n in names], ups, allow_unused=True)
    gr_our = torch.autograd.grad(ours, [t[n] for n in names], ups, allow_unused=True)
    for n, a, b in zip(names, gr_our, gr_ref):
        if b is None or b.numel() == 0:
            continue
        err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
        assert err <= 1e-6, (n, err)
    # the clamp mask: zero gradient strictly outside [0,1], pass-through on the bounds (torch semantics)
    o = t["opacity"].detach()
    assert torch.equal(gr_our[4] != 0, ((o >= 0) & (o <= 1)) & (ups[3] != 0))


def test_activation_of_empty_store():
    t = {k: v for k, v in _raw(1, 16).items()}
    e = {k: v[:0].detach().requires_grad_(True) for k, v in t.items()}
    out = activate_gaussians(e["dc"], e["rest"], e["scaling"], e["rotation"], e["opacity"])
    assert [tuple(o.shape) for o in out] == [(0, 16, 3), (0, 3), (0, 4), (0, 1)]


def _reference_optimizer(params, lrs):
    groups = [{"params": [p], "lr": lr, "name": str(i)} for i, (p, lr) in enumerate(zip(params, lrs))]
    return torch.optim.Adam(groups, lr=0.0, eps=1e-15, foreach=False, fused=False)


@pytest.mark.parametrize("clip", [None, 0.02])
def test_fused_adam_matches_torch_adam(clip):
    g = torch.Generator(device="cpu").manual_seed(3)
    shapes = [(5000, 3), (5000, 1, 3), (5000, 15, 3), (5000, 1), (5000, 3), (5000, 4)]   # the reference's 6 groups
    lrs = [0.00016, 0.0025, 0.0025 / 20.0, 0.05, 0.005, 0.001]
    init = [torch.randn(*s, generator=g).cuda() for s in shapes]
    p_ref = [x.clone().requires_grad_(True) for x in init]
    p_our = [x.clone().requires_grad_(True) for x in init]
    o_ref = _reference_optimizer(p_ref, lrs)
    o_our = FusedAdam([{"params": [p], "lr": lr, "name": str(i)} for i, (p, lr) in enumerate(zip(p_our, lrs))],
                      lr=0.0, eps=1e-15)
    for it in range(25):
        grads = [torch.randn(*s, generator=g).cuda() * (10.0 ** (-(it % 5))) for s in shapes]
        if it == 7:
            grads[2] = None   # a parameter without gradient is skipped and keeps its step count
        for p, q, gr in zip(p_ref, p_our, grads):
            p.grad = None if gr is None else gr.clone()
            q.grad = None if gr is None else gr.clone()
        if it == 12:   # learning-rate schedule between steps (update_learning_rate)
            o_ref.param_groups[0]["lr"] = o_our.param_groups[0]["lr"] = 0.0001
        if clip is not None:
            torch.nn.utils.clip_grad_value_([p for p in p_ref if p.grad is not None], clip)
        o_ref.step()
        o_our.step(clip_grad_value=clip)
        o_ref.zero_grad(set_to_none=True)
        o_our.zero_grad(set_to_none=True)
    for i, (p, q) in enumerate(zip(p_ref, p_our)):
        err = (p - q).abs().max().item() / p.abs().max().item()
        assert err <= 1e-6, (i, err)
        s_ref, s_our = o_ref.state[p], o_our.state[q]
        assert int(s_ref["step"]) == s_our["step"]
        assert (s_ref["exp_avg"] - s_our["exp_avg"]).abs().max().item() <= 1e-6 * s_ref["exp_avg"].abs().max().item()
        assert (s_ref["exp_avg_sq"] - s_our["exp_avg_sq"]).abs().max().item() <= \
            1e-6 * s_ref["exp_avg_sq"].abs().max().item()


def test_fused_adam_state_dict_round_trip_and_odd_sizes():
    g = torch.Generator(device="cpu").manual_seed(4)
    # sizes around the 2048-element block chunk, plus more tensors than one launch holds (8)
    sizes = [1, 2047, 2048, 2049, 4096, 7, 100003, 3, 5, 11]
    ps = [torch.randn(n, generator=g).cuda().requires_grad_(True) for n in sizes]
    ref = [p.detach().clone().requires_grad_(True) for p in ps]
    o1 = FusedAdam(ps, lr=0.01, eps=1e-15)
    o2 = torch.optim.Adam(ref, lr=0.01, eps=1e-15, foreach=False, fused=False)
    for it in range(3):
        for p, q in zip(ps, ref):
            gr = torch.randn(p.shape, generator=g).cuda()
            p.grad, q.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    for p, q in zip(ps, ref):
        assert (p - q).abs().max().item() <= 1e-6 * max(q.abs().max().item(), 1.0)
    sd = o1.state_dict()
    assert set(sd.keys()) == {"state", "param_groups"} and len(sd["state"]) == len(sizes)
    o3 = FusedAdam([p.detach().clone().requires_grad_(True) for p in ps], lr=0.01, eps=1e-15)
    o3.load_state_dict(sd)
    flat3 = [p for gr in o3.param_groups for p in gr["params"]]
    for p, q in zip(ps, flat3):
        gr = torch.randn(p.shape, generator=g).cuda()
        p.grad, q.grad = gr.clone(), gr.clone()
    o1.step()
    o3.step()
    for p, q in zip(ps, flat3):
        assert torch.equal(p, q)


def test_training_step_through_the_store_reduces_the_loss():
    """render_blurry reading the fused activations, backward through them, FusedAdam.step: a few iterations
    of the reference's inner loop (train.py:134-208 without densification) must lower the L1 loss."""
    from tests import parity_utils as pu
    from deblurgs_b200 import renderer
    from deblurgs_b200.motion import GaussianParams
    cam, scene, traj, bg, view, proj, campos = pu.make_inputs("tiny")
    gauss = GaussianParams.from_scene(scene)
    opt = gauss.training_setup(position_lr_init=0.0, feature_lr=0.02, opacity_lr=0.02, scaling_lr=0.005,
                               rotation_lr=0.001)

    class ref_cam:   # noqa: N801  (what render_blurry reads from the reference camera)
        image_width, image_height, FoVx, FoVy = cam.width, cam.height, cam.fovx, cam.fovy
    gt = torch.rand(3, cam.height, cam.width, generator=torch.Generator().manual_seed(2)).cuda()
    losses = []
    for it in range(12):
        pkg = renderer.render_blurry(view, proj, campos, ref_cam, gauss, bg)
        loss = (pkg["blurred"] - gt).abs().mean()
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(loss.item())
    assert math.isfinite(losses[-1]) and losses[-1] < losses[0]


def test_regularisers_match_the_reference_definitions():
    """tv_loss / hinge_l2 (utils/loss_utils.py:66-78, 95-104) and the weighted training loss (train.py:147-163)."""
    from deblurgs_b200.loss import blur_photometric_loss, hinge_l2, training_loss, tv_loss
    g = torch.Generator().manual_seed(9)

    def ref_tv(x):
        return ((x[:, :, :-1, :] - x[:, :, 1:, :]) ** 2).mean() + ((x[:, :, :, :-1] - x[:, :, :, 1:]) ** 2).mean()

    def ref_hinge(x):
        loss = torch.zeros_like(x)
        loss = torch.where(x <= 0.0, x ** 2, loss)
        loss = torch.where(x >= 1.0, (x - 1.0) ** 2, loss)
        return loss.mean()

    for shape in [(5, 1, 37, 53), (2, 3, 16, 16), (1, 1, 2, 2)]:
        x = (torch.rand(shape, generator=g) * 10).cuda().requires_grad_(True)
        a, b = tv_loss(x), ref_tv(x)
        ga, = torch.autograd.grad(a * 1.3, x)
        gb, = torch.autograd.grad(b * 1.3, x)
        assert abs(a.item() - b.item()) <= 1e-5 * abs(b.item()) and torch.allclose(ga, gb, rtol=1e-4, atol=1e-7)
    o = (torch.rand(20_000, 1, generator=g) * 1.4 - 0.2).cuda()
    o[:5, 0] = torch.tensor([0.0, 1.0, -0.5, 1.5, 0.5])
    o.requires_grad_(True)
    a, b = hinge_l2(o), ref_hinge(o)
    ga, = torch.autograd.grad(a * 0.7, o)
    gb, = torch.autograd.grad(b * 0.7, o)
    assert abs(a.item() - b.item()) <= 1e-6 and torch.allclose(ga, gb, rtol=1e-5, atol=1e-10)
    # the reference's weighted sum
    F, H, W = 4, 24, 40
    sub = torch.rand(F, 3, H, W, generator=g).cuda().requires_grad_(True)
    dep = (torch.rand(F, 1, H, W, generator=g) * 5).cuda().requires_grad_(True)
    blur = sub.mean(0)
    gt = torch.rand(3, H, W, generator=g).cuda()
    mine = training_loss(blur, sub, gt, dep, o, lambda_t_smooth=0.01, lambda_depth_tv=0.1, lambda_hinge=0.1)
    ref = (blur - gt).abs().mean() + 0.01 * (sub[1:] - sub[:-1]).abs().mean() + 0.1 * ref_tv(dep) + 0.1 * ref_hinge(o)
    assert abs(mine.item() - ref.item()) <= 1e-5 * abs(ref.item())
    gm = torch.autograd.grad(mine, [sub, dep, o])
    gr = torch.autograd.grad(ref, [sub, dep, o])
    for x, y in zip(gm, gr):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-9)
    # a zero weight gates its term off like the reference
    assert training_loss(blur, sub, gt, None, None, 0.0, 0.0, 0.0).item() == blur_photometric_loss(blur, sub, gt, 0.0).item()


def test_fused_densify_rebuild_equals_the_cpu_plan():
    """One densify_and_prune round on the GPU (one row-gather launch per rebuild, csrc/dgs_params.cu) against the same
    round executed with torch indexing on CPU tensors, which tests/test_densify.py pins to the reference's semantics
    (scene/gaussian_model.py:300-454).  The split's random offsets come from different generators on the two
    devices, so the new positions are compared through the CPU's samples fed to both."""
    from tests.test_densify import NAMES, ATTR, _store
    from deblurgs_b200.params import FusedAdam
    lib = _lib.load()
    stores = {}
    for dev in ("cpu", "cuda"):
        gp, opt, g = _store(P=3000, M=16, seed=5)
        if dev == "cuda":
            for name in NAMES:
                p = getattr(gp, ATTR[name])
                q = torch.nn.Parameter(p.detach().cuda())
                st = opt.state.pop(p)
                opt.state[q] = {"step": st["step"], "exp_avg": st["exp_avg"].cuda(), "exp_avg_sq": st["exp_avg_sq"].cuda()}
                [grp for grp in opt.param_groups if grp["name"] == name][0]["params"][0] = q
                setattr(gp, ATTR[name], q)
            gp.xyz_gradient_accum, gp.denom, gp.max_radii2D = gp.xyz_gradient_accum.cuda(), gp.denom.cuda(), gp.max_radii2D.cuda()
        stores[dev] = (gp, opt)
    # same random offsets on both devices: draw on the CPU, replay on the GPU
    drawn = []
    real_normal = torch.normal

    def record(mean, std, generator=None):
        out = real_normal(mean=mean.cpu(), std=std.cpu(), generator=torch.Generator().manual_seed(len(drawn)))
        drawn.append(out)
        return out.to(std.device)
    before = lib.dgs_launch_count(0)
    torch.normal = record
    try:
        for dev in ("cpu", "cuda"):
            drawn.clear()
            gp, opt = stores[dev]
            gp.percent_dense = 0.01
            gp.densify_and_prune(max_grad=0.5, extent=20.0)
    finally:
        torch.normal = real_normal
    assert lib.dgs_launch_count(0) - before == 3          # clone, split, prune: one launch each
    (gc, oc), (gg, og) = stores["cpu"], stores["cuda"]
    assert gg._xyz.shape[0] == gc._xyz.shape[0] and gg._xyz.shape[0] != 3000
    for name in NAMES:
        pc, pg = getattr(gc, ATTR[name]), getattr(gg, ATTR[name])
        assert pg.is_cuda and isinstance(pg, torch.nn.Parameter) and pg.requires_grad
        # (the split's new positions go through a rotation matmul: CPU and GPU round it differently)
        tol = 1e-4 if name == "xyz" else 1e-6
        assert torch.allclose(pg.detach().cpu(), pc.detach(), rtol=tol, atol=tol), name
        sc, sg = oc.state[pc], og.state[pg]
        assert torch.equal(sg["exp_avg"].cpu(), sc["exp_avg"]) and torch.equal(sg["exp_avg_sq"].cpu(), sc["exp_avg_sq"])
        assert float(sg["step"]) == float(sc["step"])
    assert torch.equal(gg.max_radii2D.cpu(), gc.max_radii2D) and torch.equal(gg.denom.cpu(), gc.denom)
    # and the rebuilt store still trains: one fused Adam step over the new tensors
    for name in NAMES:
        p = getattr(gg, ATTR[name])
        p.grad = torch.ones_like(p)
    og.step()

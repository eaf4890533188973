"""GPU: the blurry-view step replayed as a CUDA graph (deblurgs_b200.graph.BlurryViewGraph) gives the loss and the
gradients of the kernel-by-kernel step, survives a changed input, and recovers from an exceeded binning capacity."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _eager(w, gt, lam):
    from deblurgs_b200.loss import blur_photometric_loss
    params = w["gaussians"].parameters() + w["cmm"].parameters()
    for p in params:
        p.grad = None
    out = w["cmm"].query(0, "all", background=w["bg"])
    loss = blur_photometric_loss(out["blurred"], out["subframes"], gt, lam)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), [p.grad.clone() for p in params]


@pytest.mark.parametrize("lam", [0.0, 1e-3])
def test_graph_replay_equals_eager_step(lam):
    import bench
    from deblurgs_b200.graph import BlurryViewGraph
    dev = torch.device("cuda", 0)
    w = bench.build_workload("small", 0, dev)
    params = w["gaussians"].parameters() + w["cmm"].parameters()
    gt = w["gt_host"].to(dev)
    g = BlurryViewGraph(w["cmm"], 0, w["bg"], (3, w["H"], w["W"]), lam)
    assert g.launches_per_replay >= 20
    for gt_k in (gt, 1.0 - gt, gt):           # a changed input, then the first one again
        l_ref, g_ref = _eager(w, gt_k, lam)   # (replaces the .grad tensors: the graph re-attaches its own on replay)
        val = g.step(gt_k)
        assert abs(val - l_ref) <= 1e-6
        for p, r in zip(params, g_ref):
            assert ((p.grad - r).abs().max() / r.abs().max().clamp_min(1e-30)).item() <= 1e-3
    assert g.recaptures == 0 and g.num_rendered() > 0


def test_graph_recovers_from_exceeded_capacity():
    import bench
    from deblurgs_b200.graph import BlurryViewGraph
    dev = torch.device("cuda", 0)
    w = bench.build_workload("small", 0, dev)
    gt = w["gt_host"].to(dev)
    params = w["gaussians"].parameters() + w["cmm"].parameters()
    l_ref, g_ref = _eager(w, gt, 0.0)
    g = BlurryViewGraph(w["cmm"], 0, w["bg"], (3, w["H"], w["W"]))
    g.step(gt)
    need = g.num_rendered()
    g._capture(max(need // 3, 1))              # a capacity the scene has outgrown
    g.gt.copy_(gt)
    g.replay()
    torch.cuda.synchronize()
    assert int(g.status_host[4]) == 1          # overflow seen on the device, reported through the graph's memcpy node
    assert g.check() is True and g.recaptures == 1 and g.capacity >= need
    assert abs(g.loss.item() - l_ref) <= 1e-6
    for p, r in zip(params, g_ref):
        assert ((p.grad - r).abs().max() / r.abs().max().clamp_min(1e-30)).item() <= 1e-3

"""Multi-GPU check of the sub-frame-sharded blurry view (run under torchrun on a GPU box):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gpu_check.py [config]

Every rank renders its contiguous block of sub-frames (deblurgs_b200.dist.render_blurry_sharded), the partial
blurry images are all-reduced over NCCL, every rank evaluates the same (L2) loss, and the Gaussian / control-point
gradients are all-reduced.  Rank 0 compares the result with the unsharded computation on its own GPU.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deblurgs_b200 import dist as dd  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "small"
    rank, world, local = bench.dist_setup(0)
    dev = torch.device("cuda", local)
    w = bench.build_workload(cfg, 0, dev)          # same view on every rank (seed independent of rank)
    cmm, g = w["cmm"], w["gaussians"]
    gt = w["gt_host"].to(dev)
    params = g.parameters() + cmm.parameters()

    def run(sharded, use_sink=False):
        for p in params:
            p.grad = None
        sink = None
        if use_sink:      # gradient sink: finished rows are all-reduced while the backward still runs
            sink = dd.FlatGradBuffer.for_gaussians(g)
        g.grad_sink = sink
        if sharded:
            blurred, pkg, (a, b) = dd.render_blurry_sharded(cmm, 0, w["bg"])
        else:
            out = cmm.query(0, "all", background=w["bg"])
            blurred = out["blurred"]
        # smooth (L2) loss for the equality check: with L1, sign(blurred - gt) flips on the handful of
        # pixels where the two summation orders of the blurred image straddle gt (seen at 1080p: a few
        # Gaussians move by ~1e-3 of the max gradient although both paths are exact)
        loss = ((blurred - gt) ** 2).mean()
        loss.backward()
        if sink is not None:
            sink.wait()
        grads = [p.grad.clone() for p in params]
        if sharded and world > 1:
            for t, p in zip(grads, params):
                if sink is None or not any(p is q for q in g.parameters()):    # the sink has reduced the Gaussians' already
                    dist.all_reduce(t)
        stats = (pkg if sharded else out["batched"])["densification"]
        if sharded:
            dd.all_reduce_densification_stats(stats)
        return blurred.detach(), loss.detach(), grads, stats

    b1, l1, g1, s1 = run(True)
    # timing of the sharded step (max over ranks)
    torch.cuda.synchronize(); dist.barrier() if world > 1 else None
    t0 = time.perf_counter()
    for _ in range(10):
        run(True)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / 10], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    b2, l2, g2, s2 = run(True, use_sink=True)

    # the bench's steps, kernel by kernel vs as ONE CUDA-graph replay with the NCCL collectives as graph nodes
    # (deblurgs_b200.nccl_direct): same loss, same gradients
    def bench_step(make, graph_mode):
        step = make(graph_mode)
        for _ in range(2):
            loss = step(gt)
        torch.cuda.synchronize()
        return loss.detach().clone(), [p.grad.detach().clone() for p in params]
    graph_report = []
    for name, make in (("subframes", lambda gm: bench.make_step_subframe_sharded(w, world, gm)),
                       ("views", lambda gm: bench.make_step_ours(w, world, 0.0, gm))):
        le, ge = bench_step(make, False)
        lg, gg = bench_step(make, True)
        relg = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(gg, ge))
        assert abs(lg.item() - le.item()) <= 1e-6 * max(1.0, abs(le.item())) and relg <= 1e-3, (name, lg.item(), le.item(), relg)
        graph_report.append("%s %.1e" % (name, relg))
    g.grad_sink = None
    if rank == 0:
        print("graph replay with in-graph collectives vs kernel by kernel, max gradient rel diff: " + ", ".join(graph_report))
        # the overlapped (gradient sink) path gives the plain path's gradients
        rel_sink = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(g2, g1))
        assert (b2 - b1).abs().max().item() <= 1e-6 and rel_sink <= 1e-3, rel_sink
        b0, l0, g0, s0 = run(False)
        # densification statistics of the sharded view (all-reduced) = those of the unsharded render
        assert s1.num_subframes == s0.num_subframes
        assert torch.equal(s1.visible_count, s0.visible_count) and torch.equal(s1.max_radius, s0.max_radius)
        assert ((s1.grad_norm_sum - s0.grad_norm_sum).abs().max() / s0.grad_norm_sum.abs().max()).item() <= 1e-3
        err_img = (b0 - b1).abs().max().item()
        rel = max(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(g1, g0))
        print("config %s world %d: sharded step %.2f ms; blurred max-abs diff %.2e; loss %.6f vs %.6f; "
              "max gradient rel diff %.2e" % (cfg, world, dt.item() * 1e3, err_img, l1.item(), l0.item(), rel))
        assert err_img <= 1e-5 and rel <= 1e-3
        print("OK")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""BASELINE.json config 5 (reduced): fwd+bwd blurry-view time for num_subframes x Gaussians sweep points, this
library vs the reference extension (baseline/_ref) on the same GPU.  Prints one JSON line per point.

  python tests/sweep.py > gpurun_out/sweep.jsonl                                                  # 1 GPU, both arms
  python -m torch.distributed.run --nproc-per-node 8 ... tests/sweep.py > gpurun_out/sweep_n8.jsonl  # N GPUs: this
      library with one view per rank (gradient all-reduces inside the replayed graph); the reference is single-GPU and is
      timed at N = 1 only
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = bench.dist_setup(1)
    dev = torch.device("cuda", local)
    have_ref = world == 1 and os.path.isdir(os.path.join(bench.ROOT, "baseline", "_ref", "diff_gaussian_rasterization"))
    points = [(100_000, 2), (100_000, 4), (100_000, 16), (100_000, 64), (1_000_000, 2), (1_000_000, 4),
              (1_000_000, 16), (1_000_000, 64), (5_000_000, 16)]
    for P, F in points:
        cfg = "P=%d,F=%d" % (P, F)
        if F < 3:
            cfg += ",C=3"
        w = bench.build_workload(cfg, rank, dev)
        gt = w["gt_host"].to(dev)
        step = bench.make_step_ours(w, world, 0.0, use_graph=True)
        for _ in range(3):
            step(gt)
        ours = bench.time_steps(step, gt, 5, world, dev)
        w["gaussians"].grad_sink = None
        ref = None
        if have_ref and P * F <= 64_000_000:
            rstep = bench.make_step_reference(w)
            rstep(gt)
            ref = bench.time_steps(rstep, gt, 2, 1, dev)
        if rank == 0:
            line = {"P": P, "F": F, "W": w["W"], "H": w["H"], "n_gpus": world, "ours_ms_per_step": round(ours, 3),
                    "ours_views_per_s": round(world * 1000.0 / ours, 2), "launch_mode": step.mode,
                    "ref_ms": None if ref is None else round(ref, 3),
                    "speedup_vs_ref_1gpu": None if ref is None else round(ref / ours, 1),
                    "mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        del w, gt, step
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

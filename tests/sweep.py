"""BASELINE.json config 5 (reduced): fwd+bwd blurry-view time for num_subframes x Gaussians sweep points, this
library vs the reference extension (baseline/_ref) on the same GPU.  Prints one JSON line per point.
  python tests/sweep.py > gpurun_out/sweep.jsonl"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def timeit(step, gt, warm, n):
    for _ in range(warm):
        step(gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(gt)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    bench.dist_setup(1)
    dev = torch.device("cuda", 0)
    have_ref = os.path.isdir(os.path.join(bench.ROOT, "baseline", "_ref", "diff_gaussian_rasterization"))
    points = [(100_000, 2), (100_000, 4), (100_000, 16), (100_000, 64), (1_000_000, 2), (1_000_000, 4),
              (1_000_000, 16), (1_000_000, 64), (5_000_000, 16)]
    for P, F in points:
        cfg = "P=%d,F=%d" % (P, F)
        if F < 3:
            cfg += ",C=3"
        w = bench.build_workload(cfg, 0, dev)
        gt = w["gt_host"].to(dev)
        ours = timeit(bench.make_step_ours(w, 1), gt, 3, 5)
        ref = None
        if have_ref and P * F <= 64_000_000:
            ref = timeit(bench.make_step_reference(w), gt, 1, 2)
        print(json.dumps({"P": P, "F": F, "W": w["W"], "H": w["H"], "ours_ms": round(ours, 3),
                          "ref_ms": None if ref is None else round(ref, 3),
                          "speedup": None if ref is None else round(ref / ours, 1),
                          "mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}), flush=True)
        del w, gt
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()

"""Single-GPU step time of the eight views the 8-rank views-dp bench renders (one per rank): how much of the
N = 8 step (max over ranks) is the spread between views rather than the collectives."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
out = []
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    w = bench.build_workload("c2", r, dev)
    step = bench.make_step_ours(w, 1, 0.0, True)
    gt = w["gt_host"].to(dev)
    for _ in range(5):
        step(gt)
    ms = bench.time_steps(step, gt, 20, 1, dev)
    out.append(round(ms, 4))
    del w, step
    torch.cuda.empty_cache()
print(json.dumps({"ms_per_view": out, "max": max(out), "mean": sum(out) / len(out)}))

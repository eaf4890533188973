import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from deblurgs_b200 import _lib
rank, world, local = bench.dist_setup(1)
dev = torch.device("cuda", 0)
for cfg in sys.argv[1:]:
    w = bench.build_workload(cfg, 0, dev)
    step = bench.make_step_ours(w, 1)
    gt = w["gt_host"].to(dev)
    for i in range(3): step(gt)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 5
    for i in range(n): l = step(gt)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e3
    st = bench.workload_stats(w)
    print(cfg, "ms/step %.2f" % dt, "loss %.5f" % l.item(), st, "mem GB %.1f" % (torch.cuda.max_memory_allocated() / 2**30), flush=True)
    del w, step, gt
    torch.cuda.empty_cache()

"""torchrun-launched check of deblurgs_b200.nccl_direct (run by tests/test_gpu_dist.py when >= 2 GPUs are visible):
an eager all-reduce, then a CUDA graph whose side-stream branch carries two all-reduces while the capturing stream
computes, replayed three times."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from deblurgs_b200.nccl_direct import DirectComm
    comm = DirectComm(dev)
    x = torch.full((1 << 20,), float(rank + 1), device=dev)
    comm.all_reduce_(x)
    torch.cuda.synchronize()
    want = world * (world + 1) / 2
    assert float(x[0]) == want and float(x[-1]) == want, (float(x[0]), want)

    a = torch.zeros(1 << 22, device=dev)
    b = torch.zeros(1 << 22, device=dev)
    c = torch.zeros(1 << 20, device=dev)
    side = torch.cuda.Stream(dev)
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        with torch.cuda.graph(g, stream=cap):
            a.fill_(float(rank + 1))
            side.wait_stream(cap)
            comm.all_reduce_(a, stream=side)           # overlaps the work below
            b.fill_(2.0 * (rank + 1))
            c.add_(1.0)
            side.wait_stream(cap)
            comm.all_reduce_(b, stream=side)
            cap.wait_stream(side)
    torch.cuda.current_stream().wait_stream(cap)
    for i in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert float(a[5]) == want and float(b[7]) == 2 * want and float(c[0]) == 3.0, (float(a[5]), float(b[7]), float(c[0]))
    dist.barrier()
    if rank == 0:
        print("nccl_direct OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

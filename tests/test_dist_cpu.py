"""CPU, world_size 2 over gloo: host-side sharding and gradient all-reduce logic of deblurgs_b200.dist."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deblurgs_b200 import dist as dd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # sub-frame blocks: contiguous, disjoint, cover [0, F)
        F = 7
        a, b = dd.subframe_shard(F, rank, world)
        blocks = [None] * world
        dist.all_gather_object(blocks, (a, b))
        assert blocks[0][0] == 0 and blocks[-1][1] == F
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        # partial blurry images sum to the full mean; backward is the identity
        g = torch.Generator().manual_seed(0)
        frames = torch.rand(F, 3, 4, 5, generator=g)
        part = (frames[a:b].sum(0) / F).requires_grad_(True)
        full = dd.all_reduce_sum(part)
        assert torch.allclose(full, frames.mean(0), atol=1e-6)
        w = torch.rand(3, 4, 5, generator=g)
        (full * w).sum().backward()
        assert torch.allclose(part.grad, w)
        # flat gradient buffer: per-rank grads are summed over ranks, parameters keep their views
        p1 = torch.nn.Parameter(torch.zeros(5, 3))
        p2 = torch.nn.Parameter(torch.zeros(4))
        fb = dd.FlatGradBuffer([p1, p2])
        fb.zero()
        loss = (p1 * (rank + 1)).sum() + (p2 * 10 * (rank + 1)).sum()
        loss.backward()
        assert p1.grad.data_ptr() == fb.flat.data_ptr()
        fb.all_reduce()
        tot = sum(r + 1 for r in range(world))
        assert torch.allclose(p1.grad, torch.full((5, 3), float(tot)))
        assert torch.allclose(p2.grad, torch.full((4,), 10.0 * tot))
        # densification statistics of a sub-frame-sharded view: every rank ends up with the whole view's statistics
        # (sums add, radii take the maximum), normalised by the view's FULL sub-frame count on every rank
        from deblurgs_b200.rasterizer import DensificationStats
        from deblurgs_b200.motion import GaussianParams
        P = 6
        gen = torch.Generator().manual_seed(3)
        per_sub = torch.rand(F, P, 3, generator=gen)                    # (norm, visible, radius) of every sub-frame
        per_sub[..., 1] = (per_sub[..., 1] > 0.4).float()
        per_sub[..., 2] = (per_sub[..., 2] * 50).floor()
        local = torch.stack([per_sub[a:b, :, 0].sum(0), per_sub[a:b, :, 1].sum(0), per_sub[a:b, :, 2].max(0).values], 1)
        st = DensificationStats()
        st.fill(local, float(F))
        dd.all_reduce_densification_stats(st)
        assert torch.allclose(st.grad_norm_sum[:, 0], per_sub[:, :, 0].sum(0), atol=1e-6)
        assert torch.equal(st.visible_count[:, 0], per_sub[:, :, 1].sum(0))
        assert torch.equal(st.max_radius, per_sub[:, :, 2].max(0).values.to(torch.int32))
        gp = GaussianParams(torch.zeros(P, 3), torch.zeros(P, 1, 3), torch.zeros(P, 3, 3), torch.zeros(P, 3),
                            torch.ones(P, 4), torch.zeros(P, 1), 1)
        gp.add_densification_stats_blurry({"densification": st})
        assert torch.allclose(gp.denom[:, 0], per_sub[:, :, 1].sum(0) / F)      # 1 / len(render_pkgs) of the whole view
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_subframe_shard_balanced():
    for F in (1, 2, 7, 16, 21, 32):
        for ws in (1, 2, 4, 8):
            blocks = [dd.subframe_shard(F, r, ws) for r in range(ws)]
            sizes = [b - a for a, b in blocks]
            assert sum(sizes) == F and max(sizes) - min(sizes) <= 1
            assert blocks[0][0] == 0 and blocks[-1][1] == F


def test_direct_nccl_wrapper_binds_the_library_torch_ships_and_stays_off_without_nccl():
    """deblurgs_b200.nccl_direct: the libnccl it would call exports every entry point it binds (no GPU needed to load
    it), and the dispatcher in dist.py keeps to torch.distributed when there is no NCCL process group (CPU / gloo)."""
    import ctypes
    import os
    from deblurgs_b200 import dist as dd, nccl_direct
    path = nccl_direct._find_library()
    if os.path.isabs(path) and os.path.exists(path):
        lib = ctypes.CDLL(path)
        for name in ("ncclGetUniqueId", "ncclCommInitRank", "ncclAllReduce", "ncclGroupStart", "ncclGroupEnd",
                     "ncclCommDestroy", "ncclGetErrorString"):
            assert hasattr(lib, name), name
    assert dd.direct_comm(torch.device("cpu")) is None
    assert dd.direct_comm(torch.device("cuda", 0)) is None          # no process group in this process
    buf = dd.FlatGradBuffer([torch.nn.Parameter(torch.zeros(4, 3))])
    assert buf.comm is None and buf.comm_stream is None

"""`distCUDA2` drop-in (reference: submodules/simple-knn/spatial.cu:15-26, called from
scene/gaussian_model.py:158)."""
import ctypes as C

import torch

from . import _lib


def distCUDA2(points):
    lib = _lib.load()
    if not points.is_cuda:
        raise _lib.DgsError("points must be a CUDA tensor: libdgs_b200 has no CPU path")
    pts = points.detach().float().contiguous()
    P = pts.shape[0]
    out = torch.zeros(P, dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    scratch = torch.empty(int(lib.dgs_knn_scratch_bytes(P)), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        rc = lib.dgs_knn_mean_dist2(P, _lib.ptr(pts), _lib.ptr(out), _lib.ptr(scratch),
                                    C.c_void_p(torch.cuda.current_stream(pts.device).cuda_stream))
    _lib.check(rc, "dgs_knn_mean_dist2")
    return out

"""NCCL called directly (ctypes on the libnccl.so.2 that torch itself loads), for the collectives that sit INSIDE a
captured CUDA graph.

torch.distributed's ProcessGroupNCCL routes every collective through its own stream, event and work bookkeeping; inside a
stream capture that hung on this build (round 2, first half: the graph-replayed step therefore all-reduced the gradients
AFTER the replay, serially).  A plain `ncclAllReduce(sendbuf, recvbuf, count, type, op, comm, stream)` on a stream that
has joined the capture is an ordinary captured kernel node: the gradient all-reduces of finished rows run on a side
stream of the same graph while the rest of the per-Gaussian backward is still computing, and the sub-frame-sharded
step (image all-reduce in the middle of it) becomes one graph replay as well.

The communicator is created once per process (rank / world size / the unique id travel over the already initialised
torch.distributed group); torch.distributed stays the plumbing for everything that is not on the hot path.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

_NCCL_FLOAT32 = 7      # ncclDataType_t / ncclRedOp_t values of nccl.h
_NCCL_SUM = 0


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def _find_library():
    try:
        import nvidia.nccl as pkg       # the wheel torch depends on: the copy torch's own NCCL backend has loaded
        for base in list(pkg.__path__):
            p = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:
        pass
    return "libnccl.so.2"


class NcclError(RuntimeError):
    pass


class DirectComm:
    """One NCCL communicator over the ranks of `group` (default: the world), bound to this process's CUDA device."""

    def __init__(self, device, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise NcclError("torch.distributed is not initialised")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.lib = C.CDLL(_find_library())
        L = self.lib
        L.ncclGetErrorString.restype = C.c_char_p
        L.ncclGetErrorString.argtypes = [C.c_int]
        L.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        L.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        L.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ncclCommDestroy.argtypes = [C.c_void_p]
        L.ncclGroupStart.argtypes = []
        L.ncclGroupEnd.argtypes = []
        uid = _UniqueId()
        if self.rank == 0:
            self._check(L.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        box = [bytes(uid.internal) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        C.memmove(C.byref(uid), box[0], 128)
        self.comm = C.c_void_p()
        with torch.cuda.device(self.device):
            self._check(L.ncclCommInitRank(C.byref(self.comm), self.world, uid, self.rank), "ncclCommInitRank")
            # first collectives eagerly, over the size classes the step uses: connections, channels and protocol
            # buffers are set up outside any capture
            for n in (1 << 10, 1 << 18, 1 << 23):
                t = torch.zeros(n, dtype=torch.float32, device=self.device)
                self.all_reduce_(t)
            torch.cuda.synchronize(self.device)

    def _check(self, rc, what):
        if rc != 0:
            raise NcclError("%s failed: %s" % (what, self.lib.ncclGetErrorString(rc).decode()))

    def all_reduce_(self, tensor, stream=None):
        """In-place fp32 SUM on `stream` (default: the current stream of the tensor's device)."""
        if tensor.dtype != torch.float32 or not tensor.is_contiguous() or tensor.device != self.device:
            raise NcclError("all_reduce_: contiguous float32 tensor on %s expected" % (self.device,))
        if tensor.numel() == 0:
            return tensor
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        p = C.c_void_p(tensor.data_ptr())
        self._check(self.lib.ncclAllReduce(p, p, tensor.numel(), _NCCL_FLOAT32, _NCCL_SUM, self.comm,
                                           C.c_void_p(st.cuda_stream)), "ncclAllReduce")
        return tensor

    def all_reduce_group_(self, tensors, stream=None):
        """Several in-place all-reduces fused into one NCCL launch (ncclGroupStart / ncclGroupEnd)."""
        tensors = [t for t in tensors if t.numel() > 0]
        if not tensors:
            return
        self._check(self.lib.ncclGroupStart(), "ncclGroupStart")
        try:
            for t in tensors:
                self.all_reduce_(t, stream)
        finally:
            self._check(self.lib.ncclGroupEnd(), "ncclGroupEnd")

"""Fused photometric loss of the blurry-view training step (reference: train.py:147-163,
utils/loss_utils.py:17-18, 80-93): loss = l1_loss(blurred, gt) + lambda_t_smooth *
batchwise_smoothness_loss(subframes), forward and backward in one kernel each."""
import ctypes as C

import torch

from . import _lib


class _BlurLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, blurred, subframes, gt, lambda_t_smooth):
        lib = _lib.load()
        if not blurred.is_cuda:
            raise _lib.DgsError("blurred must be a CUDA tensor: libdgs_b200 has no CPU path")
        ctx.set_materialize_grads(False)
        lam = float(lambda_t_smooth)
        b = blurred.detach().float().contiguous()
        g = gt.detach().float().contiguous()
        F = subframes.shape[0]
        chw = b.numel()
        if subframes.numel() != F * chw or g.numel() != chw:
            raise _lib.DgsError("shape mismatch: subframes [F,3,H,W], blurred / gt [3,H,W]")
        # lambda_t_smooth == 0 (plain L1, e.g. pose refinement): the sub-frame stack is neither read nor given a gradient
        s = subframes.detach().float().contiguous() if lam != 0.0 else None
        out = torch.empty(3, dtype=torch.float32, device=b.device)
        scratch = torch.empty(2, dtype=torch.float64, device=b.device)
        with torch.cuda.device(b.device):
            rc = lib.dgs_blur_loss_forward(F, chw, _lib.ptr(s), _lib.ptr(b), _lib.ptr(g), float(lambda_t_smooth),
                                           _lib.ptr(out), _lib.ptr(scratch),
                                           C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream))
        _lib.check(rc, "dgs_blur_loss_forward")
        ctx.save_for_backward(b, g) if s is None else ctx.save_for_backward(b, g, s)
        ctx.lam = lam
        ctx.F = F
        ctx.shapes = (blurred.shape, subframes.shape)
        ctx.parts = out
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        if grad_out is None:
            return None, None, None, None
        saved = ctx.saved_tensors
        b, g = saved[0], saved[1]
        s = saved[2] if len(saved) > 2 else None
        F, chw = ctx.F, b.numel()
        go = grad_out.detach().float().contiguous().reshape(1)
        db = torch.empty_like(b)
        ds = torch.empty_like(s) if s is not None else None
        with torch.cuda.device(b.device):
            rc = lib.dgs_blur_loss_backward(F, chw, _lib.ptr(s), _lib.ptr(b), _lib.ptr(g), ctx.lam, _lib.ptr(go),
                                            _lib.ptr(db), _lib.ptr(ds),
                                            C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream))
        _lib.check(rc, "dgs_blur_loss_backward")
        return db.view(ctx.shapes[0]), (ds.view(ctx.shapes[1]) if ds is not None else None), None, None


def blur_photometric_loss(blurred, subframes, gt, lambda_t_smooth=0.0):
    """l1_loss(blurred, gt) + lambda_t_smooth * batchwise_smoothness_loss(subframes) as a 0-dim tensor."""
    return _BlurLoss.apply(blurred, subframes, gt, lambda_t_smooth)

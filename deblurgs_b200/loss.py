"""Loss block of the blurry-view training step (reference: train.py:147-163, utils/loss_utils.py):
`blur_photometric_loss` = l1_loss(blurred, gt) + lambda_t_smooth * batchwise_smoothness_loss(subframes), forward and
backward in one kernel each; `tv_loss` (depth smoothness) and `hinge_l2` (opacity range penalty), one kernel each way;
`training_loss` = the reference's weighted sum of the four."""
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


def _st(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _TV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.DgsError("tv_loss: CUDA tensor required (libdgs_b200 has no CPU path)")
        if x.dim() != 4:
            raise _lib.DgsError("tv_loss: x must be [B,C,H,W]")
        xc = x.detach().float().contiguous()
        B, Cc, H, W = xc.shape
        out = torch.empty(1, dtype=torch.float32, device=xc.device)
        scratch = torch.empty(2, dtype=torch.float64, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.check(lib.dgs_tv_loss_forward(B * Cc, H, W, _lib.ptr(xc), _lib.ptr(out), _lib.ptr(scratch), _st(xc)),
                       "dgs_tv_loss_forward")
        ctx.save_for_backward(xc)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        (xc,) = ctx.saved_tensors
        B, Cc, H, W = xc.shape
        go = grad_out.detach().float().contiguous().reshape(1)
        dx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(lib.dgs_tv_loss_backward(B * Cc, H, W, _lib.ptr(xc), _lib.ptr(go), _lib.ptr(dx), _st(xc)),
                       "dgs_tv_loss_backward")
        return dx


class _Hinge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.DgsError("hinge_l2: CUDA tensor required (libdgs_b200 has no CPU path)")
        xc = x.detach().float().contiguous()
        out = torch.empty(1, dtype=torch.float32, device=xc.device)
        scratch = torch.empty(2, dtype=torch.float64, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.check(lib.dgs_hinge_l2_forward(xc.numel(), _lib.ptr(xc), _lib.ptr(out), _lib.ptr(scratch), _st(xc)),
                       "dgs_hinge_l2_forward")
        ctx.save_for_backward(xc)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        (xc,) = ctx.saved_tensors
        go = grad_out.detach().float().contiguous().reshape(1)
        dx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(lib.dgs_hinge_l2_backward(xc.numel(), _lib.ptr(xc), _lib.ptr(go), _lib.ptr(dx), _st(xc)),
                       "dgs_hinge_l2_backward")
        return dx


def tv_loss(x):
    """utils/loss_utils.py:66-78 on a [B,C,H,W] tensor: l2 of the vertical + l2 of the horizontal neighbour differences."""
    return _TV.apply(x)


def hinge_l2(x):
    """utils/loss_utils.py:95-104: mean of x^2 where x <= 0 and (x - 1)^2 where x >= 1."""
    return _Hinge.apply(x)


def training_loss(blurred, subframes, gt, depths=None, opacity=None, lambda_t_smooth=0.0, lambda_depth_tv=0.0,
                  lambda_hinge=0.0):
    """The reference's loss (train.py:147-163): Ll1 + lambda_t_smooth * L_t_smooth + lambda_depth_tv * L_depth_tv +
    lambda_hinge * L_hinge, with the reference's gating (a term whose weight is 0 is not evaluated).
    depths [F,1,H,W] (query()['depths']), opacity = gaussians._opacity."""
    loss = blur_photometric_loss(blurred, subframes, gt, lambda_t_smooth)
    if lambda_depth_tv > 0.0:
        loss = loss + lambda_depth_tv * tv_loss(depths)
    if lambda_hinge > 0.0:
        loss = loss + lambda_hinge * hinge_l2(opacity)
    return loss

"""Sub-frame camera poses from Bezier control points in se(3), on device, differentiable.

Replaces (for curve_type == "se3") BezierModel.forward (scene/bezier.py:54-83), se3_exp_map
(utils/pytorch3d_functions.py:373-457), CameraMotionModule._c2w_to_minicam (scene/motion.py:258-294)
and MiniCam.__init__ (scene/cameras.py:63-74) of taekkii/deblurgs with one kernel launch each way.
"""
import ctypes as C

import torch

from . import _lib


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _BezierSE3Poses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ctrl_trans, ctrl_rot, nu, proj_t):
        lib = _lib.load()
        if not ctrl_trans.is_cuda:
            raise _lib.DgsError("control points must be CUDA tensors: libdgs_b200 has no CPU path")
        dev = ctrl_trans.device
        ct = ctrl_trans.detach().float().contiguous()
        cr = ctrl_rot.detach().float().contiguous()
        nu_c = nu.detach().float().contiguous()
        pt = proj_t.detach().float().contiguous().to(dev)
        F = nu_c.shape[0]
        order = ct.shape[0] - 1
        view = torch.empty((F, 4, 4), dtype=torch.float32, device=dev)
        proj = torch.empty((F, 4, 4), dtype=torch.float32, device=dev)
        campos = torch.empty((F, 3), dtype=torch.float32, device=dev)
        jac = torch.empty((F, 35, 7), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = lib.dgs_pose_forward(F, order, _lib.ptr(ct), _lib.ptr(cr), _lib.ptr(nu_c), _lib.ptr(pt),
                                      _lib.ptr(view), _lib.ptr(proj), _lib.ptr(campos), _lib.ptr(jac),
                                      _stream_ptr(dev))
        _lib.check(rc, "dgs_pose_forward")
        ctx.save_for_backward(ct, cr, nu_c, jac)
        ctx.mark_non_differentiable(campos)   # the reference gives camera_center no gradient
        return view, proj, campos

    @staticmethod
    def backward(ctx, dview, dproj, _dcampos):
        lib = _lib.load()
        ct, cr, nu_c, jac = ctx.saved_tensors
        dev = ct.device
        F = nu_c.shape[0]
        order = ct.shape[0] - 1
        z = torch.zeros((F, 4, 4), dtype=torch.float32, device=dev)
        dv = dview.float().contiguous() if dview is not None else z
        dp = dproj.float().contiguous() if dproj is not None else z
        dct = torch.empty_like(ct)
        dcr = torch.empty_like(cr)
        dnu = torch.empty_like(nu_c)
        with torch.cuda.device(dev):
            rc = lib.dgs_pose_backward(F, order, _lib.ptr(ct), _lib.ptr(cr), _lib.ptr(nu_c), _lib.ptr(jac),
                                       _lib.ptr(dv), _lib.ptr(dp), _lib.ptr(dct), _lib.ptr(dcr), _lib.ptr(dnu),
                                       _stream_ptr(dev))
        _lib.check(rc, "dgs_pose_backward")
        return dct, dcr, dnu, None


def bezier_se3_poses(ctrl_trans, ctrl_rot, nu, projection_matrix_t):
    """ctrl_trans/ctrl_rot [C+1,3], nu [F] in [0,1], projection_matrix_t [4,4] (= the reference
    camera's `projection_matrix`). Returns world_view_transform [F,4,4], full_proj_transform
    [F,4,4], camera_center [F,3] -- the three MiniCam tensors of every sub-frame."""
    return _BezierSE3Poses.apply(ctrl_trans, ctrl_rot, nu, projection_matrix_t)

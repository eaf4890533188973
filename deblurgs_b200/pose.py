"""Sub-frame camera poses from Bezier control points in se(3), on device, differentiable.

Replaces (for curve_type == "se3") BezierModel.forward (scene/bezier.py:54-83), se3_exp_map
(utils/pytorch3d_functions.py:373-457), CameraMotionModule._c2w_to_minicam (scene/motion.py:258-294)
and MiniCam.__init__ (scene/cameras.py:63-74) of taekkii/deblurgs with one kernel launch each way.
"""
import ctypes as C
import math

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


# ---------------------------------------------------------------------------------------------------
# Host-side pose algebra around the kernel (plain torch on tiny tensors, like the reference's Python):
# initialisation of the se(3) control points from camera poses, and the quaternion curve type.
# ---------------------------------------------------------------------------------------------------
def _skew(v):
    """[N,3] -> [N,3,3] cross-product matrices."""
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2)


def _acos_with_linear_tails(x, bound):
    """acos(x) for |x| < bound, continued by its tangent at +-bound outside (finite value and slope at
    |x| >= 1) -- the reference's `acos_linear_extrapolation` (utils/pytorch3d_functions.py:26-81)."""
    def tail(v, b):
        return math.acos(b) - (v - b) / math.sqrt(1.0 - b * b)
    inner = torch.acos(x.clamp(-bound, bound))
    return torch.where(x >= bound, tail(x, bound), torch.where(x <= -bound, tail(x, -bound), inner))


def so3_log_map(R, eps=1e-4, cos_bound=1e-4):
    """Rotation matrices [N,3,3] -> axis-angle vectors [N,3]; semantics of the reference's so3_log_map
    (utils/pytorch3d_functions.py:250-303): angle from the clamped/extrapolated acos of (tr R - 1)/2,
    factor phi / (2 sin phi) with the series 1/2 + phi^2/12 where |sin phi| <= eps/2."""
    if R.dim() != 3 or R.shape[1:] != (3, 3):
        raise ValueError("Input has to be a batch of 3x3 Tensors.")
    trace = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    if bool(((trace < -1.0 - eps) | (trace > 3.0 + eps)).any()):
        raise ValueError("A matrix has trace outside valid range [-1-eps,3+eps].")
    phi = _acos_with_linear_tails((trace - 1.0) * 0.5, 1.0 - cos_bound)
    sin_phi = torch.sin(phi)
    small = sin_phi.abs() <= 0.5 * eps
    factor = torch.where(small, 0.5 + phi * phi * (1.0 / 12), phi / (2.0 * torch.where(small, torch.ones_like(phi), sin_phi)))
    A = factor[:, None, None] * (R - R.transpose(1, 2))
    if float((A + A.transpose(1, 2)).abs().max()) > 1e-5:
        raise ValueError("One of input matrices is not skew-symmetric.")
    return torch.stack((A[:, 2, 1], A[:, 0, 2], A[:, 1, 0]), dim=1)


def se3_log_map(transform, eps=1e-4, cos_bound=1e-4):
    """[N,4,4] rigid transforms in the reference's transposed layout (R^T in [:3,:3], translation in row 3)
    -> [N,6] = [log_translation | log_rotation]; inverse of the pose kernel's exponential map
    (reference: utils/pytorch3d_functions.py:462-541, used by CameraMotionModule._set_initial_parameters,
    scene/motion.py:196-204)."""
    if transform.dim() != 3 or transform.shape[1:] != (4, 4):
        raise ValueError("Input tensor shape has to be (N, 4, 4).")
    if not torch.allclose(transform[:, :3, 3], torch.zeros_like(transform[:, :3, 3])):
        raise ValueError("All elements of `transform[:, :3, 3]` should be 0.")
    log_rot = so3_log_map(transform[:, :3, :3].transpose(1, 2), eps=eps, cos_bound=cos_bound)
    theta = (log_rot * log_rot).sum(-1).clamp(min=eps).sqrt()
    K = _skew(log_rot)
    V = torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None] \
        + K * ((1 - torch.cos(theta)) / theta ** 2)[:, None, None] \
        + torch.bmm(K, K) * ((theta - torch.sin(theta)) / theta ** 3)[:, None, None]
    log_trans = torch.linalg.solve(V, transform[:, 3, :3][:, :, None])[:, :, 0]
    return torch.cat((log_trans, log_rot), dim=1)


def unitquat_to_rotmat(q):
    """Unit quaternions [N,4] in XYZW order (the convention of `roma`, which the reference uses for
    curve_type == "quarternion_cartesian", scene/motion.py:191-194, 242-246) -> [N,3,3]."""
    x, y, z, w = q.unbind(-1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


def rotmat_to_unitquat(R):
    """[N,3,3] -> unit quaternions [N,4] XYZW, as `roma.rotmat_to_unitquat` (scene/motion.py:192, test.py:66) returns
    them: the row of the largest of (m00, m11, m22, trace) is normalised and its sign is NOT canonicalised, so the
    leading component of the chosen row is positive and w may be negative near a half turn (q and -q are the same
    rotation; the sign only decides which representative a curve's control points start from)."""
    m = R
    t = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    cand = torch.stack([
        torch.stack([1 + m[:, 0, 0] - m[:, 1, 1] - m[:, 2, 2], m[:, 0, 1] + m[:, 1, 0], m[:, 0, 2] + m[:, 2, 0],
                     m[:, 2, 1] - m[:, 1, 2]], -1),
        torch.stack([m[:, 0, 1] + m[:, 1, 0], 1 - m[:, 0, 0] + m[:, 1, 1] - m[:, 2, 2], m[:, 1, 2] + m[:, 2, 1],
                     m[:, 0, 2] - m[:, 2, 0]], -1),
        torch.stack([m[:, 0, 2] + m[:, 2, 0], m[:, 1, 2] + m[:, 2, 1], 1 - m[:, 0, 0] - m[:, 1, 1] + m[:, 2, 2],
                     m[:, 1, 0] - m[:, 0, 1]], -1),
        torch.stack([m[:, 2, 1] - m[:, 1, 2], m[:, 0, 2] - m[:, 2, 0], m[:, 1, 0] - m[:, 0, 1], 1 + t], -1)], 1)
    best = torch.stack([m[:, 0, 0], m[:, 1, 1], m[:, 2, 2], t], -1).argmax(-1)
    q = cand[torch.arange(R.shape[0], device=R.device), best]
    return q / q.norm(dim=-1, keepdim=True)


def c2w_to_minicam_tensors(rots, transes, projection_matrix_t):
    """Batched form of CameraMotionModule._c2w_to_minicam + MiniCam.__init__ (scene/motion.py:258-294,
    scene/cameras.py:63-74) for c2w rotations [F,3,3] / translations [F,3] that already live in torch
    (the quaternion curve type): world_view_transform [F,4,4], full_proj_transform [F,4,4] and
    camera_center [F,3] (= the c2w translation; the reference gets it from a matrix inverse, without
    gradient)."""
    F = rots.shape[0]
    view = torch.zeros((F, 4, 4), dtype=torch.float32, device=rots.device)
    view[:, :3, :3] = rots.float()
    view[:, 3, :3] = -torch.einsum("fi,fij->fj", transes.float(), rots.float())
    view[:, 3, 3] = 1.0
    proj = view @ projection_matrix_t.to(view.device).float()[None]
    return view, proj, transes.detach().float()

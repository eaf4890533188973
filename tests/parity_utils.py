"""Shared helpers for the GPU parity tests: run libdgs_b200 (through the package's C-ABI binding) and
the reference's own CUDA rasterizer (oracle/_ref, test infrastructure) on identical device buffers
and compare every intermediate the reference defines."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from deblurgs_b200 import _lib, synthetic  # noqa: E402
from deblurgs_b200 import rasterizer as rz  # noqa: E402
from deblurgs_b200.pose import bezier_se3_poses  # noqa: E402


def make_inputs(name, device="cuda", sh_degree=3, P=None, F=None):
    Pn, W, H, Fn, Cn = synthetic.CONFIGS[name]
    P = Pn if P is None else P
    F = Fn if F is None else F
    cam = synthetic.make_camera(W, H)
    scene = synthetic.make_scene(P, cam, sh_degree=sh_degree).to(device)
    traj = synthetic.make_trajectory(F, Cn).to(device)
    bg = synthetic.make_background().to(device)
    view, proj, campos = bezier_se3_poses(traj.ctrl_trans, traj.ctrl_rot, traj.nu,
                                          cam.projection_matrix_t().to(device))
    return cam, scene, traj, bg, view.contiguous(), proj.contiguous(), campos.contiguous()


def ours_forward(cam, scene, bg, view, proj, campos, sh_degree=None, use_sigmoid=False, colors_precomp=None,
                 cov3D_precomp=None, scale_modifier=1.0, want_blur=True, exact=False):
    sh_degree = scene.sh_degree if sh_degree is None else sh_degree
    shs = None if colors_precomp is not None else scene.shs
    scales = None if cov3D_precomp is not None else scene.scales
    rots = None if cov3D_precomp is not None else scene.rotations
    out = rz._forward_batched(scene.means3D, shs, colors_precomp, scene.opacities, scales, rots, cov3D_precomp,
                              view, proj, campos, bg, cam.height, cam.width, cam.tanfovx, cam.tanfovy,
                              scale_modifier, 0.2, 100.0, sh_degree, False, use_sigmoid, want_blur,
                              float(view.shape[0]), exact=exact)
    color, depth, radii, blur, D, geom, binning, img = out
    return dict(color=color, depth=depth, radii=radii, blur=blur, num_rendered=D, geom=geom, binning=binning,
                img=img)


def ours_decode(fw, P, F, W, H):
    lib = _lib.load()
    dev = fw["color"].device
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    N = P * F
    D = fw["num_rendered"]
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    d = dict(depths=torch.zeros((F, P), **f32), means2D=torch.zeros((F, P, 2), **f32),
             conic_opacity=torch.zeros((F, P, 4), **f32), rgb=torch.zeros((F, P, 3), **f32),
             clamped=torch.zeros((F, P, 3), **f32), tiles_touched=torch.zeros((F, P), **i32),
             point_offsets=torch.zeros((F, P), **i32),
             keys=torch.zeros(D, dtype=torch.int64, device=dev), point_list=torch.zeros(D, **i32),
             ranges=torch.zeros((F, tiles, 2), **i32), final_T=torch.zeros((F, H, W), **f32),
             n_contrib=torch.zeros((F, H, W), **i32))
    p = _lib.ptr
    if N:
        _lib.check(lib.dgs_debug_geometry(p(fw["geom"]), P, F, p(d["depths"]), p(d["means2D"]),
                                          p(d["conic_opacity"]), p(d["rgb"]), p(d["clamped"]),
                                          p(d["tiles_touched"]), p(d["point_offsets"]), st), "debug_geometry")
    if N:
        _lib.check(lib.dgs_debug_binning(p(fw["geom"]), p(fw["binning"]), p(fw["img"]), P, F, W, H, D, p(d["keys"]),
                                         p(d["point_list"]), p(d["ranges"]), st), "debug_binning")
    _lib.check(lib.dgs_debug_image(p(fw["img"]), F, W, H, p(d["final_T"]), p(d["n_contrib"]), st), "debug_image")
    tb, sb = C.c_int(0), C.c_int(0)
    lib.dgs_key_bits(W, H, F, C.byref(tb), C.byref(sb))
    d["tile_bits"], d["subframe_bits"] = tb.value, sb.value
    torch.cuda.synchronize(dev)
    return d


def ours_backward(cam, scene, bg, view, proj, campos, fw, dL_dpix, dL_ddepth, sh_degree=None, use_sigmoid=False,
                  colors_precomp=None, cov3D_precomp=None, scale_modifier=1.0):
    sh_degree = scene.sh_degree if sh_degree is None else sh_degree
    P, F = scene.means3D.shape[0], view.shape[0]
    shs = None if colors_precomp is not None else scene.shs
    scales = None if cov3D_precomp is not None else scene.scales
    rots = None if cov3D_precomp is not None else scene.rotations
    M = shs.shape[1] if shs is not None else 0
    names = ["dL_dmeans2D", "dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dcolors",
             "dL_dcov3D", "dL_dviewmatrix", "dL_dprojmatrix"]
    out = rz._backward_batched(P, F, M, fw["num_rendered"], scene.means3D, shs, colors_precomp, scene.opacities,
                               scales, rots, cov3D_precomp, view, proj, campos, bg, cam.height, cam.width,
                               cam.tanfovx, cam.tanfovy, scale_modifier, 0.2, 100.0, sh_degree, use_sigmoid,
                               fw["radii"], fw["geom"], fw["binning"], fw["img"], dL_dpix, dL_ddepth, True,
                               want_stats=True)
    return dict(zip(names + ["densify_stats"], out))


def rel_err(a, b, floor_frac=1e-3):
    """max |a-b| / max(|b|, floor) with floor = floor_frac * max|b| (relative error that does not blow
    up on entries that are ~0 by cancellation)."""
    a, b = a.double(), b.double()
    scale = b.abs().max().item()
    if scale == 0:
        return (a - b).abs().max().item()
    denom = torch.clamp(b.abs(), min=floor_frac * scale)
    return ((a - b).abs() / denom).max().item()

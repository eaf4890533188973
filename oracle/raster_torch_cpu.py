"""TEST INFRASTRUCTURE ONLY -- never imported by the product (deblurgs_b200/).

PyTorch-CPU evaluation of the blurry-view path on the host cores (BASELINE.json north_star: "a PyTorch-CPU
evaluation of the same projection+composite path on the box's host cores"):
pose chain -> projection (cov3D, EWA cov2D, SH colour, radii) -> tile keys -> sort -> front-to-back composite per
tile -> mean over the sub-frames -> L1 loss -> torch autograd backward down to the Gaussian parameters and the
Bezier control points.  Vectorised torch ops on `torch.get_num_threads()` threads; it is bench.py's `cpu_baseline`
leg and is checked against the numpy oracle (tests/test_oracle_torch_cpu.py).  It restates, for taekkii/deblurgs:

  projection        submodules/diff-gaussian-rasterization/cuda_rasterizer/forward.cu:20-163, 194-268
  tile keys / sort  cuda_rasterizer/rasterizer_impl.cu:88-108, 122-137, 306-314
  composite         cuda_rasterizer/forward.cu:341-391
  pose chain        scene/bezier.py:54-83, utils/pytorch3d_functions.py (through oracle/pose_torch.py)
  blur mean / loss  scene/motion.py:148, utils/loss_utils.py:17-18

Arithmetic is fp32 without the FMA-contraction emulation of oracle/raster_np.py, so radii can differ from the
CUDA code in the last-ulp cases; gradients are plain autograd (the true derivative, without the reference
backward's view/projection-matrix conventions), which is what a CPU evaluation of this path computes.
"""
import math

import torch

from . import pose_torch as pt

TILE = 16
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def _sh_color(shs, deg, dirs):
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * shs[:, 0]
    if deg > 0:
        res = res - C1 * y * shs[:, 1] + C1 * z * shs[:, 2] - C1 * x * shs[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = res + C2[0] * xy * shs[:, 4] + C2[1] * yz * shs[:, 5] + C2[2] * (2 * zz - xx - yy) * shs[:, 6] \
            + C2[3] * xz * shs[:, 7] + C2[4] * (xx - yy) * shs[:, 8]
    if deg > 2:
        res = res + C3[0] * y * (3 * xx - yy) * shs[:, 9] + C3[1] * xy * z * shs[:, 10] \
            + C3[2] * y * (4 * zz - xx - yy) * shs[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 12] \
            + C3[4] * x * (4 * zz - xx - yy) * shs[:, 13] + C3[5] * z * (xx - yy) * shs[:, 14] \
            + C3[6] * x * (xx - 3 * yy) * shs[:, 15]
    return torch.clamp_min(res + 0.5, 0.0)


def project(means, scales, rots, opac, shs, deg, view, proj, campos, W, H, tanx, tany):
    """view / proj: [4,4] torch (row-vector convention: p_view = [p,1] @ view). Differentiable."""
    P = means.shape[0]
    hom = torch.cat([means, torch.ones_like(means[:, :1])], 1)
    pv = hom @ view
    ph = hom @ proj
    pw = 1.0 / (ph[:, 3] + 1e-7)
    ndc = ph[:, :2] * pw[:, None]
    # cov3D = (S R)^T (S R), un-normalised quaternion (r,x,y,z)
    r, x, y, z = rots.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).view(P, 3, 3)
    L = R * scales[:, None, :]
    Sigma = L @ L.transpose(1, 2)
    # EWA projection
    fx, fy = W / (2 * tanx), H / (2 * tany)
    tz = pv[:, 2]
    limx, limy = 1.3 * tanx, 1.3 * tany
    tx = torch.clamp(pv[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(pv[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], 1).view(P, 2, 3)
    Wm = view[:3, :3].t()                       # world -> view rotation
    T = J @ Wm[None]
    cov = T @ Sigma @ T.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    det_inv = 1.0 / det
    conic = torch.stack([c * det_inv, -b * det_inv, a * det_inv], 1)
    mid = 0.5 * (a + c)
    disc = torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(mid + disc, mid - disc))).detach()
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    tiles_x, tiles_y = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    with torch.no_grad():
        x0 = ((px - radius) / TILE).to(torch.int64).clamp(0, tiles_x)
        y0 = ((py - radius) / TILE).to(torch.int64).clamp(0, tiles_y)
        x1 = ((px + radius + TILE - 1) / TILE).to(torch.int64).clamp(0, tiles_x)
        y1 = ((py + radius + TILE - 1) / TILE).to(torch.int64).clamp(0, tiles_y)
        vis = (tz > 0.2) & (det != 0) & ((x1 - x0) * (y1 - y0) > 0)
    dirs = means - campos[None]
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    rgb = _sh_color(shs, deg, dirs)
    return dict(px=px, py=py, depth=tz, conic=conic, opac=opac[:, 0], rgb=rgb, radius=radius, vis=vis,
                rect=(x0, y0, x1, y1), tiles=(tiles_x, tiles_y))


@torch.no_grad()
def bin_tiles(pre):
    """(point_list [D], ranges [tiles+1]): duplicates sorted by (tile, depth bits), stable in the Gaussian index."""
    x0, y0, x1, y1 = pre["rect"]
    tiles_x, tiles_y = pre["tiles"]
    idx = torch.nonzero(pre["vis"]).squeeze(1)
    w, h = (x1 - x0)[idx], (y1 - y0)[idx]
    cnt = w * h
    owner = torch.repeat_interleave(torch.arange(idx.numel()), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    j = torch.arange(owner.numel()) - start[owner]
    ty = y0[idx][owner] + j // w[owner]
    tx = x0[idx][owner] + j % w[owner]
    tile = ty * tiles_x + tx
    gid = idx[owner]
    depth_bits = pre["depth"].detach()[gid].contiguous().view(torch.int32).to(torch.int64)
    key = (tile << 32) | depth_bits
    order = torch.sort(key, stable=True).indices
    tile_sorted = tile[order]
    bounds = torch.searchsorted(tile_sorted, torch.arange(tiles_x * tiles_y + 1))
    return gid[order], bounds


def composite(pre, point_list, bounds, bg, W, H, chunk_pairs=6_000_000):
    """Front-to-back alpha compositing of every tile's list (forward.cu:341-391), tiles processed in padded
    chunks of [tiles, K, 256 pixels]."""
    tiles_x, tiles_y = pre["tiles"]
    n_tiles = tiles_x * tiles_y
    lens = bounds[1:] - bounds[:-1]
    lx = torch.arange(TILE).repeat(TILE).float()
    ly = torch.arange(TILE).repeat_interleave(TILE).float()
    t0 = 0
    outs = []
    while t0 < n_tiles:
        t1 = t0 + 1
        kmax = int(lens[t0])
        while t1 < n_tiles and max(kmax, int(lens[t1])) * (t1 + 1 - t0) * 256 <= chunk_pairs:
            kmax = max(kmax, int(lens[t1]))
            t1 += 1
        nt = t1 - t0
        K = max(kmax, 1)
        pos = bounds[t0:t1, None] + torch.arange(K)[None]
        valid = torch.arange(K)[None] < lens[t0:t1, None]
        if point_list.numel():
            gid = point_list[torch.where(valid, pos, torch.zeros_like(pos)).clamp(max=point_list.numel() - 1)]
        else:
            gid = torch.zeros_like(pos)
        tiles = torch.arange(t0, t1)
        pixx = ((tiles % tiles_x) * TILE)[:, None].float() + lx[None]          # [nt,256]
        pixy = ((tiles // tiles_x) * TILE)[:, None].float() + ly[None]
        dx = pre["px"][gid][:, :, None] - pixx[:, None, :]                       # [nt,K,256]
        dy = pre["py"][gid][:, :, None] - pixy[:, None, :]
        con = pre["conic"][gid]
        power = -0.5 * (con[:, :, 0:1] * dx * dx + con[:, :, 2:3] * dy * dy) - con[:, :, 1:2] * dx * dy
        alpha = torch.clamp_max(pre["opac"][gid][:, :, None] * torch.exp(power), 0.99)
        ok = valid[:, :, None] & (power <= 0) & (alpha >= 1.0 / 255.0)
        inside = ((pixx < W) & (pixy < H))[:, None, :]
        a = torch.where(ok & inside, alpha, torch.zeros_like(alpha))
        with torch.no_grad():
            T_incl = torch.cumprod(1 - a, dim=1)
            stopped = torch.cummax(((T_incl < 1e-4) & ok).to(torch.uint8), dim=1).values.bool()
        a = torch.where(stopped, torch.zeros_like(a), a)
        T_incl = torch.cumprod(1 - a, dim=1)
        T_excl = T_incl / (1 - a)
        wgt = a * T_excl                                                          # [nt,K,256]
        col = torch.einsum("tkp,tkc->tcp", wgt, pre["rgb"][gid]) + T_incl[:, -1, :][:, None, :] * bg[None, :, None]
        outs.append((t0, t1, col))
        t0 = t1
    full = torch.cat([c for (_, _, c) in outs], 0)                                # [tiles,3,256]
    full = full.view(tiles_y, tiles_x, 3, TILE, TILE).permute(2, 0, 3, 1, 4).reshape(3, tiles_y * TILE, tiles_x * TILE)
    return full[:, :H, :W]


def render_view(means, scales, rots, opac, shs, deg, view, proj, campos, bg, W, H, tanx, tany):
    pre = project(means, scales, rots, opac, shs, deg, view, proj, campos, W, H, tanx, tany)
    point_list, bounds = bin_tiles(pre)
    return composite(pre, point_list, bounds, bg, W, H), pre, point_list, bounds


def blurry_view_step(params, ctrl_trans, ctrl_rot, nu, proj_t, bg, gt, W, H, tanx, tany, deg=3):
    """One bench step on the CPU: poses from the control points, F renders, mean, L1, autograd backward.
    params = (means, scales, rots, opac, shs) leaf tensors (activated values, as the rasterizer takes them).
    Returns (loss, blurred)."""
    leaves = list(params) + [ctrl_trans, ctrl_rot]
    for t in leaves:
        t.grad = None
        t.requires_grad_(True)
    poses = pt.trajectory(ctrl_trans, ctrl_rot, nu, proj_t)
    imgs = [render_view(*params, deg, v.float(), p.float(), c.float(), bg, W, H, tanx, tany)[0] for (v, p, c) in poses]
    blurred = torch.stack(imgs).mean(0)
    loss = (blurred - gt).abs().mean()
    loss.backward()
    return loss.detach(), blurred.detach()

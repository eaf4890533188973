"""TEST INFRASTRUCTURE ONLY -- never imported by the product (deblurgs_b200/).

CPU (numpy) restatement of the reference's differentiable Gaussian rasterizer for ONE view, used as
the checker in tests/, in __graft_entry__.smoke() and as bench.py's `cpu_baseline` ("port").
Each function cites the reference code it follows (taekkii/deblurgs,
submodules/diff-gaussian-rasterization/cuda_rasterizer/):

  preprocess            forward.cu:166-268, :85-163 (cov3D / EWA cov2D), :20-82 (SH), auxiliary.h:41-56
  bin_tiles             rasterizer_impl.cu:70-111 (keys), :306-314 (stable sort), :116-138 (ranges)
  render_forward        forward.cu:273-392
  render_backward       backward.cu:463-640
  preprocess_backward   backward.cu:145-295 (cov2D), :367-460 (projection + quirks), :20-140 (SH),
                        :299-362 (cov3D -> scale / rotation)

Forward arithmetic is float32 with the fused multiply-adds nvcc emits for these expression trees
(`a*b + c*d` -> fma(a, b, c*d); `x + c*d` -> fma(c, d, x)), emulated exactly through float64, so
that radii / tile rectangles / depth keys are comparable bit for bit; exp() is numpy's float32 exp,
which is NOT bit-identical to CUDA's expf, so alpha-threshold decisions can differ for isolated
pixels.  Backward is evaluated in float64 on the float32 forward state: it is the exact sum that
the reference's float atomics approximate in arbitrary order.

Parity status: pinned against outputs of the reference's own CUDA extension (tests/golden/raster_*.npz,
generated on a B200 by tests/golden/make_raster_golden.py through oracle/_ref).
"""
import numpy as np

f32 = np.float32
f64 = np.float64

SH_C0 = f32(0.28209479177387814)
SH_C1 = f32(0.4886025119029199)
SH_C2 = [f32(1.0925484305920792), f32(-1.0925484305920792), f32(0.31539156525252005),
         f32(-1.0925484305920792), f32(0.5462742152960396)]
SH_C3 = [f32(-0.5900435899266435), f32(2.890611442640554), f32(-0.4570457994644658),
         f32(0.3731763325901154), f32(-0.4570457994644658), f32(1.445305721320277),
         f32(-0.5900435899266435)]
TILE = 16


def fma(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64."""
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def dot3(a0, b0, a1, b1, a2, b2):
    """a0*b0 + a1*b1 + a2*b2 as nvcc contracts it: fma(a2, b2, fma(a0, b0, a1*b1))."""
    return fma(a2, b2, fma(a0, b0, (a1 * b1).astype(f32)))


def xform(p, m, rows, with_w=True):
    """matrix[4*col+row] applied to points [P,3]; per row: fma(z,m8, fma(x,m0, y*m4)) + m12."""
    out = []
    for r in rows:
        v = dot3(p[:, 0], m[r], p[:, 1], m[4 + r], p[:, 2], m[8 + r])
        out.append((v + m[12 + r]).astype(f32) if with_w else v)
    return out


def mat3_mul(A, B):
    """Column-major 3x3 product in the reference algebra library's association:
    R[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]; entries are arrays [P]."""
    return [[dot3(A[0][r], B[c][0], A[1][r], B[c][1], A[2][r], B[c][2]) for r in range(3)] for c in range(3)]


def mat3_t(A):
    return [[A[r][c] for r in range(3)] for c in range(3)]


def cov3d(scales, rots, mod):
    P = scales.shape[0]
    z = np.zeros(P, f32)
    mod = f32(mod)
    s = [(mod * scales[:, i]).astype(f32) for i in range(3)]
    r, x, y, zz = rots[:, 0], rots[:, 1], rots[:, 2], rots[:, 3]
    one, two = f32(1.0), f32(2.0)

    def m1(a, b, c, d):   # 1 - 2*(a*b + c*d)
        return (one - two * fma(a, b, (c * d).astype(f32))).astype(f32)

    def p2(a, b, c, d, sign):  # 2*(a*b +- c*d)
        return (two * fma(a, b, (sign * (c * d)).astype(f32))).astype(f32)

    # Which product of each `a*b +- c*d` gets fused depends on how nvcc shares the products between
    # the nine entries (read off the reference's sm_100 SASS): y*y + z*z is a plain add of two rounded
    # products; x*z +- r*y fuses r*y; the others fuse their first product.
    yy_zz = ((y * y).astype(f32) + (zz * zz).astype(f32)).astype(f32)
    xz = (x * zz).astype(f32)
    R = [[(one - two * yy_zz).astype(f32), p2(x, y, r, zz, f32(-1)), (two * fma(r, y, xz)).astype(f32)],
         [p2(x, y, r, zz, f32(1)), m1(x, x, zz, zz), p2(y, zz, r, x, f32(-1))],
         [(two * fma(-r, y, xz)).astype(f32), p2(y, zz, r, x, f32(1)), m1(x, x, y, y)]]
    S = [[s[0], z, z], [z, s[1], z], [z, z, s[2]]]
    M = mat3_mul(S, R)
    Sigma = mat3_mul(mat3_t(M), M)
    return np.stack([Sigma[0][0], Sigma[0][1], Sigma[0][2], Sigma[1][1], Sigma[1][2], Sigma[2][2]], axis=1), M, R, s


def ewa(means, cov, view, fx, fy, tanx, tany):
    tx, ty, tz = xform(means, view, (0, 1, 2))
    limx, limy = (f32(1.3) * f32(tanx)).astype(f32), (f32(1.3) * f32(tany)).astype(f32)
    with np.errstate(all="ignore"):
        txtz, tytz = (tx / tz).astype(f32), (ty / tz).astype(f32)
        cx = (np.minimum(limx, np.maximum(-limx, txtz)) * tz).astype(f32)
        cy = (np.minimum(limy, np.maximum(-limy, tytz)) * tz).astype(f32)
        z = np.zeros_like(tz)
        fx, fy = f32(fx), f32(fy)
        tz2 = (tz * tz).astype(f32)
        J = [[(fx / tz).astype(f32), z, (-(fx * cx).astype(f32) / tz2).astype(f32)],
             [z, (fy / tz).astype(f32), (-(fy * cy).astype(f32) / tz2).astype(f32)],
             [z, z, z]]
        b = lambda v: np.full_like(tz, v)
        W = [[b(view[0]), b(view[4]), b(view[8])], [b(view[1]), b(view[5]), b(view[9])],
             [b(view[2]), b(view[6]), b(view[10])]]
        T = mat3_mul(W, J)
        Vrk = [[cov[:, 0], cov[:, 1], cov[:, 2]], [cov[:, 1], cov[:, 3], cov[:, 4]], [cov[:, 2], cov[:, 4], cov[:, 5]]]
        c2 = mat3_mul(mat3_mul(mat3_t(T), mat3_t(Vrk)), T)
    a = (c2[0][0] + f32(0.3)).astype(f32)
    bb = c2[0][1]
    c = (c2[1][1] + f32(0.3)).astype(f32)
    return dict(t=(cx, cy, tz), txtz=txtz, tytz=tytz, T=T, W=W, a=a, b=bb, c=c, limx=limx, limy=limy)


def sh_basis(d, deg):
    """Basis values b_k(dir) with colour = sum_k b_k * sh[k] (+0.5 / sigmoid), float32, same grouping
    as the reference expressions. d: [P,3] unit directions."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    B = [np.full_like(x, SH_C0)]
    if deg > 0:
        B += [(-SH_C1 * y).astype(f32), (SH_C1 * z).astype(f32), (-SH_C1 * x).astype(f32)]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        B += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (f32(2) * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if deg > 2:
        B += [SH_C3[0] * y * (f32(3) * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (f32(4) * zz - xx - yy),
              SH_C3[3] * z * (f32(2) * zz - f32(3) * xx - f32(3) * yy), SH_C3[4] * x * (f32(4) * zz - xx - yy),
              SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - f32(3) * yy)]
    return [np.asarray(b, f32) for b in B]


def preprocess(means, scales, rots, opac, shs, deg, view, proj, campos, W, H, tanx, tany, mod=1.0,
               use_sigmoid=False, colors_precomp=None, cov3D_precomp=None):
    means, view, proj = np.asarray(means, f32), np.asarray(view, f32).reshape(16), np.asarray(proj, f32).reshape(16)
    P = means.shape[0]
    fx, fy = f32(W) / (f32(2.0) * f32(tanx)), f32(H) / (f32(2.0) * f32(tany))
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    hx, hy, hz, hw = xform(means, proj, (0, 1, 2, 3))
    with np.errstate(all="ignore"):
        p_w = (f32(1.0) / (hw + f32(0.0000001)).astype(f32)).astype(f32)
        projx, projy = (hx * p_w).astype(f32), (hy * p_w).astype(f32)
        vz = xform(means, view, (2,))[0]
        alive = vz > f32(0.2)
        if cov3D_precomp is None:
            cov, _, _, _ = cov3d(np.asarray(scales, f32), np.asarray(rots, f32), mod)
        else:
            cov = np.asarray(cov3D_precomp, f32)
        e = ewa(means, cov, view, fx, fy, tanx, tany)
        a, b, c = e["a"], e["b"], e["c"]
        det = fma(a, c, -(b * b).astype(f32))
        alive &= det != 0
        det_inv = (f32(1.0) / det).astype(f32)
        conic = np.stack([(c * det_inv).astype(f32), (-b * det_inv).astype(f32), (a * det_inv).astype(f32)], axis=1)
        mid = (f32(0.5) * (a + c).astype(f32)).astype(f32)
        disc = np.sqrt(np.maximum(f32(0.1), fma(mid, mid, -det))).astype(f32)
        lam = np.maximum((mid + disc).astype(f32), (mid - disc).astype(f32))
        radius = np.ceil((f32(3.0) * np.sqrt(lam).astype(f32)).astype(f32)).astype(f32)
        px = (((projx.astype(f64) + 1.0) * W - 1.0) * 0.5).astype(f32)
        py = (((projy.astype(f64) + 1.0) * H - 1.0) * 0.5).astype(f32)
        radius = np.where(alive & np.isfinite(radius), radius, 0).astype(f32)
        ri = radius.astype(np.int64)

        def trunc_div(v):
            q = (v / f32(TILE)).astype(f32)
            q = np.where(np.isfinite(q), q, 0)
            return np.trunc(q).astype(np.int64)
        rf = ri.astype(f32)
        xmin = np.clip(trunc_div((px - rf).astype(f32)), 0, gx)
        ymin = np.clip(trunc_div((py - rf).astype(f32)), 0, gy)
        xmax = np.clip(trunc_div((((px + rf).astype(f32) + f32(TILE)).astype(f32) - f32(1)).astype(f32)), 0, gx)
        ymax = np.clip(trunc_div((((py + rf).astype(f32) + f32(TILE)).astype(f32) - f32(1)).astype(f32)), 0, gy)
    tiles = (xmax - xmin) * (ymax - ymin)
    alive &= tiles > 0
    tiles = np.where(alive, tiles, 0)
    radii = np.where(alive, ri, 0).astype(np.int32)

    if colors_precomp is None:
        d = (means - np.asarray(campos, f32)[None]).astype(f32)
        ln = np.sqrt(dot3(d[:, 0], d[:, 0], d[:, 1], d[:, 1], d[:, 2], d[:, 2])).astype(f32)
        dirs = (d / ln[:, None]).astype(f32)
        B = sh_basis(dirs, deg)
        res = np.zeros((P, 3), f32)
        for k, bk in enumerate(B):
            res = (res + bk[:, None] * np.asarray(shs, f32)[:, k, :]).astype(f32)
        if use_sigmoid:
            pre = res
            rgb = (1.0 / (1.0 + np.exp(-res.astype(f64)))).astype(f32)
            clamped = np.zeros((P, 3), f32)
        else:
            res = (res + f32(0.5)).astype(f32)
            clamped = (res >= 0).astype(f32)
            pre = res
            rgb = np.maximum(res, 0).astype(f32)
    else:
        rgb = np.asarray(colors_precomp, f32)
        clamped = np.ones((P, 3), f32)
        pre = rgb
        dirs = None
    return dict(radii=radii, tiles_touched=tiles.astype(np.int64), depths=vz, means2D=np.stack([px, py], 1),
                cov3D=cov, conic=conic, opacity=np.asarray(opac, f32).reshape(P), rgb=rgb, clamped=clamped,
                pre_act=pre, rect=(xmin, ymin, xmax, ymax), alive=alive, ewa=e, grid=(gx, gy),
                m_hom=(hx, hy, hz, hw), p_w=p_w, focal=(fx, fy))


def bin_tiles(pre, W, H):
    """keys = (tile << 32) | depth bits, stable sort, per-tile [start, end)."""
    gx, gy = pre["grid"]
    xmin, ymin, xmax, ymax = pre["rect"]
    idx = np.nonzero(pre["radii"] > 0)[0]
    keys, vals = [], []
    depth_bits = pre["depths"].view(np.uint32).astype(np.uint64)
    cnt = pre["tiles_touched"][idx]
    rep = np.repeat(idx, cnt)
    if rep.size:
        start = np.cumsum(cnt) - cnt
        j = np.arange(rep.size) - np.repeat(start, cnt)
        w = (xmax - xmin)[rep]
        ty = ymin[rep] + j // w
        tx = xmin[rep] + j % w
        keys = ((ty * gx + tx).astype(np.uint64) << np.uint64(32)) | depth_bits[rep]
        vals = rep.astype(np.uint32)
        order = np.argsort(keys, kind="stable")
        keys, vals = keys[order], vals[order]
    else:
        keys, vals = np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    ranges = np.zeros((gx * gy, 2), np.int64)
    if keys.size:
        t = (keys >> np.uint64(32)).astype(np.int64)
        first = np.nonzero(np.r_[True, t[1:] != t[:-1]])[0]
        last = np.r_[first[1:], t.size]
        ranges[t[first], 0] = first
        ranges[t[first], 1] = last
    return keys, vals, ranges


def _tile_pixels(W, H, gx, gy):
    """pixel coordinates per tile: [tiles,256] x / y and validity."""
    ty, tx = np.divmod(np.arange(gx * gy), gx)
    ly, lx = np.divmod(np.arange(TILE * TILE), TILE)
    X = tx[:, None] * TILE + lx[None]
    Y = ty[:, None] * TILE + ly[None]
    return X, Y, (X < W) & (Y < H)


def _power(cx, cy, cz, dx, dy):
    # -0.5f*(cx*dx*dx + cz*dy*dy) - cy*dx*dy with nvcc's contraction
    s = fma((cx * dx).astype(f32), dx, ((cz * dy).astype(f32) * dy).astype(f32))
    return fma(f32(-0.5), s, -((cy * dx).astype(f32) * dy).astype(f32))


def render_forward(pre, vals, ranges, bg, W, H, z_far=100.0):
    gx, gy = pre["grid"]
    X, Y, valid = _tile_pixels(W, H, gx, gy)
    nt = gx * gy
    T = np.ones((nt, 256), f32)
    C = np.zeros((nt, 256, 3), f32)
    D = np.zeros((nt, 256), f32)
    done = ~valid
    ncontrib = np.zeros((nt, 256), np.int64)
    lens = ranges[:, 1] - ranges[:, 0]
    Xf, Yf = X.astype(f32), Y.astype(f32)
    m2, con, op, rgb, dep = pre["means2D"], pre["conic"], pre["opacity"], pre["rgb"], pre["depths"]
    for k in range(int(lens.max()) if nt else 0):
        act = np.nonzero((lens > k) & (~done).any(axis=1))[0]
        if act.size == 0:
            break
        g = vals[ranges[act, 0] + k].astype(np.int64)
        dx = (m2[g, 0][:, None] - Xf[act]).astype(f32)
        dy = (m2[g, 1][:, None] - Yf[act]).astype(f32)
        power = _power(con[g, 0][:, None], con[g, 1][:, None], con[g, 2][:, None], dx, dy)
        with np.errstate(all="ignore"):
            alpha = np.minimum(f32(0.99), (op[g][:, None] * np.exp(power).astype(f32)).astype(f32))
        ok = (~done[act]) & (power <= 0) & (alpha >= f32(1.0 / 255.0))
        testT = (T[act] * (f32(1) - alpha).astype(f32)).astype(f32)
        stop = ok & (testT < f32(0.0001))
        ok &= ~stop
        w = ((alpha * T[act]).astype(f32))
        for ch in range(3):
            C[act, :, ch] = np.where(ok, fma(rgb[g, ch][:, None] * alpha, T[act], C[act, :, ch]), C[act, :, ch])
        D[act] = np.where(ok, fma((dep[g][:, None] * alpha).astype(f32), T[act], D[act]), D[act])
        T[act] = np.where(ok, testT, T[act])
        ncontrib[act] = np.where(ok, k + 1, ncontrib[act])
        d2 = done[act]
        d2 |= stop
        done[act] = d2
        del w
    bg = np.asarray(bg, f32)
    color = np.zeros((3, H, W), f32)
    depth = np.zeros((1, H, W), f32)
    finalT = np.zeros((H, W), f32)
    ncon = np.zeros((H, W), np.int64)
    v = valid
    for ch in range(3):
        color[ch][Y[v], X[v]] = fma(T[v], bg[ch], C[..., ch][v])
    depth[0][Y[v], X[v]] = fma(T[v], f32(z_far), D[v])
    finalT[Y[v], X[v]] = T[v]
    ncon[Y[v], X[v]] = ncontrib[v]
    return color, depth, finalT, ncon


def render_backward(pre, vals, ranges, bg, W, H, finalT, ncon, dL_dpix, dL_ddepthpix, z_far=100.0):
    """float64 evaluation of backward.cu:463-640. Returns per-Gaussian dL_dmean2D [P,2] (w.r.t. NDC),
    dL_dconic [P,3] (x, y, w components), dL_dopacity [P], dL_dcolor [P,3], dL_ddepth [P]."""
    gx, gy = pre["grid"]
    P = pre["radii"].shape[0]
    X, Y, valid = _tile_pixels(W, H, gx, gy)
    nt = gx * gy
    Xc, Yc = np.minimum(X, W - 1), np.minimum(Y, H - 1)
    Tfin = np.where(valid, finalT[Yc, Xc], 0).astype(f64)
    last = np.where(valid, ncon[Yc, Xc], 0)
    dpix = np.where(valid[..., None], np.moveaxis(np.asarray(dL_dpix, f64), 0, -1)[Yc, Xc], 0.0)
    ddep = np.where(valid, np.asarray(dL_ddepthpix, f64).reshape(H, W)[Yc, Xc], 0.0)
    bg = np.asarray(bg, f64)
    bgdot = dpix @ bg + f64(z_far) * ddep
    T = Tfin.copy()
    acc = np.zeros((nt, 256, 3), f64)
    accd = np.zeros((nt, 256), f64)
    last_alpha = np.zeros((nt, 256), f64)
    last_c = np.zeros((nt, 256, 3), f64)
    last_d = np.zeros((nt, 256), f64)
    g_mean = np.zeros((P, 2), f64)
    g_conic = np.zeros((P, 3), f64)
    g_op = np.zeros(P, f64)
    g_col = np.zeros((P, 3), f64)
    g_dep = np.zeros(P, f64)
    Xf, Yf = X.astype(f32), Y.astype(f32)
    m2, con, op, rgb, dep = pre["means2D"], pre["conic"], pre["opacity"], pre["rgb"], pre["depths"]
    maxc = last.max(axis=1) if nt else np.zeros(0, np.int64)
    for k in range(int(maxc.max()) - 1 if nt and maxc.size else -1, -1, -1):
        act = np.nonzero(maxc > k)[0]
        if act.size == 0:
            continue
        g = vals[ranges[act, 0] + k].astype(np.int64)
        dx32 = (m2[g, 0][:, None] - Xf[act]).astype(f32)
        dy32 = (m2[g, 1][:, None] - Yf[act]).astype(f32)
        power = _power(con[g, 0][:, None], con[g, 1][:, None], con[g, 2][:, None], dx32, dy32)
        with np.errstate(all="ignore"):
            G32 = np.exp(power).astype(f32)
            alpha32 = np.minimum(f32(0.99), (op[g][:, None] * G32).astype(f32))
        ok = (k < last[act]) & (power <= 0) & (alpha32 >= f32(1.0 / 255.0))
        if not ok.any():
            continue
        alpha, G = alpha32.astype(f64), G32.astype(f64)
        dx, dy = dx32.astype(f64), dy32.astype(f64)
        Tn = np.where(ok, T[act] / (1.0 - alpha), T[act])
        wgt = alpha * Tn
        la = last_alpha[act]
        c = rgb[g].astype(f64)[:, None, :]
        acc_n = np.where(ok[..., None], la[..., None] * last_c[act] + (1 - la[..., None]) * acc[act], acc[act])
        accd_n = np.where(ok, la * last_d[act] + (1 - la) * accd[act], accd[act])
        cd = dep[g].astype(f64)[:, None]
        dL_dalpha = ((c - acc_n) * dpix[act]).sum(-1) + (cd - accd_n) * ddep[act]
        dL_dalpha = dL_dalpha * Tn + (-Tfin[act] / (1.0 - alpha)) * bgdot[act]
        o = op[g].astype(f64)[:, None]
        dL_dG = o * dL_dalpha
        gdx, gdy = G * dx, G * dy
        cx, cy, cz = (con[g, i].astype(f64)[:, None] for i in range(3))
        dG_ddelx = -gdx * cx - gdy * cy
        dG_ddely = -gdy * cz - gdx * cy
        z = lambda a: np.where(ok, a, 0.0).sum(axis=1)
        np.add.at(g_mean, (g, 0), z(dL_dG * dG_ddelx * (0.5 * W)))
        np.add.at(g_mean, (g, 1), z(dL_dG * dG_ddely * (0.5 * H)))
        np.add.at(g_conic, (g, 0), z(-0.5 * gdx * dx * dL_dG))
        np.add.at(g_conic, (g, 1), z(-0.5 * gdx * dy * dL_dG))
        np.add.at(g_conic, (g, 2), z(-0.5 * gdy * dy * dL_dG))
        np.add.at(g_op, g, z(G * dL_dalpha))
        for ch in range(3):
            np.add.at(g_col, (g, ch), z(wgt * dpix[act][..., ch]))
        np.add.at(g_dep, g, z(wgt * ddep[act]))
        T[act] = Tn
        acc[act] = acc_n
        accd[act] = accd_n
        last_alpha[act] = np.where(ok, alpha, la)
        last_c[act] = np.where(ok[..., None], c, last_c[act])
        last_d[act] = np.where(ok, cd, last_d[act])
    return g_mean, g_conic, g_op, g_col, g_dep


def preprocess_backward(pre, means, scales, rots, shs, deg, view, proj, campos, W, H, tanx, tany, g2d, mod=1.0,
                        use_sigmoid=False):
    """float64 evaluation of the per-Gaussian backward incl. the reference's view/projection-matrix
    gradient conventions. g2d = outputs of render_backward. Returns dict of gradients."""
    g_mean2d, g_conic, g_op, g_col, g_dep = g2d
    means64 = np.asarray(means, f64)
    view = np.asarray(view, f64).reshape(16)
    proj = np.asarray(proj, f64).reshape(16)
    P = means64.shape[0]
    vis = pre["radii"] > 0
    e = pre["ewa"]
    a, b, c = (e[k].astype(f64) for k in ("a", "b", "c"))
    T = [[e["T"][i][j].astype(f64) for j in range(3)] for i in range(3)]
    Wm = [[e["W"][i][j].astype(f64) for j in range(3)] for i in range(3)]
    cov = pre["cov3D"].astype(f64)
    tx, ty, tz = (v.astype(f64) for v in e["t"])
    limx, limy = f64(e["limx"]), f64(e["limy"])
    xg = np.where((e["txtz"] < -e["limx"]) | (e["txtz"] > e["limx"]), 0.0, 1.0)
    yg = np.where((e["tytz"] < -e["limy"]) | (e["tytz"] > e["limy"]), 0.0, 1.0)
    del limx, limy
    dcx, dcy, dcw = g_conic[:, 0], g_conic[:, 1], g_conic[:, 2]
    denom = a * c - b * b
    with np.errstate(all="ignore"):
        d2i = 1.0 / (denom * denom + 0.0000001)
        dL_da = d2i * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcw)
        dL_dc = d2i * (-a * a * dcw + 2 * a * b * dcy + (denom - a * c) * dcx)
        dL_db = d2i * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcw)
        dcov = np.zeros((P, 6), f64)
        dcov[:, 0] = T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc
        dcov[:, 3] = T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc
        dcov[:, 5] = T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc
        dcov[:, 1] = 2 * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][1] * dL_dc
        dcov[:, 2] = 2 * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][2] * dL_dc
        dcov[:, 4] = 2 * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db + 2 * T[1][1] * T[1][2] * dL_dc
        V = [[cov[:, 0], cov[:, 1], cov[:, 2]], [cov[:, 1], cov[:, 3], cov[:, 4]], [cov[:, 2], cov[:, 4], cov[:, 5]]]
        r0 = [T[0][0] * V[j][0] + T[0][1] * V[j][1] + T[0][2] * V[j][2] for j in range(3)]
        r1 = [T[1][0] * V[j][0] + T[1][1] * V[j][1] + T[1][2] * V[j][2] for j in range(3)]
        dT0 = [2 * r0[j] * dL_da + r1[j] * dL_db for j in range(3)]
        dT1 = [2 * r1[j] * dL_dc + r0[j] * dL_db for j in range(3)]
        dJ00 = Wm[0][0] * dT0[0] + Wm[0][1] * dT0[1] + Wm[0][2] * dT0[2]
        dJ02 = Wm[2][0] * dT0[0] + Wm[2][1] * dT0[1] + Wm[2][2] * dT0[2]
        dJ11 = Wm[1][0] * dT1[0] + Wm[1][1] * dT1[1] + Wm[1][2] * dT1[2]
        dJ12 = Wm[2][0] * dT1[0] + Wm[2][1] * dT1[1] + Wm[2][2] * dT1[2]
        fx, fy = (f64(v) for v in pre["focal"])
        iz = 1.0 / tz
        iz2, iz3 = iz * iz, iz * iz * iz
        dtx = xg * -fx * iz2 * dJ02
        dty = yg * -fy * iz2 * dJ12
        dtz = -fx * iz2 * dJ00 - fy * iz2 * dJ11 + (2 * fx * tx) * iz3 * dJ02 + (2 * fy * ty) * iz3 * dJ12
    dt = np.stack([dtx, dty, dtz], 1)
    dt[~vis] = 0
    dcov[~vis] = 0
    dmean = np.stack([view[0] * dt[:, 0] + view[1] * dt[:, 1] + view[2] * dt[:, 2],
                      view[4] * dt[:, 0] + view[5] * dt[:, 1] + view[6] * dt[:, 2],
                      view[8] * dt[:, 0] + view[9] * dt[:, 1] + view[10] * dt[:, 2]], 1)
    dview = np.zeros(16, f64)
    for col in range(3):
        for row in range(3):
            dview[4 * col + row] = (dt[:, row] * means64[:, col]).sum()
    for row in range(3):
        dview[12 + row] = dt[:, row].sum()

    # projection of the mean (backward.cu:396-457)
    hx, hy, _, hw = (v.astype(f64) for v in pre["m_hom"])
    m_w = 1.0 / (hw + 0.0000001)
    gm = np.where(vis[:, None], g_mean2d, 0.0)
    gd = np.where(vis, g_dep, 0.0)
    mul1 = hx * m_w * m_w
    mul2 = hy * m_w * m_w
    for i, (p0, p1, p3) in enumerate(((0, 1, 3), (4, 5, 7), (8, 9, 11))):
        dmean[:, i] += (proj[p0] * m_w - proj[p3] * mul1) * gm[:, 0] + (proj[p1] * m_w - proj[p3] * mul2) * gm[:, 1] \
            + gd * view[2 + 4 * i]
    dproj = np.zeros(16, f64)
    lastcol = (hx * W * gm[:, 0] + hy * H * gm[:, 1]) * m_w * m_w
    mh = np.concatenate([means64, np.ones((P, 1))], 1)
    for col in range(4):
        dproj[4 * col + 0] = (0.5 * gm[:, 0] * mh[:, col] * W * m_w)[vis].sum()
        dproj[4 * col + 1] = (0.5 * gm[:, 1] * mh[:, col] * H * m_w)[vis].sum()
        dproj[4 * col + 3] = (-0.5 * lastcol)[vis].sum()
        dview[4 * col + 2] += (gd * mh[:, col])[vis].sum()

    out = dict(dL_dviewmatrix=dview.reshape(4, 4), dL_dprojmatrix=dproj.reshape(4, 4), dL_dcov3D=dcov,
               dL_dopacity=np.where(vis, g_op, 0.0)[:, None], dL_dmeans2D=gm)
    # SH backward
    if shs is not None:
        shs64 = np.asarray(shs, f64)
        M = shs64.shape[1]
        d0 = means64 - np.asarray(campos, f64)[None]
        n = np.linalg.norm(d0, axis=1)
        dirs = d0 / n[:, None]
        dRGB = np.where(vis[:, None], g_col, 0.0)
        if use_sigmoid:
            sg = 1.0 / (1.0 + np.exp(-pre["pre_act"].astype(f64)))
            dRGB = dRGB * sg * (1 - sg)
        else:
            dRGB = dRGB * pre["clamped"].astype(f64)
        x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
        B, dBx, dBy, dBz = _sh_basis_grad(x, y, z, deg)
        dsh = np.zeros((P, M, 3), f64)
        for k in range(len(B)):
            dsh[:, k, :] = B[k][:, None] * dRGB
        ddir = np.zeros((P, 3), f64)
        for k in range(len(B)):
            s = (shs64[:, k, :] * dRGB).sum(1)
            ddir[:, 0] += dBx[k] * s
            ddir[:, 1] += dBy[k] * s
            ddir[:, 2] += dBz[k] * s
        # through normalisation: (I - d d^T)/|v| applied to ddir
        dmean += np.where(vis[:, None], (ddir - dirs * (dirs * ddir).sum(1, keepdims=True)) / n[:, None], 0.0)
        out["dL_dsh"] = dsh
    else:
        out["dL_dcolors"] = np.where(vis[:, None], g_col, 0.0)
    dmean[~vis] = 0
    out["dL_dmeans3D"] = dmean
    # cov3D -> scale, rotation (no normalisation Jacobian; scale gradient not multiplied by mod)
    if scales is not None:
        _, M32, R32, s32 = cov3d(np.asarray(scales, f32), np.asarray(rots, f32), mod)
        R = [[R32[i][j].astype(f64) for j in range(3)] for i in range(3)]      # column-major R[c][r]
        s = [v.astype(f64) for v in s32]
        Mm = [[s[r] * R[c][r] for r in range(3)] for c in range(3)]
        dS = [[dcov[:, 0], 0.5 * dcov[:, 1], 0.5 * dcov[:, 2]], [0.5 * dcov[:, 1], dcov[:, 3], 0.5 * dcov[:, 4]],
              [0.5 * dcov[:, 2], 0.5 * dcov[:, 4], dcov[:, 5]]]
        dM = [[sum(2.0 * Mm[k][r] * dS[cc][k] for k in range(3)) for r in range(3)] for cc in range(3)]
        Rt = [[R[r][cc] for r in range(3)] for cc in range(3)]
        dMt = [[dM[r][cc] for r in range(3)] for cc in range(3)]
        dscale = np.stack([sum(Rt[i][k] * dMt[i][k] for k in range(3)) for i in range(3)], 1)
        dMt = [[dMt[i][k] * s[i] for k in range(3)] for i in range(3)]
        rots64 = np.asarray(rots, f64)
        r, x, y, z = rots64[:, 0], rots64[:, 1], rots64[:, 2], rots64[:, 3]
        dq = np.stack([
            2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]),
            2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]),
            2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]),
            2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0])], 1)
        out["dL_dscales"] = np.where(vis[:, None], dscale, 0.0)
        out["dL_drotations"] = np.where(vis[:, None], dq, 0.0)
    return out


def _sh_basis_grad(x, y, z, deg):
    """float64 SH basis b_k(x,y,z) (as polynomials, without re-normalising) and its partial derivatives."""
    C0, C1 = f64(SH_C0), f64(SH_C1)
    C2 = [f64(v) for v in SH_C2]
    C3 = [f64(v) for v in SH_C3]
    o = np.zeros_like(x)
    B, dx, dy, dz = [o + C0], [o], [o], [o]
    if deg > 0:
        B += [-C1 * y, C1 * z, -C1 * x]
        dx += [o, o, o - C1]
        dy += [o - C1, o, o]
        dz += [o, o + C1, o]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        B += [C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy)]
        dx += [C2[0] * y, o, C2[2] * -2 * x, C2[3] * z, C2[4] * 2 * x]
        dy += [C2[0] * x, C2[1] * z, C2[2] * -2 * y, o, C2[4] * -2 * y]
        dz += [o, C2[1] * y, C2[2] * 4 * z, C2[3] * x, o]
    if deg > 2:
        B += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
              C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
              C3[6] * x * (xx - 3 * yy)]
        dx += [C3[0] * 6 * xy, C3[1] * yz, C3[2] * -2 * xy, C3[3] * -6 * xz, C3[4] * (-3 * xx + 4 * zz - yy),
               C3[5] * 2 * xz, C3[6] * 3 * (xx - yy)]
        dy += [C3[0] * 3 * (xx - yy), C3[1] * xz, C3[2] * (-3 * yy + 4 * zz - xx), C3[3] * -6 * yz, C3[4] * -2 * xy,
               C3[5] * -2 * yz, C3[6] * -6 * xy]
        dz += [o, C3[1] * xy, C3[2] * 8 * yz, C3[3] * 3 * (2 * zz - xx - yy), C3[4] * 8 * xz, C3[5] * (xx - yy), o]
    return B, dx, dy, dz


def forward(means, scales, rots, opac, shs, deg, view, proj, campos, bg, W, H, tanx, tany, mod=1.0,
            use_sigmoid=False, z_far=100.0, colors_precomp=None, cov3D_precomp=None):
    pre = preprocess(means, scales, rots, opac, shs, deg, view, proj, campos, W, H, tanx, tany, mod, use_sigmoid,
                     colors_precomp, cov3D_precomp)
    keys, vals, ranges = bin_tiles(pre, W, H)
    color, depth, finalT, ncon = render_forward(pre, vals, ranges, bg, W, H, z_far)
    return dict(pre=pre, keys=keys, point_list=vals, ranges=ranges, color=color, depth=depth, final_T=finalT,
                n_contrib=ncon)


def backward(fw, means, scales, rots, shs, deg, view, proj, campos, bg, W, H, tanx, tany, dL_dpix, dL_ddepth,
             mod=1.0, use_sigmoid=False, z_far=100.0):
    g2d = render_backward(fw["pre"], fw["point_list"], fw["ranges"], bg, W, H, fw["final_T"], fw["n_contrib"],
                          dL_dpix, dL_ddepth, z_far)
    return preprocess_backward(fw["pre"], means, scales, rots, shs, deg, view, proj, campos, W, H, tanx, tany,
                               g2d, mod, use_sigmoid)

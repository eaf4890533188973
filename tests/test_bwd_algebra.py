"""CPU: the two algebraic regroupings the blend backward kernel (csrc/dgs_backward.cu, k_render_bwd) makes, checked in
float64 numpy against the reference's formulation (cuda_rasterizer/backward.cu:566-637) on random data:

 (1) phase 1 replays dL/dalpha on UN-NORMALISED suffix sums,
         dL/dalpha_i = T_i (c_i . g) - R_i / (1 - alpha_i),   R_i = sum_{j behind i} alpha_j T_j (c_j . g) + T_final (bg . g),
     where the reference carries the normalised recurrences  accum_rec <- last_alpha * last_color + (1 - last_alpha) * accum_rec
     and  T <- T / (1 - alpha)  back to front;
 (2) phase 2 accumulates moments about the pixel rectangle's corner and shifts them to the Gaussian's centre,
         sum w dx^2 = bx^2 M0 - 2 bx Mx + Mxx  etc.  with  dx = bx - px.
"""
import numpy as np


def _reference_dalpha(alpha, col, g, bg, T_final):
    """Back-to-front replay exactly as the reference writes it (4 channels: r, g, b, depth)."""
    n = len(alpha)
    T = T_final
    accum = np.zeros(4)
    last_alpha, last_col = 0.0, np.zeros(4)
    out = np.zeros(n)
    bg_dot = float(bg @ g)
    for i in range(n - 1, -1, -1):
        T = T / (1.0 - alpha[i])
        d = 0.0
        for ch in range(4):
            accum[ch] = last_alpha * last_col[ch] + (1.0 - last_alpha) * accum[ch]
            last_col[ch] = col[i, ch]
            d += (col[i, ch] - accum[ch]) * g[ch]
        d *= T
        last_alpha = alpha[i]
        d += (-T_final / (1.0 - alpha[i])) * bg_dot
        out[i] = d
    return out


def _suffix_sum_dalpha(alpha, col, g, bg, T_final):
    n = len(alpha)
    T = T_final
    R = T_final * float(bg @ g)
    out = np.zeros(n)
    for i in range(n - 1, -1, -1):
        inv = 1.0 / (1.0 - alpha[i])
        T = T * inv
        cdot = float(col[i] @ g)
        out[i] = T * cdot - R * inv
        R = R + alpha[i] * T * cdot
    return out


def test_suffix_sum_replay_equals_the_reference_recurrences():
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 90))
        alpha = np.minimum(0.99, rng.uniform(1 / 255, 1.0, n) ** rng.uniform(0.5, 3.0))
        col = rng.uniform(0, 1, (n, 4)) * np.array([1, 1, 1, 8.0])
        g = rng.standard_normal(4) * np.array([1, 1, 1, 0.1])
        bg = np.append(rng.uniform(0, 1, 3), 100.0)
        T_final = float(np.prod(1.0 - alpha))
        a = _reference_dalpha(alpha, col, g, bg, T_final)
        b = _suffix_sum_dalpha(alpha, col, g, bg, T_final)
        # and both equal the analytic derivative of C = sum_i alpha_i T_i c_i + T_final bg
        T_front = np.concatenate([[1.0], np.cumprod(1.0 - alpha)[:-1]])
        w = alpha * T_front
        behind = np.array([float((w[i + 1:, None] * col[i + 1:]).sum(0) @ g) for i in range(n)])
        exact = T_front * (col @ g) - (behind + T_final * float(bg @ g)) / (1.0 - alpha)
        scale = np.abs(exact).max() + 1e-30
        assert np.abs(a - exact).max() <= 1e-9 * scale
        assert np.abs(b - exact).max() <= 1e-9 * scale


def test_corner_relative_moments_shift_to_the_centre():
    rng = np.random.default_rng(1)
    px = np.tile(np.arange(8.0), 4)
    py = np.repeat(np.arange(4.0), 8)
    for trial in range(200):
        w = rng.standard_normal(32) * (rng.uniform(size=32) < 0.4)     # ~13 of 32 pixels contribute, mixed signs
        cx, cy = rng.uniform(-20, 28), rng.uniform(-20, 24)            # Gaussian centre relative to the corner
        dx, dy = cx - px, cy - py
        direct = np.array([w.sum(), (w * dx).sum(), (w * dy).sum(), (w * dx * dx).sum(), (w * dx * dy).sum(),
                           (w * dy * dy).sum()])
        # the kernel's two half-warps: rows {0,1} and {2,3}; inside a half py is 0 or 1, so Myy = My
        total = np.zeros(6)
        for half in range(2):
            sel = (py >= 2 * half) & (py < 2 * half + 2)
            ww, x, y = w[sel], px[sel], py[sel] - 2 * half
            M0, Mx, My, Mxx, Mxy = ww.sum(), (ww * x).sum(), (ww * y).sum(), (ww * x * x).sum(), (ww * x * y).sum()
            bx, by = cx, cy - 2 * half
            total += np.array([M0, bx * M0 - Mx, by * M0 - My, bx * (bx * M0 - 2 * Mx) + Mxx,
                               bx * (by * M0 - My) - by * Mx + Mxy, by * (by * M0 - 2 * My) + My])
        scale = np.abs(w).sum() * (abs(cx) + 8) ** 2 + 1e-30
        assert np.abs(total - direct).max() <= 1e-12 * scale

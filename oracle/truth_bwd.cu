// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Float64 evaluation of the tile-blend backward (what the reference's renderCUDA backward,
// submodules/diff-gaussian-rasterization/cuda_rasterizer/backward.cu:463-640, approximates with float
// arithmetic and float atomics in scheduling order).  It runs on the forward state of a reference run (the
// decoded GeometryState / BinningState / ImageState arrays): which (pixel, Gaussian) pairs blend is decided
// exactly as the float forward decided it (same float exponent / expf / thresholds, and the pixel's
// n_contrib), every VALUE -- alpha, transmittance, the recurrences, the sums over pixels -- is float64.
// One thread per pixel, double atomics.  Used by tests to tell whose rounding a difference between the
// library and the reference is.
#include <cuda_runtime.h>
#include <cstdint>

__global__ void k_truth_blend_bwd(int W, int H, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                  const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
                                  const float* __restrict__ colors, const float* __restrict__ depths,
                                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ bg, float z_far,
                                  const float* __restrict__ dL_dpix, const float* __restrict__ dL_ddepth,
                                  double* __restrict__ dmean2D, double* __restrict__ dconic, double* __restrict__ dopacity,
                                  double* __restrict__ dcolor, double* __restrict__ ddepth)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int tiles_x = (W + 15) / 16;
    const uint2 range = ranges[(y / 16) * tiles_x + (x / 16)];
    const size_t pix = (size_t)y * W + x, HW = (size_t)H * W;
    const int last = (int)n_contrib[pix];
    const float pxf = (float)x, pyf = (float)y;
    const double dp[4] = {dL_dpix[pix], dL_dpix[HW + pix], dL_dpix[2 * HW + pix], dL_ddepth ? dL_ddepth[pix] : 0.0};

    // does the float forward blend this pair?  (forward.cu:341-373, same expression trees)
    auto blends = [&](uint32_t g) {
        const float2 xy = means2D[g];
        const float4 co = conic_opacity[g];
        const float dx = xy.x - pxf, dy = xy.y - pyf;
        const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
        if (power > 0.0f) return false;
        const float alpha = min(0.99f, co.w * expf(power));
        return !(alpha < 1.0f / 255.0f);
    };
    auto alpha64 = [&](uint32_t g, double& G, double& dx, double& dy) {
        const float2 xy = means2D[g];
        const float4 co = conic_opacity[g];
        dx = (double)xy.x - (double)pxf;
        dy = (double)xy.y - (double)pyf;
        G = exp(-0.5 * ((double)co.x * dx * dx + (double)co.z * dy * dy) - (double)co.y * dx * dy);
        return fmin(0.99, (double)co.w * G);
    };

    // transmittance behind the last contributor, in float64
    double T = 1.0;
    for (int i = 0; i < last; i++) {
        const uint32_t g = point_list[range.x + i];
        if (!blends(g)) continue;
        double G, dx, dy;
        T *= 1.0 - alpha64(g, G, dx, dy);
    }
    const double T_final = T;
    const double bg_dot = bg[0] * dp[0] + bg[1] * dp[1] + bg[2] * dp[2] + (double)z_far * dp[3];

    double accum[4] = {0, 0, 0, 0}, last_alpha = 0, last_c[4] = {0, 0, 0, 0};
    for (int i = last - 1; i >= 0; i--) {
        const uint32_t g = point_list[range.x + i];
        if (!blends(g)) continue;
        double G, dx, dy;
        const double alpha = alpha64(g, G, dx, dy);
        T = T / (1.0 - alpha);
        const double w = alpha * T;
        const double c[4] = {colors[3 * g], colors[3 * g + 1], colors[3 * g + 2], depths[g]};
        double dL_dalpha = 0;
        for (int ch = 0; ch < 4; ch++) {
            accum[ch] = last_alpha * last_c[ch] + (1.0 - last_alpha) * accum[ch];
            last_c[ch] = c[ch];
            dL_dalpha += (c[ch] - accum[ch]) * dp[ch];
        }
        for (int ch = 0; ch < 3; ch++) atomicAdd(dcolor + 3 * g + ch, w * dp[ch]);
        atomicAdd(ddepth + g, w * dp[3]);
        dL_dalpha *= T;
        last_alpha = alpha;
        dL_dalpha += (-T_final / (1.0 - alpha)) * bg_dot;
        const float4 co = conic_opacity[g];
        const double dL_dG = (double)co.w * dL_dalpha;
        const double gdx = G * dx, gdy = G * dy;
        const double dG_ddelx = -gdx * co.x - gdy * co.y, dG_ddely = -gdy * co.z - gdx * co.y;
        atomicAdd(dmean2D + 2 * g, dL_dG * dG_ddelx * (0.5 * W));
        atomicAdd(dmean2D + 2 * g + 1, dL_dG * dG_ddely * (0.5 * H));
        atomicAdd(dconic + 3 * g, -0.5 * gdx * dx * dL_dG);
        atomicAdd(dconic + 3 * g + 1, -0.5 * gdx * dy * dL_dG);
        atomicAdd(dconic + 3 * g + 2, -0.5 * gdy * dy * dL_dG);
        atomicAdd(dopacity + g, G * dL_dalpha);
    }
}

extern "C" void truth_blend_backward(int W, int H, const void* ranges, const void* point_list, const void* means2D,
                                     const void* conic_opacity, const float* colors, const float* depths,
                                     const void* n_contrib, const float* bg, float z_far, const float* dL_dpix,
                                     const float* dL_ddepth, double* dmean2D, double* dconic, double* dopacity,
                                     double* dcolor, double* ddepth)
{
    dim3 block(16, 16), grid((W + 15) / 16, (H + 15) / 16);
    k_truth_blend_bwd<<<grid, block>>>(W, H, (const uint2*)ranges, (const uint32_t*)point_list, (const float2*)means2D,
                                       (const float4*)conic_opacity, colors, depths, (const uint32_t*)n_contrib, bg,
                                       z_far, dL_dpix, dL_ddepth, dmean2D, dconic, dopacity, dcolor, ddepth);
}

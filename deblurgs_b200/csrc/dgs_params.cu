// The two steps either side of the rasterizer on the Gaussian parameter store:
// the activations `render` reads, and the Adam update that consumes the rasterizer's gradients.
//
// Reference behaviour restated here (taekkii/deblurgs):
//   get_scaling   = exp(_scaling) + scale_lb  (isotropic: column 0 expanded)   scene/gaussian_model.py:114-119,
//                                                                              scene/gaussian_activation.py:43-52
//   get_rotation  = _rotation / max(|_rotation|_2, 1e-12)  (F.normalize)       scene/gaussian_model.py:50,121-123
//   get_opacity   = clamp(_opacity, 0, 1)                                      scene/gaussian_activation.py:29-34
//   get_features  = cat(_features_dc, _features_rest, dim=1)                   scene/gaussian_model.py:129-133
//   optimizer     = torch.optim.Adam(groups, lr=0.0, eps=1e-15)                scene/gaussian_model.py:181-190
// The reference evaluates the four getters once per SUB-FRAME render (~8 elementwise launches and a
// 12*M-byte cat copy each, again in backward) and steps Adam with one multi-kernel foreach pass per
// parameter tensor.  Here: one launch activates the whole store, one launch back-propagates through
// it, one launch steps every parameter tensor of the model.  All three are pure streaming kernels
// (HBM-bound): flat, coalesced indexing; the per-Gaussian maths rides on the first P threads.
#include "dgs_b200.h"
#include "dgs_internal.cuh"

#include <cmath>
#include <cstring>

namespace dgs {

// sh [P, M, 3] = cat(dc [P,1,3], rest [P,M-1,3]);  per Gaussian: scales, rotations, opacities.
__global__ void __launch_bounds__(256) k_activate_fwd(int P, int M, const float* __restrict__ dc,
                                                      const float* __restrict__ rest,
                                                      const float* __restrict__ scaling,
                                                      const float* __restrict__ rotation,
                                                      const float* __restrict__ opacity, float scale_lb,
                                                      int isotropic, float* __restrict__ sh,
                                                      float* __restrict__ scales, float* __restrict__ rot,
                                                      float* __restrict__ opac)
{
    const size_t row = (size_t)3 * M, total = (size_t)P * row;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (total < 0xF0000000ull) {   // 32-bit index arithmetic (a 64-bit divide per element would make this ALU-bound); headroom for i += stride
        const uint32_t row32 = (uint32_t)row;
        for (uint32_t i = (uint32_t)t0; i < (uint32_t)total; i += (uint32_t)stride) {
            const uint32_t g = i / row32, r = i - g * row32;
            sh[i] = r < 3 ? dc[g * 3 + r] : rest[(size_t)g * (row32 - 3) + (r - 3)];
        }
    } else {
        for (size_t i = t0; i < total; i += stride) {
            const size_t g = i / row, r = i - g * row;
            sh[i] = r < 3 ? dc[g * 3 + r] : rest[g * (row - 3) + (r - 3)];
        }
    }
    for (size_t g = t0; g < (size_t)P; g += stride) {
        const float s0 = scaling[3 * g], s1 = scaling[3 * g + 1], s2 = scaling[3 * g + 2];
        scales[3 * g] = expf(s0) + scale_lb;
        scales[3 * g + 1] = expf(isotropic ? s0 : s1) + scale_lb;
        scales[3 * g + 2] = expf(isotropic ? s0 : s2) + scale_lb;
        const float4 q = reinterpret_cast<const float4*>(rotation)[g];
        const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
        reinterpret_cast<float4*>(rot)[g] = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
        opac[g] = fminf(fmaxf(opacity[g], 0.0f), 1.0f);
    }
}

// Chain rule of the above.  Any of the incoming gradients may be NULL (treated as zero).
__global__ void __launch_bounds__(256) k_activate_bwd(int P, int M, const float* __restrict__ scaling,
                                                      const float* __restrict__ rotation,
                                                      const float* __restrict__ opacity, int isotropic,
                                                      const float* __restrict__ dsh,
                                                      const float* __restrict__ dscales,
                                                      const float* __restrict__ drot,
                                                      const float* __restrict__ dopac, float* __restrict__ ddc,
                                                      float* __restrict__ drest, float* __restrict__ dscaling,
                                                      float* __restrict__ drotation, float* __restrict__ dopacity)
{
    const size_t row = (size_t)3 * M, total = (size_t)P * row;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (total < 0xF0000000ull) {
        const uint32_t row32 = (uint32_t)row;
        for (uint32_t i = (uint32_t)t0; i < (uint32_t)total; i += (uint32_t)stride) {
            const uint32_t g = i / row32, r = i - g * row32;
            const float v = dsh ? dsh[i] : 0.f;
            if (r < 3) ddc[g * 3 + r] = v; else drest[(size_t)g * (row32 - 3) + (r - 3)] = v;
        }
    } else {
        for (size_t i = t0; i < total; i += stride) {
            const size_t g = i / row, r = i - g * row;
            const float v = dsh ? dsh[i] : 0.f;
            if (r < 3) ddc[g * 3 + r] = v; else drest[g * (row - 3) + (r - 3)] = v;
        }
    }
    for (size_t g = t0; g < (size_t)P; g += stride) {
        // exp: d/dx = exp(x) (torch multiplies by the saved result; the lower bound is an additive constant)
        const float s0 = scaling[3 * g], s1 = scaling[3 * g + 1], s2 = scaling[3 * g + 2];
        const float g0 = dscales ? dscales[3 * g] : 0.f, g1 = dscales ? dscales[3 * g + 1] : 0.f,
                    g2 = dscales ? dscales[3 * g + 2] : 0.f;
        if (isotropic) {
            // the expanded column 0 collects the three gradients; columns 1, 2 are unused parameters
            const float e0 = expf(s0);
            dscaling[3 * g] = g0 * e0 + g1 * e0 + g2 * e0;
            dscaling[3 * g + 1] = 0.f;
            dscaling[3 * g + 2] = 0.f;
        } else {
            dscaling[3 * g] = g0 * expf(s0);
            dscaling[3 * g + 1] = g1 * expf(s1);
            dscaling[3 * g + 2] = g2 * expf(s2);
        }
        // normalize: r = q / n  =>  dq = (g - r (r.g)) / n   (n clamped at 1e-12: then dq = g / 1e-12)
        const float4 q = reinterpret_cast<const float4*>(rotation)[g];
        const float4 gr = drot ? reinterpret_cast<const float4*>(drot)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float nn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        float4 dq;
        if (nn > 1e-12f) {
            const float inv = 1.0f / nn;
            const float rx = q.x * inv, ry = q.y * inv, rz = q.z * inv, rw = q.w * inv;
            const float dot = rx * gr.x + ry * gr.y + rz * gr.z + rw * gr.w;
            dq = make_float4((gr.x - rx * dot) * inv, (gr.y - ry * dot) * inv, (gr.z - rz * dot) * inv,
                             (gr.w - rw * dot) * inv);
        } else {
            dq = make_float4(gr.x * 1e12f, gr.y * 1e12f, gr.z * 1e12f, gr.w * 1e12f);
        }
        reinterpret_cast<float4*>(drotation)[g] = dq;
        // clamp: gradient passes where 0 <= x <= 1
        const float o = opacity[g];
        dopacity[g] = (o >= 0.0f && o <= 1.0f) ? (dopac ? dopac[g] : 0.f) : 0.f;
    }
}

// ---- Adam over up to DGS_ADAM_MAX_TENSORS parameter tensors in one launch ---------------------
struct AdamTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    long long first_block;   // blocks [first_block, first_block + ceil(n / ADAM_CHUNK)) belong to this tensor
    float step_size;         // lr / (1 - beta1^t)
    float bc2_sqrt;          // sqrt(1 - beta2^t)
};
struct AdamBatch {
    AdamTensor t[DGS_ADAM_MAX_TENSORS];
    int count;
    float w1, beta2, w2, eps;   // w1 = 1 - beta1, w2 = 1 - beta2 (rounded from double, like torch's scalars)
    float clip;                 // clip_grad_value_ bound, <= 0: off (train.py:204-205)
};
#define ADAM_PER_THREAD 8
#define ADAM_CHUNK (256 * ADAM_PER_THREAD)   // elements per block

// torch.optim.Adam (single-tensor path, amsgrad / weight_decay / maximize off), same operation order:
//   m <- lerp(m, g, 1 - beta1) = m + (1 - beta1) (g - m)
//   v <- (v * beta2) + ((1 - beta2) * g) * g
//   p <- p + (-step_size) * (m / (sqrt(v) / sqrt(bc2) + eps))
// 28 B of HBM traffic per element and nothing else: every thread first issues all of its loads
// (4 x ADAM_PER_THREAD independent requests in flight), then computes, then stores.
__global__ void __launch_bounds__(256) k_adam(const AdamBatch b)
{
    int ti = 0;
#pragma unroll
    for (int k = 1; k < DGS_ADAM_MAX_TENSORS; k++)
        if (k < b.count && (long long)blockIdx.x >= b.t[k].first_block) ti = k;
    const AdamTensor& t = b.t[ti];
    const long long base = ((long long)blockIdx.x - t.first_block) * ADAM_CHUNK + threadIdx.x;
    float g[ADAM_PER_THREAD], m[ADAM_PER_THREAD], v[ADAM_PER_THREAD], p[ADAM_PER_THREAD];
#pragma unroll
    for (int k = 0; k < ADAM_PER_THREAD; k++) {
        const long long i = base + k * 256;
        if (i < t.n) { g[k] = __ldg(t.g + i); m[k] = t.m[i]; v[k] = t.v[i]; p[k] = t.p[i]; }
    }
#pragma unroll
    for (int k = 0; k < ADAM_PER_THREAD; k++) {
        const long long i = base + k * 256;
        if (i < t.n) {
            float gk = g[k];
            if (b.clip > 0.f) gk = fminf(fmaxf(gk, -b.clip), b.clip);
            const float mk = fmaf(b.w1, gk - m[k], m[k]);
            const float vk = fmaf(b.w2 * gk, gk, __fmul_rn(v[k], b.beta2));
            const float denom = __fdiv_rn(sqrtf(vk), t.bc2_sqrt) + b.eps;
            t.m[i] = mk;
            t.v[i] = vk;
            t.p[i] = fmaf(-t.step_size, __fdiv_rn(mk, denom), p[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Densify / prune on the SoA parameter store (reference: scene/gaussian_model.py:300-454 -- prune_points,
// cat_tensors_to_optimizer / densification_postfix, densify_and_clone, densify_and_split).  Every one of those
// operations is "new row r of EVERY per-Gaussian tensor = old row src[r]" with the Adam moments of appended rows
// zeroed; the reference runs it as ~40 torch index / cat launches over the 6 parameter tensors, their 12 moment
// tensors and the 3 statistics vectors.  Here one launch rebuilds all of them: a thread per output float, rows of
// one tensor contiguous across the block so loads and stores are coalesced along the row.
// ---------------------------------------------------------------------------------------
struct GatherTensor { const float* in; float* out; int width; int zero_new; long long first_block; };
struct GatherBatch { GatherTensor t[DGS_GATHER_MAX_TENSORS]; int count; long long n_out; };

__global__ void __launch_bounds__(256) k_rows_gather(const GatherBatch b, const long long* __restrict__ src,
                                                     const unsigned char* __restrict__ carry)
{
    int ti = 0;
#pragma unroll 1
    for (int k = 1; k < b.count; k++)
        if ((long long)blockIdx.x >= b.t[k].first_block) ti = k;
    const GatherTensor& t = b.t[ti];
    const long long i = ((long long)blockIdx.x - t.first_block) * 256 + threadIdx.x;     // output float index
    if (i >= b.n_out * t.width) return;
    const long long row = i / t.width;
    const int c = (int)(i - row * t.width);
    float v = 0.f;
    if (!(t.zero_new && carry != nullptr && carry[row] == 0)) v = __ldg(t.in + src[row] * t.width + c);
    t.out[i] = v;
}

}  // namespace dgs

extern "C" {

int dgs_rows_gather(int n_tensors, const float* const* in, float* const* out, const int* width, const int* zero_new,
                    int64_t n_out, const int64_t* src_rows, const uint8_t* carry, void* stream)
{
    if (n_tensors < 0 || n_tensors > DGS_GATHER_MAX_TENSORS || n_out < 0)
        return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_rows_gather: invalid argument");
    if (n_tensors == 0 || n_out == 0) return DGS_OK;
    if (!in || !out || !width || !zero_new || !src_rows) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_rows_gather: null argument");
    dgs::GatherBatch b;
    memset(&b, 0, sizeof(b));
    b.n_out = n_out;
    long long blocks = 0;
    for (int k = 0; k < n_tensors; k++) {
        if (width[k] <= 0 || !in[k] || !out[k]) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_rows_gather: bad tensor");
        dgs::GatherTensor& t = b.t[b.count++];
        t.in = in[k]; t.out = out[k]; t.width = width[k]; t.zero_new = zero_new[k]; t.first_block = blocks;
        blocks += ((long long)n_out * width[k] + 255) / 256;
    }
    if (blocks > 0x7fffffffLL) return dgs::fail(DGS_ERR_UNSUPPORTED, "dgs_rows_gather: size not supported");
    {
        dgs::StageTimer timer(dgs::ST_DENSIFY, (cudaStream_t)stream, 1);
        dgs::k_rows_gather<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b, (const long long*)src_rows, carry);
    }
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_rows_gather"); }
}

int dgs_activate_forward(int P, int sh_coeffs, const float* features_dc, const float* features_rest,
                         const float* scaling, const float* rotation, const float* opacity,
                         float scale_lower_bound, int isotropic,
                         float* shs, float* scales, float* rotations, float* opacities, void* stream)
{
    if (P < 0 || sh_coeffs < 1) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_activate_forward: invalid argument");
    if (P == 0) return DGS_OK;
    if (!features_dc || (sh_coeffs > 1 && !features_rest) || !scaling || !rotation || !opacity || !shs ||
        !scales || !rotations || !opacities)
        return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_activate_forward: invalid argument");
    const size_t total = (size_t)P * 3 * sh_coeffs;
    const size_t want = (total + 255) / 256;
    const int blocks = (int)(want < (size_t)148 * 16 ? want : (size_t)148 * 16);
    {
        dgs::StageTimer timer(dgs::ST_ACTIVATE_FWD, (cudaStream_t)stream, 1);
        dgs::k_activate_fwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, sh_coeffs, features_dc, features_rest, scaling,
                                                                      rotation, opacity, scale_lower_bound, isotropic,
                                                                      shs, scales, rotations, opacities);
    }
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_activate_forward"); }
}

int dgs_activate_backward(int P, int sh_coeffs, const float* scaling, const float* rotation, const float* opacity,
                          int isotropic, const float* dL_dshs, const float* dL_dscales, const float* dL_drotations,
                          const float* dL_dopacities, float* dL_dfeatures_dc, float* dL_dfeatures_rest,
                          float* dL_dscaling, float* dL_drotation, float* dL_dopacity, void* stream)
{
    if (P < 0 || sh_coeffs < 1) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_activate_backward: invalid argument");
    if (P == 0) return DGS_OK;
    if (!scaling || !rotation || !opacity || !dL_dfeatures_dc || (sh_coeffs > 1 && !dL_dfeatures_rest) ||
        !dL_dscaling || !dL_drotation || !dL_dopacity)
        return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_activate_backward: invalid argument");
    const size_t total = (size_t)P * 3 * sh_coeffs;
    const size_t want = (total + 255) / 256;
    const int blocks = (int)(want < (size_t)148 * 16 ? want : (size_t)148 * 16);
    {
        dgs::StageTimer timer(dgs::ST_ACTIVATE_BWD, (cudaStream_t)stream, 1);
        dgs::k_activate_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, sh_coeffs, scaling, rotation, opacity, isotropic,
                                                                      dL_dshs, dL_dscales, dL_drotations, dL_dopacities,
                                                                      dL_dfeatures_dc, dL_dfeatures_rest, dL_dscaling,
                                                                      dL_drotation, dL_dopacity);
    }
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_activate_backward"); }
}

int dgs_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const int64_t* numel, const double* lr, const int64_t* step,
                  double beta1, double beta2, double eps, double clip_grad_value, void* stream)
{
    if (n_tensors < 0 || n_tensors > DGS_ADAM_MAX_TENSORS) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_adam_step: invalid argument");
    if (n_tensors == 0) return DGS_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !lr || !step) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_adam_step: invalid argument");
    dgs::AdamBatch b;
    memset(&b, 0, sizeof(b));
    // scalars are formed in double and rounded once, as torch does with its Python floats
    b.w1 = (float)(1.0 - beta1); b.beta2 = (float)beta2; b.w2 = (float)(1.0 - beta2); b.eps = (float)eps;
    b.clip = (float)clip_grad_value;
    long long blocks = 0;
    for (int k = 0; k < n_tensors; k++) {
        if (numel[k] < 0 || step[k] < 1) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_adam_step: invalid argument");
        if (numel[k] == 0) continue;
        if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k]) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_adam_step: invalid argument");
        dgs::AdamTensor& t = b.t[b.count++];
        t.p = params[k]; t.g = grads[k]; t.m = exp_avg[k]; t.v = exp_avg_sq[k];
        t.n = numel[k];
        t.first_block = blocks;
        const double bc1 = 1.0 - pow(beta1, (double)step[k]);
        const double bc2 = 1.0 - pow(beta2, (double)step[k]);
        t.step_size = (float)(lr[k] / bc1);
        t.bc2_sqrt = (float)sqrt(bc2);
        blocks += (numel[k] + ADAM_CHUNK - 1) / ADAM_CHUNK;
    }
    if (blocks == 0) return DGS_OK;
    if (blocks > 0x7fffffffLL) return dgs::fail(DGS_ERR_UNSUPPORTED, "dgs_adam_step: size not supported");
    {
        dgs::StageTimer timer(dgs::ST_ADAM, (cudaStream_t)stream, 1);
        dgs::k_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b);
    }
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_adam_step"); }
}

}  // extern "C"

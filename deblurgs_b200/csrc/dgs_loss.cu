// Fused photometric loss of the blurry-view training step and its gradient.
//
// Reference behaviour restated here (taekkii/deblurgs):
//   Ll1        = l1_loss(blur, gt)                       utils/loss_utils.py:17-18, train.py:147
//   L_t_smooth = batchwise_smoothness_loss(subframes)    utils/loss_utils.py:80-93, train.py:148
//              = mean |subframes[1:] - subframes[:-1]|   (0 when there is a single sub-frame)
//   loss       = Ll1 + lambda_t_smooth * L_t_smooth      train.py:160-163 (photometric part)
// The reference evaluates this with ~8 elementwise / reduction launches over the [F,3,H,W] stack
// forward and as many backward; here one kernel reads every sub-frame pixel once and produces both
// sums, and one kernel writes dL/dblurred and dL/dsubframes (sign() = 0 at 0, like torch.abs).
#include "dgs_b200.h"
#include "dgs_internal.cuh"

namespace dgs {

__global__ void __launch_bounds__(256) k_blur_loss_fwd(int F, size_t chw, const float* __restrict__ sub,
                                                       const float* __restrict__ blur,
                                                       const float* __restrict__ gt, double* __restrict__ sums)
{
    double a1 = 0.0, a2 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chw; i += (size_t)gridDim.x * blockDim.x) {
        a1 += (double)fabsf(blur[i] - gt[i]);
        if (sub == nullptr) continue;   // lambda_t_smooth == 0: the [F,3,H,W] stack is not read at all
        float prev = sub[i];
        float acc = 0.f;
        for (int s = 1; s < F; s++) {
            const float cur = sub[(size_t)s * chw + i];
            acc += fabsf(cur - prev);
            prev = cur;
        }
        a2 += (double)acc;
    }
    for (int d = 16; d >= 1; d >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, d);
        a2 += __shfl_xor_sync(0xffffffffu, a2, d);
    }
    __shared__ double s1[8], s2[8];
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a1; s2[threadIdx.x >> 5] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; w++) { t1 += s1[w]; t2 += s2[w]; }
        atomicAdd(sums, t1);
        atomicAdd(sums + 1, t2);
    }
}

__global__ void k_blur_loss_finalize(int F, size_t chw, float lambda_t, const double* __restrict__ sums,
                                     float* __restrict__ out)
{
    const double l1 = sums[0] / (double)chw;
    const double sm = F > 1 ? sums[1] / ((double)(F - 1) * (double)chw) : 0.0;
    out[0] = (float)(l1 + (double)lambda_t * sm);
    out[1] = (float)l1;
    out[2] = (float)sm;
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(256) k_blur_loss_bwd(int F, size_t chw, const float* __restrict__ sub,
                                                       const float* __restrict__ blur,
                                                       const float* __restrict__ gt, float lambda_t,
                                                       const float* __restrict__ grad_out,
                                                       float* __restrict__ dblur, float* __restrict__ dsub)
{
    const float go = grad_out ? grad_out[0] : 1.0f;
    const float k1 = go / (float)chw;
    const float k2 = F > 1 ? go * lambda_t / ((float)(F - 1) * (float)chw) : 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chw; i += (size_t)gridDim.x * blockDim.x) {
        dblur[i] = sgn(blur[i] - gt[i]) * k1;
        if (dsub == nullptr) continue;   // lambda_t_smooth == 0: no gradient for the sub-frame stack
        float prev = sub[i];
        float sprev = 0.f;                       // sign(sub[s] - sub[s-1])
        for (int s = 0; s < F; s++) {
            float snext = 0.f, nxt = prev;
            if (s + 1 < F) {
                nxt = sub[(size_t)(s + 1) * chw + i];
                snext = sgn(nxt - prev);
            }
            dsub[(size_t)s * chw + i] = (sprev - snext) * k2;
            sprev = snext;
            prev = nxt;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Regularisers of the training loss (train.py:150-163):
//   tv_loss(x [B,C,H,W]) = mean (x[..,h,w] - x[..,h+1,w])^2 + mean (x[..,h,w] - x[..,h,w+1])^2      utils/loss_utils.py:66-78
//   hinge_l2(x)          = mean of x^2 where x <= 0, (x-1)^2 where x >= 1, 0 elsewhere                utils/loss_utils.py:95-104
// One kernel each way, double accumulation of the sums.  (The reference's call site feeds tv_loss the depth stack
// with an extra singleton axis, `subframe_depths[:,None,:,:]` of a [F,1,H,W] tensor, so that its "h" axis has
// length 1 and the first term is the mean of an empty tensor; lambda_depth_tv defaults to 0 and the term is dead
// there.  Here tv_loss is what its docstring says: over the image axes of a [B,C,H,W] tensor.)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_to(double a, double b, double* sums)
{
    for (int d = 16; d >= 1; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    __shared__ double s1[8], s2[8];
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; w++) { t1 += s1[w]; t2 += s2[w]; }
        atomicAdd(sums, t1);
        atomicAdd(sums + 1, t2);
    }
}

__global__ void __launch_bounds__(256) k_tv_fwd(size_t planes, int H, int W, const float* __restrict__ x,
                                                double* __restrict__ sums)
{
    const size_t n = planes * H * W;
    double a = 0.0, b = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const float v = x[i];
        if (h + 1 < H) { const float d = v - x[i + W]; a += (double)(d * d); }
        if (w + 1 < W) { const float d = v - x[i + 1]; b += (double)(d * d); }
    }
    block_sum_to(a, b, sums);
}
__global__ void k_tv_finalize(size_t planes, int H, int W, const double* __restrict__ sums, float* __restrict__ out)
{
    const double nh = (double)planes * (H - 1) * W, nw = (double)planes * H * (W - 1);
    out[0] = (float)(sums[0] / nh + sums[1] / nw);    // (0/0 = NaN for H == 1 or W == 1, like torch's mean of an empty tensor)
}
__global__ void __launch_bounds__(256) k_tv_bwd(size_t planes, int H, int W, const float* __restrict__ x,
                                                const float* __restrict__ grad_out, float* __restrict__ dx)
{
    const size_t n = planes * H * W;
    const float go = grad_out ? grad_out[0] : 1.0f;
    const float kh = 2.0f * go / ((float)planes * (float)(H - 1) * (float)W);
    const float kw = 2.0f * go / ((float)planes * (float)H * (float)(W - 1));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const float v = x[i];
        float g = 0.f;
        if (h + 1 < H) g += kh * (v - x[i + W]);
        if (h > 0) g -= kh * (x[i - W] - v);
        if (w + 1 < W) g += kw * (v - x[i + 1]);
        if (w > 0) g -= kw * (x[i - 1] - v);
        dx[i] = g;
    }
}

__global__ void __launch_bounds__(256) k_hinge_fwd(size_t n, const float* __restrict__ x, double* __restrict__ sums)
{
    double a = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        if (v <= 0.0f) a += (double)(v * v);
        else if (v >= 1.0f) a += (double)((v - 1.0f) * (v - 1.0f));
    }
    block_sum_to(a, 0.0, sums);
}
__global__ void k_hinge_finalize(size_t n, const double* __restrict__ sums, float* __restrict__ out)
{
    out[0] = (float)(sums[0] / (double)n);
}
__global__ void __launch_bounds__(256) k_hinge_bwd(size_t n, const float* __restrict__ x, const float* __restrict__ grad_out,
                                                   float* __restrict__ dx)
{
    const float k = 2.0f * (grad_out ? grad_out[0] : 1.0f) / (float)n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        dx[i] = v <= 0.0f ? k * v : (v >= 1.0f ? k * (v - 1.0f) : 0.0f);
    }
}

}  // namespace dgs

extern "C" {

int dgs_tv_loss_forward(int64_t planes, int H, int W, const float* x, float* loss_out, double* scratch, void* stream)
{
    if (planes <= 0 || H <= 0 || W <= 0 || !x || !loss_out || !scratch)
        return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_tv_loss_forward: invalid argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st) != cudaSuccess) return dgs::fail_cuda(cudaGetLastError(), "dgs_tv_loss_forward");
    const size_t n = (size_t)planes * H * W;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    dgs::k_tv_fwd<<<blocks, 256, 0, st>>>((size_t)planes, H, W, x, scratch);
    dgs::k_tv_finalize<<<1, 1, 0, st>>>((size_t)planes, H, W, scratch, loss_out);
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_tv_loss_forward"); }
}
int dgs_tv_loss_backward(int64_t planes, int H, int W, const float* x, const float* grad_out, float* dL_dx, void* stream)
{
    if (planes <= 0 || H <= 0 || W <= 0 || !x || !dL_dx)
        return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_tv_loss_backward: invalid argument");
    const size_t n = (size_t)planes * H * W;
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    dgs::k_tv_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>((size_t)planes, H, W, x, grad_out, dL_dx);
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_tv_loss_backward"); }
}
int dgs_hinge_l2_forward(int64_t n, const float* x, float* loss_out, double* scratch, void* stream)
{
    if (n <= 0 || !x || !loss_out || !scratch) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_hinge_l2_forward: invalid argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st) != cudaSuccess) return dgs::fail_cuda(cudaGetLastError(), "dgs_hinge_l2_forward");
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    dgs::k_hinge_fwd<<<blocks, 256, 0, st>>>((size_t)n, x, scratch);
    dgs::k_hinge_finalize<<<1, 1, 0, st>>>((size_t)n, scratch, loss_out);
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_hinge_l2_forward"); }
}
int dgs_hinge_l2_backward(int64_t n, const float* x, const float* grad_out, float* dL_dx, void* stream)
{
    if (n <= 0 || !x || !dL_dx) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_hinge_l2_backward: invalid argument");
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    dgs::k_hinge_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>((size_t)n, x, grad_out, dL_dx);
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_hinge_l2_backward"); }
}

int dgs_blur_loss_forward(int F, int64_t chw, const float* subframes, const float* blurred, const float* gt,
                          float lambda_t_smooth, float* loss_out, double* scratch, void* stream)
{
    if (F <= 0 || chw <= 0) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_blur_loss_forward: invalid argument");
    if (!blurred || !gt || !loss_out || !scratch) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_blur_loss_forward: invalid argument");
    if (!subframes && lambda_t_smooth != 0.0f) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_blur_loss_forward: invalid argument");   // NULL stack only with lambda = 0
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st) != cudaSuccess) return dgs::fail_cuda(cudaGetLastError(), "dgs_blur_loss_forward");
    const int blocks = (int)((chw + 255) / 256 < 148 * 8 ? (chw + 255) / 256 : 148 * 8);
    dgs::k_blur_loss_fwd<<<blocks, 256, 0, st>>>(F, (size_t)chw, subframes, blurred, gt, scratch);
    dgs::k_blur_loss_finalize<<<1, 1, 0, st>>>(F, (size_t)chw, lambda_t_smooth, scratch, loss_out);
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_blur_loss_forward"); }
}

int dgs_blur_loss_backward(int F, int64_t chw, const float* subframes, const float* blurred, const float* gt,
                           float lambda_t_smooth, const float* grad_out, float* dL_dblurred,
                           float* dL_dsubframes, void* stream)
{
    if (F <= 0 || chw <= 0) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_blur_loss_backward: invalid argument");
    if (!blurred || !gt || !dL_dblurred) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_blur_loss_backward: invalid argument");
    if ((!subframes || !dL_dsubframes) && lambda_t_smooth != 0.0f) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_blur_loss_backward: invalid argument");
    if (!subframes) dL_dsubframes = nullptr;
    const int blocks = (int)((chw + 255) / 256 < 148 * 8 ? (chw + 255) / 256 : 148 * 8);
    dgs::k_blur_loss_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(F, (size_t)chw, subframes, blurred, gt,
                                                                    lambda_t_smooth, grad_out, dL_dblurred,
                                                                    dL_dsubframes);
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_blur_loss_backward"); }
}

}  // extern "C"

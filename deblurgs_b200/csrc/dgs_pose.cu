// Sub-frame camera poses from Bezier control points in se(3), with their Jacobian.
//
// Reference behaviour restated here (taekkii/deblurgs):
//   Bernstein weights   scene/bezier.py:54-64  (control point k <-> binom(C,k) t^(C-k) (1-t)^k,
//                       powers and their product in fp32, times an fp64 binomial => fp64)
//   curve sample        scene/bezier.py:81      (fp64 sum over control points)
//   se3 exponential     utils/pytorch3d_functions.py:218-247, 373-457, 546-573
//                       (theta = sqrt(max(|omega|^2, 1e-4)); no Taylor branch)
//   view / projection   scene/motion.py:277-282 (wvt[:3,:3]=R, wvt[3,:3]=-t@R cast to fp32,
//                       full_proj = wvt @ projection_matrix in fp32)
//   camera centre       scene/cameras.py:63-74  (inverse(wvt)[3,:3] == t analytically)
// The reference evaluates this with ~30 tiny torch kernels and a Python loop per sub-frame and
// differentiates it with autograd; here eight threads per sub-frame evaluate it in fp64 with
// forward-mode dual numbers (one tangent direction per thread: the six se(3) components and the curve
// parameter), which yields the Jacobian the backward pass needs without any tape.  The kernels are pure
// latency (a few hundred dependent fp64 operations): spreading the tangents and the Bernstein weights over
// threads is what shortens them (0.030 + 0.047 ms -> see DESIGN.md at F = 16).
#include "dgs_b200.h"
#include "dgs_internal.cuh"
#include <string>

namespace dgs {

// One tangent per thread: the 7 tangent directions of a sub-frame (6 se(3) components + the curve parameter nu, whose
// seed is d se3 / d nu) are carried by 7 threads that each repeat the (cheap) value arithmetic -- the same fp64 operations in
// the same order on every thread, so the values do not depend on the split.
struct Dual {
    double v;
    double d;
};
__device__ __forceinline__ Dual dconst(double c) { return Dual{c, 0.0}; }
__device__ __forceinline__ Dual operator+(const Dual& a, const Dual& b) { return Dual{a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(const Dual& a, const Dual& b) { return Dual{a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator-(const Dual& a) { return Dual{-a.v, -a.d}; }
__device__ __forceinline__ Dual operator*(const Dual& a, const Dual& b) { return Dual{a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(const Dual& a, const Dual& b)
{
    const double inv = 1.0 / b.v, q = a.v * inv;
    return Dual{q, (a.d - q * b.d) * inv};
}
__device__ __forceinline__ Dual dsqrt(const Dual& a)
{
    const double r = sqrt(a.v);
    return Dual{r, a.d * (0.5 / r)};
}
__device__ __forceinline__ void dsincos(const Dual& a, Dual& s, Dual& c)
{
    double sv, cv;
    sincos(a.v, &sv, &cv);
    s = Dual{sv, a.d * cv};
    c = Dual{cv, a.d * -sv};
}

__device__ __forceinline__ double binom(int n, int k)
{
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (double)(n - k + i) / (double)i;
    return rint(r);
}
__device__ __forceinline__ double ipow(double x, int n)      // x^n, n >= 0 (square and multiply)
{
    double r = 1.0;
    while (n > 0) {
        if (n & 1) r *= x;
        x *= x;
        n >>= 1;
    }
    return r;
}

// Bernstein weight of control point k at t (powers and their product in fp32 like the reference, times an fp64 binomial)
__device__ __forceinline__ double bezier_coeff(float t, int C, int k, float& a, float& b, double& bn)
{
    a = powf(t, (float)(C - k));
    b = powf(1.0f - t, (float)k);
    bn = binom(C, k);
    return (double)(a * b) * bn;
}
// ... and its derivative w.r.t. t
__device__ __forceinline__ double bezier_dcoeff(float t, int C, int k, float a, float b, double bn)
{
    const double td = (double)t, omt = (double)(1.0f - t);
    const double da = (C - k) > 0 ? (double)(C - k) * ipow(td, C - k - 1) : 0.0;
    const double db = k > 0 ? -(double)k * ipow(omt, k - 1) : 0.0;
    return bn * (da * (double)b + (double)a * db);
}

#define POSE_ROWS 35
#define POSE_COLS 7
#define POSE_LANES 8          // threads per sub-frame (7 tangents + 1 idle): a power of two for the width-limited shuffles
#define POSE_FWD_THREADS 64

// thread (s, d): sub-frame s, tangent d.
__global__ void __launch_bounds__(POSE_FWD_THREADS) k_pose_forward(int F, int C, const float* __restrict__ ctrl_trans,
                               const float* __restrict__ ctrl_rot, const float* __restrict__ nu,
                               const float* __restrict__ proj_t, float* __restrict__ view,
                               float* __restrict__ proj, float* __restrict__ campos,
                               double* __restrict__ jac)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int s_raw = gt / POSE_LANES, d = gt % POSE_LANES;
    const bool live = s_raw < F;
    const int s = live ? s_raw : F - 1;         // (every lane takes part in the shuffles)
    const float t = nu[s];
    // Bezier sample: the lanes of a sub-frame evaluate the Bernstein weights of different control points (the powf /
    // binomial work), then every lane accumulates all of them in control-point order -- the reference's fp64 sum.
    double se3[6] = {0, 0, 0, 0, 0, 0}, dse3_dnu[6] = {0, 0, 0, 0, 0, 0};
    for (int k0 = 0; k0 <= C; k0 += POSE_LANES) {
        double c = 0.0, dc = 0.0;
        if (k0 + d <= C) {
            float a, b; double bn;
            c = bezier_coeff(t, C, k0 + d, a, b, bn);
            dc = bezier_dcoeff(t, C, k0 + d, a, b, bn);
        }
        for (int j = 0; j < POSE_LANES; j++) {
            const double cj = __shfl_sync(0xffffffffu, c, j, POSE_LANES);
            const double dcj = __shfl_sync(0xffffffffu, dc, j, POSE_LANES);
            const int k = k0 + j;
            if (k <= C) {
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double ct = (double)ctrl_trans[3 * k + i], cr = (double)ctrl_rot[3 * k + i];
                    se3[i] += cj * ct;       dse3_dnu[i] += dcj * ct;
                    se3[3 + i] += cj * cr;   dse3_dnu[3 + i] += dcj * cr;
                }
            }
        }
    }
    Dual u[3], w[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        u[i] = Dual{se3[i], d == 6 ? dse3_dnu[i] : (d == i ? 1.0 : 0.0)};
        w[i] = Dual{se3[3 + i], d == 6 ? dse3_dnu[3 + i] : (d == 3 + i ? 1.0 : 0.0)};
    }

    Dual nrms = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (nrms.v < 1e-4) nrms = dconst(1e-4);   // clamp: constant value, zero gradient
    const Dual theta = dsqrt(nrms);
    const Dual one = dconst(1.0);
    const Dual inv = one / theta;
    Dual st, ct;
    dsincos(theta, st, ct);
    const Dual fac1 = inv * st;
    const Dual fac2 = inv * inv * (one - ct);
    const Dual facV1 = (one - ct) / (theta * theta);
    const Dual facV2 = (theta - st) / (theta * theta * theta);

    // K = hat(w), K2 = K K
    Dual K[3][3], K2[3][3];
    const Dual zero = dconst(0.0);
    K[0][0] = zero;  K[0][1] = -w[2]; K[0][2] = w[1];
    K[1][0] = w[2];  K[1][1] = zero;  K[1][2] = -w[0];
    K[2][0] = -w[1]; K[2][1] = w[0];  K[2][2] = zero;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) K2[i][j] = K[i][0] * K[0][j] + K[i][1] * K[1][j] + K[i][2] * K[2][j];
    Dual R[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const Dual eye = dconst(i == j ? 1.0 : 0.0);
            R[i][j] = fac1 * K[i][j] + fac2 * K2[i][j] + eye;
            V[i][j] = eye + K[i][j] * facV1 + K2[i][j] * facV2;
        }
    Dual T[3];
#pragma unroll
    for (int i = 0; i < 3; i++) T[i] = V[i][0] * u[0] + V[i][1] * u[1] + V[i][2] * u[2];

    // wvt (row-major, row-vector convention): [:3,:3] = R, [3,:3] = -T @ R, last column (0,0,0,1)
    Dual wvt[4][4];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) wvt[i][j] = R[i][j];
#pragma unroll
    for (int j = 0; j < 3; j++) wvt[3][j] = (-T[0]) * R[0][j] + (-T[1]) * R[1][j] + (-T[2]) * R[2][j];
#pragma unroll
    for (int i = 0; i < 3; i++) wvt[i][3] = zero;
    wvt[3][3] = one;
    if (!live) return;

    if (d == 0) {
        float wf[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                wf[i][j] = (float)wvt[i][j].v;
                view[16 * s + 4 * i + j] = wf[i][j];
            }
        // full_proj = wvt @ proj_t in fp32 (ascending-k fused multiply-add chain)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) acc = fmaf(wf[i][k], proj_t[4 * k + j], acc);
                proj[16 * s + 4 * i + j] = acc;
            }
#pragma unroll
        for (int i = 0; i < 3; i++) campos[3 * s + i] = (float)T[i].v;
    }

    if (jac != nullptr && d < POSE_COLS) {      // column d of the sub-frame's Jacobian
        double* J = jac + (size_t)s * POSE_ROWS * POSE_COLS + d;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                J[(4 * i + j) * POSE_COLS] = wvt[i][j].d;
                // d full_proj[i][j] = sum_k d wvt[i][k] * proj_t[k][j]
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < 4; k++) a += wvt[i][k].d * (double)proj_t[4 * k + j];
                J[(16 + 4 * i + j) * POSE_COLS] = a;
            }
        for (int e = 32; e < POSE_ROWS; e++) J[e * POSE_COLS] = 0.0;   // campos: no gradient in the reference
    }
}

// One block. dL/dse3[s] = J_s^T [dL/dview_s ; dL/dproj_s]; dL/dctrl[k] = sum_s coeff[s,k] dL/dse3[s].
// Shared memory: acc [F][7] doubles, then (when it fits: `table` != 0) the Bernstein weights [F][C+1].
#define POSE_BWD_THREADS 256
__global__ void __launch_bounds__(POSE_BWD_THREADS) k_pose_backward(int F, int C, int table, const float* __restrict__ nu,
                                const double* __restrict__ jac,
                                const float* __restrict__ dview, const float* __restrict__ dproj,
                                float* __restrict__ dctrl_trans, float* __restrict__ dctrl_rot,
                                float* __restrict__ dnu)
{
    extern __shared__ double sm[];   // [F][7] (+ [F][C+1])
    double* coef = sm + (size_t)F * POSE_COLS;
    for (int i = threadIdx.x; i < F * POSE_COLS; i += blockDim.x) {
        const int s = i / POSE_COLS, d = i % POSE_COLS;
        const double* J = jac + (size_t)s * POSE_ROWS * POSE_COLS + d;
        double acc = 0.0;
        for (int e = 0; e < 32; e++) {
            const double gsrc = e < 16 ? (double)dview[16 * s + e] : (double)dproj[16 * s + e - 16];
            acc += J[e * POSE_COLS] * gsrc;
        }
        sm[i] = acc;
        if (d == 6 && dnu) dnu[s] = (float)acc;
    }
    if (table) {
        for (int i = threadIdx.x; i < F * (C + 1); i += blockDim.x) {
            float a, b; double bn;
            coef[i] = bezier_coeff(nu[i / (C + 1)], C, i % (C + 1), a, b, bn);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (C + 1) * 6; i += blockDim.x) {
        const int k = i / 6, d = i % 6;
        double acc = 0.0;
        for (int s = 0; s < F; s++) {
            double c;
            if (table) {
                c = coef[s * (C + 1) + k];
            } else {
                float a, b; double bn;
                c = bezier_coeff(nu[s], C, k, a, b, bn);
            }
            acc += c * sm[s * POSE_COLS + d];
        }
        if (d < 3) dctrl_trans[3 * k + d] = (float)acc;
        else dctrl_rot[3 * k + d - 3] = (float)acc;
    }
}

}  // namespace dgs

extern "C" {

int dgs_pose_forward(int F, int curve_order, const float* ctrl_trans, const float* ctrl_rot,
                     const float* nu, const float* proj_t, float* viewmatrix, float* projmatrix,
                     float* campos, double* jacobian, void* stream)
{
    if (F <= 0) return DGS_OK;
    if (curve_order < 0 || curve_order > 64 || !ctrl_trans || !ctrl_rot || !nu || !proj_t || !viewmatrix ||
        !projmatrix || !campos)
        return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_pose_forward: invalid argument");
    {
        dgs::StageTimer timer(dgs::ST_POSE_FWD, (cudaStream_t)stream, 1);
        const int threads = F * POSE_LANES;
        dgs::k_pose_forward<<<(threads + POSE_FWD_THREADS - 1) / POSE_FWD_THREADS, POSE_FWD_THREADS, 0, (cudaStream_t)stream>>>(
            F, curve_order, ctrl_trans, ctrl_rot, nu, proj_t, viewmatrix, projmatrix, campos, jacobian);
    }
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_pose_forward"); }
}

int dgs_pose_backward(int F, int curve_order, const float* ctrl_trans, const float* ctrl_rot,
                      const float* nu, const double* jacobian, const float* dL_dviewmatrix,
                      const float* dL_dprojmatrix, float* dL_dctrl_trans, float* dL_dctrl_rot,
                      float* dL_dnu, void* stream)
{
    (void)ctrl_trans; (void)ctrl_rot;
    if (curve_order < 0 || curve_order > 64 || F < 0 || F > DGS_MAX_SUBFRAMES) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_pose_backward: invalid argument");
    if (!dL_dctrl_trans || !dL_dctrl_rot) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_pose_backward: invalid argument");
    if (F > 0 && (!nu || !jacobian || !dL_dviewmatrix || !dL_dprojmatrix)) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_pose_backward: invalid argument");
    {
        dgs::StageTimer timer(dgs::ST_POSE_BWD, (cudaStream_t)stream, 1);
        const size_t acc_bytes = (size_t)(F > 0 ? F : 1) * POSE_COLS * sizeof(double);
        const size_t tab_bytes = (size_t)F * (curve_order + 1) * sizeof(double);
        const int table = acc_bytes + tab_bytes <= 40 * 1024;       // else the weights are re-evaluated in the sum
        dgs::k_pose_backward<<<1, POSE_BWD_THREADS, acc_bytes + (table ? tab_bytes : 0), (cudaStream_t)stream>>>(
            F, curve_order, table, nu, jacobian, dL_dviewmatrix, dL_dprojmatrix, dL_dctrl_trans, dL_dctrl_rot, dL_dnu);
    }
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_pose_backward"); }
}

}  // extern "C"

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
// differentiates it with autograd; here one thread per sub-frame evaluates it in fp64 with
// forward-mode dual numbers (6 tangents), which yields the Jacobian the backward pass needs
// without any tape.
#include "dgs_b200.h"
#include "dgs_internal.cuh"
#include <string>

namespace dgs {

struct Dual {
    double v;
    double d[6];
};
__device__ __forceinline__ Dual dconst(double c)
{
    Dual r; r.v = c;
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = 0.0;
    return r;
}
__device__ __forceinline__ Dual dvar(double c, int idx)
{
    Dual r = dconst(c); r.d[idx] = 1.0; return r;
}
__device__ __forceinline__ Dual operator+(const Dual& a, const Dual& b)
{
    Dual r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
__device__ __forceinline__ Dual operator-(const Dual& a, const Dual& b)
{
    Dual r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
__device__ __forceinline__ Dual operator-(const Dual& a)
{
    Dual r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = -a.d[i];
    return r;
}
__device__ __forceinline__ Dual operator*(const Dual& a, const Dual& b)
{
    Dual r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
__device__ __forceinline__ Dual operator/(const Dual& a, const Dual& b)
{
    Dual r; const double inv = 1.0 / b.v; r.v = a.v * inv;
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
__device__ __forceinline__ Dual dsqrt(const Dual& a)
{
    Dual r; r.v = sqrt(a.v); const double k = 0.5 / r.v;
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] * k;
    return r;
}
__device__ __forceinline__ Dual dsin(const Dual& a)
{
    Dual r; r.v = sin(a.v); const double k = cos(a.v);
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] * k;
    return r;
}
__device__ __forceinline__ Dual dcos(const Dual& a)
{
    Dual r; r.v = cos(a.v); const double k = -sin(a.v);
#pragma unroll
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] * k;
    return r;
}

__device__ __forceinline__ double binom(int n, int k)
{
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (double)(n - k + i) / (double)i;
    return rint(r);
}

// Bernstein weight of control point k at t, and its derivative w.r.t. t.
__device__ __forceinline__ void bezier_coeff(float t, int C, int k, double& coeff, double& dcoeff)
{
    const float a = powf(t, (float)(C - k));
    const float b = powf(1.0f - t, (float)k);
    const double bn = binom(C, k);
    coeff = (double)(a * b) * bn;
    const double td = (double)t, omt = (double)(1.0f - t);
    double da = (C - k) > 0 ? (double)(C - k) * pow(td, (double)(C - k - 1)) : 0.0;
    double db = k > 0 ? -(double)k * pow(omt, (double)(k - 1)) : 0.0;
    dcoeff = bn * (da * (double)b + (double)a * db);
}

#define POSE_ROWS 35
#define POSE_COLS 7

__global__ void k_pose_forward(int F, int C, const float* __restrict__ ctrl_trans,
                               const float* __restrict__ ctrl_rot, const float* __restrict__ nu,
                               const float* __restrict__ proj_t, float* __restrict__ view,
                               float* __restrict__ proj, float* __restrict__ campos,
                               double* __restrict__ jac)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= F) return;
    const float t = nu[s];
    double se3[6] = {0, 0, 0, 0, 0, 0}, dse3_dnu[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k <= C; k++) {
        double c, dc;
        bezier_coeff(t, C, k, c, dc);
        for (int d = 0; d < 3; d++) {
            const double ct = (double)ctrl_trans[3 * k + d], cr = (double)ctrl_rot[3 * k + d];
            se3[d] += c * ct;       dse3_dnu[d] += dc * ct;
            se3[3 + d] += c * cr;   dse3_dnu[3 + d] += dc * cr;
        }
    }
    Dual u[3], w[3];
    for (int d = 0; d < 3; d++) { u[d] = dvar(se3[d], d); w[d] = dvar(se3[3 + d], 3 + d); }

    Dual nrms = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    if (nrms.v < 1e-4) nrms = dconst(1e-4);   // clamp: constant value, zero gradient
    const Dual theta = dsqrt(nrms);
    const Dual one = dconst(1.0);
    const Dual inv = one / theta;
    const Dual st = dsin(theta), ct = dcos(theta);
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
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K2[i][j] = K[i][0] * K[0][j] + K[i][1] * K[1][j] + K[i][2] * K[2][j];
    Dual R[3][3], V[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const Dual eye = dconst(i == j ? 1.0 : 0.0);
            R[i][j] = fac1 * K[i][j] + fac2 * K2[i][j] + eye;
            V[i][j] = eye + K[i][j] * facV1 + K2[i][j] * facV2;
        }
    Dual T[3];
    for (int i = 0; i < 3; i++) T[i] = V[i][0] * u[0] + V[i][1] * u[1] + V[i][2] * u[2];

    // wvt (row-major, row-vector convention): [:3,:3] = R, [3,:3] = -T @ R, last column (0,0,0,1)
    Dual wvt[4][4];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) wvt[i][j] = R[i][j];
    for (int j = 0; j < 3; j++) wvt[3][j] = (-T[0]) * R[0][j] + (-T[1]) * R[1][j] + (-T[2]) * R[2][j];
    for (int i = 0; i < 3; i++) wvt[i][3] = zero;
    wvt[3][3] = one;

    float wf[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            wf[i][j] = (float)wvt[i][j].v;
            view[16 * s + 4 * i + j] = wf[i][j];
        }
    // full_proj = wvt @ proj_t in fp32 (ascending-k fused multiply-add chain)
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float acc = 0.f;
            for (int k = 0; k < 4; k++) acc = fmaf(wf[i][k], proj_t[4 * k + j], acc);
            proj[16 * s + 4 * i + j] = acc;
        }
    for (int d = 0; d < 3; d++) campos[3 * s + d] = (float)T[d].v;

    if (jac != nullptr) {
        double* J = jac + (size_t)s * POSE_ROWS * POSE_COLS;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                double* row = J + (4 * i + j) * POSE_COLS;
                double dn = 0.0;
                for (int d = 0; d < 6; d++) { row[d] = wvt[i][j].d[d]; dn += wvt[i][j].d[d] * dse3_dnu[d]; }
                row[6] = dn;
                // d full_proj[i][j] = sum_k d wvt[i][k] * proj_t[k][j]
                double* prow = J + (16 + 4 * i + j) * POSE_COLS;
                double pn = 0.0;
                for (int d = 0; d < 6; d++) {
                    double a = 0.0;
                    for (int k = 0; k < 4; k++) a += wvt[i][k].d[d] * (double)proj_t[4 * k + j];
                    prow[d] = a;
                    pn += a * dse3_dnu[d];
                }
                prow[6] = pn;
            }
        for (int e = 32; e < POSE_ROWS; e++)
            for (int d = 0; d < POSE_COLS; d++) J[e * POSE_COLS + d] = 0.0;   // campos: no gradient in the reference
    }
}

// One block. dL/dse3[s] = J_s^T [dL/dview_s ; dL/dproj_s]; dL/dctrl[k] = sum_s coeff[s,k] dL/dse3[s].
__global__ void k_pose_backward(int F, int C, const float* __restrict__ nu, const double* __restrict__ jac,
                                const float* __restrict__ dview, const float* __restrict__ dproj,
                                float* __restrict__ dctrl_trans, float* __restrict__ dctrl_rot,
                                float* __restrict__ dnu)
{
    extern __shared__ double sm[];   // [F][7]
    for (int s = threadIdx.x; s < F; s += blockDim.x) {
        const double* J = jac + (size_t)s * POSE_ROWS * POSE_COLS;
        double acc[POSE_COLS];
        for (int d = 0; d < POSE_COLS; d++) acc[d] = 0.0;
        for (int e = 0; e < 32; e++) {
            const double gsrc = e < 16 ? (double)dview[16 * s + e] : (double)dproj[16 * s + e - 16];
            for (int d = 0; d < POSE_COLS; d++) acc[d] += J[e * POSE_COLS + d] * gsrc;
        }
        for (int d = 0; d < POSE_COLS; d++) sm[s * POSE_COLS + d] = acc[d];
        if (dnu) dnu[s] = (float)acc[6];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (C + 1) * 6; i += blockDim.x) {
        const int k = i / 6, d = i % 6;
        double a = 0.0;
        for (int s = 0; s < F; s++) {
            double c, dc;
            bezier_coeff(nu[s], C, k, c, dc);
            a += c * sm[s * POSE_COLS + d];
        }
        if (d < 3) dctrl_trans[3 * k + d] = (float)a;
        else dctrl_rot[3 * k + d - 3] = (float)a;
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
        dgs::k_pose_forward<<<(F + 63) / 64, 64, 0, (cudaStream_t)stream>>>(F, curve_order, ctrl_trans, ctrl_rot, nu,
                                                                            proj_t, viewmatrix, projmatrix, campos,
                                                                            jacobian);
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
        dgs::k_pose_backward<<<1, 128, (size_t)(F > 0 ? F : 1) * POSE_COLS * sizeof(double), (cudaStream_t)stream>>>(
            F, curve_order, nu, jacobian, dL_dviewmatrix, dL_dprojmatrix, dL_dctrl_trans, dL_dctrl_rot, dL_dnu);
    }
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_pose_backward"); }
}

}  // extern "C"

// Backward kernels of the batched blurry-view rasterizer (sm_100a).
//
// Reference behaviour restated here (taekkii/deblurgs, submodules/diff-gaussian-rasterization):
//   tile blending backward    cuda_rasterizer/backward.cu:463-640
//   cov2D backward            cuda_rasterizer/backward.cu:145-295
//   preprocess / SH / cov3D   cuda_rasterizer/backward.cu:20-140, 299-362, 367-460
// including its gradient conventions ("quirks"): mean2D gradients
// are w.r.t. NDC; the view-matrix gradient flows through the view-space mean t and the
// depth only; the projection-matrix gradient carries the extra 0.5*W / 0.5*H factor and one
// shared last-row term; quaternion and scale gradients ignore normalisation / scale_modifier.
//
// Design differences (B200-first): the reference issues 10 global float atomics per
// contributing (pixel, Gaussian) pair.  Here the blend backward queues, per warp, the two
// per-pixel scalars of every contributing list entry in shared memory and lets each lane
// accumulate one queued entry's gradient moments over the warp's pixels (k_render_bwd below),
// ending in three 16-B vector REDs per (warp, Gaussian) into a 48-B AoS gradient record.  The
// per-Gaussian stage is two kernels split by register budget, not by the reference's stages --
// a geometry half (cov2D, projection, cov3D, pose gradients, densification statistics) and an
// SH half -- whose threads loop over the F sub-frames and keep the Gaussian's parameter
// gradients in registers, writing them once per blurry view; view/projection-matrix gradients
// (21 values per sub-frame) are reduced over the warp with a 23-shuffle halving exchange (each
// step trades half of the remaining components with the partner lane) and accumulated in fp64.
#include "dgs_internal.cuh"

namespace dgs {

#define FULL_MASK 0xffffffffu

// Halving exchange: reduce N per-lane values over the 32 lanes of a warp.  After the call,
// out holds the total of component `warp_reduce_owner_component(lane)` (or garbage when that
// is < 0).  Costs ceil(N/2)+ceil(N/4)+... shuffles instead of 5*N.
template <int N>
struct HalvingReduce {
    // one exchange step over distance DIST on an array of CNT live values
    template <int CNT, int DIST>
    static __device__ __forceinline__ void step(float (&v)[N], unsigned lane)
    {
        constexpr int KEEP = (CNT + 1) / 2;  // lower half keeps [0,KEEP), upper keeps [KEEP,CNT)
        const bool up = (lane & DIST) != 0;
#pragma unroll
        for (int i = 0; i < KEEP; i++) {
            const float hi = (i + KEEP < CNT) ? v[i + KEEP] : 0.0f;
            const float send = up ? v[i] : hi;
            const float keep = up ? hi : v[i];
            v[i] = keep + __shfl_xor_sync(FULL_MASK, send, DIST);
        }
    }
    static __device__ __forceinline__ float run(float (&v)[N], unsigned lane)
    {
        constexpr int C1 = N, C2 = (C1 + 1) / 2, C3 = (C2 + 1) / 2, C4 = (C3 + 1) / 2, C5 = (C4 + 1) / 2;
        if (C1 > 1) step<C1, 16>(v, lane); else v[0] += __shfl_xor_sync(FULL_MASK, v[0], 16);
        if (C2 > 1) step<C2, 8>(v, lane); else v[0] += __shfl_xor_sync(FULL_MASK, v[0], 8);
        if (C3 > 1) step<C3, 4>(v, lane); else v[0] += __shfl_xor_sync(FULL_MASK, v[0], 4);
        if (C4 > 1) step<C4, 2>(v, lane); else v[0] += __shfl_xor_sync(FULL_MASK, v[0], 2);
        if (C5 > 1) step<C5, 1>(v, lane); else v[0] += __shfl_xor_sync(FULL_MASK, v[0], 1);
        return v[0];
    }
    // which component's total does `lane` hold after run()?  (-1: none / padding)
    static __device__ __forceinline__ int owner(unsigned lane)
    {
        int base = 0, live = N, cnt = N;  // cnt follows the compile-time sequence used by run()
#pragma unroll
        for (int dist = 16; dist >= 1; dist >>= 1) {
            if (cnt > 1) {
                const int keep = (cnt + 1) / 2;
                if (lane & dist) { base += keep; live -= keep; } else { live = min(live, keep); }
                if (live <= 0) return -1;   // this lane ended up on zero padding
                cnt = keep;
            } else if (lane & dist) {
                return -1;  // duplicate holder; only the lane with the bit clear reports
            }
        }
        return base;
    }
};

// ---------------------------------------------------------------------------------------
// tile blending, backward.  grid = (tiles_x, tiles_y * 8/BWD_WARPS, F), BWD_THREADS threads = one 16 x 8 strip of a
// tile; each warp owns an 8x4 pixel rectangle -- one half (rows y..y+3 of the 8 columns) of a forward warp's 8x8 block -- and
// works on its own: there is no block-level barrier in this kernel.
// gradient scratch = three planes of N float4:
//   g0: dmean2D(x,y), dconic(x,y) | g1: dconic(w), dopacity, ddepth, - | g2: dcolor(r,g,b), -
// (the geometry half of the per-Gaussian backward reads g0 and g1, the colour half g2: every sector it touches is
// fully used)
//
// Which entries: the forward's hand-over byte says exactly which list entries a warp blended.  Per step a warp reads 32
// hand-over bytes + list ids (coalesced, back to front), and only the lanes whose entry it blended gather that entry's
// three 16-B records -- straight into the warp's own shared-memory slots with cp.async (no registers, no barrier),
// compacted back to front.  The gather of step k+1 and the byte / id loads of step k+2 are in flight while step k is
// replayed.  (Round 2's first version staged EVERY entry of the tile list for a 4-warp strip behind two
// __syncthreads per 256 entries: 16 % of the warp cycles waited at those barriers, ~8x more records were gathered
// than used, and the replay depth was the strip's instead of the warp's: 2.17 -> 1.79 ms at c2.)
//
// Gradient accumulation is split in two phases so that no per-entry cross-lane reduction is
// needed (the reference issues 10 atomics per contributing pixel; a per-entry warp reduction
// costs ~50 shuffle/select/add instructions per (warp, entry)):
//   phase 1 (lane = pixel): replay the gathered entries back to front (same pairs, same alpha as the forward;
//           dL/dalpha from un-normalised suffix sums, see the loop) and append ONE column per entry to a per-warp
//           queue in shared memory: per pixel the two scalars w1 = opacity*G*dL/dalpha and w2 = alpha*T
//           (zeros where the pixel does not contribute), plus the entry's Gaussian index.
//   phase 2 (lane = queued entry x pixel row, when BWD_QN entries are queued): accumulates the moments
//           sum w1*{1,px,px^2} about the row's first pixel and sum w2*dL/dpix{r,g,b,depth} in registers with every
//           lane busy, shifts the moments to the Gaussian's centre (= sum w1*{1,dx,dy,dx^2,dx*dy,dy^2}), sums the
//           four rows, converts to the 10 gradient components and issues three 16-B vector REDs.
// The sums are the reference's, regrouped.
// ---------------------------------------------------------------------------------------
#define BWD_WARPS 4               // warps per block: a block owns a 16 x (4*BWD_WARPS/2) strip of a tile
#define BWD_THREADS (32 * BWD_WARPS)
#define BWD_QN 8                  // queued entries per flush
#define BWD_QSTRIDE 33            // float2 row stride of the queue (bank-conflict-free both ways)

struct __align__(16) BwdWarpSmem {
    float4 rec[2][32][3];                 // [buffer][compacted entry] -> geo0 | geo1 | geo2 records (cp.async targets)
    uint32_t idr[2][32];                  // Gaussian index | (list position relative to the step's lowest) << 27
    float2 qw[BWD_QN][BWD_QSTRIDE];       // [queued entry][pixel] -> (w1, w2)
    float4 dpix[32];                      // dL/dpix r,g,b,depth of the warp's 32 pixels
};
#define BWD_ID_BITS 27
#define BWD_ID_MASK ((1u << BWD_ID_BITS) - 1u)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// Phase 2.  Lane (e, quarter) owns queued entry e and the 8 pixels of row `quarter` of the warp's 8x4
// rectangle (8 queued entries x 4 rows = 32 busy lanes).  The weighted moments are accumulated about the row's first
// pixel with the pixel offsets as compile-time constants (sum w, sum w*px, sum w*px^2), and shifted to the
// Gaussian's centre afterwards: the entry's own data (centre, conic, opacity) is therefore not needed
// until after the loop, so it is simply re-read from the geometry records (an L1/L2 hit: the gather
// fetched it moments ago) while the loop runs, instead of being copied into the queue by phase 1.
template <bool DEPTH>
__device__ __forceinline__ void bwd_flush(BwdWarpSmem& sm, unsigned lane, int qn, uint32_t myqid, float wx0f, float wy0f,
                                            float ddelx_dx, float ddely_dy, const float4* __restrict__ geo0,
                                            const float4* __restrict__ geo1, float4* __restrict__ gp0,
                                            float4* __restrict__ gp1, float4* __restrict__ gp2)
{
    __syncwarp();
    const unsigned e = lane & 7u, quarter = lane >> 3;
    const bool live = (int)e < qn;
    // Gaussian index of queued entry e: lane e kept it in a register when the entry was queued (a shared-memory
    // array costs one store wavefront per entry, and this kernel is bound by the L1 / shared-memory data pipe)
    const uint32_t id = __shfl_sync(FULL_MASK, myqid, e) & BWD_ID_MASK;
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
    if (live) {
        g0 = __ldg(geo0 + id);
        g1 = __ldg(geo1 + id);
    }
    float M0 = 0.f, Mx = 0.f, Mxx = 0.f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Cd = 0.f;
#pragma unroll
    for (int pp = 0; pp < 8; pp++) {
        const float2 w = sm.qw[e][quarter * 8 + pp];
        const float px = (float)pp;
        M0 += w.x;
        Mx = fmaf(w.x, px, Mx);
        Mxx = fmaf(w.x, px * px, Mxx);
        if (DEPTH) {
            const float4 dp = sm.dpix[quarter * 8 + pp];
            Cr = fmaf(w.y, dp.x, Cr); Cg = fmaf(w.y, dp.y, Cg); Cb = fmaf(w.y, dp.z, Cb); Cd = fmaf(w.y, dp.w, Cd);
        } else {
            // no depth gradient: 12 of the 16 bytes (an 8-byte + a 4-byte load: 3 shared-memory wavefronts instead of 4)
            const float* dp = reinterpret_cast<const float*>(&sm.dpix[quarter * 8 + pp]);
            const float2 rg = *reinterpret_cast<const float2*>(dp);
            Cr = fmaf(w.y, rg.x, Cr); Cg = fmaf(w.y, rg.y, Cg); Cb = fmaf(w.y, dp[2], Cb);
        }
    }
    const float bx = g0.x - wx0f, by = g0.y - (wy0f + (float)quarter);
    float S0 = M0;
    float Sx = bx * M0 - Mx;
    float Sy = by * M0;
    float Sxx = bx * (bx * M0 - 2.0f * Mx) + Mxx;
    float Sxy = by * Sx;
    float Syy = by * Sy;
#pragma unroll
    for (int d = 8; d <= 16; d <<= 1) {
        S0 += __shfl_xor_sync(FULL_MASK, S0, d);   Sx += __shfl_xor_sync(FULL_MASK, Sx, d);
        Sy += __shfl_xor_sync(FULL_MASK, Sy, d);   Sxx += __shfl_xor_sync(FULL_MASK, Sxx, d);
        Sxy += __shfl_xor_sync(FULL_MASK, Sxy, d); Syy += __shfl_xor_sync(FULL_MASK, Syy, d);
        Cr += __shfl_xor_sync(FULL_MASK, Cr, d);   Cg += __shfl_xor_sync(FULL_MASK, Cg, d);
        Cb += __shfl_xor_sync(FULL_MASK, Cb, d);
        if (DEPTH) Cd += __shfl_xor_sync(FULL_MASK, Cd, d);
    }
    if (quarter == 0 && live) {
        const float A = g1.x, B = g1.y, Cc = g1.z, o = g1.w;
        red_add_v4(reinterpret_cast<float*>(gp0 + id), -(A * Sx + B * Sy) * ddelx_dx, -(Cc * Sy + B * Sx) * ddely_dy,
                   -0.5f * Sxx, -0.5f * Sxy);
        red_add_v4(reinterpret_cast<float*>(gp1 + id), -0.5f * Syy, o != 0.f ? S0 / o : 0.f, Cd, 0.f);
        red_add_v4(reinterpret_cast<float*>(gp2 + id), Cr, Cg, Cb, 0.f);
    }
    __syncwarp();
}

// DEPTH: a gradient of the depth image is supplied.  Without one (the photometric losses of train.py) the depth
// of an entry is not read, its record costs 5 instead of 6 shared-memory wavefronts per replay, and the flush
// skips the fourth dL/dpix component -- this kernel is bound by the L1 / shared-memory data pipe (ncu:
// l1tex__data_pipe_lsu_wavefronts at 95 % of peak), so wavefronts, not instructions, are what it pays for.
template <bool DEPTH>
__global__ void __launch_bounds__(BWD_THREADS) k_render_bwd(const BwdParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FwdParams& f = p.f;
    const int s = blockIdx.z;
    constexpr unsigned BWD_STRIPS = 8 / BWD_WARPS, STRIP_H = DGS_TILE_Y / BWD_STRIPS;
    const unsigned tile_y = blockIdx.y / BWD_STRIPS, strip = blockIdx.y % BWD_STRIPS;
    const int tile = tile_y * f.tiles_x + blockIdx.x;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BwdWarpSmem& sm = reinterpret_cast<BwdWarpSmem*>(smem_raw)[warp];
    const unsigned wx0 = blockIdx.x * DGS_TILE_X + (warp & 1) * 8;
    const unsigned wy0 = tile_y * DGS_TILE_Y + strip * STRIP_H + (warp >> 1) * 4;
    const unsigned pixx = wx0 + (lane & 7), pixy = wy0 + (lane >> 3);
    const bool inside = pixx < (unsigned)f.W && pixy < (unsigned)f.H;
    const size_t HW = (size_t)f.H * f.W;
    const size_t pix_id = (size_t)f.W * pixy + pixx;
    const float pixfx = (float)pixx, pixfy = (float)pixy;

    const uint2 range = decode_range(p.ranges[(size_t)s * f.tiles_x * f.tiles_y + tile]);
    const unsigned wbit = strip * BWD_WARPS + warp;      // this rectangle's bit in the hand-over byte: 2 * (4-row band) + column half

    const float T_final = inside ? p.final_T[(size_t)s * HW + pix_id] : 0.f;
    const int last_contributor = inside ? (int)p.n_contrib[(size_t)s * HW + pix_id] : 0;
    // nothing behind the warp's deepest contributor can receive gradient from it
    const int list_len = min((int)(range.y - range.x), __reduce_max_sync(FULL_MASK, last_contributor));
    if (list_len <= 0) return;      // warp-uniform; the kernel has no block-level barrier

    float dpix0 = 0.f, dpix1 = 0.f, dpix2 = 0.f, dpixd = 0.f;
    if (inside) {
        if (p.dL_dpix) {
            const float* d = p.dL_dpix + (size_t)s * 3 * HW;
            dpix0 = d[pix_id]; dpix1 = d[HW + pix_id]; dpix2 = d[2 * HW + pix_id];
        }
        if (DEPTH) dpixd = p.dL_dpixdepth[(size_t)s * HW + pix_id];
        if (p.dL_dblur) {   // backward of blurred = sum_s color_s / denominator
            dpix0 += p.dL_dblur[pix_id] / p.blur_denominator;
            dpix1 += p.dL_dblur[HW + pix_id] / p.blur_denominator;
            dpix2 += p.dL_dblur[2 * HW + pix_id] / p.blur_denominator;
        }
    }
    sm.dpix[lane] = make_float4(dpix0, dpix1, dpix2, dpixd);
    float bg_dot_dpixel = 0.f;
    bg_dot_dpixel += f.background[0] * dpix0;
    bg_dot_dpixel += f.background[1] * dpix1;
    bg_dot_dpixel += f.background[2] * dpix2;
    bg_dot_dpixel += f.z_far * dpixd;
    float T = T_final;
    float R = T_final * bg_dot_dpixel;      // suffix sum of the replay (see `replay` below): behind the last contributor there is only the background
    const float ddelx_dx = 0.5f * f.W, ddely_dy = 0.5f * f.H;
    const float rx0 = (float)wx0, ry0 = (float)wy0;

    const float4* __restrict__ geo0 = f.geo0 + (size_t)s * f.P;
    const float4* __restrict__ geo1 = f.geo1 + (size_t)s * f.P;
    const float4* __restrict__ geo2 = f.geo2 + (size_t)s * f.P;
    float4* __restrict__ gp0 = p.g0 + (size_t)s * f.P;
    float4* __restrict__ gp1 = p.g1 + (size_t)s * f.P;
    float4* __restrict__ gp2 = p.g2 + (size_t)s * f.P;
    const uint32_t* __restrict__ plist = p.point_list + range.x;
    const uint8_t* __restrict__ wmask =
        reinterpret_cast<const uint8_t*>(p.bin_header) + p.bin_header->wmask_offset + range.x;

    uint32_t a_rec, a_idr;
    asm volatile("mov.u32 %0, %1;" : "=r"(a_rec) : "r"(smem_addr(sm.rec)));
    asm volatile("mov.u32 %0, %1;" : "=r"(a_idr) : "r"(smem_addr(sm.idr)));

    // Step k covers the list positions hi_k - lane, hi_k = list_len - 1 - 32 k (back to front); r = 31 - lane is the
    // position relative to the step's lowest one, low_k = hi_k - 31.  A pixel replays the entries in front of its last
    // contributor: low_k + r < last_contributor  <=>  r < rel_k, rel_k = last_contributor - low_k (grows by 32 per step).
    const int steps = (list_len + 31) >> 5;
    int rel = last_contributor - (list_len - 32);
    auto load_step = [&](int k, uint32_t& wmb, uint32_t& id) {
        const int ps = list_len - 1 - 32 * k - (int)lane;
        wmb = 0u; id = 0u;
        if (ps >= 0 && k < steps) { wmb = wmask[ps]; id = plist[ps]; }
    };
    auto issue_gather = [&](int buf, uint32_t wmb, uint32_t id) -> int {
        const bool keep = ((wmb >> wbit) & 1u) != 0u;
        const unsigned mask = __ballot_sync(FULL_MASK, keep);
        if (keep) {
            const unsigned rank = __popc(mask & ((1u << lane) - 1u));
            const uint32_t dst = a_rec + (uint32_t)buf * (32u * 48u) + rank * 48u;
            cp_async16(dst, geo0 + id);
            cp_async16(dst + 16u, geo1 + id);
            cp_async16(dst + 32u, geo2 + id);
            sts_u32(a_idr + (uint32_t)buf * 128u + rank * 4u, id | ((31u - lane) << BWD_ID_BITS));
        }
        cp_async_commit();
        return __popc(mask);
    };

    uint32_t wmb, idn;
    load_step(0, wmb, idn);
    int n_cur = issue_gather(0, wmb, idn);
    load_step(1, wmb, idn);
    int qn = 0;   // queued entries (warp-uniform)
    uint32_t myqid = 0;   // lane q: id of queued entry q
    for (int k = 0; k < steps; k++, rel += 32) {
        const int n_next = issue_gather((k + 1) & 1, wmb, idn);    // (past the end: no copies, an empty group)
        load_step(k + 2, wmb, idn);
        cp_async_wait<1>();
        __syncwarp();
        uint32_t rb = a_rec + (uint32_t)(k & 1) * (32u * 48u), ib = a_idr + (uint32_t)(k & 1) * 128u;
        // The body is branch-free: a hit entry has at least one contributing lane (the forward said so), so skipping
        // on a warp-wide "nobody contributes" never happens and per-lane decisions are selects.  A lane that does not
        // contribute carries G = alpha = 0 through the same arithmetic (1 / (1 - 0) = 1: T, R unchanged, w1 = w2 = 0).
        // Two entries per iteration: their loads, exponents and colour dot products are independent and overlap; only
        // the T / R recurrences are sequential.
        auto weigh = [&](uint32_t idr, const float2& xy, const float4& con_o, float& G, float& alpha) {
            const float dx = xy.x - pixfx, dy = xy.y - pixfy;
            // Exponent and exp() are the forward's (= the reference's), operation for operation: the replay must
            // classify every pair exactly as the forward did.  One pair whose alpha falls on the other side of 1/255
            // injects a bogus w * (c . dL/dpix) into R and moves dL/dalpha of EVERY entry in front of it at that pixel
            // by up to ~10 % (measured with ex2.approx: ~100 such pixels per c2 view, for 0.14 ms) -- not worth it.
            const float power = -0.5f * (con_o.x * dx * dx + con_o.z * dy * dy) - con_o.y * dx * dy;
            const float Gx = expf(power);
            const float ax = min(0.99f, con_o.w * Gx);
            const bool ok = (int)(idr >> BWD_ID_BITS) < rel && power <= 0.0f && ax >= 1.0f / 255.0f;
            G = ok ? Gx : 0.f;
            alpha = ok ? ax : 0.f;
        };
        // The reference replays T and the colour accumulated BEHIND the entry as normalised recurrences (accum_rec,
        // backward.cu:585-600) and divides twice by (1 - alpha).  The same derivative written on un-normalised sums
        // needs one dot product and one reciprocal: with T_i the transmittance in front of entry i, w_j = alpha_j T_j
        // and R_i = sum_{j behind i} w_j (c_j . dL/dpix) + T_final (bg . dL/dpix),
        //   dL/dalpha_i = T_i (c_i . dL/dpix) - R_i / (1 - alpha_i),   R_{i-1} = R_i + w_i (c_i . dL/dpix).
        // 1 - alpha is in [0.01, 1]: the approximate reciprocal (1 ulp) is far inside the 1e-3 gradient tolerance.
        auto replay = [&](float G, float alpha, float opac, float cdot, float& w1, float& w2) {
            const float inv_1ma = rcp_approx(1.f - alpha);
            T = T * inv_1ma;
            w2 = alpha * T;
            const float dL_dalpha = T * cdot - R * inv_1ma;
            R = fmaf(w2, cdot, R);
            w1 = opac * G * dL_dalpha;
        };
        auto enqueue = [&](float w1, float w2, uint32_t idr) {
            sm.qw[qn][lane] = make_float2(w1, w2);
            if ((int)lane == qn) myqid = idr;            // lane qn remembers the entry's id; the flush strips the position bits
            if (++qn == BWD_QN) {
                bwd_flush<DEPTH>(sm, lane, qn, myqid, rx0, ry0, ddelx_dx, ddely_dy, geo0, geo1, gp0, gp1, gp2);
                qn = 0;
            }
        };
        // centre (+ depth when its gradient is wanted) of the entry whose record starts at byte OFF
        auto cdot_of = [&](const float4& c, float depth) {
            float d = c.x * dpix0 + c.y * dpix1 + c.z * dpix2;
            if (DEPTH) d += depth * dpixd;
            return d;
        };
        int j = 0;
        for (; j + 2 <= n_cur; j += 2, rb += 96u, ib += 8u) {
            const uint32_t idrA = lds_u32(ib), idrB = lds_u32(ib + 4u);
            float2 xyA, xyB;
            float zA = 0.f, zB = 0.f;
            if (DEPTH) {
                const float4 g0A = lds_f4_off<0>(rb), g0B = lds_f4_off<48>(rb);
                xyA = make_float2(g0A.x, g0A.y); zA = g0A.z;
                xyB = make_float2(g0B.x, g0B.y); zB = g0B.z;
            } else {
                xyA = lds_f2_off<0>(rb);
                xyB = lds_f2_off<48>(rb);
            }
            const float4 conA = lds_f4_off<16>(rb), cA = lds_f4_off<32>(rb);
            const float4 conB = lds_f4_off<64>(rb), cB = lds_f4_off<80>(rb);
            float GA, aA, GB, aB;
            weigh(idrA, xyA, conA, GA, aA);
            weigh(idrB, xyB, conB, GB, aB);
            const float cdotA = cdot_of(cA, zA);
            const float cdotB = cdot_of(cB, zB);
            float w1A, w2A, w1B, w2B;
            replay(GA, aA, conA.w, cdotA, w1A, w2A);
            replay(GB, aB, conB.w, cdotB, w1B, w2B);
            enqueue(w1A, w2A, idrA);
            enqueue(w1B, w2B, idrB);
        }
        if (j < n_cur) {
            const uint32_t idrA = lds_u32(ib);
            float2 xyA;
            float zA = 0.f;
            if (DEPTH) {
                const float4 g0A = lds_f4_off<0>(rb);
                xyA = make_float2(g0A.x, g0A.y); zA = g0A.z;
            } else {
                xyA = lds_f2_off<0>(rb);
            }
            const float4 conA = lds_f4_off<16>(rb), cA = lds_f4_off<32>(rb);
            float GA, aA, w1A, w2A;
            weigh(idrA, xyA, conA, GA, aA);
            replay(GA, aA, conA.w, cdot_of(cA, zA), w1A, w2A);
            enqueue(w1A, w2A, idrA);
        }
        __syncwarp();      // every lane is done with this buffer before step k+2 is gathered into it
        n_cur = n_next;
    }
    cp_async_wait<0>();
    if (qn > 0) bwd_flush<DEPTH>(sm, lane, qn, myqid, rx0, ry0, ddelx_dx, ddely_dy, geo0, geo1, gp0, gp1, gp2);
}

void launch_render_bwd(const BwdParams& p, cudaStream_t st)
{
    const FwdParams& f = p.f;
    if (f.F == 0 || f.W == 0 || f.H == 0) return;
    static_assert(sizeof(BwdWarpSmem) * BWD_WARPS <= 48 * 1024, "fits the default dynamic shared-memory limit: no per-device opt-in needed");
    dim3 grid(f.tiles_x, f.tiles_y * (8 / BWD_WARPS), f.F), block(BWD_THREADS);
    if (p.dL_dpixdepth != nullptr) k_render_bwd<true><<<grid, block, sizeof(BwdWarpSmem) * BWD_WARPS, st>>>(p);
    else k_render_bwd<false><<<grid, block, sizeof(BwdWarpSmem) * BWD_WARPS, st>>>(p);
}

// ---------------------------------------------------------------------------------------
// per-Gaussian backward: cov2D, projection, SH and cov3D stages for all F sub-frames.
// ---------------------------------------------------------------------------------------
#define NPOSE 21
#define BWD_PF 3    // prefetch depth of the SH half of the per-Gaussian backward (iterations ahead; register-limited)
#define GEO_PF 5    // prefetch depth of the geometry half (106 -> ~125 registers, still 4 blocks of 128 threads per SM)
#define PRE_BWD_THREADS 128
// pose component order: view {0,1,2,4,5,6,8,9,10,12,13,14} -> 0..11, proj {0,1,4,5,8,9,12,13}
// -> 12..19, proj last-row term (entries 3,7,11,15) -> 20
__device__ static const int kPoseSlot[NPOSE] = {0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14,
                                                16, 17, 20, 21, 24, 25, 28, 29, 19};

// Geometry half (cov2D / EWA, projection of the mean, cov3D -> scale / rotation, view / projection-matrix
// gradients, densification statistics).  PRECOMP: colours were supplied precomputed, their gradient
// is a plain sum over sub-frames and rides along here; otherwise the SH half (k_sh_bwd) handles colour.
template <bool PRECOMP>
__global__ void __launch_bounds__(PRE_BWD_THREADS) k_preprocess_bwd(const BwdParams p)
{
    const FwdParams& f = p.f;
    const int g = p.g_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const bool live = g < p.g_end;
    const int gi = live ? g : 0;

    const float3 mean = {f.means3D[3 * gi], f.means3D[3 * gi + 1], f.means3D[3 * gi + 2]};
    float cov3D[6];
    float3 scale = {0.f, 0.f, 0.f};
    float4 q = {0.f, 0.f, 0.f, 0.f};
    if (f.cov3D_precomp != nullptr) {
#pragma unroll
        for (int i = 0; i < 6; i++) cov3D[i] = f.cov3D_precomp[6 * (size_t)gi + i];
    } else {
        q = reinterpret_cast<const float4*>(f.rotations)[gi];
        scale = {f.scales[3 * gi], f.scales[3 * gi + 1], f.scales[3 * gi + 2]};
        cov3d_from_scale_rot(scale.x, scale.y, scale.z, f.scale_modifier, q, cov3D);
    }

    float3 dmean = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dopac = 0.f;
    float3 dcolor_acc = {0.f, 0.f, 0.f};
    float st_norm = 0.f, st_count = 0.f;   // densification statistics (train.py:188-193 of the reference)
    int st_radius = 0;
    const int my_comp = HalvingReduce<NPOSE>::owner(lane);
    const int my_slot = my_comp >= 0 ? kPoseSlot[my_comp] : -1;

    // Software pipeline, GEO_PF iterations deep: the per-(sub-frame, Gaussian) record (radius, 48-B
    // gradient, clamp mask) of iteration s+GEO_PF is requested while iteration s computes.  The kernel
    // is latency-bound (16 warps per SM, 43 % of warp cycles waiting on these loads at depth 3), so the bytes
    // in flight per SM -- warps x lanes x 36 B x depth -- are what sets its speed.  The loads of a stage are
    // independent (the gradient record of an invisible entry is the zero-fill), so no stage waits for its
    // own radius.
    struct Stage { int radius; float4 a, b, c; };
    auto fetch = [&](int s) {
        Stage t;
        const size_t n = (size_t)s * f.P + gi;
        t.radius = live ? f.radii[n] : 0;
        t.a = p.g0[n];
        t.b = p.g1[n];
        t.c = PRECOMP ? p.g2[n] : make_float4(0.f, 0.f, 0.f, 0.f);
        return t;
    };
    Stage pf[GEO_PF];
#pragma unroll
    for (int k = 0; k < GEO_PF; k++) {
        pf[k].radius = 0;
        if (k < f.F) pf[k] = fetch(k);
    }

    for (int s = 0; s < f.F; s++) {
        const size_t n = (size_t)s * f.P + gi;
        const Stage cur = pf[0];
#pragma unroll
        for (int k = 0; k + 1 < GEO_PF; k++) pf[k] = pf[k + 1];
        if (s + GEO_PF < f.F) pf[GEO_PF - 1] = fetch(s + GEO_PF);
        const bool vis = cur.radius > 0;
        const int nx_radius_cur = cur.radius;
        const float4 ga = cur.a, gb = cur.b, gc = cur.c;
        if (p.dL_dmeans2D != nullptr && live) {
            p.dL_dmeans2D[n * 3] = vis ? ga.x : 0.f;
            p.dL_dmeans2D[n * 3 + 1] = vis ? ga.y : 0.f;
            p.dL_dmeans2D[n * 3 + 2] = 0.f;
        }
        if (!__any_sync(FULL_MASK, vis)) continue;
        float pv[NPOSE];
#pragma unroll
        for (int k = 0; k < NPOSE; k++) pv[k] = 0.f;
        if (vis) {
            const float* __restrict__ V = f.view + 16 * s;
            const float* __restrict__ PM = f.proj + 16 * s;
            const float dm2x = ga.x, dm2y = ga.y;
            const float3 dconic = {ga.z, ga.w, gb.x};
            dopac += gb.y;
            st_norm += sqrtf(ga.x * ga.x + ga.y * ga.y);
            st_count += 1.0f;
            st_radius = max(st_radius, nx_radius_cur);
            const float ddepth = gb.z;
            float3 dcol = {gc.x, gc.y, gc.z};

            // ---- EWA projection backward (what the reference's computeCov2DCUDA backward computes, backward.cu:145-295),
            // derived here from the matrix identities instead of the expanded scalar formulas:
            //   Sigma2 = [[a, b], [b, c]] = (J Mv) Sigma3 (J Mv)^T + 0.3 I,   K = Sigma2^-1 = the conic.
            //   The blend backward delivers Gm = dL/dK as a full symmetric matrix ((x, y; y, w): the off-diagonal
            //   entry counted once per position).  d(K) = -K d(Sigma2) K  =>  dL/dSigma2 = -K Gm K, and b occupies both
            //   off-diagonal positions, so dL/db is twice that entry.
            const Ewa e = ewa_project(mean, f.focal_x, f.focal_y, f.tan_fovx, f.tan_fovy, cov3D, V);
            const float inv_det = 1.0f / (e.a * e.c - e.b * e.b);        // Sigma2 >= 0.3 I: never singular
            const float k00 = e.c * inv_det, k01 = -e.b * inv_det, k11 = e.a * inv_det;
            const float kg00 = k00 * dconic.x + k01 * dconic.y, kg01 = k00 * dconic.y + k01 * dconic.z;   // K Gm, row 0
            const float kg10 = k01 * dconic.x + k11 * dconic.y, kg11 = k01 * dconic.y + k11 * dconic.z;   //       row 1
            const float da = -(kg00 * k00 + kg01 * k01);
            const float dc = -(kg10 * k01 + kg11 * k11);
            const float db = -2.0f * (kg00 * k01 + kg01 * k11);
            //   With u, v the two rows of J Mv (the Ewa struct keeps them as T.m[0][.], T.m[1][.]):
            //   a = u.S3.u, b = u.S3.v, c = v.S3.v.  Gradient w.r.t. the six stored entries of the symmetric Sigma3
            //   (an off-diagonal entry stands for two positions, hence the doubled terms there):
            const float u[3] = {e.T.m[0][0], e.T.m[0][1], e.T.m[0][2]};
            const float v[3] = {e.T.m[1][0], e.T.m[1][1], e.T.m[1][2]};
            {
                constexpr int slot[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = i; j < 3; j++) {
                        const float g = u[i] * u[j] * da + 0.5f * (u[i] * v[j] + u[j] * v[i]) * db + v[i] * v[j] * dc;
                        dcov[slot[i][j]] += (i == j) ? g : 2.0f * g;
                    }
            }
            //   Gradient w.r.t. the rows themselves: du = 2 da S3 u + db S3 v,  dv = 2 dc S3 v + db S3 u.
            const float S3[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
            float du[3], dv[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float su = S3[i][0] * u[0] + S3[i][1] * u[1] + S3[i][2] * u[2];
                const float sv = S3[i][0] * v[0] + S3[i][1] * v[1] + S3[i][2] * v[2];
                du[i] = 2.0f * da * su + db * sv;
                dv[i] = 2.0f * dc * sv + db * su;
            }
            //   u = j00 Mv[0] + j02 Mv[2],  v = j11 Mv[1] + j12 Mv[2]  with the perspective Jacobian
            //   j00 = fx / tz, j02 = -fx tx / tz^2, j11 = fy / tz, j12 = -fy ty / tz^2  (Mv rows = e.W.m[k][.]).
            const Mat3& Mv = e.W;
            const float d_j00 = du[0] * Mv.m[0][0] + du[1] * Mv.m[0][1] + du[2] * Mv.m[0][2];
            const float d_j02 = du[0] * Mv.m[2][0] + du[1] * Mv.m[2][1] + du[2] * Mv.m[2][2];
            const float d_j11 = dv[0] * Mv.m[1][0] + dv[1] * Mv.m[1][1] + dv[2] * Mv.m[1][2];
            const float d_j12 = dv[0] * Mv.m[2][0] + dv[1] * Mv.m[2][1] + dv[2] * Mv.m[2][2];
            //   ... and the Jacobian entries depend on the (clamped) view-space mean t: with iz = 1 / tz,
            //   kx = -fx iz^2, ky = -fy iz^2:  d j02/d tx = kx,  d j00/d tz = kx,  d j02/d tz = -2 iz kx tx  (same in y).
            //   Where the forward clamped tx/tz (ty/tz) to the frustum guard band, t no longer follows the mean.
            const float limx = 1.3f * f.tan_fovx, limy = 1.3f * f.tan_fovy;
            const bool free_x = !(e.txtz < -limx || e.txtz > limx), free_y = !(e.tytz < -limy || e.tytz > limy);
            const float iz = 1.0f / e.t.z;
            const float kx = -f.focal_x * iz * iz, ky = -f.focal_y * iz * iz;
            const float dL_dtx = free_x ? kx * d_j02 : 0.f;
            const float dL_dty = free_y ? ky * d_j12 : 0.f;
            const float dL_dtz = kx * d_j00 + ky * d_j11 - 2.0f * iz * (kx * e.t.x * d_j02 + ky * e.t.y * d_j12);
            // dL/dmean through t (V^T applied to the 3-vector)
            float3 dm;
            dm.x = V[0] * dL_dtx + V[1] * dL_dty + V[2] * dL_dtz;
            dm.y = V[4] * dL_dtx + V[5] * dL_dty + V[6] * dL_dtz;
            dm.z = V[8] * dL_dtx + V[9] * dL_dty + V[10] * dL_dtz;
            // view-matrix gradient through t only (reference quirk, backward.cu:279-293)
            pv[0] = dL_dtx * mean.x; pv[1] = dL_dty * mean.x; pv[2] = dL_dtz * mean.x;
            pv[3] = dL_dtx * mean.y; pv[4] = dL_dty * mean.y; pv[5] = dL_dtz * mean.y;
            pv[6] = dL_dtx * mean.z; pv[7] = dL_dty * mean.z; pv[8] = dL_dtz * mean.z;
            pv[9] = dL_dtx; pv[10] = dL_dty; pv[11] = dL_dtz;

            // ---- projection of the mean (reference backward.cu:396-457)
            const float4 m_hom = xform_point_4x4(mean, PM);
            const float m_w = 1.0f / (m_hom.w + 0.0000001f);
            const float mul1 = (PM[0] * mean.x + PM[4] * mean.y + PM[8] * mean.z + PM[12]) * m_w * m_w;
            const float mul2 = (PM[1] * mean.x + PM[5] * mean.y + PM[9] * mean.z + PM[13]) * m_w * m_w;
            dm.x += (PM[0] * m_w - PM[3] * mul1) * dm2x + (PM[1] * m_w - PM[3] * mul2) * dm2y + ddepth * V[2];
            dm.y += (PM[4] * m_w - PM[7] * mul1) * dm2x + (PM[5] * m_w - PM[7] * mul2) * dm2y + ddepth * V[6];
            dm.z += (PM[8] * m_w - PM[11] * mul1) * dm2x + (PM[9] * m_w - PM[11] * mul2) * dm2y + ddepth * V[10];

            const float lastcol = (m_hom.x * f.W * dm2x + m_hom.y * f.H * dm2y) * m_w * m_w;
            const float wx = 0.5f * dm2x * f.W * m_w, wy = 0.5f * dm2y * f.H * m_w;
            pv[12] = wx * mean.x; pv[13] = wy * mean.x;
            pv[14] = wx * mean.y; pv[15] = wy * mean.y;
            pv[16] = wx * mean.z; pv[17] = wy * mean.z;
            pv[18] = wx;          pv[19] = wy;
            pv[20] = -0.5f * lastcol;
            // depth part of the view-matrix gradient (backward.cu:454-457)
            pv[2] += ddepth * mean.x; pv[5] += ddepth * mean.y; pv[8] += ddepth * mean.z; pv[11] += ddepth;

            if (PRECOMP) { dcolor_acc.x += dcol.x; dcolor_acc.y += dcol.y; dcolor_acc.z += dcol.z; }
            dmean.x += dm.x; dmean.y += dm.y; dmean.z += dm.z;
        }
        // pose gradients of this sub-frame: warp reduce, then fp64 atomics
        const float total = HalvingReduce<NPOSE>::run(pv, lane);
        if (my_slot >= 0) atomicAdd(p.pose_acc + (size_t)s * 32 + my_slot, (double)total);
    }
    if (!live) return;

    p.dL_dmeans3D[3 * g] = dmean.x; p.dL_dmeans3D[3 * g + 1] = dmean.y; p.dL_dmeans3D[3 * g + 2] = dmean.z;
    p.dL_dopacity[g] = dopac;
    if (p.densify_stats != nullptr) {
        p.densify_stats[3 * g] = st_norm; p.densify_stats[3 * g + 1] = st_count; p.densify_stats[3 * g + 2] = (float)st_radius;
    }
    if (PRECOMP && p.dL_dcolors_precomp) {
        p.dL_dcolors_precomp[3 * g] = dcolor_acc.x; p.dL_dcolors_precomp[3 * g + 1] = dcolor_acc.y;
        p.dL_dcolors_precomp[3 * g + 2] = dcolor_acc.z;
    }
    if (f.cov3D_precomp != nullptr) {
        if (p.dL_dcov3D_precomp) {
#pragma unroll
            for (int i = 0; i < 6; i++) p.dL_dcov3D_precomp[6 * (size_t)g + i] = dcov[i];
        }
    } else {
        // Sigma3 -> scale / rotation (what the reference's computeCov3D backward computes, backward.cu:299-362);
        // linear in dL/dSigma3, so it is applied once to the sum over the sub-frames.  Derivation:
        //   Sigma3 = Rq S^2 Rq^T,  Rq = rotation of the (un-normalised) quaternion, S = diag(mod * scale).
        //   D = dL/dSigma3 as a full symmetric matrix (stored off-diagonal entries stand for two positions: halve them).
        //   H = 2 D Rq;  dL/dRq = H S^2;  dL/ds_j = s_j sum_i Rq_ij H_ij  (the reference differentiates w.r.t.
        //   s = mod * scale and reports that as the scale gradient; reproduced).
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const float Rq[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float D[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                               {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                               {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
        const float sv[3] = {f.scale_modifier * scale.x, f.scale_modifier * scale.y, f.scale_modifier * scale.z};
        float G[3][3];      // dL/dRq
#pragma unroll
        for (int j = 0; j < 3; j++) {
            float ds = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const float h = 2.0f * (D[i][0] * Rq[0][j] + D[i][1] * Rq[1][j] + D[i][2] * Rq[2][j]);
                ds += Rq[i][j] * h;
                G[i][j] = h * sv[j] * sv[j];
            }
            p.dL_dscales[3 * g + j] = sv[j] * ds;
        }
        // Rq is linear in the products of quaternion components; split G into its antisymmetric part (an axial
        // vector w, which pairs with r) and its symmetric part (which pairs with x, y, z):
        const float wx = G[2][1] - G[1][2], wy = G[0][2] - G[2][0], wz = G[1][0] - G[0][1];
        const float sxy = G[0][1] + G[1][0], sxz = G[0][2] + G[2][0], syz = G[1][2] + G[2][1];
        float4 dq;
        dq.x = 2.f * (x * wx + y * wy + z * wz);
        dq.y = 2.f * (r * wx + y * sxy + z * sxz) - 4.f * x * (G[1][1] + G[2][2]);
        dq.z = 2.f * (r * wy + x * sxy + z * syz) - 4.f * y * (G[0][0] + G[2][2]);
        dq.w = 2.f * (r * wz + x * sxz + y * syz) - 4.f * z * (G[0][0] + G[1][1]);
        reinterpret_cast<float4*>(p.dL_drotations)[g] = dq;
    }
}


// ---------------------------------------------------------------------------------------
// Colour half of the per-Gaussian backward: spherical harmonics (reference backward.cu:20-140).
// Thread = Gaussian, loop over the F sub-frames, dL/dsh accumulated in registers and written once.
// The Gaussian's own coefficients are parked in shared memory (one float4 column per thread) and
// streamed through the loop instead of living in 48 more registers: with the geometry half split
// off, this runs at 4 blocks of 128 threads per SM instead of one 256-thread block, which is what a
// latency-bound kernel needs (the fused version sat at 2 warps per scheduler, 33 % issue slots busy).
// The view-direction gradient uses  dL/ddir = sum_k (sh_k . dL/dRGB) grad b_k(dir): one dot product
// per coefficient instead of three derivative accumulations per channel (same terms, regrouped).
// Adds its dL/dmean contribution to the value the geometry half wrote (stream order).
// ---------------------------------------------------------------------------------------
#define SH_BWD_THREADS 128

template <int DEG>
__global__ void __launch_bounds__(SH_BWD_THREADS, 4) k_sh_bwd(const BwdParams p)
{
    const FwdParams& f = p.f;
    const int g = p.g_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = g < p.g_end;
    const int gi = live ? g : 0;
    constexpr int NC = (DEG + 1) * (DEG + 1);
    constexpr int NV = (3 * NC + 3) / 4;          // float4 slots per Gaussian
    __shared__ float4 s_sh[NV][SH_BWD_THREADS];

    const float3 mean = {f.means3D[3 * gi], f.means3D[3 * gi + 1], f.means3D[3 * gi + 2]};
    // The block's coefficients are one contiguous run of global memory when every allocated coefficient is active
    // (M == NC): the block loads it with coalesced 16-B loads and transposes it into the per-thread columns; per-thread
    // loads of a 192-B row touch 32 different sectors per instruction.
    const bool dense = f.M == NC && (3 * NC) % 4 == 0 && (reinterpret_cast<uintptr_t>(f.shs) & 15) == 0 &&
                       ((size_t)p.g_begin * 3 * NC) % 4 == 0;
    if (dense) {
        const size_t g0 = (size_t)p.g_begin + (size_t)blockIdx.x * SH_BWD_THREADS;
        const int rows = (int)min((size_t)SH_BWD_THREADS, (size_t)p.g_end - g0);
        const float4* src4 = reinterpret_cast<const float4*>(f.shs + g0 * 3 * NC);
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const int i = k * SH_BWD_THREADS + threadIdx.x;          // float4 index inside the block's run
            if (i < rows * NV) s_sh[i % NV][i / NV] = __ldg(src4 + i);
        }
        __syncthreads();
    } else {
        const float* src = f.shs + (size_t)gi * f.M * 3;
        float tmp[4 * NV];
#pragma unroll
        for (int k = 0; k < 4 * NV; k++) tmp[k] = k < 3 * NC ? __ldg(src + k) : 0.f;
#pragma unroll
        for (int v = 0; v < NV; v++)
            s_sh[v][threadIdx.x] = make_float4(tmp[4 * v], tmp[4 * v + 1], tmp[4 * v + 2], tmp[4 * v + 3]);
    }
    // (non-dense path: only this thread reads its column back, no barrier needed)
    const uint32_t a_sh = smem_addr(&s_sh[0][threadIdx.x]);
    // coefficient k, channel ch = float 3k+ch of the column; SHV(v) re-reads float4 slot v (volatile: the
    // compiler must not hoist the 48 values back into registers)
#define SHV(v) lds_f4(a_sh + (uint32_t)(v) * (uint32_t)(SH_BWD_THREADS * 16))

    float dsh[NC][3];
#pragma unroll
    for (int k = 0; k < NC; k++) dsh[k][0] = dsh[k][1] = dsh[k][2] = 0.f;
    float3 dmean = {0.f, 0.f, 0.f};

    struct Stage { int radius; float cr, cg, cb; unsigned mask; };
    auto fetch = [&](int s) {
        Stage t;
        const size_t n = (size_t)s * f.P + gi;
        t.radius = live ? f.radii[n] : 0;
        const float4 c = p.g2[n];   // (dL/dr, dL/dg, dL/db, -): its own plane, fully used sectors
        t.cr = c.x; t.cg = c.y; t.cb = c.z;
        t.mask = f.cmask[n];
        return t;
    };
    Stage pf[BWD_PF];
#pragma unroll
    for (int k = 0; k < BWD_PF; k++) {
        pf[k].radius = 0;
        if (k < f.F) pf[k] = fetch(k);
    }

    for (int s = 0; s < f.F; s++) {
        const Stage cur = pf[0];
#pragma unroll
        for (int k = 0; k + 1 < BWD_PF; k++) pf[k] = pf[k + 1];
        if (s + BWD_PF < f.F) pf[BWD_PF - 1] = fetch(s + BWD_PF);
        if (cur.radius <= 0) continue;

        const float3 cam = {f.campos[3 * s], f.campos[3 * s + 1], f.campos[3 * s + 2]};
        const float3 vv = {mean.x - cam.x, mean.y - cam.y, mean.z - cam.z};
        const float len = sqrtf(vv.x * vv.x + vv.y * vv.y + vv.z * vv.z);
        const float x = vv.x / len, y = vv.y / len, z = vv.z / len;
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;

        // basis values b_k(dir), in the reference's expression order
        float bs[NC];
        bs[0] = kSH0;
        if (DEG > 0) { bs[1] = -kSH1 * y; bs[2] = kSH1 * z; bs[3] = -kSH1 * x; }
        if (DEG > 1) {
            bs[4] = kSH2[0] * xy; bs[5] = kSH2[1] * yz; bs[6] = kSH2[2] * (2.f * zz - xx - yy);
            bs[7] = kSH2[3] * xz; bs[8] = kSH2[4] * (xx - yy);
        }
        if (DEG > 2) {
            bs[9] = kSH3[0] * y * (3.f * xx - yy); bs[10] = kSH3[1] * xy * z;
            bs[11] = kSH3[2] * y * (4.f * zz - xx - yy); bs[12] = kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
            bs[13] = kSH3[4] * x * (4.f * zz - xx - yy); bs[14] = kSH3[5] * z * (xx - yy);
            bs[15] = kSH3[6] * x * (xx - 3.f * yy);
        }

        float dRGB[3] = {cur.cr, cur.cg, cur.cb};
        if (f.use_sigmoid) {
            // colour = sigmoid(pre): recompute the pre-activation value (reference stores it, forward.cu:72-76);
            // one pass over the coefficient stream (float i of the column = coefficient i/3, channel i%3)
            float pre[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int v4 = 0; v4 < NV; v4++) {
                const float4 q = SHV(v4);
                const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int i = 4 * v4 + j;
                    if (i < 3 * NC) pre[i % 3] = fmaf(bs[i / 3], e[j], pre[i % 3]);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const float sg = 1.0f / (1.0f + expf(-pre[ch]));
                dRGB[ch] *= sg * (1.0f - sg);
            }
        } else {
            dRGB[0] *= (cur.mask & 1u) ? 1.f : 0.f;
            dRGB[1] *= (cur.mask & 2u) ? 1.f : 0.f;
            dRGB[2] *= (cur.mask & 4u) ? 1.f : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NC; k++) {
            dsh[k][0] = fmaf(bs[k], dRGB[0], dsh[k][0]);
            dsh[k][1] = fmaf(bs[k], dRGB[1], dsh[k][1]);
            dsh[k][2] = fmaf(bs[k], dRGB[2], dsh[k][2]);
        }
        if (DEG > 0) {
            // v_k = sh_k . dL/dRGB, accumulated while the coefficient stream passes (one float4 live at a
            // time);  dL/ddir = sum_k v_k grad b_k
            float v[NC];
#pragma unroll
            for (int k = 0; k < NC; k++) v[k] = 0.f;
#pragma unroll
            for (int v4 = 0; v4 < NV; v4++) {
                const float4 q = SHV(v4);
                const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int i = 4 * v4 + j;
                    if (i >= 3 && i < 3 * NC) v[i / 3] = fmaf(e[j], dRGB[i % 3], v[i / 3]);
                }
            }
            float dx = -kSH1 * v[3], dy = -kSH1 * v[1], dz = kSH1 * v[2];
            if (DEG > 1) {
                dx += kSH2[0] * y * v[4] + kSH2[2] * 2.f * -x * v[6] + kSH2[3] * z * v[7] + kSH2[4] * 2.f * x * v[8];
                dy += kSH2[0] * x * v[4] + kSH2[1] * z * v[5] + kSH2[2] * 2.f * -y * v[6] + kSH2[4] * 2.f * -y * v[8];
                dz += kSH2[1] * y * v[5] + kSH2[2] * 2.f * 2.f * z * v[6] + kSH2[3] * x * v[7];
            }
            if (DEG > 2) {
                dx += kSH3[0] * v[9] * 3.f * 2.f * xy + kSH3[1] * v[10] * yz + kSH3[2] * v[11] * -2.f * xy +
                      kSH3[3] * v[12] * -3.f * 2.f * xz + kSH3[4] * v[13] * (-3.f * xx + 4.f * zz - yy) +
                      kSH3[5] * v[14] * 2.f * xz + kSH3[6] * v[15] * 3.f * (xx - yy);
                dy += kSH3[0] * v[9] * 3.f * (xx - yy) + kSH3[1] * v[10] * xz +
                      kSH3[2] * v[11] * (-3.f * yy + 4.f * zz - xx) + kSH3[3] * v[12] * -3.f * 2.f * yz +
                      kSH3[4] * v[13] * -2.f * xy + kSH3[5] * v[14] * -2.f * yz + kSH3[6] * v[15] * -3.f * 2.f * xy;
                dz += kSH3[1] * v[10] * xy + kSH3[2] * v[11] * 4.f * 2.f * yz +
                      kSH3[3] * v[12] * 3.f * (2.f * zz - xx - yy) + kSH3[4] * v[13] * 4.f * 2.f * xz +
                      kSH3[5] * v[14] * (xx - yy);
            }
            // gradient through the normalisation of the view direction (reference dnormvdv, auxiliary.h)
            const float sum2 = vv.x * vv.x + vv.y * vv.y + vv.z * vv.z;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean.x += ((+sum2 - vv.x * vv.x) * dx - vv.y * vv.x * dy - vv.z * vv.x * dz) * invsum32;
            dmean.y += (-vv.x * vv.y * dx + (sum2 - vv.y * vv.y) * dy - vv.z * vv.y * dz) * invsum32;
            dmean.z += (-vv.x * vv.z * dx - vv.y * vv.z * dy + (sum2 - vv.z * vv.z) * dz) * invsum32;
        }
    }
#undef SHV
    if (dense && (reinterpret_cast<uintptr_t>(p.dL_dsh) & 15) == 0) {
        // the same transposition on the way out: dL/dsh rows through the (now free) columns, coalesced 16-B stores
        __syncthreads();
        float flat[4 * NV];
#pragma unroll
        for (int k = 0; k < NC; k++) { flat[3 * k] = dsh[k][0]; flat[3 * k + 1] = dsh[k][1]; flat[3 * k + 2] = dsh[k][2]; }
#pragma unroll
        for (int v = 0; v < NV; v++)
            s_sh[v][threadIdx.x] = make_float4(flat[4 * v], flat[4 * v + 1], flat[4 * v + 2], flat[4 * v + 3]);
        __syncthreads();
        const size_t g0 = (size_t)p.g_begin + (size_t)blockIdx.x * SH_BWD_THREADS;
        const int rows = (int)min((size_t)SH_BWD_THREADS, (size_t)p.g_end - g0);
        float4* dst4 = reinterpret_cast<float4*>(p.dL_dsh + g0 * 3 * NC);
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const int i = k * SH_BWD_THREADS + threadIdx.x;
            if (i < rows * NV) dst4[i] = s_sh[i % NV][i / NV];
        }
    } else if (live) {
        float* dst = p.dL_dsh + (size_t)g * f.M * 3;
#pragma unroll
        for (int k = 0; k < NC; k++) { dst[3 * k] = dsh[k][0]; dst[3 * k + 1] = dsh[k][1]; dst[3 * k + 2] = dsh[k][2]; }
        for (int k = NC; k < f.M; k++) { dst[3 * k] = 0.f; dst[3 * k + 1] = 0.f; dst[3 * k + 2] = 0.f; }
    }
    if (live && DEG > 0) {
        p.dL_dmeans3D[3 * g] += dmean.x; p.dL_dmeans3D[3 * g + 1] += dmean.y; p.dL_dmeans3D[3 * g + 2] += dmean.z;
    }
}

// pose_acc [F,32] fp64 -> dL_dview [F,16], dL_dproj [F,16] fp32
__global__ void k_pose_finalize(const double* __restrict__ acc, int F, float* __restrict__ dview,
                                float* __restrict__ dproj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * 32) return;
    const int s = i / 32, e = i % 32;
    if (e < 16) {
        dview[s * 16 + e] = (float)acc[s * 32 + e];
    } else {
        const int k = e - 16;
        // entries 3,7,11,15 of the projection gradient all carry the shared last-row term
        const double v = ((k & 3) == 3) ? acc[s * 32 + 19] : acc[s * 32 + e];
        dproj[s * 16 + k] = (float)v;
    }
}

void launch_preprocess_bwd(const BwdParams& p, int sh_degree, cudaStream_t st)
{
    const FwdParams& f = p.f;
    const int n = p.g_end - p.g_begin;
    if (n <= 0 || f.F <= 0) return;
    dim3 grid((n + PRE_BWD_THREADS - 1) / PRE_BWD_THREADS), block(PRE_BWD_THREADS);
    if (f.colors_precomp != nullptr) {
        k_preprocess_bwd<true><<<grid, block, 0, st>>>(p);
    } else {
        k_preprocess_bwd<false><<<grid, block, 0, st>>>(p);
        dim3 sgrid((n + SH_BWD_THREADS - 1) / SH_BWD_THREADS), sblock(SH_BWD_THREADS);
        switch (sh_degree) {
            case 0: k_sh_bwd<0><<<sgrid, sblock, 0, st>>>(p); break;
            case 1: k_sh_bwd<1><<<sgrid, sblock, 0, st>>>(p); break;
            case 2: k_sh_bwd<2><<<sgrid, sblock, 0, st>>>(p); break;
            default: k_sh_bwd<3><<<sgrid, sblock, 0, st>>>(p); break;
        }
    }
}

void launch_pose_finalize(const BwdParams& p, cudaStream_t st)
{
    const FwdParams& f = p.f;
    if (f.F > 0) k_pose_finalize<<<(f.F * 32 + 255) / 256, 256, 0, st>>>(p.pose_acc, f.F, p.dL_dview, p.dL_dproj);
}

}  // namespace dgs

// Mean squared distance to the three nearest neighbours (initialisation of the Gaussian
// scales).  Replaces `distCUDA2` / SimpleKNN::knn of the reference
// (submodules/simple-knn/simple_knn.cu:45-221, spatial.cu:15-26): bounding box -> 30-bit
// Morton codes -> sort -> per-box AABBs -> per point an exact pruned 3-NN search; result
// (d0 + d1 + d2) / 3 with d = dx*dx + dy*dy + dz*dz, self excluded by index.
//
// The search is exact, so values equal the reference's up to nothing but the order of the
// three additions (which is the same: ascending distance).  Design differences: points are
// gathered once into Morton order (coalesced float4), boxes hold 256 points instead of 1024
// (tighter AABBs), and a block of 256 Morton-consecutive queries walks the boxes together:
// a box is staged in shared memory once per block and only when some query of the block
// cannot reject it, instead of every thread re-reading every accepted box from global memory;
// runs of 32 boxes are rejected as a whole through a coarse AABB level (the reference tests every
// box for every point: O(P * P / 1024)).
#include "dgs_b200.h"
#include "dgs_internal.cuh"
#include <cfloat>

namespace dgs {

#define KNN_BOX 256
#define KNN_SUPER 32      // boxes per super-box (the coarse level of the search's pruning)

struct KnnLayout {
    size_t bbox, codes, keys_a, keys_b, vals_a, idx_sorted, pts, boxes, supers, sort_scratch, ticket, total;
    size_t stride;   // P rounded up to the sort's chunk
};
static KnnLayout knn_layout(size_t P)
{
    KnnLayout L;
    size_t o = 0;
    const size_t nb = (P + KNN_BOX - 1) / KNN_BOX;
    L.bbox = o; o = align_up(o + 8 * sizeof(float));
    const size_t Pp = (P + SORT_CHUNK - 1) / SORT_CHUNK * SORT_CHUNK;
    L.stride = Pp;
    L.codes = o; o = align_up(o + P * sizeof(uint32_t));
    L.keys_a = o; o = align_up(o + Pp * sizeof(uint32_t));
    L.keys_b = o; o = align_up(o + Pp * sizeof(uint32_t));
    L.vals_a = o; o = align_up(o + Pp * sizeof(uint32_t));
    L.idx_sorted = o; o = align_up(o + Pp * sizeof(uint32_t));
    L.pts = o; o = align_up(o + P * sizeof(float4));
    L.boxes = o; o = align_up(o + nb * 2 * sizeof(float4));
    L.supers = o; o = align_up(o + (nb + KNN_SUPER - 1) / KNN_SUPER * 2 * sizeof(float4));
    L.sort_scratch = o; o = align_up(o + sort_scratch_bytes((uint32_t)(Pp / SORT_CHUNK), 8));
    L.ticket = o; o = align_up(o + sizeof(uint32_t));
    L.total = o + 128;
    return L;
}

// order-preserving float <-> uint mapping for atomicMin/Max
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_knn_bbox_init(unsigned* bbox)
{
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;       // min
    else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;           // max
}
__global__ void k_knn_bbox(int P, const float* __restrict__ pts, unsigned* bbox)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
        for (int d = 0; d < 3; d++) {
            const float v = pts[3 * (size_t)i + d];
            mn[d] = fminf(mn[d], v); mx[d] = fmaxf(mx[d], v);
        }
    for (int d = 0; d < 3; d++) {
        for (int o = 16; o >= 1; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bbox + d, f2ord(mn[d]));
            atomicMax(bbox + 3 + d, f2ord(mx[d]));
        }
    }
}

__device__ __forceinline__ uint32_t spread10(uint32_t x)
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}
__global__ void k_knn_morton(int P, const float* __restrict__ pts, const unsigned* __restrict__ bbox,
                             uint32_t* __restrict__ codes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t c = 0;
    for (int d = 0; d < 3; d++) {
        const float mn = ord2f(bbox[d]), mx = ord2f(bbox[3 + d]);
        const float ext = mx - mn;
        float nrm = ext > 0.f ? (pts[3 * (size_t)i + d] - mn) / ext : 0.f;
        nrm = fminf(fmaxf(nrm, 0.f), 1.f);
        c |= spread10((uint32_t)(nrm * 1023.0f)) << d;
    }
    codes[i] = c;
}

__global__ void __launch_bounds__(KNN_BOX) k_knn_gather_boxes(int P, const float* __restrict__ pts,
                                                             const uint32_t* __restrict__ idx_sorted,
                                                             float4* __restrict__ sorted, float4* __restrict__ boxes)
{
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const uint32_t src = idx_sorted[i];
        const float x = pts[3 * (size_t)src], y = pts[3 * (size_t)src + 1], z = pts[3 * (size_t)src + 2];
        sorted[i] = make_float4(x, y, z, __uint_as_float(src));
        mn[0] = mx[0] = x; mn[1] = mx[1] = y; mn[2] = mx[2] = z;
    }
    __shared__ float s_mn[3][KNN_BOX / 32], s_mx[3][KNN_BOX / 32];
    for (int d = 0; d < 3; d++) {
        for (int o = 16; o >= 1; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
        if ((threadIdx.x & 31) == 0) { s_mn[d][threadIdx.x >> 5] = mn[d]; s_mx[d][threadIdx.x >> 5] = mx[d]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a[3], b[3];
        for (int d = 0; d < 3; d++) {
            a[d] = s_mn[d][0]; b[d] = s_mx[d][0];
            for (int w = 1; w < KNN_BOX / 32; w++) { a[d] = fminf(a[d], s_mn[d][w]); b[d] = fmaxf(b[d], s_mx[d][w]); }
        }
        boxes[2 * blockIdx.x] = make_float4(a[0], a[1], a[2], 0.f);
        boxes[2 * blockIdx.x + 1] = make_float4(b[0], b[1], b[2], 0.f);
    }
}

__device__ __forceinline__ void knn_insert(float dist, float (&best)[3])
{
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (best[j] > dist) { const float t = best[j]; best[j] = dist; dist = t; }
    }
}
__device__ __forceinline__ float sqdist(const float4& ref, const float4& p)
{
    const float dx = p.x - ref.x, dy = p.y - ref.y, dz = p.z - ref.z;
    return dx * dx + dy * dy + dz * dz;
}
__device__ __forceinline__ float box_dist(const float4& mn, const float4& mx, const float4& p)
{
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (p.x < mn.x || p.x > mx.x) dx = fminf(fabsf(p.x - mn.x), fabsf(p.x - mx.x));
    if (p.y < mn.y || p.y > mx.y) dy = fminf(fabsf(p.y - mn.y), fabsf(p.y - mx.y));
    if (p.z < mn.z || p.z > mx.z) dz = fminf(fabsf(p.z - mn.z), fabsf(p.z - mx.z));
    return dx * dx + dy * dy + dz * dz;
}

// AABB of every run of KNN_SUPER consecutive boxes (one warp per super-box).
__global__ void __launch_bounds__(32) k_knn_super(int nb, const float4* __restrict__ boxes, float4* __restrict__ supers)
{
    const int b = blockIdx.x * KNN_SUPER + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (b < nb) {
        const float4 a = boxes[2 * b], c = boxes[2 * b + 1];
        mn[0] = a.x; mn[1] = a.y; mn[2] = a.z; mx[0] = c.x; mx[1] = c.y; mx[2] = c.z;
    }
    for (int d = 0; d < 3; d++)
        for (int o = 16; o >= 1; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if (threadIdx.x == 0) {
        supers[2 * blockIdx.x] = make_float4(mn[0], mn[1], mn[2], 0.f);
        supers[2 * blockIdx.x + 1] = make_float4(mx[0], mx[1], mx[2], 0.f);
    }
}

// A block of 256 Morton-consecutive queries walks the boxes together: its own box first (it almost always contains
// the neighbours, so the rejection radius is small from the start), then the super-boxes outward in Morton order;
// a super-box no query of the block can reach is skipped as a whole (one vote), so that a block looks at
// nb / KNN_SUPER coarse boxes plus the few fine ones around it instead of at all nb (3 M points: 12 k boxes).
__global__ void __launch_bounds__(KNN_BOX) k_knn_search(int P, const float4* __restrict__ sorted,
                                                       const float4* __restrict__ boxes,
                                                       const float4* __restrict__ supers,
                                                       float* __restrict__ out)
{
    __shared__ float4 s_pts[KNN_BOX];
    const int nb = (P + KNN_BOX - 1) / KNN_BOX, nsb = (nb + KNN_SUPER - 1) / KNN_SUPER;
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    const bool live = i < P;
    const float4 me = live ? sorted[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    auto scan_box = [&](int b) {           // block-uniform b
        bool need = false;
        if (live) need = box_dist(boxes[2 * b], boxes[2 * b + 1], me) <= best[2];
        if (!__syncthreads_or(need)) return;
        const int j = b * KNN_BOX + threadIdx.x;
        if (j < P) s_pts[threadIdx.x] = sorted[j];
        __syncthreads();
        if (need) {
            const int cnt = min(KNN_BOX, P - b * KNN_BOX);
            for (int q = 0; q < cnt; q++) {
                if (b * KNN_BOX + q == i) continue;
                knn_insert(sqdist(me, s_pts[q]), best);
            }
        }
        __syncthreads();
    };
    const int own = blockIdx.x, own_sb = own / KNN_SUPER;
    scan_box(own);
    for (int step = 0; step < 2 * nsb; step++) {
        int sb;
        if (step == 0) sb = own_sb;
        else {
            const int k = (step + 1) >> 1;
            sb = (step & 1) ? own_sb + k : own_sb - k;
        }
        if (sb < 0 || sb >= nsb) continue;   // block-uniform
        bool reach = false;
        if (live) reach = box_dist(supers[2 * sb], supers[2 * sb + 1], me) <= best[2];
        if (!__syncthreads_or(reach)) continue;
        const int b0 = sb * KNN_SUPER, b1 = min(nb, b0 + KNN_SUPER);
        for (int b = b0; b < b1; b++)
            if (b != own) scan_box(b);
    }
    if (live) out[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

}  // namespace dgs

extern "C" {

size_t dgs_knn_scratch_bytes(int P) { return dgs::knn_layout((size_t)(P > 0 ? P : 0)).total; }

int dgs_knn_mean_dist2(int P, const float* points, float* mean_dist2, char* scratch, void* stream)
{
    using namespace dgs;
    if (P <= 0) return DGS_OK;
    if (!points || !mean_dist2 || !scratch) return dgs::fail(DGS_ERR_INVALID_ARGUMENT, "dgs_knn_mean_dist2: invalid argument");
    cudaStream_t st = (cudaStream_t)stream;
    const KnnLayout L = knn_layout((size_t)P);
    char* sc = (char*)(((uintptr_t)scratch + 127) & ~(uintptr_t)127);
    unsigned* bbox = (unsigned*)(sc + L.bbox);
    uint32_t* codes = (uint32_t*)(sc + L.codes);
    uint32_t* idx_sorted = (uint32_t*)(sc + L.idx_sorted);
    float4* pts = (float4*)(sc + L.pts);
    float4* boxes = (float4*)(sc + L.boxes);
    float4* supers = (float4*)(sc + L.supers);
    const int nb = (P + KNN_BOX - 1) / KNN_BOX;
    k_knn_bbox_init<<<1, 32, 0, st>>>(bbox);
    k_knn_bbox<<<min(1024, (P + 255) / 256), 256, 0, st>>>(P, points, bbox);
    k_knn_morton<<<(P + 255) / 256, 256, 0, st>>>(P, points, bbox, codes);
    // Morton order: the library's own segmented radix sort (dgs_binning.cu) with one segment; values = point index
    if (cudaMemsetAsync(sc + L.ticket, 0, sizeof(uint32_t), st) != cudaSuccess) return dgs::fail_cuda(cudaGetLastError(), "dgs_knn_mean_dist2");
    const SortScratch ss = bind_sort_scratch(sc + L.sort_scratch, (uint32_t)(L.stride / SORT_CHUNK), 8, (uint32_t*)(sc + L.ticket));
    sort_uniform_u32(1, (uint32_t)P, (uint32_t)L.stride, codes, (uint32_t*)(sc + L.keys_a), (uint32_t*)(sc + L.vals_a),
                     (uint32_t*)(sc + L.keys_b), idx_sorted, ss, 30, st);
    k_knn_gather_boxes<<<nb, KNN_BOX, 0, st>>>(P, points, idx_sorted, pts, boxes);
    k_knn_super<<<(nb + KNN_SUPER - 1) / KNN_SUPER, 32, 0, st>>>(nb, boxes, supers);
    k_knn_search<<<nb, KNN_BOX, 0, st>>>(P, pts, boxes, supers, mean_dist2);
    { const cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? DGS_OK : dgs::fail_cuda(e, "dgs_knn_mean_dist2"); }
}

}  // extern "C"

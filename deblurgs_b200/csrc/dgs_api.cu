// C-ABI of libdgs_b200.so (see include/dgs_b200.h for the contract and the reference
// interfaces each entry point replaces).  Host-side orchestration only: buffer carving,
// the scan / radix-sort library calls and kernel launches, all on the caller's stream.
#include "dgs_b200.h"
#include "dgs_internal.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace dgs {

static thread_local std::string g_last_error;

static int fail(int code, const char* what)
{
    g_last_error = what;
    return code;
}
static int fail_cuda(cudaError_t e, const char* where)
{
    g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
    return DGS_ERR_CUDA;
}
#define DGS_CUDA(call, where)                                   \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return fail_cuda(e__, where);   \
    } while (0)

// ---- stage profiling ---------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int stage; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static double g_stage_ms[ST_COUNT];
static long long g_stage_calls[ST_COUNT];
static long long g_own_launches = 0;
static const char* kStageNames[ST_COUNT] = {"preprocess_fwd", "scan", "duplicate", "sort", "tile_ranges",
                                            "render_fwd", "blur_mean", "bwd_memset", "render_bwd",
                                            "preprocess_bwd", "pose_fwd", "pose_bwd", "activate_fwd",
                                            "activate_bwd", "adam"};
static cudaEvent_t get_event()
{
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
StageTimer::StageTimer(int stage, cudaStream_t st_, int own_kernels) : idx(-1), st(st_)
{
    g_own_launches += own_kernels;
    if (!g_prof_on) return;
    ProfRec r;
    r.a = get_event(); r.b = get_event(); r.stage = stage;
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
    idx = (int)g_prof.size() - 1;
}
StageTimer::~StageTimer()
{
    if (idx >= 0) cudaEventRecord(g_prof[idx].b, st);
}

static int bits_for(uint32_t n)  // number of bits needed to hold values 0..n-1 (>= 1)
{
    int b = 1;
    while ((1ull << b) < n) b++;
    return b;
}
// The reference sorts bits [0, 32 + getHigherMsb(tiles)) (rasterizer_impl.cu:35-50,306):
// getHigherMsb(n) = position of the MSB of n, plus one.
static int ref_tile_bits(uint32_t tiles)
{
    int b = 0;
    while (tiles >> b) b++;
    return b < 1 ? 1 : b;
}

GeomLayout geom_layout(size_t N)
{
    GeomLayout L;
    L.n_entries = N;
    size_t o = 0;
    L.geo0 = o; o = align_up(o + N * sizeof(float4));
    L.geo1 = o; o = align_up(o + N * sizeof(float4));
    L.geo2 = o; o = align_up(o + N * sizeof(float4));
    L.tiles = o; o = align_up(o + N * sizeof(uint32_t));
    L.offsets = o; o = align_up(o + (N + 1) * sizeof(uint32_t));   // [N] scan + 1 word: the key-overflow flag
    size_t tmp = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)N);
    tmp += 1024;   // the permuted-input scan may ask for a little more than the plain one
    L.scan_temp_bytes = tmp;
    L.scan_temp = o; o = align_up(o + tmp);
    L.dkeys = o; o = align_up(o + N * sizeof(uint64_t));
    L.dkeys_sorted = o; o = align_up(o + N * sizeof(uint64_t));
    L.order_in = o; o = align_up(o + N * sizeof(uint32_t));
    L.order = o; o = align_up(o + N * sizeof(uint32_t));
    size_t tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp2, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)N);
    L.sort_temp_bytes = tmp2;
    L.sort_temp = o; o = align_up(o + tmp2);
    L.total = o + 128;
    return L;
}
BinLayout bin_layout(size_t D)
{
    BinLayout L;
    size_t o = 0;
    L.point_list = o; o = align_up(o + D * sizeof(uint32_t));
    L.keys = o; o = align_up(o + D * sizeof(uint32_t));
    L.keys_unsorted = o; o = align_up(o + D * sizeof(uint32_t));
    L.vals_unsorted = o; o = align_up(o + D * sizeof(uint32_t));
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)D);
    L.sort_temp_bytes = tmp;
    L.sort_temp = o; o = align_up(o + tmp);
    L.total = o + 128;
    return L;
}
ImgLayout img_layout(size_t F, size_t tiles, size_t pixels)
{
    ImgLayout L;
    size_t o = 0;
    L.ranges = o; o = align_up(o + F * tiles * sizeof(uint2));
    L.final_T = o; o = align_up(o + F * pixels * sizeof(float));
    L.n_contrib = o; o = align_up(o + F * pixels * sizeof(uint32_t));
    L.total = o + 128;
    return L;
}

static char* aligned128(char* p) { return (char*)(((uintptr_t)p + 127) & ~(uintptr_t)127); }

static int fill_params(FwdParams& p, int P, int F, int M, const float* background, int W, int H,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* opacities, const float* scales, float scale_modifier,
                       const float* rotations, const float* cov3D_precomp, const float* view,
                       const float* proj, const float* campos, float tan_fovx, float tan_fovy,
                       float z_near, float z_far, int prefiltered, int use_sigmoid)
{
    if (P < 0 || F < 0 || W < 0 || H < 0) return fail(DGS_ERR_INVALID_ARGUMENT, "negative size");
    if (F > DGS_MAX_SUBFRAMES) return fail(DGS_ERR_INVALID_ARGUMENT, "too many sub-frames (max 256)");
    if (P > 0 && F > 0) {
        if ((shs == nullptr) == (colors_precomp == nullptr))
            return fail(DGS_ERR_INVALID_ARGUMENT, "provide exactly one of SHs or precomputed colors");
        const bool has_sr = scales != nullptr && rotations != nullptr;
        if (has_sr == (cov3D_precomp != nullptr) || ((scales != nullptr) != (rotations != nullptr)))
            return fail(DGS_ERR_INVALID_ARGUMENT,
                        "provide exactly one of scale/rotation pair or precomputed 3D covariance");
        if (!means3D || !opacities || !view || !proj || !background)
            return fail(DGS_ERR_INVALID_ARGUMENT, "null required input");
        if (shs != nullptr && campos == nullptr) return fail(DGS_ERR_INVALID_ARGUMENT, "campos required with SHs");
    }
    p.P = P; p.F = F; p.M = M; p.W = W; p.H = H;
    p.tiles_x = (W + DGS_TILE_X - 1) / DGS_TILE_X;
    p.tiles_y = (H + DGS_TILE_Y - 1) / DGS_TILE_Y;
    p.tile_bits = ref_tile_bits((uint32_t)(p.tiles_x * p.tiles_y));
    p.tan_fovx = tan_fovx; p.tan_fovy = tan_fovy;
    p.focal_y = H / (2.0f * tan_fovy);
    p.focal_x = W / (2.0f * tan_fovx);
    p.scale_modifier = scale_modifier;
    p.z_near = z_near; p.z_far = z_far;
    p.prefiltered = prefiltered; p.use_sigmoid = use_sigmoid;
    p.means3D = means3D; p.shs = shs; p.colors_precomp = colors_precomp; p.opacities = opacities;
    p.scales = scales; p.rotations = rotations; p.cov3D_precomp = cov3D_precomp;
    p.view = view; p.proj = proj; p.campos = campos; p.background = background;
    return DGS_OK;
}

static void bind_geom(FwdParams& p, char* geom, const GeomLayout& G)
{
    p.geo0 = (float4*)(geom + G.geo0);
    p.geo1 = (float4*)(geom + G.geo1);
    p.geo2 = (float4*)(geom + G.geo2);
    p.tiles = (uint32_t*)(geom + G.tiles);
    p.offsets = (uint32_t*)(geom + G.offsets);
    p.dkeys = (uint64_t*)(geom + G.dkeys);
    p.order_in = (uint32_t*)(geom + G.order_in);
    p.order = (uint32_t*)(geom + G.order);
    p.key_overflow = p.offsets + G.n_entries;
}

// tiles[order[i]]: input of the scan over the depth-sorted entry order
struct PermutedTiles {
    const uint32_t* tiles;
    const uint32_t* order;
    __host__ __device__ uint32_t operator()(uint32_t i) const { return tiles[order[i]]; }
};

}  // namespace dgs

using namespace dgs;

extern "C" {

const char* dgs_last_error(void) { return g_last_error.c_str(); }
int dgs_version(void) { return 100; }
int dgs_compiled_arch(void) { return 1000; }

int dgs_key_bits(int width, int height, int F, int* tile_bits, int* subframe_bits)
{
    const int tx = (width + DGS_TILE_X - 1) / DGS_TILE_X, ty = (height + DGS_TILE_Y - 1) / DGS_TILE_Y;
    if (tile_bits) *tile_bits = ref_tile_bits((uint32_t)(tx * ty));
    if (subframe_bits) *subframe_bits = F > 1 ? bits_for((uint32_t)F) : 0;
    return DGS_OK;
}

int dgs_blur_forward(
    dgs_alloc_fn geom_alloc, void* geom_ctx, dgs_alloc_fn binning_alloc, void* binning_ctx,
    dgs_alloc_fn image_alloc, void* image_ctx,
    int P, int F, int sh_degree, int sh_coeffs,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float z_near, float z_far,
    int prefiltered, int use_sigmoid,
    float* out_color, float* out_depth, int* radii,
    float* out_blur, float blur_denominator,
    int64_t* num_rendered, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    int rc = fill_params(p, P, F, sh_coeffs, background, width, height, means3D, shs, colors_precomp,
                         opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                         projmatrix, campos, tan_fovx, tan_fovy, z_near, z_far, prefiltered, use_sigmoid);
    if (rc != DGS_OK) return rc;
    if (!geom_alloc || !binning_alloc || !image_alloc) return fail(DGS_ERR_INVALID_ARGUMENT, "null allocator");
    if (F > 0 && width * height > 0 && (!out_color || !out_depth)) return fail(DGS_ERR_INVALID_ARGUMENT, "null output");
    if (shs != nullptr && (sh_degree < 0 || sh_degree > 3 || sh_coeffs < (sh_degree + 1) * (sh_degree + 1)))
        return fail(DGS_ERR_INVALID_ARGUMENT, "bad SH degree / coefficient count");
    if (P > 0 && F > 0 && !radii) return fail(DGS_ERR_INVALID_ARGUMENT, "null radii");
    p.radii = radii;

    const size_t N = (size_t)P * F;
    const size_t tiles = (size_t)p.tiles_x * p.tiles_y;
    const size_t pixels = (size_t)width * height;
    const int sf_bits = F > 1 ? bits_for((uint32_t)F) : 0;
    if (N >= (1ull << 32)) return fail(DGS_ERR_UNSUPPORTED, "P*F must be < 2^32");

    const GeomLayout G = geom_layout(N);
    const ImgLayout I = img_layout(F, tiles, pixels);
    char* geom = geom_alloc(geom_ctx, G.total);
    char* img = image_alloc(image_ctx, I.total);
    if (!geom || !img) return fail(DGS_ERR_ALLOC, "state buffer allocation failed");
    geom = aligned128(geom);
    img = aligned128(img);
    bind_geom(p, geom, G);
    uint2* ranges = (uint2*)(img + I.ranges);
    float* final_T = (float*)(img + I.final_T);
    uint32_t* n_contrib = (uint32_t*)(img + I.n_contrib);

    int64_t D = 0;
    if (N > 0) {
        // Depth sort on a 32-bit key [sub-frame | depth code] (4 radix passes over 8-B pairs) whenever the sub-frame
        // id leaves >= 27 bits for the depth code; a scene with a visible depth beyond the code's range (reported by
        // preprocess through key_overflow, read in the one host synchronisation below) is re-sorted on the 64-bit key
        // [sub-frame | depth bits] (5 passes over 12-B pairs at F = 16), which is also the path for F > 32.
        p.depth_key_bits = sf_bits <= 5 ? (sf_bits >= 1 ? 32 - sf_bits : 31) : 0;
        DGS_CUDA(cudaMemsetAsync(p.key_overflow, 0, sizeof(uint32_t), st), "flag memset");
        { StageTimer t(ST_PREPROCESS_FWD, st, 1); launch_preprocess_fwd(p, sh_degree, st); }
        uint32_t host_vals[2] = {0u, 0u};   // total duplicates, overflow flag
        for (int attempt = 0; attempt < 2; attempt++) {
            {
                // stage 1 of the binning: depth order of the (sub-frame, Gaussian) entries
                StageTimer t(ST_SORT, st, 0);
                size_t tmp1 = G.sort_temp_bytes;
                if (p.depth_key_bits != 0)
                    DGS_CUDA(cub::DeviceRadixSort::SortPairs(geom + G.sort_temp, tmp1, (const uint32_t*)p.dkeys,
                                                             (uint32_t*)(geom + G.dkeys_sorted), p.order_in, p.order,
                                                             (int64_t)N, 0, 32, st),
                             "depth sort (32-bit keys)");
                else
                    DGS_CUDA(cub::DeviceRadixSort::SortPairs(geom + G.sort_temp, tmp1, p.dkeys,
                                                             (uint64_t*)(geom + G.dkeys_sorted), p.order_in, p.order,
                                                             (int64_t)N, 0, 32 + sf_bits, st),
                             "depth sort");
            }
            size_t tmp = G.scan_temp_bytes;
            {
                StageTimer t(ST_SCAN, st, 0);
                auto in = thrust::make_transform_iterator(thrust::make_counting_iterator<uint32_t>(0u),
                                                          PermutedTiles{p.tiles, p.order});
                DGS_CUDA(cub::DeviceScan::InclusiveSum(geom + G.scan_temp, tmp, in, p.offsets, (int64_t)N, st), "scan");
            }
            // The one host synchronisation of the batched forward (the reference does one per
            // sub-frame, rasterizer_impl.cu:287): the binning buffer is sized from it.
            // (offsets[N-1] and the flag word are adjacent: one copy)
            DGS_CUDA(cudaMemcpyAsync(host_vals, p.offsets + N - 1, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "num_rendered copy");
            DGS_CUDA(cudaStreamSynchronize(st), "num_rendered sync");
            if (p.depth_key_bits == 0 || host_vals[1] == 0u) break;
            p.depth_key_bits = 0;             // rare: a depth beyond the compact code -> exact 64-bit keys, sort again
            launch_rebuild_depth_keys(p, st);
        }
        D = (int64_t)host_vals[0];
    }
    if (num_rendered) *num_rendered = D;

    const BinLayout B = bin_layout((size_t)D);
    char* bin = binning_alloc(binning_ctx, B.total);
    if (!bin) return fail(DGS_ERR_ALLOC, "binning buffer allocation failed");
    bin = aligned128(bin);
    uint32_t* point_list = (uint32_t*)(bin + B.point_list);
    uint32_t* keys = (uint32_t*)(bin + B.keys);
    uint32_t* keys_unsorted = (uint32_t*)(bin + B.keys_unsorted);
    uint32_t* vals_unsorted = (uint32_t*)(bin + B.vals_unsorted);

    if (F > 0 && tiles > 0)
        DGS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)F * tiles * sizeof(uint2), st), "ranges memset");
    if (D > 0) {
        { StageTimer t(ST_DUPLICATE, st, 1); launch_duplicate(p, keys_unsorted, vals_unsorted, st); }
        size_t tmp = B.sort_temp_bytes;
        {
            // stage 2: stable sort of the duplicates on the short [sub-frame | tile] key
            StageTimer t(ST_SORT, st, 0);
            DGS_CUDA(cub::DeviceRadixSort::SortPairs(bin + B.sort_temp, tmp, keys_unsorted, keys, vals_unsorted,
                                                     point_list, (int64_t)D, 0, p.tile_bits + sf_bits, st),
                     "tile sort");
        }
        { StageTimer t(ST_TILE_RANGES, st, 1); launch_tile_ranges(D, keys, p.tile_bits, (int)tiles, ranges, st); }
    }
    if (F > 0 && pixels > 0) {
        { StageTimer t(ST_RENDER_FWD, st, 1); launch_render_fwd(p, ranges, point_list, final_T, n_contrib, out_color, out_depth, st); }
        if (out_blur) { StageTimer t(ST_BLUR_MEAN, st, 1); launch_blur_mean(out_color, F, 3 * pixels, blur_denominator, out_blur, st); }
    }
    DGS_CUDA(cudaGetLastError(), "forward launch");
    return DGS_OK;
}

size_t dgs_blur_backward_scratch_bytes(int P, int F)
{
    const size_t N = (size_t)P * (size_t)F;
    return align_up(N * 12 * sizeof(float)) + align_up((size_t)F * 32 * sizeof(double)) + 256;
}

int dgs_blur_backward(
    int P, int F, int sh_degree, int sh_coeffs, int64_t num_rendered,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float z_near, float z_far, int use_sigmoid,
    const int* radii,
    const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
    const float* dL_dpix, const float* dL_dpixdepth, const float* dL_dblur, float blur_denominator,
    char* scratch,
    float* dL_dmeans2D, float* dL_dmeans3D, float* dL_dsh, float* dL_dopacity,
    float* dL_dscales, float* dL_drotations, float* dL_dcolors_precomp, float* dL_dcov3D_precomp,
    float* dL_dviewmatrix, float* dL_dprojmatrix, float* densify_stats, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BwdParams b;
    memset(&b, 0, sizeof(b));
    int rc = fill_params(b.f, P, F, sh_coeffs, background, width, height, means3D, shs, colors_precomp,
                         opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                         projmatrix, campos, tan_fovx, tan_fovy, z_near, z_far, 0, use_sigmoid);
    if (rc != DGS_OK) return rc;
    if (F == 0) return DGS_OK;
    if (!dL_dviewmatrix || !dL_dprojmatrix) return fail(DGS_ERR_INVALID_ARGUMENT, "null pose gradient output");
    if (P == 0) {
        DGS_CUDA(cudaMemsetAsync(dL_dviewmatrix, 0, (size_t)F * 16 * sizeof(float), st), "memset");
        DGS_CUDA(cudaMemsetAsync(dL_dprojmatrix, 0, (size_t)F * 16 * sizeof(float), st), "memset");
        return DGS_OK;
    }
    if (!geom_buffer || !binning_buffer || !image_buffer || !scratch || !radii)
        return fail(DGS_ERR_INVALID_ARGUMENT, "null state buffer");
    if (!dL_dmeans3D || !dL_dopacity) return fail(DGS_ERR_INVALID_ARGUMENT, "null gradient output");
    if (shs && !dL_dsh) return fail(DGS_ERR_INVALID_ARGUMENT, "null dL_dsh");
    if (scales && (!dL_dscales || !dL_drotations)) return fail(DGS_ERR_INVALID_ARGUMENT, "null dL_dscales/dL_drotations");

    const size_t N = (size_t)P * F;
    const size_t tiles = (size_t)b.f.tiles_x * b.f.tiles_y;
    const size_t pixels = (size_t)width * height;
    const GeomLayout G = geom_layout(N);
    const ImgLayout I = img_layout(F, tiles, pixels);
    const BinLayout B = bin_layout((size_t)num_rendered);
    char* geom = aligned128((char*)geom_buffer);
    char* img = aligned128((char*)image_buffer);
    char* bin = aligned128((char*)binning_buffer);
    bind_geom(b.f, geom, G);
    b.f.radii = (int*)radii;
    b.ranges = (const uint2*)(img + I.ranges);
    b.final_T = (const float*)(img + I.final_T);
    b.n_contrib = (const uint32_t*)(img + I.n_contrib);
    b.point_list = (const uint32_t*)(bin + B.point_list);
    b.dL_dpix = dL_dpix;
    b.dL_dpixdepth = dL_dpixdepth;
    b.dL_dblur = dL_dblur;
    b.blur_denominator = blur_denominator;
    char* sc = aligned128(scratch);
    b.g0 = (float4*)sc;
    b.pose_acc = (double*)(sc + align_up(N * 12 * sizeof(float)));
    b.dL_dmeans2D = dL_dmeans2D; b.dL_dmeans3D = dL_dmeans3D; b.dL_dsh = dL_dsh; b.dL_dopacity = dL_dopacity;
    b.dL_dscales = dL_dscales; b.dL_drotations = dL_drotations;
    b.dL_dcolors_precomp = dL_dcolors_precomp; b.dL_dcov3D_precomp = dL_dcov3D_precomp;
    b.dL_dview = dL_dviewmatrix; b.dL_dproj = dL_dprojmatrix;
    b.densify_stats = densify_stats;

    {
        StageTimer t(ST_BWD_MEMSET, st, 0);
        DGS_CUDA(cudaMemsetAsync(sc, 0, align_up(N * 12 * sizeof(float)) + (size_t)F * 32 * sizeof(double), st), "grad memset");
    }
    if (num_rendered > 0 && pixels > 0) { StageTimer t(ST_RENDER_BWD, st, 1); launch_render_bwd(b, st); }
    { StageTimer t(ST_PREPROCESS_BWD, st, colors_precomp ? 2 : 3); launch_preprocess_bwd(b, sh_degree, st); }
    DGS_CUDA(cudaGetLastError(), "backward launch");
    return DGS_OK;
}

int dgs_forward(
    dgs_alloc_fn geom_alloc, void* geom_ctx, dgs_alloc_fn binning_alloc, void* binning_ctx,
    dgs_alloc_fn image_alloc, void* image_ctx,
    int P, int sh_degree, int sh_coeffs,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float z_near, float z_far,
    int prefiltered, int use_sigmoid,
    float* out_color, float* out_depth, int* radii,
    int64_t* num_rendered, void* stream)
{
    return dgs_blur_forward(geom_alloc, geom_ctx, binning_alloc, binning_ctx, image_alloc, image_ctx, P, 1,
                            sh_degree, sh_coeffs, background, width, height, means3D, shs, colors_precomp,
                            opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                            projmatrix, campos, tan_fovx, tan_fovy, z_near, z_far, prefiltered, use_sigmoid,
                            out_color, out_depth, radii, nullptr, 1.0f, num_rendered, stream);
}

int dgs_backward(
    int P, int sh_degree, int sh_coeffs, int64_t num_rendered,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float z_near, float z_far, int use_sigmoid,
    const int* radii,
    const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
    const float* dL_dpix, const float* dL_dpixdepth,
    char* scratch,
    float* dL_dmeans2D, float* dL_dmeans3D, float* dL_dsh, float* dL_dopacity,
    float* dL_dscales, float* dL_drotations, float* dL_dcolors_precomp, float* dL_dcov3D_precomp,
    float* dL_dviewmatrix, float* dL_dprojmatrix, void* stream)
{
    return dgs_blur_backward(P, 1, sh_degree, sh_coeffs, num_rendered, background, width, height, means3D, shs,
                             colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp,
                             viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, z_near, z_far, use_sigmoid,
                             radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dpixdepth, nullptr, 1.0f, scratch,
                             dL_dmeans2D, dL_dmeans3D, dL_dsh, dL_dopacity, dL_dscales, dL_drotations,
                             dL_dcolors_precomp, dL_dcov3D_precomp, dL_dviewmatrix, dL_dprojmatrix, nullptr, stream);
}

// ---- profiling / measurement ------------------------------------------------------------
int dgs_profile_enable(int on)
{
    g_prof_on = on != 0;
    return DGS_OK;
}
int dgs_profile_num_stages(void) { return ST_COUNT; }
const char* dgs_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
int dgs_profile_read(double* ms, int64_t* calls, int n, int reset)
{
    for (auto& r : g_prof) {
        float t = 0.f;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
        if (e != cudaSuccess) return fail_cuda(e, "profile read");
        g_stage_ms[r.stage] += t;
        g_stage_calls[r.stage] += 1;
        g_event_pool.push_back(r.a);
        g_event_pool.push_back(r.b);
    }
    g_prof.clear();
    for (int i = 0; i < n && i < ST_COUNT; i++) {
        if (ms) ms[i] = g_stage_ms[i];
        if (calls) calls[i] = g_stage_calls[i];
    }
    if (reset) {
        for (int i = 0; i < ST_COUNT; i++) { g_stage_ms[i] = 0.0; g_stage_calls[i] = 0; }
    }
    return DGS_OK;
}
int64_t dgs_launch_count(int reset)
{
    const int64_t v = g_own_launches;
    if (reset) g_own_launches = 0;
    return v;
}
void dgs_profile_note(int stage, void* stream, int own_kernels, int begin, int* token)
{
    // used by translation units that cannot see StageTimer's storage (pose kernels)
    if (begin) {
        StageTimer* t = new StageTimer(stage, (cudaStream_t)stream, own_kernels);
        *token = t->idx;
        t->idx = -1;   // do not record the end event on destruction
        delete t;
    } else if (*token >= 0 && *token < (int)g_prof.size()) {
        cudaEventRecord(g_prof[*token].b, (cudaStream_t)stream);
    }
}

int dgs_debug_workload(const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
                       int P, int F, int width, int height, int64_t num_rendered, uint64_t* out_dev,
                       void* stream)
{
    if (F == 0 || P == 0) return DGS_OK;
    if (!geom_buffer || !binning_buffer || !image_buffer || !out_dev) return fail(DGS_ERR_INVALID_ARGUMENT, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    FwdParams p;
    memset(&p, 0, sizeof(p));
    p.P = P; p.F = F; p.W = width; p.H = height;
    p.tiles_x = (width + DGS_TILE_X - 1) / DGS_TILE_X;
    p.tiles_y = (height + DGS_TILE_Y - 1) / DGS_TILE_Y;
    const size_t N = (size_t)P * F, tiles = (size_t)p.tiles_x * p.tiles_y, pixels = (size_t)width * height;
    const GeomLayout G = geom_layout(N);
    const ImgLayout I = img_layout(F, tiles, pixels);
    const BinLayout B = bin_layout((size_t)num_rendered);
    char* geom = aligned128((char*)geom_buffer);
    char* img = aligned128((char*)image_buffer);
    char* bin = aligned128((char*)binning_buffer);
    bind_geom(p, geom, G);
    DGS_CUDA(cudaMemsetAsync(out_dev, 0, 3 * sizeof(uint64_t), st), "workload memset");
    launch_workload(p, (const uint2*)(img + I.ranges), (const uint32_t*)(bin + B.point_list),
                    (const uint32_t*)(img + I.n_contrib), (unsigned long long*)out_dev, st);
    DGS_CUDA(cudaGetLastError(), "workload");
    return DGS_OK;
}

// ---- debug / parity accessors ---------------------------------------------------------
__global__ void k_debug_geometry(size_t N, const float4* g0, const float4* g1, const float4* g2,
                                 const uint32_t* tiles, const uint32_t* offsets, float* depths,
                                 float* means2D, float* conic_opacity, float* rgb, float* clamped,
                                 uint32_t* tiles_out, uint32_t* offsets_out)
{
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const bool vis = tiles[n] > 0;
    float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
    if (vis) { a = g0[n]; b = g1[n]; c = g2[n]; }
    if (depths) depths[n] = a.z;
    if (means2D) { means2D[2 * n] = a.x; means2D[2 * n + 1] = a.y; }
    if (conic_opacity) { conic_opacity[4 * n] = b.x; conic_opacity[4 * n + 1] = b.y; conic_opacity[4 * n + 2] = b.z; conic_opacity[4 * n + 3] = b.w; }
    if (rgb) { rgb[3 * n] = c.x; rgb[3 * n + 1] = c.y; rgb[3 * n + 2] = c.z; }
    if (clamped) {
        const unsigned m = vis ? __float_as_uint(c.w) : 0u;
        clamped[3 * n] = (m & 1u) ? 1.f : 0.f; clamped[3 * n + 1] = (m & 2u) ? 1.f : 0.f; clamped[3 * n + 2] = (m & 4u) ? 1.f : 0.f;
    }
    if (tiles_out) tiles_out[n] = tiles[n];
    if (offsets_out) offsets_out[n] = offsets[n];
}

int dgs_debug_geometry(const char* geom_buffer, int P, int F, float* depths, float* means2D,
                       float* conic_opacity, float* rgb, float* clamped, uint32_t* tiles_touched,
                       uint32_t* point_offsets, void* stream)
{
    const size_t N = (size_t)P * F;
    if (N == 0) return DGS_OK;
    if (!geom_buffer) return fail(DGS_ERR_INVALID_ARGUMENT, "null geometry buffer");
    const GeomLayout G = geom_layout(N);
    char* geom = aligned128((char*)geom_buffer);
    k_debug_geometry<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        N, (const float4*)(geom + G.geo0), (const float4*)(geom + G.geo1), (const float4*)(geom + G.geo2),
        (const uint32_t*)(geom + G.tiles), (const uint32_t*)(geom + G.offsets), depths, means2D,
        conic_opacity, rgb, clamped, tiles_touched, point_offsets);
    DGS_CUDA(cudaGetLastError(), "debug geometry");
    return DGS_OK;
}

int dgs_debug_binning(const char* geom_buffer, const char* binning_buffer, int P, int F, int width, int height,
                      int64_t num_rendered, uint64_t* keys, uint32_t* point_list, void* stream)
{
    if (num_rendered <= 0) return DGS_OK;
    if (!binning_buffer || !geom_buffer) return fail(DGS_ERR_INVALID_ARGUMENT, "null state buffer");
    const BinLayout B = bin_layout((size_t)num_rendered);
    char* bin = aligned128((char*)binning_buffer);
    cudaStream_t st = (cudaStream_t)stream;
    if (keys) {
        FwdParams p;
        memset(&p, 0, sizeof(p));
        p.P = P; p.F = F; p.W = width; p.H = height;
        p.tiles_x = (width + DGS_TILE_X - 1) / DGS_TILE_X;
        p.tiles_y = (height + DGS_TILE_Y - 1) / DGS_TILE_Y;
        p.tile_bits = ref_tile_bits((uint32_t)(p.tiles_x * p.tiles_y));
        bind_geom(p, aligned128((char*)geom_buffer), geom_layout((size_t)P * F));
        launch_rebuild_keys(p, num_rendered, (const uint32_t*)(bin + B.keys), (const uint32_t*)(bin + B.point_list),
                            keys, st);
        DGS_CUDA(cudaGetLastError(), "rebuild keys");
    }
    if (point_list) DGS_CUDA(cudaMemcpyAsync(point_list, bin + B.point_list, num_rendered * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st), "copy list");
    return DGS_OK;
}

int dgs_debug_image(const char* image_buffer, int F, int width, int height, uint32_t* ranges,
                    float* final_T, uint32_t* n_contrib, void* stream)
{
    const size_t tiles = (size_t)((width + DGS_TILE_X - 1) / DGS_TILE_X) * ((height + DGS_TILE_Y - 1) / DGS_TILE_Y);
    const size_t pixels = (size_t)width * height;
    if (F == 0) return DGS_OK;
    if (!image_buffer) return fail(DGS_ERR_INVALID_ARGUMENT, "null image buffer");
    const ImgLayout I = img_layout(F, tiles, pixels);
    char* img = aligned128((char*)image_buffer);
    cudaStream_t st = (cudaStream_t)stream;
    if (ranges && tiles) DGS_CUDA(cudaMemcpyAsync(ranges, img + I.ranges, F * tiles * sizeof(uint2), cudaMemcpyDeviceToDevice, st), "copy ranges");
    if (final_T && pixels) DGS_CUDA(cudaMemcpyAsync(final_T, img + I.final_T, F * pixels * sizeof(float), cudaMemcpyDeviceToDevice, st), "copy T");
    if (n_contrib && pixels) DGS_CUDA(cudaMemcpyAsync(n_contrib, img + I.n_contrib, F * pixels * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st), "copy n_contrib");
    return DGS_OK;
}

// ---- mark_visible ----------------------------------------------------------------------
__global__ void k_mark_visible(int P, const float* __restrict__ means, const float* __restrict__ V,
                               uint8_t* __restrict__ present)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P) return;
    const float3 m = {means[3 * g], means[3 * g + 1], means[3 * g + 2]};
    const float3 pv = xform_point_4x3(m, V);
    present[g] = pv.z > 0.2f ? 1 : 0;
}

int dgs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream)
{
    (void)projmatrix;
    if (P == 0) return DGS_OK;
    if (!means3D || !viewmatrix || !present) return fail(DGS_ERR_INVALID_ARGUMENT, "null argument");
    k_mark_visible<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, means3D, viewmatrix, present);
    DGS_CUDA(cudaGetLastError(), "mark_visible");
    return DGS_OK;
}

}  // extern "C"

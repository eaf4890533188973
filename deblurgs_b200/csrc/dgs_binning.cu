// Binning of the batched blurry-view rasterizer (sm_100a): everything between preprocess and the tile
// blend -- the reference's InclusiveSum + duplicateWithKeys + SortPairs + identifyTileRanges
// (cuda_rasterizer/rasterizer_impl.cu:70-138, 283-320) -- as hand-written kernels with NO host
// synchronisation: the number of duplicates D stays on the device.
//
// Design (B200-first, not the reference's single 64-bit sort of D duplicates):
//   stage 1  segmented LSD radix sort of the (sub-frame, Gaussian) entries on the 32 depth bits, one segment
//            per sub-frame (4 passes x 8 bits over N = F*P items; values = Gaussian index, generated in pass 1).
//            Visible depths are positive floats, so their bit patterns order like the values: exact for ANY
//            depth, culled entries carry 0xFFFFFFFF and end up last.
//   scan     k_entry_gather / k_seg_scan / k_entry_offsets: tile counts in depth order -> per-sub-frame
//            inclusive offsets, segment table (start of every sub-frame's list, padded to the sort's chunk
//            size), D as a 64-bit total, overflow flag against the binning capacity.
//   stage 2  segmented LSD radix sort of the duplicates on the TILE id only (the sub-frame is the segment, not
//            a key field: 10 bits at 600x400 = 2 passes x 5 bits, 13 bits at 1080p = 7 + 6).  Pass 1 GENERATES
//            its items from the depth-ordered entries (duplicateWithKeys fused into the sort: no unsorted
//            key/value arrays are ever written), the last pass writes only the Gaussian indices
//            (point_list) and derives the per-tile ranges from the run boundaries it sees in shared memory
//            (identifyTileRanges fused into the sort: sorted keys are never written either).
//   Stable sorts + emission in (depth, Gaussian index) order give exactly the reference's lists.
//
// Every pass is reduce-then-scan (k_sort_upsweep -> k_scan_counters -> k_sort_downsweep): per-chunk digit
// histograms, one exclusive scan over the [segment][digit][chunk] counter matrix, then a stable scatter that
// ranks 4096 items per block with warp-level match_any, reorders them in shared memory and writes runs.
// No spin-waiting anywhere (the scan uses a last-block-done ticket), so a kernel can never hang on
// block-scheduling order.
#include "dgs_internal.cuh"

namespace dgs {

#define FULL_MASK 0xffffffffu

// ---------------------------------------------------------------------------------------------------
// chunk -> segment lookup
// ---------------------------------------------------------------------------------------------------
struct ChunkInfo {
    uint32_t s;        // segment
    uint32_t c;        // chunk index inside the segment
    uint32_t nv;       // valid items in the chunk
    uint32_t nch;      // chunks of the segment
    uint32_t cbase;    // flat index of the segment's first chunk
    uint32_t start;    // seg_start[s]
    uint32_t adj;      // seg_start[s] - (number of items in earlier segments)
};

__device__ __forceinline__ ChunkInfo locate_chunk(const SegTable& t, uint32_t b)
{
    ChunkInfo ci;
    if (t.seg_start == nullptr) {
        const uint32_t per = t.uni_stride / SORT_CHUNK;
        ci.s = b / per;
        ci.c = b - ci.s * per;
        ci.nch = per;
        ci.cbase = ci.s * per;
        ci.start = ci.s * t.uni_stride;
        ci.adj = ci.s * (t.uni_stride - t.uni_len);
        const uint32_t done = ci.c * SORT_CHUNK;
        ci.nv = t.uni_len > done ? min((uint32_t)SORT_CHUNK, t.uni_len - done) : 0u;
    } else {
        const uint32_t pos = b * SORT_CHUNK;
        int lo = 0, hi = t.nseg - 1;          // last s with seg_start[s] <= pos (empty segments share a start
        while (lo < hi) {                      // with their successor and are skipped by taking the last)
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(t.seg_start + mid) <= pos) lo = mid; else hi = mid - 1;
        }
        ci.s = (uint32_t)lo;
        ci.start = __ldg(t.seg_start + lo);
        const uint32_t next = __ldg(t.seg_start + lo + 1);
        ci.cbase = ci.start / SORT_CHUNK;
        ci.nch = (next - ci.start) / SORT_CHUNK;
        ci.c = b - ci.cbase;
        ci.adj = __ldg(t.seg_adj + lo);
        const uint32_t len = __ldg(t.seg_len + lo), done = ci.c * SORT_CHUNK;
        ci.nv = len > done ? min((uint32_t)SORT_CHUNK, len - done) : 0u;
    }
    return ci;
}

__device__ __forceinline__ uint32_t total_chunks(const SegTable& t)
{
    return t.n_chunks ? *t.n_chunks : (uint32_t)t.nseg * (t.uni_stride / SORT_CHUNK);
}

// ---------------------------------------------------------------------------------------------------
// Generation of the stage-2 items (duplicateWithKeys, rasterizer_impl.cu:70-111, fused into the sort).
// The chunk covers duplicates [r0, r1) of its sub-frame's depth-ordered emission; entry i owns
// [off[i-1], off[i]) and enumerates the tiles of its rectangle row-major, like the reference.  A warp takes
// 32 consecutive entries at a time (coalesced loads of offsets and packed rectangles) and then walks them
// one after the other with all lanes spread over the entry's tiles.
// ---------------------------------------------------------------------------------------------------
template <class Emit>
__device__ __forceinline__ void expand_chunk(const GenParams& gp, const ChunkInfo& ci, uint32_t chunk_flat,
                                             unsigned warp, unsigned lane, Emit emit)
{
    const uint32_t r0 = ci.c * SORT_CHUNK, r1 = r0 + ci.nv;
    const size_t seg = (size_t)ci.s * gp.entry_stride;
    const uint32_t* __restrict__ off = gp.off + seg;
    const uint2* __restrict__ rec = gp.rec + seg;
    const uint32_t i0 = gp.chunk_first[chunk_flat];
    for (uint32_t base = i0 + warp * 32u;; base += SORT_WARPS * 32u) {
        const uint32_t i = base + lane;
        uint32_t incl = 0xFFFFFFFFu, excl = 0xFFFFFFFFu;
        uint2 r = make_uint2(0u, 0u);
        if (i < gp.entries_per_seg) {
            incl = off[i];
            excl = i ? off[i - 1] : 0u;
            r = rec[i];
        }
        if (__shfl_sync(FULL_MASK, excl, 0) >= r1) break;     // this and every later group start behind the chunk
        unsigned m = __ballot_sync(FULL_MASK, excl < r1 && incl > r0 && incl > excl);
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t e_excl = __shfl_sync(FULL_MASK, excl, l), e_incl = __shfl_sync(FULL_MASK, incl, l);
            const uint32_t rx = __shfl_sync(FULL_MASK, r.x, l), g = __shfl_sync(FULL_MASK, r.y, l);
            const uint32_t x0 = rx & 1023u, y0 = (rx >> 10) & 1023u, w = ((rx >> 20) & 1023u) + 1u;
            const uint32_t ja = max(e_excl, r0) - e_excl, jb = min(e_incl, r1) - e_excl;
            const float wf = (float)w;
            for (uint32_t j = ja + lane; j < jb; j += 32u) {
                // row-major over the rectangle; (j + 0.5) / w is never within 0.5 / w of an integer, so the float
                // quotient (2 ulp) truncates to floor(j / w) exactly for any rectangle of a <= 16K x 16K image
                const uint32_t row = (uint32_t)__fdividef((float)j + 0.5f, wf);
                const uint32_t tile = (y0 + row) * (uint32_t)gp.tiles_x + x0 + (j - row * w);
                emit(e_excl + j - r0, tile, g);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// upsweep: digit histogram of one chunk -> counters[(cbase * BINS) + digit * nch + c]
// ---------------------------------------------------------------------------------------------------
template <int BITS, bool GEN>
__global__ void __launch_bounds__(SORT_THREADS) k_sort_upsweep(const SegTable t, const uint32_t* __restrict__ keys_in,
                                                               uint32_t in_seg_stride, int shift,
                                                               uint32_t* __restrict__ counters, const GenParams gp)
{
    constexpr int BINS = 1 << BITS;
    __shared__ uint32_t hist[SORT_WARPS][BINS];
    if (blockIdx.x >= total_chunks(t)) return;
    const ChunkInfo ci = locate_chunk(t, blockIdx.x);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (int k = tid; k < SORT_WARPS * BINS; k += SORT_THREADS) (&hist[0][0])[k] = 0u;
    __syncthreads();
    if (GEN) {
        expand_chunk(gp, ci, blockIdx.x, warp, lane, [&](uint32_t, uint32_t tile, uint32_t) {
            atomicAdd(&hist[warp][(tile >> shift) & (BINS - 1)], 1u);
        });
    } else {
        const uint32_t* __restrict__ src =
            keys_in + (in_seg_stride ? (size_t)ci.s * in_seg_stride : (size_t)ci.start) + (size_t)ci.c * SORT_CHUNK;
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
            const uint32_t j = i * SORT_THREADS + tid;
            if (j < ci.nv) atomicAdd(&hist[warp][(src[j] >> shift) & (BINS - 1)], 1u);
        }
    }
    __syncthreads();
    for (int d = tid; d < BINS; d += SORT_THREADS) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) sum += hist[w][d];
        counters[(size_t)ci.cbase * BINS + (size_t)d * ci.nch + ci.c] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------
// exclusive scan over the counter matrix (n = chunks * bins elements): every block scans one slice of
// SCAN_SLICE elements in place and publishes the slice total; the last block to finish (ticket) turns the
// slice totals into exclusive slice bases.  Consumers add slice_base[index / SCAN_SLICE].
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_warp /*[9]*/, unsigned tid,
                                                             uint32_t& total)
{
    const unsigned lane = tid & 31u, warp = tid >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { const uint32_t x = s_warp[w]; s_warp[w] = acc; acc += x; }
        s_warp[SORT_WARPS] = acc;
    }
    __syncthreads();
    total = s_warp[SORT_WARPS];
    const uint32_t res = inc - v + s_warp[warp];
    __syncthreads();          // s_warp may be reused by the caller
    return res;
}

__global__ void __launch_bounds__(SORT_THREADS) k_scan_counters(uint32_t* __restrict__ data, const uint32_t* n_chunks_dev,
                                                                uint32_t n_chunks_host, uint32_t bins,
                                                                uint32_t* __restrict__ slice_base, uint32_t* ticket)
{
    __shared__ uint32_t s_warp[SORT_WARPS + 1];
    __shared__ bool s_last;
    const uint32_t n = (n_chunks_dev ? *n_chunks_dev : n_chunks_host) * bins;
    const uint32_t n_slices = (n + SCAN_SLICE - 1) / SCAN_SLICE;
    if (blockIdx.x >= n_slices) return;
    const unsigned tid = threadIdx.x;
    const uint32_t base = blockIdx.x * SCAN_SLICE + tid * SORT_ITEMS;     // blocked: 16 consecutive counters per thread
    uint32_t v[SORT_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < SORT_ITEMS / 4; q++) {
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t i = base + 4 * q;
        if (i + 3 < n) x = *reinterpret_cast<const uint4*>(data + i);
        else {
            if (i < n) x.x = data[i];
            if (i + 1 < n) x.y = data[i + 1];
            if (i + 2 < n) x.z = data[i + 2];
        }
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) { const uint32_t x = v[i]; v[i] = sum; sum += x; }
    uint32_t total;
    const uint32_t ex = block_exclusive_scan_256(sum, s_warp, tid, total);
#pragma unroll
    for (int q = 0; q < SORT_ITEMS / 4; q++) {
        const uint32_t i = base + 4 * q;
        const uint4 x = make_uint4(v[4 * q] + ex, v[4 * q + 1] + ex, v[4 * q + 2] + ex, v[4 * q + 3] + ex);
        if (i + 3 < n) *reinterpret_cast<uint4*>(data + i) = x;
        else {
            if (i < n) data[i] = x.x;
            if (i + 1 < n) data[i + 1] = x.y;
            if (i + 2 < n) data[i + 2] = x.z;
        }
    }
    if (tid == 0) {
        slice_base[blockIdx.x] = total;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == n_slices - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: exclusive scan of the slice totals, in place
    uint32_t carry = 0;
    for (uint32_t s0 = 0; s0 < n_slices; s0 += SORT_THREADS) {
        const uint32_t i = s0 + tid;
        const uint32_t x = i < n_slices ? __ldcg(slice_base + i) : 0u;
        uint32_t tot;
        const uint32_t e = block_exclusive_scan_256(x, s_warp, tid, tot);
        if (i < n_slices) slice_base[i] = carry + e;
        carry += tot;
    }
    if (tid == 0) *ticket = 0u;   // ready for the next pass (stream order)
}

// ---------------------------------------------------------------------------------------------------
// downsweep: stable scatter of one chunk.  Items sit in registers in warp-striped order (warp w owns the 512
// consecutive items w*512 .. w*512+511, item i of lane l is w*512 + i*32 + l), so that "rank order" =
// (warp, round, lane) = input order.  Ranking: per round, lanes with equal digits find each other with
// match_any; the lowest of them bumps the warp's private counter of that digit.  Then one pass over the
// bins turns the per-warp counts into block-local start positions, the items are written to shared memory
// at their sorted position and read back in position order, which makes the global stores runs of
// consecutive addresses per digit.
//   LAST: only the values are written (point_list) and the per-tile ranges are derived from the tile-id
//   changes between neighbours in shared memory: the chunk's input is ordered by the lower digits, the
//   ranking is stable, so inside a digit the items of one tile are contiguous both here and in the output.
//   ranges[t] = (~start, end) accumulated with atomicMax over the chunks that hold a piece of the run.
// ---------------------------------------------------------------------------------------------------
template <int BITS, bool GEN, bool LAST>
__global__ void __launch_bounds__(SORT_THREADS) k_sort_downsweep(
    const SegTable t, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t in_seg_stride,
    int shift, const uint32_t* __restrict__ counters, const uint32_t* __restrict__ slice_base,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, const GenParams gp, uint2* __restrict__ ranges,
    uint32_t tiles_per_seg)
{
    constexpr int BINS = 1 << BITS;
    __shared__ uint32_t s_keys[SORT_CHUNK];
    __shared__ uint32_t s_vals[SORT_CHUNK];
    __shared__ uint32_t s_wc[SORT_WARPS][BINS + 1];      // per-warp digit counters (+1: invalid items)
    __shared__ uint32_t s_gdelta[BINS];                  // global position - block-local sorted position, per digit
    __shared__ uint32_t s_scan[SORT_WARPS + 1];
    if (blockIdx.x >= total_chunks(t)) return;
    const ChunkInfo ci = locate_chunk(t, blockIdx.x);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    for (int k = tid; k < SORT_WARPS * (BINS + 1); k += SORT_THREADS) (&s_wc[0][0])[k] = 0u;

    uint32_t key[SORT_ITEMS], val[SORT_ITEMS];
    if (GEN) {
        expand_chunk(gp, ci, blockIdx.x, warp, lane, [&](uint32_t pos, uint32_t tile, uint32_t g) {
            s_keys[pos] = tile;
            s_vals[pos] = g;
        });
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
            const uint32_t j = warp * (32 * SORT_ITEMS) + i * 32 + lane;
            key[i] = s_keys[j];
            val[i] = s_vals[j];
        }
    } else {
        const size_t in0 = (in_seg_stride ? (size_t)ci.s * in_seg_stride : (size_t)ci.start) + (size_t)ci.c * SORT_CHUNK;
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
            const uint32_t j = warp * (32 * SORT_ITEMS) + i * 32 + lane;
            key[i] = 0u;
            val[i] = 0u;
            if (j < ci.nv) {
                key[i] = keys_in[in0 + j];
                val[i] = vals_in ? vals_in[in0 + j] : ci.c * SORT_CHUNK + j;     // pass 1 of stage 1: index in the segment
            }
        }
    }
    __syncthreads();     // counters zeroed; GEN: every thread has read its items (s_keys / s_vals are reused below)

    // ---- rank inside the warp
    uint32_t rank[SORT_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t j = warp * (32 * SORT_ITEMS) + i * 32 + lane;
        const uint32_t d = j < ci.nv ? ((key[i] >> shift) & (BINS - 1)) : (uint32_t)BINS;
        const unsigned peers = __match_any_sync(FULL_MASK, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if ((int)lane == leader) {
            old = s_wc[warp][d];
            s_wc[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(FULL_MASK, old, leader);
        rank[i] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: counts of the warps -> exclusive prefix over the warps; digit totals -> block-local starts
    uint32_t tot = 0;
    if (tid < BINS) {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) { const uint32_t x = s_wc[w][tid]; s_wc[w][tid] = tot; tot += x; }
    }
    uint32_t chunk_total;
    const uint32_t bin_start = block_exclusive_scan_256(tot, s_scan, tid, chunk_total);
    if (tid < BINS) {
        const size_t ci_idx = (size_t)ci.cbase * BINS + (size_t)tid * ci.nch + ci.c;
        const uint32_t g = counters[ci_idx] + slice_base[ci_idx / SCAN_SLICE] + ci.adj;
        s_gdelta[tid] = g - bin_start;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) s_wc[w][tid] += bin_start;
    }
    __syncthreads();

    // ---- scatter to the block-local sorted position
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t j = warp * (32 * SORT_ITEMS) + i * 32 + lane;
        if (j < ci.nv) {
            const uint32_t d = (key[i] >> shift) & (BINS - 1);
            const uint32_t p = s_wc[warp][d] + rank[i];
            s_keys[p] = key[i];
            s_vals[p] = val[i];
        }
    }
    __syncthreads();

    // ---- write runs
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t p = i * SORT_THREADS + tid;
        if (p < ci.nv) {
            const uint32_t k = s_keys[p];
            const uint32_t out = s_gdelta[(k >> shift) & (BINS - 1)] + p;
            vals_out[out] = s_vals[p];
            if (!LAST) {
                keys_out[out] = k;
            } else {
                uint2* r = ranges + (size_t)ci.s * tiles_per_seg + k;
                if (p == 0 || s_keys[p - 1] != k) atomicMax(&r->x, ~out);
                if (p == ci.nv - 1 || s_keys[p + 1] != k) atomicMax(&r->y, out + 1u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side of one segmented sort
// ---------------------------------------------------------------------------------------------------
template <int BITS, bool GEN>
static void launch_upsweep(uint32_t grid, const SegTable& t, const uint32_t* keys_in, uint32_t in_stride, int shift,
                           uint32_t* counters, const GenParams& gp, cudaStream_t st)
{
    k_sort_upsweep<BITS, GEN><<<grid, SORT_THREADS, 0, st>>>(t, keys_in, in_stride, shift, counters, gp);
}
template <int BITS, bool GEN, bool LAST>
static void launch_downsweep(uint32_t grid, const SegTable& t, const uint32_t* keys_in, const uint32_t* vals_in,
                             uint32_t in_stride, int shift, const uint32_t* counters, const uint32_t* slice_base,
                             uint32_t* keys_out, uint32_t* vals_out, const GenParams& gp, uint2* ranges,
                             uint32_t tiles_per_seg, cudaStream_t st)
{
    k_sort_downsweep<BITS, GEN, LAST><<<grid, SORT_THREADS, 0, st>>>(t, keys_in, vals_in, in_stride, shift, counters,
                                                                     slice_base, keys_out, vals_out, gp, ranges,
                                                                     tiles_per_seg);
}

#define DGS_BITS_SWITCH(bits, CALL)          \
    switch (bits) {                          \
        case 5: { constexpr int B = 5; CALL; } break; \
        case 6: { constexpr int B = 6; CALL; } break; \
        case 7: { constexpr int B = 7; CALL; } break; \
        default: { constexpr int B = 8; CALL; } break; \
    }

int sort_pass_plan(int key_bits, int* bits_per_pass)
{
    if (key_bits < 1) key_bits = 1;
    const int passes = (key_bits + 7) / 8;
    int bits = (key_bits + passes - 1) / passes;
    if (bits < 5) bits = 5;
    *bits_per_pass = bits;
    return passes;
}

// One pass.  gen != nullptr: items are generated (stage 2, pass 1); ranges != nullptr: last pass of stage 2.
void sort_pass(const SegTable& t, uint32_t max_chunks, int bits, int shift, const uint32_t* keys_in,
               const uint32_t* vals_in, uint32_t in_stride, uint32_t* keys_out, uint32_t* vals_out,
               const SortScratch& sc, const GenParams* gen, uint2* ranges, uint32_t tiles_per_seg, cudaStream_t st)
{
    if (max_chunks == 0) return;
    GenParams gp;
    memset(&gp, 0, sizeof(gp));
    if (gen) gp = *gen;
    const uint32_t max_slices = (uint32_t)(((size_t)max_chunks * (1u << bits) + SCAN_SLICE - 1) / SCAN_SLICE);
    if (gen) {
        DGS_BITS_SWITCH(bits, (launch_upsweep<B, true>(max_chunks, t, keys_in, in_stride, shift, sc.counters, gp, st)));
    } else {
        DGS_BITS_SWITCH(bits, (launch_upsweep<B, false>(max_chunks, t, keys_in, in_stride, shift, sc.counters, gp, st)));
    }
    k_scan_counters<<<max_slices, SORT_THREADS, 0, st>>>(sc.counters, t.n_chunks,
                                                         (uint32_t)t.nseg * (t.uni_stride / SORT_CHUNK), 1u << bits,
                                                         sc.slice_base, sc.ticket);
    if (gen && ranges) {
        DGS_BITS_SWITCH(bits, (launch_downsweep<B, true, true>(max_chunks, t, keys_in, vals_in, in_stride, shift, sc.counters, sc.slice_base, keys_out, vals_out, gp, ranges, tiles_per_seg, st)));
    } else if (gen) {
        DGS_BITS_SWITCH(bits, (launch_downsweep<B, true, false>(max_chunks, t, keys_in, vals_in, in_stride, shift, sc.counters, sc.slice_base, keys_out, vals_out, gp, ranges, tiles_per_seg, st)));
    } else if (ranges) {
        DGS_BITS_SWITCH(bits, (launch_downsweep<B, false, true>(max_chunks, t, keys_in, vals_in, in_stride, shift, sc.counters, sc.slice_base, keys_out, vals_out, gp, ranges, tiles_per_seg, st)));
    } else {
        DGS_BITS_SWITCH(bits, (launch_downsweep<B, false, false>(max_chunks, t, keys_in, vals_in, in_stride, shift, sc.counters, sc.slice_base, keys_out, vals_out, gp, ranges, tiles_per_seg, st)));
    }
}

size_t sort_scratch_bytes(uint32_t max_chunks, int bits)
{
    const size_t counters = (size_t)max_chunks * (1u << bits);
    const size_t slices = (counters + SCAN_SLICE - 1) / SCAN_SLICE;
    return align_up(counters * 4) + align_up(slices * 4 + 4) + 128;
}
SortScratch bind_sort_scratch(char* base, uint32_t max_chunks, int bits, uint32_t* ticket)
{
    SortScratch sc;
    const size_t counters = (size_t)max_chunks * (1u << bits);
    sc.counters = (uint32_t*)base;
    sc.slice_base = (uint32_t*)(base + align_up(counters * 4));
    sc.ticket = ticket;
    return sc;
}

// Uniform segments (stage 1 and the debug entry): keys [nseg][len] -> sorted values (index inside the segment)
// in the padded layout [nseg][stride]; all 32 key bits, 4 passes of 8.  tmp: 3 arrays of nseg*stride u32 + scratch.
void sort_uniform_u32(int nseg, uint32_t len, uint32_t stride, const uint32_t* keys, uint32_t* keys_a, uint32_t* vals_a,
                      uint32_t* keys_b, uint32_t* vals_b, const SortScratch& sc, int key_bits, cudaStream_t st)
{
    SegTable t;
    memset(&t, 0, sizeof(t));
    t.nseg = nseg; t.uni_len = len; t.uni_stride = stride;
    const uint32_t chunks = (uint32_t)nseg * (stride / SORT_CHUNK);
    const int passes = (key_bits + 7) / 8;
    // ping-pong so that the LAST pass lands in (keys_b, vals_b)
    const uint32_t* kin = keys;
    const uint32_t* vin = nullptr;
    uint32_t in_stride = len;
    for (int p = 0; p < passes; p++) {
        const bool to_b = ((passes - 1 - p) & 1) == 0;
        uint32_t* ko = to_b ? keys_b : keys_a;
        uint32_t* vo = to_b ? vals_b : vals_a;
        sort_pass(t, chunks, 8, 8 * p, kin, vin, in_stride, ko, vo, sc, nullptr, nullptr, 0, st);
        kin = ko; vin = vo; in_stride = 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// scan stage: tile counts in depth order -> offsets, segment table, D
// ---------------------------------------------------------------------------------------------------
#define ENT_BLOCK 1024   // entries per block of the gather / offsets kernels (256 threads x 4)

// rect[n] = (x0, y0, w, h) as 4 x u16 (w = h = 0: culled) -> packed record of the duplicate generator
__global__ void __launch_bounds__(256) k_entry_gather(int P, uint32_t stride, const uint32_t* __restrict__ order,
                                                      const uint2* __restrict__ rect, uint32_t* __restrict__ cnt_sorted,
                                                      uint2* __restrict__ rec, unsigned long long* __restrict__ block_sums)
{
    __shared__ unsigned long long s_red[8];
    const int s = blockIdx.y;
    const uint32_t i0 = blockIdx.x * ENT_BLOCK;
    unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t i = i0 + k * 256 + threadIdx.x;
        if (i < (uint32_t)P) {
            const uint32_t g = order[(size_t)s * stride + i];
            const uint2 r = __ldg(rect + (size_t)s * P + g);
            const uint32_t x0 = r.x & 0xFFFFu, y0 = r.x >> 16, w = r.y & 0xFFFFu, h = r.y >> 16;
            const uint32_t cnt = w * h;
            cnt_sorted[(size_t)s * stride + i] = cnt;
            rec[(size_t)s * stride + i] = make_uint2(x0 | (y0 << 10) | ((w ? w - 1u : 0u) << 20), g);
            sum += cnt;
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0;
        for (int w = 0; w < 8; w++) a += s_red[w];
        block_sums[(size_t)s * gridDim.x + blockIdx.x] = a;
    }
}

// one block: per-segment exclusive scan of the block sums, then the segment table and the status words
__global__ void __launch_bounds__(1024) k_seg_scan(int nseg, uint32_t nb, const unsigned long long* __restrict__ block_sums,
                                                   uint32_t* __restrict__ block_excl, BinStatus* __restrict__ status,
                                                   uint32_t* __restrict__ seg_start, uint32_t* __restrict__ seg_len,
                                                   uint32_t* __restrict__ seg_adj, unsigned long long capacity)
{
    __shared__ unsigned long long s_warp[33];
    __shared__ unsigned long long s_total[DGS_MAX_SUBFRAMES];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (int s = 0; s < nseg; s++) {
        unsigned long long carry = 0;
        for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
            const uint32_t b = b0 + tid;
            const unsigned long long v = b < nb ? block_sums[(size_t)s * nb + b] : 0ull;
            unsigned long long inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(FULL_MASK, inc, d);
                if (lane >= (unsigned)d) inc += o;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            if (tid == 0) {
                unsigned long long acc = 0;
                for (int w = 0; w < 32; w++) { const unsigned long long x = s_warp[w]; s_warp[w] = acc; acc += x; }
                s_warp[32] = acc;
            }
            __syncthreads();
            if (b < nb) block_excl[(size_t)s * nb + b] = (uint32_t)(carry + s_warp[warp] + inc - v);
            carry += s_warp[32];
            __syncthreads();
        }
        if (tid == 0) s_total[s] = carry;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long D = 0, padded = 0;
        bool too_big = false;
        for (int s = 0; s < nseg; s++) {
            const unsigned long long len = s_total[s];
            if (len >= 0xFFFFFFFFull) too_big = true;
            seg_start[s] = (uint32_t)padded;
            seg_len[s] = (uint32_t)len;
            seg_adj[s] = (uint32_t)(padded - D);
            D += len;
            padded += (len + SORT_CHUNK - 1) / SORT_CHUNK * SORT_CHUNK;
        }
        seg_start[nseg] = (uint32_t)padded;
        const bool overflow = too_big || padded > capacity || padded >= 0xFFFFFFFFull;
        status->num_rendered = D;
        status->padded = padded;
        status->overflow = overflow ? 1u : 0u;
        status->n_chunks = overflow ? 0u : (uint32_t)(padded / SORT_CHUNK);   // overflow: stage 2 does nothing
    }
}

__global__ void __launch_bounds__(256) k_entry_offsets(int P, uint32_t stride, const uint32_t* __restrict__ cnt_sorted,
                                                       const uint32_t* __restrict__ block_excl,
                                                       const uint32_t* __restrict__ seg_start,
                                                       const BinStatus* __restrict__ status, uint32_t* __restrict__ off,
                                                       uint32_t* __restrict__ chunk_first)
{
    __shared__ uint32_t s_scan[SORT_WARPS + 1];
    if (status->overflow) return;
    const int s = blockIdx.y;
    const unsigned tid = threadIdx.x;
    const uint32_t i0 = blockIdx.x * ENT_BLOCK + tid * 4;             // blocked: 4 consecutive entries per thread
    uint32_t c[4] = {0u, 0u, 0u, 0u};
    const uint32_t* src = cnt_sorted + (size_t)s * stride;
    if (i0 + 3 < (uint32_t)P) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i0);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k < (uint32_t)P) c[k] = src[i0 + k];
    }
    const uint32_t mine = c[0] + c[1] + c[2] + c[3];
    uint32_t total;
    uint32_t run = block_exclusive_scan_256(mine, s_scan, tid, total) + block_excl[(size_t)s * gridDim.x + blockIdx.x];
    const uint32_t cbase = seg_start[s] / SORT_CHUNK;
    uint32_t inc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t excl = run;
        run += c[k];
        inc[k] = run;
        if (c[k] != 0u) {
            // chunk boundaries q * CHUNK inside [excl, run): this entry owns the first duplicate of chunk q
            for (uint32_t q = (excl + SORT_CHUNK - 1) / SORT_CHUNK; (unsigned long long)q * SORT_CHUNK < run; q++)
                chunk_first[cbase + q] = i0 + k;
        }
    }
    uint32_t* dst = off + (size_t)s * stride;
    if (i0 + 3 < (uint32_t)P) {
        *reinterpret_cast<uint4*>(dst + i0) = make_uint4(inc[0], inc[1], inc[2], inc[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k < (uint32_t)P) dst[i0 + k] = inc[k];
    }
}

void launch_entry_scan(const FwdParams& p, const BinState& b, unsigned long long capacity, cudaStream_t st)
{
    const uint32_t nb = (uint32_t)((p.P + ENT_BLOCK - 1) / ENT_BLOCK);
    dim3 grid(nb, p.F);
    k_entry_gather<<<grid, 256, 0, st>>>(p.P, b.stride, b.order, p.rect, b.cnt_sorted, b.rec, b.block_sums);
    k_seg_scan<<<1, 1024, 0, st>>>(p.F, nb, b.block_sums, b.block_excl, b.status, b.seg_start, b.seg_len, b.seg_adj,
                                   capacity);
}
void launch_entry_offsets(const FwdParams& p, const BinState& b, uint32_t* chunk_first, cudaStream_t st)
{
    const uint32_t nb = (uint32_t)((p.P + ENT_BLOCK - 1) / ENT_BLOCK);
    dim3 grid(nb, p.F);
    k_entry_offsets<<<grid, 256, 0, st>>>(p.P, b.stride, b.cnt_sorted, b.block_excl, b.seg_start, b.status, b.off,
                                          chunk_first);
}

// ---------------------------------------------------------------------------------------------------
// parity accessor: the reference's sorted (key, Gaussian) list of every sub-frame, compacted (no padding),
// with the full 64-bit key [sub-frame | tile | depth bits] a single sort on the reference's layout would carry
// ---------------------------------------------------------------------------------------------------
__global__ void k_debug_lists(int P, int F, int tiles, int tile_bits, const uint2* __restrict__ ranges,
                              const uint32_t* __restrict__ point_list, const float4* __restrict__ geo0,
                              const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_adj,
                              uint64_t* __restrict__ keys64, uint32_t* __restrict__ list_out,
                              uint32_t* __restrict__ ranges_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * tiles) return;
    const int s = i / tiles, tile = i - s * tiles;
    const uint2 r = decode_range(ranges[i]);
    // compact position = padded position - (seg_start[s] - items in earlier segments)
    const uint32_t shift = seg_adj[s];
    if (ranges_out) {
        ranges_out[2 * i] = r.y > r.x ? r.x - shift : 0u;
        ranges_out[2 * i + 1] = r.y > r.x ? r.y - shift : 0u;
    }
    for (uint32_t d = r.x; d < r.y; d++) {
        const uint32_t g = point_list[d];
        const uint32_t o = d - shift;
        if (list_out) list_out[o] = g;
        if (keys64) {
            const float depth = geo0[(size_t)s * P + g].z;
            keys64[o] = ((uint64_t)s << (32 + tile_bits)) | ((uint64_t)tile << 32) | (uint64_t)__float_as_uint(depth);
        }
    }
}

void launch_debug_lists(int P, int F, int tiles, int tile_bits, const uint2* ranges, const uint32_t* point_list,
                        const float4* geo0, const uint32_t* seg_start, const uint32_t* seg_adj, uint64_t* keys64,
                        uint32_t* list_out, uint32_t* ranges_out, cudaStream_t st)
{
    const int n = F * tiles;
    if (n <= 0) return;
    k_debug_lists<<<(n + 127) / 128, 128, 0, st>>>(P, F, tiles, tile_bits, ranges, point_list, geo0, seg_start, seg_adj,
                                                   keys64, list_out, ranges_out);
}

}  // namespace dgs

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
// ranks 2048 items per block with warp-level match_any, reorders them in shared memory and writes runs.
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
    uint32_t first_entry;   // stage 2: first depth-ordered entry owning a duplicate of the chunk
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
        ci.first_entry = 0u;
        const uint32_t done = ci.c * SORT_CHUNK;
        ci.nv = t.uni_len > done ? min((uint32_t)SORT_CHUNK, t.uni_len - done) : 0u;
    } else {
        // one 8-B load of the chunk table (written by k_entry_offsets), then three independent loads
        const uint2 ct = __ldg(t.chunk_tab + b);
        ci.first_entry = ct.x;
        ci.s = ct.y;
        ci.start = __ldg(t.seg_start + ci.s);
        const uint32_t next = __ldg(t.seg_start + ci.s + 1);
        const uint32_t len = __ldg(t.seg_len + ci.s);
        ci.adj = __ldg(t.seg_adj + ci.s);
        ci.cbase = ci.start / SORT_CHUNK;
        ci.nch = (next - ci.start) / SORT_CHUNK;
        ci.c = b - ci.cbase;
        const uint32_t done = ci.c * SORT_CHUNK;
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
// [off[i-1], off[i]) and enumerates the tiles of its rectangle row-major, like the reference.  The inclusive
// offsets of the chunk's entries (at most one entry per duplicate) are staged in shared memory; every thread
// then finds the owner of each of its items by binary search.  Its items are 32 duplicates apart and every
// entry owns at least one duplicate, so the owner moves by at most 32 entries from one item to the next: after
// the first item the search window is 33 entries wide.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void generate_items(const GenParams& gp, const ChunkInfo& ci, uint32_t* s_off /*[SORT_CHUNK + 32]*/,
                                               unsigned tid, unsigned warp, unsigned lane, uint32_t (&tile)[SORT_ITEMS],
                                               uint32_t (&gid)[SORT_ITEMS])
{
    const uint32_t r0 = ci.c * SORT_CHUNK, r1 = r0 + ci.nv;
    const size_t seg = (size_t)ci.s * gp.entry_stride;
    const uint32_t* __restrict__ off = gp.off + seg;
    const uint2* __restrict__ rec = gp.rec + seg;
    const uint32_t i0 = ci.first_entry;
    const uint32_t excl0 = i0 ? __ldg(off + i0 - 1) : 0u;
    // s_off[k] = inclusive offset of entry i0 + k, staged until one reaches the end of the chunk; the 32 slots behind
    // the last staged entry are filled with "infinity" so that the window reads below need no bound checks
    uint32_t staged = SORT_CHUNK;
    for (uint32_t k0 = 0; k0 < SORT_CHUNK; k0 += SORT_THREADS) {
        const uint32_t k = k0 + tid, i = i0 + k;
        const uint32_t v = i < gp.entries_per_seg ? __ldg(off + i) : 0xFFFFFFFFu;
        s_off[k] = v;
        if (__syncthreads_or(v >= r1)) { staged = k0 + SORT_THREADS; break; }
    }
    if (tid < 32) s_off[staged + tid] = 0xFFFFFFFFu;
    __syncthreads();

    // owner of the warp's first duplicate: binary search (warp-uniform); its index is at most the duplicate's
    const uint32_t jw = warp * (32 * SORT_ITEMS);
    uint32_t eA = 0;
    if (jw < ci.nv) {
        uint32_t lo = 0, hi = min(jw, staged - 1);
        const uint32_t d = r0 + jw;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] > d) hi = mid; else lo = mid + 1;
        }
        eA = lo;
    }
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        // The 32 lanes hold 32 consecutive duplicates d0 .. d0+31; eA owns d0.  Every entry owns at least one
        // duplicate, so the owners are eA .. eA+31 at most: lane l looks at where entry eA+l+1 STARTS (= the inclusive
        // offset of eA+l); starts inside (d0, d0+31] become head bits, and a duplicate's owner is eA + the number
        // of heads at or before it.
        const uint32_t j = jw + it * 32 + lane;
        const uint32_t d0 = r0 + jw + it * 32;
        const uint32_t start_next = s_off[eA + lane];                 // start of entry eA + lane + 1
        const uint32_t rel = start_next - d0;                         // > 0 for lane 0 because eA owns d0
        const unsigned heads = __reduce_or_sync(FULL_MASK, rel < 32u ? (1u << rel) : 0u);
        const uint32_t e = eA + __popc(heads & (0xFFFFFFFFu >> (31u - lane)));
        tile[it] = 0u;
        gid[it] = 0u;
        if (j < ci.nv) {
            const uint32_t excl = e ? s_off[e - 1] : excl0;
            const uint2 r = __ldg(rec + i0 + e);
            const uint32_t x0 = r.x & 1023u, y0 = (r.x >> 10) & 1023u, w = ((r.x >> 20) & 1023u) + 1u;
            const uint32_t jj = d0 + lane - excl;
            // row-major over the rectangle; (jj + 0.5) / w is never within 0.5 / w of an integer, so the float
            // quotient (2 ulp) truncates to floor(jj / w) exactly for any rectangle of a <= 16K x 16K image
            const uint32_t row = (uint32_t)__fdividef((float)jj + 0.5f, (float)w);
            tile[it] = (y0 + row) * (uint32_t)gp.tiles_x + x0 + (jj - row * w);
            gid[it] = r.y;
        }
        // owner of the next round's first duplicate d0 + 32: the last lane's owner, or its successor if that entry ends here
        const uint32_t e31 = eA + __popc(heads);
        eA = e31 + (s_off[e31] <= d0 + 32u ? 1u : 0u);
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
    __shared__ uint32_t s_off[GEN ? SORT_CHUNK + 32 : 1];
    if (blockIdx.x >= total_chunks(t)) return;
    const ChunkInfo ci = locate_chunk(t, blockIdx.x);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (int k = tid; k < SORT_WARPS * BINS; k += SORT_THREADS) (&hist[0][0])[k] = 0u;
    if (GEN) {
        uint32_t tile[SORT_ITEMS], gid[SORT_ITEMS];
        generate_items(gp, ci, s_off, tid, warp, lane, tile, gid);      // contains barriers: hist is zeroed
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
            const uint32_t j = warp * (32 * SORT_ITEMS) + i * 32 + lane;
            if (j < ci.nv) atomicAdd(&hist[warp][(tile[i] >> shift) & (BINS - 1)], 1u);
        }
    } else {
        __syncthreads();
        const uint32_t* __restrict__ src =
            keys_in + (in_seg_stride ? (size_t)ci.s * in_seg_stride : (size_t)ci.start) + (size_t)ci.c * SORT_CHUNK;
        uint32_t k[SORT_ITEMS];
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {           // all loads in flight before the first histogram update
            const uint32_t j = i * SORT_THREADS + tid;
            k[i] = j < ci.nv ? __ldg(src + j) : 0u;
        }
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
            const uint32_t j = i * SORT_THREADS + tid;
            if (j < ci.nv) atomicAdd(&hist[warp][(k[i] >> shift) & (BINS - 1)], 1u);
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
    const uint32_t base = blockIdx.x * SCAN_SLICE + tid * SCAN_ITEMS;     // blocked: 16 consecutive counters per thread
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; q++) {
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
    for (int i = 0; i < SCAN_ITEMS; i++) { const uint32_t x = v[i]; v[i] = sum; sum += x; }
    uint32_t total;
    const uint32_t ex = block_exclusive_scan_256(sum, s_warp, tid, total);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; q++) {
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
__global__ void __launch_bounds__(SORT_THREADS, 4) k_sort_downsweep(
    const SegTable t, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t in_seg_stride,
    int shift, const uint32_t* __restrict__ counters, const uint32_t* __restrict__ slice_base,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, const GenParams gp, uint2* __restrict__ ranges,
    uint32_t tiles_per_seg)
{
    constexpr int BINS = 1 << BITS;
    __shared__ uint32_t s_keys[2 * SORT_CHUNK];          // keys | values; GEN: first the staged entry offsets
    uint32_t* const s_vals = s_keys + SORT_CHUNK;
    __shared__ uint32_t s_wc[SORT_WARPS][BINS + 1];      // per-warp digit counters (+1: invalid items)
    __shared__ uint32_t s_gdelta[BINS];                  // global position - block-local sorted position, per digit
    __shared__ uint32_t s_scan[SORT_WARPS + 1];
    if (blockIdx.x >= total_chunks(t)) return;
    const ChunkInfo ci = locate_chunk(t, blockIdx.x);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // global base of this chunk's digit `tid` (scanned counter + slice base): requested now, used after the ranking
    uint32_t gbase = 0;
    if (tid < BINS) {
        const size_t ci_idx = (size_t)ci.cbase * BINS + (size_t)tid * ci.nch + ci.c;
        gbase = __ldg(counters + ci_idx) + __ldg(slice_base + ci_idx / SCAN_SLICE) + ci.adj;
    }
    for (int k = tid; k < SORT_WARPS * (BINS + 1); k += SORT_THREADS) (&s_wc[0][0])[k] = 0u;

    uint32_t key[SORT_ITEMS], val[SORT_ITEMS];
    if (GEN) {
        generate_items(gp, ci, s_keys, tid, warp, lane, key, val);
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
    __syncthreads();     // counters zeroed; GEN: the staged offsets are dead (s_keys is reused below)

    // ---- rank inside the warp.  Which lanes share my digit: one ballot per digit bit (the match.any instruction
    // iterates over the distinct values of the warp, up to 32 of them for a random 8-bit digit; ballots do not).
    unsigned peers[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t j = warp * (32 * SORT_ITEMS) + i * 32 + lane;
        const uint32_t d = (key[i] >> shift) & (BINS - 1);
        unsigned m = __ballot_sync(FULL_MASK, j < ci.nv);          // invalid items (tail of a segment) rank nowhere
#pragma unroll
        for (int b = 0; b < BITS; b++) {
            const bool bit = (d >> b) & 1u;
            const unsigned v = __ballot_sync(FULL_MASK, bit);
            m &= bit ? v : ~v;
        }
        peers[i] = m;
    }
    // One shared-memory atomic per (round, digit group), issued by the group's lowest lane.  The rounds are issued
    // in order by the converged warp and atomics on one address are applied in issue order, so nobody has to wait
    // for a returned value before the next round goes out: the eight round trips overlap.
    uint32_t rank[SORT_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t j = warp * (32 * SORT_ITEMS) + i * 32 + lane;
        const uint32_t d = (key[i] >> shift) & (BINS - 1);
        rank[i] = 0u;
        if (j < ci.nv && lane == (unsigned)(__ffs(peers[i]) - 1))
            rank[i] = atomicAdd(&s_wc[warp][d], (uint32_t)__popc(peers[i]));
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++)   // (invalid lanes: peers may be empty, source lane 31 & garbage -- never used)
        rank[i] = __shfl_sync(FULL_MASK, rank[i], (__ffs(peers[i]) - 1) & 31) + __popc(peers[i] & lt_mask);
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
        s_gdelta[tid] = gbase - bin_start;
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

// one block per segment: exclusive scan of the segment's block sums; the last block to finish (ticket) builds the
// segment table and the status words from the segment totals
__global__ void __launch_bounds__(256) k_seg_scan(int nseg, uint32_t nb, const unsigned long long* __restrict__ block_sums,
                                                  uint32_t* __restrict__ block_excl, BinStatus* __restrict__ status,
                                                  uint32_t* __restrict__ seg_start, uint32_t* __restrict__ seg_len,
                                                  uint32_t* __restrict__ seg_adj, unsigned long long* __restrict__ seg_total,
                                                  uint32_t* ticket, unsigned long long capacity)
{
    __shared__ unsigned long long s_warp[9];
    __shared__ bool s_last;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int s = blockIdx.x;
    unsigned long long carry = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += 256) {
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
            for (int w = 0; w < 8; w++) { const unsigned long long x = s_warp[w]; s_warp[w] = acc; acc += x; }
            s_warp[8] = acc;
        }
        __syncthreads();
        if (b < nb) block_excl[(size_t)s * nb + b] = (uint32_t)(carry + s_warp[warp] + inc - v);
        carry += s_warp[8];
        __syncthreads();
    }
    if (tid == 0) {
        seg_total[s] = carry;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == (unsigned)nseg - 1u;
    }
    __syncthreads();
    if (!s_last || tid != 0) return;
    __threadfence();
    unsigned long long D = 0, padded = 0;
    bool too_big = false;
    for (int q = 0; q < nseg; q++) {
        const unsigned long long len = __ldcg(seg_total + q);
        if (len >= 0xFFFFFFFFull) too_big = true;
        seg_start[q] = (uint32_t)padded;
        seg_len[q] = (uint32_t)len;
        seg_adj[q] = (uint32_t)(padded - D);
        D += len;
        padded += (len + SORT_CHUNK - 1) / SORT_CHUNK * SORT_CHUNK;
    }
    seg_start[nseg] = (uint32_t)padded;
    const bool overflow = too_big || padded > capacity || padded >= 0xFFFFFFFFull;
    status->num_rendered = D;
    status->padded = padded;
    status->overflow = overflow ? 1u : 0u;
    status->n_chunks = overflow ? 0u : (uint32_t)(padded / SORT_CHUNK);   // overflow: stage 2 does nothing
    *ticket = 0u;
}

__global__ void __launch_bounds__(256) k_entry_offsets(int P, uint32_t stride, const uint32_t* __restrict__ cnt_sorted,
                                                       const uint32_t* __restrict__ block_excl,
                                                       const uint32_t* __restrict__ seg_start,
                                                       const BinStatus* __restrict__ status, uint32_t* __restrict__ off,
                                                       uint2* __restrict__ chunk_tab)
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
                chunk_tab[cbase + q] = make_uint2(i0 + k, (uint32_t)s);
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
    k_seg_scan<<<p.F, 256, 0, st>>>(p.F, nb, b.block_sums, b.block_excl, b.status, b.seg_start, b.seg_len, b.seg_adj,
                                    b.seg_total, b.ticket, capacity);
}
void launch_entry_offsets(const FwdParams& p, const BinState& b, uint2* chunk_tab, cudaStream_t st)
{
    const uint32_t nb = (uint32_t)((p.P + ENT_BLOCK - 1) / ENT_BLOCK);
    dim3 grid(nb, p.F);
    k_entry_offsets<<<grid, 256, 0, st>>>(p.P, b.stride, b.cnt_sorted, b.block_excl, b.seg_start, b.status, b.off,
                                          chunk_tab);
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

// binning.cu — K4 tile binning, sm_100a.  The reference has no binning: GL's rasteriser assigns
// fragments to pixels implicitly (/root/reference/gsplat_plugin/src/GSplatRenderer.C:647).  Here the
// splats of one depth chunk that still touch a live (owned, un-saturated) tile are compacted, sorted
// by depth and expanded, in that order, into one instance per 16x16 tile their pixel rectangle touches
// (SURVEY.md A.8); radix_sort.cu then stably partitions the instances by tile id, which leaves every
// tile's list in depth order.  Integer / byte work, HBM-bound.
#include "common.cuh"

namespace gsb {

namespace {

struct TileRect { int tx0, tx1, ty0, ty1; bool empty; };

__device__ __forceinline__ TileRect tile_rect(uint2 r)
{
    TileRect t;
    const int x0 = (int)(r.x & 0xffffu), x1 = (int)(r.x >> 16), y0 = (int)(r.y & 0xffffu), y1 = (int)(r.y >> 16);
    t.empty = x0 > x1;
    t.tx0 = x0 / TILE; t.tx1 = x1 / TILE; t.ty0 = y0 / TILE; t.ty1 = y1 / TILE;
    return t;
}

// tile rectangle from the packed word (has_packed false: exact rectangle of splat i).  A saturated ("wide") extent is
// resolved from the exact rectangle, or — bounded K1, which keeps no exact rectangles (rects == NULL) — widened to the
// whole screen: still a superset.
__device__ __forceinline__ TileRect tile_rect_packed(const bool has_packed, const uint32_t p, const uint2* __restrict__ rects,
                                                     const int64_t i, const int tiles_x, const int tiles_y)
{
    if (!has_packed) return tile_rect(__ldg(rects + i));
    TileRect t;
    t.empty = p == TRECT_CULLED;
    const int w = (int)((p >> 18) & 127u), h = (int)((p >> 25) & 127u);
    if (!t.empty && (w == 127 || h == 127)) {
        if (rects) return tile_rect(__ldg(rects + i));
        t.tx0 = 0; t.tx1 = tiles_x - 1; t.ty0 = 0; t.ty1 = tiles_y - 1;
        return t;
    }
    t.tx0 = (int)(p & 511u); t.ty0 = (int)((p >> 9) & 511u); t.tx1 = t.tx0 + w; t.ty1 = t.ty0 + h;
    return t;
}

// summed-area table of the live map, one CTA: row prefixes (a warp per row), then column sums (a thread per column;
// the loads do not depend on the running sum, so they pipeline)
__global__ void __launch_bounds__(1024)
live_sat_kernel(const uint32_t* __restrict__ done, int tiles_x, int tiles_y, int row_rank, int row_world, int row_group,
                uint32_t* __restrict__ sat)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int st = tiles_x + 1, wpr = done_words_per_row(tiles_x);
    for (int x = threadIdx.x; x < st; x += 1024) sat[x] = 0u;
    for (int y = warp; y < tiles_y; y += 32) {
        const bool owned = owns_row(y, row_rank, row_world, row_group);
        uint32_t carry = 0;
        if (lane == 0) sat[(y + 1) * st] = 0u;
        for (int w = 0; w < wpr; ++w) {
            const int x = w * 32 + lane;
            const uint32_t dw = done ? done[y * wpr + w] : 0u;
            const uint32_t livew = owned ? ~dw : 0u;
            const uint32_t upto = __popc(livew & (0xffffffffu >> (31 - lane)));      // live tiles of this word at columns <= x
            if (x < tiles_x) sat[(y + 1) * st + x + 1] = carry + upto;
            carry += __popc(livew);          // columns >= tiles_x are never read
        }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < st; x += 1024) {
        uint32_t run = 0;
#pragma unroll 8
        for (int y = 1; y <= tiles_y; ++y) { run += sat[y * st + x]; sat[y * st + x] = run; }
    }
}

// ---- live selection, per depth chunk, over the submitted splats (which are never moved).  Element k is selected if its
// depth key lies in the chunk's key interval [key_lo, key_hi) (the chunk plan turns its bucket boundaries into key
// boundaries, so membership is two integer compares) and it touches at least one live tile.  Three kernels, no spin-waits:
//   A  select_count:  one CTA per 2048 elements streams the keys and packed tile rectangles (8 B / splat, all loads
//      issued before any is used), compacts the selected (key, splat index) pairs order-preservingly
//      INSIDE the CTA and stores them as one contiguous run at the CTA's own slot of a staging area
//      (stage[tile * 2048 ..]); writes the tile's totals.
//   S  select_scan:   one CTA turns the per-tile totals into exclusive bases and the grand totals L and D.
//   B  select_gather: moves every tile's run to base[tile]: contiguous reads, contiguous writes, 8 B per SELECTED
//      element.  Submission order is kept, so the stable depth sort that follows breaks ties by ascending index.
// (r01 measured three other forms first.  A single-pass chained scan: with ~5000 tiles in flight the decoupled look-back
// chains grew to the number of resident CTAs, 260-320 us per pass.  Loading the rectangle only for the chunk's own
// elements: the dependent, divergent loads serialised, 130-210 us.  Ballot words + a second pass that re-reads keys and
// rectangles of the selected elements: the sparse 32-byte-sector gathers cost 81 us per chunk at 20 M, DRAM-latency bound.)
constexpr int SEL_THREADS = 256;
constexpr int SEL_WARPS   = SEL_THREADS / 32;
constexpr int SEL_ITEMS   = 8;
constexpr int SEL_TILE    = SEL_THREADS * SEL_ITEMS;     // 2048 elements per CTA

__global__ void __launch_bounds__(SEL_THREADS)
select_count_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ trects,
                    const uint2* __restrict__ rects, int64_t n, const ChunkPlan* __restrict__ plan, const int chunk,
                    int tiles_x, int tiles_y, const uint32_t* __restrict__ sat,
                    uint32_t* __restrict__ stage_k, uint32_t* __restrict__ stage_v,
                    uint32_t* __restrict__ tile_l, uint32_t* __restrict__ tile_d)
{
    __shared__ uint32_t s_wl[SEL_WARPS], s_wd[SEL_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    const int64_t base = (int64_t)tile * SEL_TILE + (int64_t)warp * (32 * SEL_ITEMS);
    const uint32_t key_lo = plan ? __ldg(&plan->key_lo[chunk]) : 0u;
    const uint32_t key_hi = plan ? __ldg(&plan->key_lo[chunk + 1]) : KEY_CULLED;

    // warp-striped items: item j of lane l is element base + j*32 + l, so (j, lane) order is element order
    uint32_t key[SEL_ITEMS], tr[SEL_ITEMS];
#pragma unroll
    for (int j = 0; j < SEL_ITEMS; ++j) {
        const int64_t k = base + j * 32 + lane;
        key[j] = (k < n) ? __ldg(keys + k) : KEY_CULLED;
        tr[j]  = (k < n && trects) ? __ldg(trects + k) : 0u;
    }
    const unsigned lt = (1u << lane) - 1u;
    uint32_t dsum = 0, wcount = 0, selbits = 0;
    uint32_t pos[SEL_ITEMS];                            // rank of item j among the warp's selected elements
#pragma unroll
    for (int j = 0; j < SEL_ITEMS; ++j) {
        uint32_t c = 0;
        if (key[j] >= key_lo && key[j] < key_hi) {      // key_hi <= KEY_CULLED: culled splats never pass
            const TileRect t = tile_rect_packed(trects != nullptr, tr[j], rects, base + j * 32 + lane, tiles_x, tiles_y);
            if (!t.empty) c = live_tiles(t.tx0, t.tx1, t.ty0, t.ty1, tiles_x, sat);
        }
        dsum += c;
        const unsigned m = __ballot_sync(0xffffffffu, c != 0u);
        pos[j] = wcount + __popc(m & lt);
        selbits |= (c != 0u ? 1u : 0u) << j;
        wcount += __popc(m);
    }
    dsum = __reduce_add_sync(0xffffffffu, dsum);
    if (lane == 0) { s_wl[warp] = wcount; s_wd[warp] = dsum; }
    __syncthreads();
    uint32_t wbase = 0, l = 0, d = 0;
#pragma unroll
    for (int w = 0; w < SEL_WARPS; ++w) { wbase += (w < warp) ? s_wl[w] : 0u; l += s_wl[w]; d += s_wd[w]; }
    if (threadIdx.x == 0) { tile_l[tile] = l; tile_d[tile] = d; }
    const size_t sb = (size_t)tile * SEL_TILE + wbase;
#pragma unroll
    for (int j = 0; j < SEL_ITEMS; ++j) {
        if ((selbits >> j) & 1u) {
            const size_t o = sb + pos[j];
            stage_k[o] = key[j]; stage_v[o] = (uint32_t)(base + j * 32 + lane);
        }
    }
}

// one CTA: tile_base = exclusive scan of tile_l; *l_total, *d_total = grand totals.  Every thread owns a run of
// consecutive entries (sum, one block-wide scan of the 1024 sums, write-back), so the kernel is three short phases
// instead of one barrier round per 1024 entries.
__global__ void __launch_bounds__(1024)
select_scan_kernel(const uint32_t* __restrict__ tile_l, const uint32_t* __restrict__ tile_d, uint32_t nt,
                   uint32_t* __restrict__ tile_base, unsigned long long* __restrict__ l_total,
                   unsigned long long* __restrict__ d_total)
{
    __shared__ uint32_t s_w[32];
    __shared__ unsigned long long s_d[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (nt + 1023u) / 1024u;
    const uint32_t b0 = min(threadIdx.x * per, nt), b1 = min(b0 + per, nt);
    uint32_t mine = 0;
    unsigned long long dacc = 0;
    for (uint32_t i = b0; i < b1; ++i) { mine += tile_l[i]; dacc += (unsigned long long)tile_d[i]; }
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
    if (lane == 31) s_w[warp] = inc;
    if (lane == 0) s_d[warp] = dacc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) { woff += (w < warp) ? s_w[w] : 0u; total += s_w[w]; }
    uint32_t run = woff + inc - mine;
    for (uint32_t i = b0; i < b1; ++i) { tile_base[i] = run; run += tile_l[i]; }
    if (threadIdx.x == 0) {
        unsigned long long d = 0;
        for (int w = 0; w < 32; ++w) d += s_d[w];
        *l_total = (unsigned long long)total; *d_total = d;
    }
}

// one warp per selection tile: copy its run of selected pairs from the staging slot to base[tile]
constexpr int GATHER_THREADS = 256;
__global__ void __launch_bounds__(GATHER_THREADS)
select_gather_kernel(const uint32_t* __restrict__ stage_k, const uint32_t* __restrict__ stage_v,
                     const uint32_t* __restrict__ tile_l, const uint32_t* __restrict__ tile_base, uint32_t nt,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.x * (GATHER_THREADS / 32) + (threadIdx.x >> 5);
    if (tile >= nt) return;
    const uint32_t l = __ldg(tile_l + tile);
    if (l == 0u) return;
    const size_t src = (size_t)tile * SEL_TILE, dst = (size_t)__ldg(tile_base + tile);
    for (uint32_t i = lane; i < l; i += 32) {
        keys_out[dst + i] = __ldg(stage_k + src + i); vals_out[dst + i] = __ldg(stage_v + src + i);
    }
}

// instance (tile id, live rank) pairs at offsets[k] .., rows ascending then columns ascending, live tiles only; the digit
// histograms of the tile partition that follows are built on the way (shared-memory bins, one flush per CTA), so the
// partition needs no pass of its own over the instance keys
struct EmitSort { int shift[SORT_MAX_PASSES]; int bits[SORT_MAX_PASSES]; int passes; };
__global__ void __launch_bounds__(256)
emit_kernel(const uint2* __restrict__ tile_rects, const uint32_t* __restrict__ offsets,
            const unsigned long long* __restrict__ total,
            int64_t n, int tiles_x, int row_rank, int row_world, int row_group,
            const uint32_t* __restrict__ tile_done, uint32_t* __restrict__ inst_keys, uint32_t* __restrict__ inst_vals,
            const EmitSort es, uint32_t* __restrict__ tile_hist)
{
    __shared__ uint32_t sh_hist[SORT_MAX_PASSES][SORT_RADIX];
    if (tile_hist) {
        for (int i = threadIdx.x; i < es.passes * SORT_RADIX; i += 256) (&sh_hist[0][0])[i] = 0u;
        __syncthreads();
    }
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t o0 = 0, o1 = 0;
    if (k < n) {
        o0 = offsets[k];
        o1 = (k + 1 < n) ? offsets[k + 1] : (uint32_t)(*total);
    }
    if (o1 != o0) {                                              // else: every tile it touches is saturated
        const uint2 tr = __ldg(tile_rects + k);
        const int tx0 = (int)(tr.x & 0xffffu), tx1 = (int)(tr.x >> 16), ty0 = (int)(tr.y & 0xffffu), ty1 = (int)(tr.y >> 16);
        size_t o = o0;
        const int wpr = done_words_per_row(tiles_x);
        for (int ty = ty0; ty <= ty1; ++ty) {
            if (!owns_row(ty, row_rank, row_world, row_group)) continue;
            for (int w = tx0 >> 5; w <= (tx1 >> 5); ++w) {
                uint32_t live;
                if (tile_done) live = live_word(tile_done, wpr, ty, w, tx0, tx1);
                else {
                    live = 0xffffffffu;
                    if (w == (tx0 >> 5)) live &= 0xffffffffu << (tx0 & 31);
                    if (w == (tx1 >> 5)) live &= 0xffffffffu >> (31 - (tx1 & 31));
                }
                while (live) {                                   // ascending columns
                    const int b = __ffs(live) - 1;
                    live &= live - 1;
                    const uint32_t id = (uint32_t)(ty * tiles_x + w * 32 + b);
                    inst_keys[o] = id; inst_vals[o] = (uint32_t)k; ++o;
                    if (tile_hist) {
#pragma unroll
                        for (int ps = 0; ps < SORT_MAX_PASSES; ++ps)
                            if (ps < es.passes) atomicAdd(&sh_hist[ps][(id >> es.shift[ps]) & ((1u << es.bits[ps]) - 1u)], 1u);
                    }
                }
            }
        }
    }
    if (tile_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < es.passes * SORT_RADIX; i += 256) {
            const uint32_t v = (&sh_hist[0][0])[i];
            if (v) atomicAdd(tile_hist + i, v);
        }
    }
}

// ranges[t] = [first, last + 1) of tile t's run in the tile-sorted instance list; four ids per thread (one 16-byte
// load plus the two neighbours)
__global__ void __launch_bounds__(256)
tile_range_kernel(const uint32_t* __restrict__ ids, const uint64_t d_max, const unsigned long long* __restrict__ d_dev,
                  uint2* __restrict__ ranges)
{
    const uint64_t d = d_dev ? min((unsigned long long)d_max, *d_dev) : d_max;      // the exact count lives on the device
    const uint64_t j0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= d) return;
    uint32_t v[6];                                           // ids[j0 - 1 .. j0 + 4]; 0xFFFFFFFF outside the list
    if (j0 + 4 <= d) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(ids + j0));
        v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[1 + k] = (j0 + k < d) ? __ldg(ids + j0 + k) : 0xFFFFFFFFu;
    }
    v[0] = j0 > 0 ? __ldg(ids + j0 - 1) : 0xFFFFFFFFu;
    v[5] = (j0 + 4 < d) ? __ldg(ids + j0 + 4) : 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t j = j0 + k;
        if (j >= d) break;
        const uint32_t t = v[1 + k];
        if (v[k] != t) ranges[t].x = (uint32_t)j;
        if (v[2 + k] != t) ranges[t].y = (uint32_t)(j + 1);
    }
}

// The spec's depth order is ascending key, ties by ascending splat index.  The live list reaches the depth
// sort in cell order, not index order, so after the sort every run of equal keys is put in index order: element j of a run
// [a, b) goes to a + (number of run members with a smaller index).  Runs are short (two or three splats at equal fp32
// distance); the scan for the run's ends is bounded at TIE_MAX on either side — a run of more than TIE_MAX bit-identical
// distances keeps the order the sort left it in (the reference's own sort leaves ties unspecified).
constexpr int TIE_MAX = 1024;
__global__ void __launch_bounds__(256)
tie_fix_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint64_t l_max,
               const unsigned long long* __restrict__ l_dev, uint32_t* __restrict__ vals_out)
{
    const uint64_t l = l_dev ? min((unsigned long long)l_max, *l_dev) : l_max;
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= l) return;
    const uint32_t k = __ldg(keys + j), v = __ldg(vals + j);
    const bool tie_l = j > 0 && __ldg(keys + j - 1) == k, tie_r = j + 1 < l && __ldg(keys + j + 1) == k;
    if (!tie_l && !tie_r) { vals_out[j] = v; return; }
    uint64_t a = j, b = j + 1;
    uint32_t smaller = 0;
    for (int t = 0; t < TIE_MAX && a > 0 && __ldg(keys + a - 1) == k; ++t) { --a; smaller += (__ldg(vals + a) < v) ? 1u : 0u; }
    for (int t = 0; t < TIE_MAX && b < l && __ldg(keys + b) == k; ++t) { smaller += (__ldg(vals + b) < v) ? 1u : 0u; ++b; }
    // b - a <= TIE_MAX: neither scan hit its bound, [a, b) is the whole run and every member sees the same run, so the
    // destinations are a permutation of it.  Otherwise the run is longer than TIE_MAX for every member: all stay put.
    vals_out[(b - a <= (uint64_t)TIE_MAX) ? a + smaller : j] = v;
}

__global__ void __launch_bounds__(256)
debug_records_kernel(const Record* __restrict__ recs, const uint32_t* __restrict__ live_splats, int64_t n_live,
                     Record* __restrict__ recs_by_splat)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_live) recs_by_splat[live_splats[j]] = recs[j];
}

__global__ void __launch_bounds__(256)
debug_instances_kernel(const uint32_t* __restrict__ inst_refs, const uint32_t* __restrict__ live_splats, uint64_t d,
                       uint32_t* __restrict__ inst_splats)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < d) inst_splats[j] = live_splats[inst_refs[j]];
}

}  // namespace

void launch_live_sat(FrameConsts fc, const uint32_t* tile_done, uint32_t* sat, cudaStream_t s)
{
    live_sat_kernel<<<1, 1024, 0, s>>>(tile_done, fc.tiles_x, fc.tiles_y, fc.row_rank, fc.row_world, fc.row_group, sat);
}

static inline size_t sel_tiles(int64_t n) { return (size_t)((n + SEL_TILE - 1) / SEL_TILE); }

size_t select_scratch_bytes(int64_t n) { return sel_tiles(n) * 3 * sizeof(uint32_t) + 64; }
size_t select_stage_elems(int64_t n) { return sel_tiles(n) * SEL_TILE; }

void launch_select_live(const uint32_t* keys, const uint32_t* trects, const uint2* rects, int64_t n,
                        const ChunkPlan* plan, int chunk,
                        FrameConsts fc, const uint32_t* sat, uint32_t* keys_out, uint32_t* vals_out,
                        uint32_t* stage_k, uint32_t* stage_v,
                        void* scratch, unsigned long long* l_total, unsigned long long* d_total,
                        cudaStream_t s)
{
    if (n <= 0) return;
    const unsigned nt = (unsigned)sel_tiles(n);
    uint32_t* tile_l = static_cast<uint32_t*>(scratch);
    uint32_t* tile_d = tile_l + nt;
    uint32_t* tile_base = tile_d + nt;
    select_count_kernel<<<nt, SEL_THREADS, 0, s>>>(keys, trects, rects, n, plan, chunk, fc.tiles_x, fc.tiles_y, sat,
                                                   stage_k, stage_v, tile_l, tile_d);
    select_scan_kernel<<<1, 1024, 0, s>>>(tile_l, tile_d, nt, tile_base, l_total, d_total);
    select_gather_kernel<<<(nt + GATHER_THREADS / 32 - 1) / (GATHER_THREADS / 32), GATHER_THREADS, 0, s>>>(
        stage_k, stage_v, tile_l, tile_base, nt, keys_out, vals_out);
}

void launch_emit(const uint2* tile_rects, const uint32_t* offsets,
                 const unsigned long long* total, int64_t n, FrameConsts fc, const uint32_t* tile_done,
                 uint32_t* inst_keys, uint32_t* inst_vals, const SortPlan& tile_plan, uint32_t* tile_hist, cudaStream_t s)
{
    if (n <= 0) return;
    EmitSort es{};
    es.passes = tile_plan.passes;
    for (int p = 0; p < tile_plan.passes; ++p) { es.shift[p] = tile_plan.shift[p]; es.bits[p] = tile_plan.bits[p]; }
    emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tile_rects, offsets, total, n, fc.tiles_x,
                                                           fc.row_rank, fc.row_world, fc.row_group, tile_done, inst_keys, inst_vals,
                                                           es, tile_hist);
}

void launch_tile_ranges(const uint32_t* sorted_tile_ids, uint64_t d_max, const unsigned long long* d_dev, uint2* ranges,
                        int num_tiles, cudaStream_t s)
{
    cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s);
    if (d_max == 0) return;
    tile_range_kernel<<<(unsigned)((d_max + 1023) / 1024), 256, 0, s>>>(sorted_tile_ids, d_max, d_dev, ranges);
}

void launch_tie_fix(const uint32_t* keys_sorted, const uint32_t* vals_sorted, uint64_t l_max, const unsigned long long* l_dev,
                    uint32_t* vals_out, cudaStream_t s)
{
    if (l_max == 0) return;
    tie_fix_kernel<<<(unsigned)((l_max + 255) / 256), 256, 0, s>>>(keys_sorted, vals_sorted, l_max, l_dev, vals_out);
}

void launch_debug_views(const Record* recs, const uint32_t* live_splats, int64_t n_live, Record* recs_by_splat,
                        const uint32_t* inst_refs, uint64_t d, uint32_t* inst_splats, cudaStream_t s)
{
    if (n_live > 0) debug_records_kernel<<<(unsigned)((n_live + 255) / 256), 256, 0, s>>>(recs, live_splats, n_live, recs_by_splat);
    if (d > 0) debug_instances_kernel<<<(unsigned)((d + 255) / 256), 256, 0, s>>>(inst_refs, live_splats, d, inst_splats);
}

}  // namespace gsb

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

// tile rectangle of element r: from the packed payload, or (wide splats / huge screens) from the exact rectangle
__device__ __forceinline__ TileRect tile_rect_of(const uint32_t* __restrict__ trects, const uint32_t* __restrict__ order,
                                                 const uint2* __restrict__ rects, int64_t r)
{
    if (!trects) {
        const uint32_t i = __ldg(order + r);
        TileRect t; t.empty = true; t.tx0 = t.ty0 = 1; t.tx1 = t.ty1 = 0;
        return i == 0xFFFFFFFFu ? t : tile_rect(__ldg(rects + i));
    }
    const uint32_t p = __ldg(trects + r);
    TileRect t;
    t.empty = p == TRECT_CULLED;
    const int w = (int)((p >> 18) & 127u), h = (int)((p >> 25) & 127u);
    if (!t.empty && (w == 127 || h == 127)) return tile_rect(__ldg(rects + __ldg(order + r)));
    t.tx0 = (int)(p & 511u); t.ty0 = (int)((p >> 9) & 511u); t.tx1 = t.tx0 + w; t.ty1 = t.ty0 + h;
    return t;
}

// counts[k] = number of live tiles touched by element r0 + k (0 for culled splats); a coalesced 4-byte stream.
// A tile is live if this rank owns its row and it is not yet saturated (tile_done, set by the blend
// of an earlier depth chunk): instances behind a saturated tile can never change a pixel.
__global__ void __launch_bounds__(256)
tile_count_kernel(const uint32_t* __restrict__ trects, const uint32_t* __restrict__ order,
                  const uint2* __restrict__ rects, int64_t r0, int64_t n,
                  int tiles_x, int row_rank, int row_world, int row_group, const uint32_t* __restrict__ tile_done,
                  uint32_t* __restrict__ counts, unsigned long long* __restrict__ d_total)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (k < n) {
        const TileRect t = tile_rect_of(trects, order, rects, r0 + k);
        if (!t.empty) {
            for (int ty = t.ty0; ty <= t.ty1; ++ty) {
                if (!owns_row(ty, row_rank, row_world, row_group)) continue;
                if (!tile_done) c += (uint32_t)(t.tx1 - t.tx0 + 1);
                else for (int tx = t.tx0; tx <= t.tx1; ++tx) c += (__ldg(tile_done + ty * tiles_x + tx) == 0u) ? 1u : 0u;
            }
        }
        counts[k] = c;
    }
    if (d_total) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, c);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(d_total, (unsigned long long)w);
    }
}

// survivors (counts != 0) keep their relative order: positions = exclusive scan of the flags
__global__ void __launch_bounds__(256)
compact_live_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ trects,
                    const uint32_t* __restrict__ counts, const uint32_t* __restrict__ positions, int64_t n,
                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t* __restrict__ trects_out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (__ldg(counts + k) != 0u) {
        const uint32_t p = __ldg(positions + k);
        keys_out[p] = __ldg(keys + k); vals_out[p] = __ldg(vals + k);
        if (trects) trects_out[p] = __ldg(trects + k);
    }
}

// instance (tile id, live rank) pairs at offsets[k] .., rows ascending then columns ascending
__global__ void __launch_bounds__(256)
emit_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ trects, const uint2* __restrict__ rects,
            const uint32_t* __restrict__ offsets, const unsigned long long* __restrict__ total,
            int64_t n, int tiles_x, int row_rank, int row_world, int row_group,
            const uint32_t* __restrict__ tile_done, uint32_t* __restrict__ inst_keys, uint32_t* __restrict__ inst_vals)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t o0 = offsets[k];
    const uint32_t o1 = (k + 1 < n) ? offsets[k + 1] : (uint32_t)(*total);
    if (o1 == o0) return;                                    // culled, or every tile it touches is saturated
    const TileRect t = tile_rect_of(trects, order, rects, k);
    size_t o = o0;
    for (int ty = t.ty0; ty <= t.ty1; ++ty) {
        if (!owns_row(ty, row_rank, row_world, row_group)) continue;
        for (int tx = t.tx0; tx <= t.tx1; ++tx) {
            const uint32_t tile = (uint32_t)(ty * tiles_x + tx);
            if (tile_done && __ldg(tile_done + tile) != 0u) continue;
            inst_keys[o] = tile;
            inst_vals[o] = (uint32_t)k;
            ++o;
        }
    }
}

__global__ void __launch_bounds__(256)
tile_range_kernel(const uint32_t* __restrict__ ids, uint64_t d, uint2* __restrict__ ranges)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const uint32_t t = __ldg(ids + j);
    if (j == 0 || __ldg(ids + j - 1) != t) ranges[t].x = (uint32_t)j;
    if (j + 1 == d || __ldg(ids + j + 1) != t) ranges[t].y = (uint32_t)(j + 1);
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

void launch_tile_counts(const uint32_t* trects, const uint32_t* order, const uint2* rects, int64_t r0, int64_t n,
                        FrameConsts fc, const uint32_t* tile_done, uint32_t* counts, unsigned long long* d_total,
                        cudaStream_t s)
{
    if (n <= 0) return;
    tile_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(trects, order, rects, r0, n, fc.tiles_x, fc.row_rank,
                                                                 fc.row_world, fc.row_group, tile_done, counts, d_total);
}

void launch_compact_live(const uint32_t* keys, const uint32_t* vals, const uint32_t* trects, const uint32_t* counts,
                         const uint32_t* positions, int64_t n, uint32_t* keys_out, uint32_t* vals_out,
                         uint32_t* trects_out, cudaStream_t s)
{
    if (n <= 0) return;
    compact_live_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys, vals, trects, counts, positions, n, keys_out, vals_out,
                                                                   trects_out);
}

void launch_emit(const uint32_t* order, const uint32_t* trects, const uint2* rects, const uint32_t* offsets,
                 const unsigned long long* total, int64_t n, FrameConsts fc, const uint32_t* tile_done,
                 uint32_t* inst_keys, uint32_t* inst_vals, cudaStream_t s)
{
    if (n <= 0) return;
    emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(order, trects, rects, offsets, total, n, fc.tiles_x,
                                                           fc.row_rank, fc.row_world, fc.row_group, tile_done, inst_keys, inst_vals);
}

void launch_tile_ranges(const uint32_t* sorted_tile_ids, uint64_t d, uint2* ranges, int num_tiles,
                        cudaStream_t s)
{
    cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s);
    if (d == 0) return;
    tile_range_kernel<<<(unsigned)((d + 255) / 256), 256, 0, s>>>(sorted_tile_ids, d, ranges);
}

void launch_debug_views(const Record* recs, const uint32_t* live_splats, int64_t n_live, Record* recs_by_splat,
                        const uint32_t* inst_refs, uint64_t d, uint32_t* inst_splats, cudaStream_t s)
{
    if (n_live > 0) debug_records_kernel<<<(unsigned)((n_live + 255) / 256), 256, 0, s>>>(recs, live_splats, n_live, recs_by_splat);
    if (d > 0) debug_instances_kernel<<<(unsigned)((d + 255) / 256), 256, 0, s>>>(inst_refs, live_splats, d, inst_splats);
}

}  // namespace gsb

// binning.cu — K4 tile binning, sm_100a.  The reference has no binning: GL's rasteriser assigns
// fragments to pixels implicitly (/root/reference/gsplat_plugin/src/GSplatRenderer.C:647).  Here each
// surviving splat is expanded, in global depth order, into one instance per 16x16 tile its pixel
// rectangle touches; radix_sort.cu then stably partitions the instances by tile id,
// which leaves every tile's list in depth order.  Integer / byte work, HBM-bound.
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

// counts[k] = number of live tiles touched by the splat of depth rank r0 + k (0 for culled splats).
// rects_sorted is in depth order (gathered by the last pass of the depth sort), so this is a coalesced stream.
// A tile is live if this rank owns its row and it is not yet saturated (tile_done, set by the blend
// of an earlier depth chunk): instances behind a saturated tile can never change a pixel.
__global__ void __launch_bounds__(256)
tile_count_kernel(const uint2* __restrict__ rects_sorted, int64_t r0, int64_t n,
                  int tiles_x, int row_rank, int row_world, int row_group, const uint32_t* __restrict__ tile_done,
                  uint32_t* __restrict__ counts)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const TileRect t = tile_rect(__ldg(rects_sorted + r0 + k));
    uint32_t c = 0;
    if (!t.empty) {
        for (int ty = t.ty0; ty <= t.ty1; ++ty) {
            if (!owns_row(ty, row_rank, row_world, row_group)) continue;
            if (tile_done) { for (int tx = t.tx0; tx <= t.tx1; ++tx) c += __ldg(tile_done + ty * tiles_x + tx) == 0u; }
            else c += (uint32_t)(t.tx1 - t.tx0 + 1);
        }
    }
    counts[k] = c;
}

// instance (tile id, splat index) pairs at offsets[k] .., rows ascending then columns ascending
__global__ void __launch_bounds__(256)
emit_kernel(const uint32_t* __restrict__ order, const uint2* __restrict__ rects_sorted,
            const uint32_t* __restrict__ offsets, const unsigned long long* __restrict__ total,
            int64_t r0, int64_t n, int tiles_x, int row_rank, int row_world, int row_group,
            const uint32_t* __restrict__ tile_done, uint32_t* __restrict__ inst_keys, uint32_t* __restrict__ inst_vals)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t o0 = offsets[k];
    const uint32_t o1 = (k + 1 < n) ? offsets[k + 1] : (uint32_t)(*total);
    if (o1 == o0) return;                                    // culled, or every tile it touches is saturated
    const uint32_t i = __ldg(order + r0 + k);
    const TileRect t = tile_rect(__ldg(rects_sorted + r0 + k));
    size_t o = o0;
    for (int ty = t.ty0; ty <= t.ty1; ++ty) {
        if (!owns_row(ty, row_rank, row_world, row_group)) continue;
        for (int tx = t.tx0; tx <= t.tx1; ++tx) {
            const uint32_t tile = (uint32_t)(ty * tiles_x + tx);
            if (tile_done && __ldg(tile_done + tile) != 0u) continue;
            inst_keys[o] = tile;
            inst_vals[o] = i;
            ++o;
        }
    }
}

// [start,end) of every tile in the tile-sorted instance list (ranges pre-zeroed: empty tiles = [0,0))
__global__ void __launch_bounds__(256)
tile_range_kernel(const uint32_t* __restrict__ ids, uint64_t d, uint2* __restrict__ ranges)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const uint32_t t = __ldg(ids + j);
    if (j == 0 || __ldg(ids + j - 1) != t) ranges[t].x = (uint32_t)j;
    if (j + 1 == d || __ldg(ids + j + 1) != t) ranges[t].y = (uint32_t)(j + 1);
}

}  // namespace

void launch_tile_counts(const uint2* rects_sorted, int64_t r0, int64_t n, FrameConsts fc,
                        const uint32_t* tile_done, uint32_t* counts, cudaStream_t s)
{
    if (n <= 0) return;
    tile_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rects_sorted, r0, n, fc.tiles_x, fc.row_rank,
                                                                 fc.row_world, fc.row_group, tile_done, counts);
}

void launch_emit(const uint32_t* order, const uint2* rects_sorted, const uint32_t* offsets,
                 const unsigned long long* total, int64_t r0, int64_t n, FrameConsts fc, const uint32_t* tile_done,
                 uint32_t* inst_keys, uint32_t* inst_vals, cudaStream_t s)
{
    if (n <= 0) return;
    emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(order, rects_sorted, offsets, total, r0, n, fc.tiles_x,
                                                           fc.row_rank, fc.row_world, fc.row_group, tile_done, inst_keys, inst_vals);
}

void launch_tile_ranges(const uint32_t* sorted_tile_ids, uint64_t d, uint2* ranges, int num_tiles,
                        cudaStream_t s)
{
    cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s);
    if (d == 0) return;
    tile_range_kernel<<<(unsigned)((d + 255) / 256), 256, 0, s>>>(sorted_tile_ids, d, ranges);
}

}  // namespace gsb

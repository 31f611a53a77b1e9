// radix_sort.cu — stable LSD radix sort of (uint32 key, uint32 value) pairs, sm_100a.
//
// Replaces the reference's CPU argsort (tbb::parallel_sort through an index indirection,
// /root/reference/gsplat_plugin/src/GSplatRenderer.C:206-208) for the depth order, and performs the
// stable tile partition of the instance list.  Keys are the raw fp32 bits of d^2
// (non-negative floats order as uint32).  LSD + stable ranking => ties keep ascending input order,
// which is the tie rule the spec fixes.
//
// Per pass (<= 8 bits): digit histogram per 4096-element block -> exclusive scan of the
// [digit][block] table -> rank + reorder in shared memory + coalesced scatter.
// HBM-bound integer work: 4 B (histogram read) + 16 B (scatter read+write) per element per pass.
// Ranking inside a warp uses match.any so equal digits (the common case for the top byte of a
// depth key, or the low bits of a tile id) cost one shared-memory update per distinct digit.
#include "common.cuh"

namespace gsb {

namespace {
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS   = RS_THREADS / 32;
constexpr int RS_ITEMS   = 16;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;   // 4096
constexpr int RS_RADIX   = 256;

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// element index of item j of this thread: warp-striped inside the warp's contiguous 512-element chunk
__device__ __forceinline__ size_t item_index(size_t base, int warp, int lane, int j)
{
    return base + (size_t)warp * (32 * RS_ITEMS) + (size_t)j * 32 + lane;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t mask, int nbins,
               uint32_t* __restrict__ block_hist, unsigned num_blocks)
{
    __shared__ uint32_t cnt[RS_WARPS][RS_RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&cnt[0][0])[i] = 0u;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    uint32_t k[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t idx = item_index(base, warp, lane, j);
        k[j] = (idx < n) ? __ldg(keys + idx) : 0u;
    }
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t idx = item_index(base, warp, lane, j);
        bool valid = idx < n;
        uint32_t d = (k[j] >> shift) & mask;
        unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0xFFFFFFFFu);
        if (valid && lane == (__ffs(peers) - 1)) cnt[warp][d] += __popc(peers);   // one lane per distinct digit
        __syncwarp();
    }
    __syncthreads();
    if ((int)threadIdx.x < nbins) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) t += cnt[w][threadIdx.x];
        block_hist[(size_t)threadIdx.x * num_blocks + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                  int shift, uint32_t mask, int nbins,
                  const uint32_t* __restrict__ block_offsets, unsigned num_blocks)
{
    __shared__ uint32_t cnt[RS_WARPS][RS_RADIX];     // per-warp digit counts -> warp-exclusive offsets
    __shared__ uint32_t local_base[RS_RADIX];        // exclusive scan of block digit totals
    __shared__ uint32_t global_delta[RS_RADIX];      // block_offsets[d][block] - local_base[d]
    __shared__ uint32_t wtot[RS_WARPS];
    __shared__ uint32_t sk[RS_TILE];
    __shared__ uint32_t sv[RS_TILE];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&cnt[0][0])[i] = 0u;

    const size_t base = (size_t)blockIdx.x * RS_TILE;
    const uint32_t tile_count = (uint32_t)((n - base < (size_t)RS_TILE) ? (n - base) : (size_t)RS_TILE);
    uint32_t k[RS_ITEMS], v[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t idx = item_index(base, warp, lane, j);
        bool valid = idx < n;
        k[j] = valid ? __ldg(keys_in + idx) : 0u;
        v[j] = valid ? __ldg(vals_in + idx) : 0u;
    }
    __syncthreads();

    // stable rank of every item among equal digits of its warp, items visited in (j, lane) order
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t idx = item_index(base, warp, lane, j);
        bool valid = idx < n;
        uint32_t d = (k[j] >> shift) & mask;
        unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0xFFFFFFFFu);
        uint32_t pre = valid ? cnt[warp][d] : 0u;
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) cnt[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[j] = (uint16_t)(pre + __popc(peers & lt));
    }
    __syncthreads();

    // thread d: exclusive prefix over warps for digit d, block total for d
    uint32_t total = 0;
    if ((int)threadIdx.x < nbins) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = total; total += c; }
    }
    // exclusive scan of the 256 digit totals across the block
    uint32_t inc = total;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += t; }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) woff += (w < warp) ? wtot[w] : 0u;
    if ((int)threadIdx.x < nbins) {
        uint32_t lb = woff + inc - total;
        local_base[threadIdx.x] = lb;
        global_delta[threadIdx.x] = block_offsets[(size_t)threadIdx.x * num_blocks + blockIdx.x] - lb;
    }
    __syncthreads();

    // reorder through shared memory so each digit's run is contiguous
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        size_t idx = item_index(base, warp, lane, j);
        if (idx < n) {
            uint32_t d = (k[j] >> shift) & mask;
            uint32_t lp = local_base[d] + cnt[warp][d] + rank[j];
            sk[lp] = k[j]; sv[lp] = v[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < RS_ITEMS; ++t) {
        uint32_t i = (uint32_t)t * RS_THREADS + threadIdx.x;
        if (i < tile_count) {
            uint32_t key = sk[i];
            uint32_t d = (key >> shift) & mask;
            size_t o = (size_t)(global_delta[d] + i);
            keys_out[o] = key; vals_out[o] = sv[i];
        }
    }
}
}  // namespace

static inline size_t rs_blocks(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

size_t sort_scratch_bytes(size_t n)
{
    size_t table = rs_blocks(n) * RS_RADIX;
    return ((table * sizeof(uint32_t) + 255) & ~size_t(255)) + scan_scratch_bytes(table) + 256;
}

int radix_sort_pairs(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, size_t n,
                     int begin_bit, int end_bit, void* scratch, cudaStream_t s, int* launches)
{
    if (n == 0 || end_bit <= begin_bit) return 0;
    const unsigned nb = (unsigned)rs_blocks(n);
    uint32_t* table = static_cast<uint32_t*>(scratch);
    void* scan_scr = static_cast<char*>(scratch) + ((((size_t)nb * RS_RADIX) * sizeof(uint32_t) + 255) & ~size_t(255));
    int bits_left = end_bit - begin_bit;
    int passes = (bits_left + 7) / 8;
    int shift = begin_bit, cur = 0;
    uint32_t* kin = k0; uint32_t* vin = v0; uint32_t* kout = k1; uint32_t* vout = v1;
    for (int p = 0; p < passes; ++p) {
        int b = (bits_left + (passes - p) - 1) / (passes - p);
        int nbins = 1 << b;
        uint32_t mask = (uint32_t)nbins - 1u;
        rs_hist_kernel<<<nb, RS_THREADS, 0, s>>>(kin, n, shift, mask, nbins, table, nb);
        if (launches) *launches += 1;
        exclusive_scan_u32(table, table, (size_t)nbins * nb, scan_scr, nullptr, s, launches);
        rs_scatter_kernel<<<nb, RS_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, mask, nbins, table, nb);
        if (launches) *launches += 1;
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        cur ^= 1; shift += b; bits_left -= b;
    }
    return cur;
}

}  // namespace gsb

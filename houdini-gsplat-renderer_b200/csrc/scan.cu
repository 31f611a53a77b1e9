// scan.cu — exclusive prefix sum over uint32 (reduce -> scan partials -> apply), sm_100a.
// HBM-bound: 12 bytes per element (read, read, write).  Used for tile counts (V entries) and for the
// radix sort's digit tables.  No spin-waits: three ordinary launches, so it cannot hang.
#include "common.cuh"

namespace gsb {

namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS   = 16;
constexpr int SCAN_CHUNK   = SCAN_THREADS * SCAN_ITEMS;   // 4096 elements per CTA

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// load 16 consecutive elements of this thread (vectorised when the chunk is full and aligned)
__device__ __forceinline__ void load_items(const uint32_t* __restrict__ in, size_t base, size_t n,
                                           uint32_t (&x)[SCAN_ITEMS])
{
    size_t first = base + (size_t)threadIdx.x * SCAN_ITEMS;
    if (first + SCAN_ITEMS <= n && ((reinterpret_cast<uintptr_t>(in + first) & 15u) == 0)) {
        const uint4* p = reinterpret_cast<const uint4*>(in + first);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            uint4 v = __ldg(p + k);
            x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) x[k] = (first + k < n) ? in[first + k] : 0u;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, size_t n,
                                                                   uint32_t* __restrict__ partial)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    uint32_t x[SCAN_ITEMS];
    load_items(in, (size_t)blockIdx.x * SCAN_CHUNK, n, x);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) s += x[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += wsum[w];
        partial[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of the per-chunk sums into 64-bit prefixes
__global__ void __launch_bounds__(1024) scan_partials_kernel(const uint32_t* __restrict__ partial, size_t m,
                                                             unsigned long long* __restrict__ prefix,
                                                             unsigned long long* __restrict__ total)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s, block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0ull;
    __syncthreads();
    for (size_t base = 0; base < m; base += 1024) {
        size_t i = base + threadIdx.x;
        unsigned long long v = (i < m) ? (unsigned long long)partial[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            wsum[lane] = winc - w;                 // exclusive warp offsets
            if (lane == 31) block_total = winc;
        }
        __syncthreads();
        unsigned long long excl = carry_s + wsum[warp] + (inc - v);
        if (i < m) prefix[i] = excl;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += block_total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in,
                                                                  uint32_t* __restrict__ out, size_t n,
                                                                  const unsigned long long* __restrict__ prefix)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
    uint32_t x[SCAN_ITEMS];
    load_items(in, base, n, x);
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) tsum += x[k];
    uint32_t inc = warp_incl_scan(tsum, lane);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) woff += (w < warp) ? wsum[w] : 0u;
    uint32_t run = (uint32_t)prefix[blockIdx.x] + woff + (inc - tsum);
    size_t first = base + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t y[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { y[k] = run; run += x[k]; }
    if (first + SCAN_ITEMS <= n && ((reinterpret_cast<uintptr_t>(out + first) & 15u) == 0)) {
        uint4* p = reinterpret_cast<uint4*>(out + first);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) p[k] = make_uint4(y[4 * k], y[4 * k + 1], y[4 * k + 2], y[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) if (first + k < n) out[first + k] = y[k];
    }
}
}  // namespace

static inline size_t scan_chunks(size_t n) { return (n + SCAN_CHUNK - 1) / SCAN_CHUNK; }

size_t scan_scratch_bytes(size_t n)
{
    size_t m = scan_chunks(n) + 1;
    return ((m * sizeof(uint32_t) + 255) & ~size_t(255)) + m * sizeof(unsigned long long) + 256;
}

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void* scratch,
                        unsigned long long* total_dev, cudaStream_t s, int* launches)
{
    if (n == 0) {
        if (total_dev) cudaMemsetAsync(total_dev, 0, sizeof(unsigned long long), s);
        return;
    }
    size_t m = scan_chunks(n);
    uint32_t* partial = static_cast<uint32_t*>(scratch);
    unsigned long long* prefix = reinterpret_cast<unsigned long long*>(
        static_cast<char*>(scratch) + (((m + 1) * sizeof(uint32_t) + 255) & ~size_t(255)));
    scan_reduce_kernel<<<(unsigned)m, SCAN_THREADS, 0, s>>>(in, n, partial);
    scan_partials_kernel<<<1, 1024, 0, s>>>(partial, m, prefix, total_dev);
    scan_apply_kernel<<<(unsigned)m, SCAN_THREADS, 0, s>>>(in, out, n, prefix);
    if (launches) *launches += 3;
}

}  // namespace gsb

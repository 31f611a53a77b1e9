// scan.cu — exclusive prefix sum over uint32, sm_100a.  Two forms:
//  * exclusive_scan_u32_onepass (the frame's count scan): ONE launch — every CTA reduces its 4096-element tile, publishes the
//    aggregate, obtains its exclusive prefix by decoupled look-back over epoch-tagged 64-bit tile states (bounded spins,
//    error flag on time-out, never cleared between scans) and writes its outputs.  8 bytes per element.  Partial sums < 2^30.
//  * exclusive_scan_u32 (reduce -> scan partials -> apply): three ordinary launches, 64-bit prefixes, no spin-waits; kept for
//    totals beyond 2^30 and as the A/B reference (GSB_SCAN=3).  12 bytes per element.
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

// ---- single pass (r02): reduce, decoupled look-back and apply in ONE launch ---------------------------------------------
// status[tile] = epoch << 32 | flag | value (value < 2^30).  Entries of another epoch mean "not published yet", so the table
// is never cleared between scans (it is zeroed once when allocated; the caller's epoch counter never repeats within 2^32
// scans).  Tile t = blockIdx.x: like CUB's scan this relies on CTAs being dispatched in index order, so every predecessor
// of a running tile is running or done; the spins are bounded all the same and a time-out raises *error_flag.
constexpr uint32_t ST_AGG = 0x40000000u, ST_INCL = 0x80000000u, ST_MASK = 0x3FFFFFFFu, ST_SPIN_LIMIT = 1u << 22;
__device__ __forceinline__ unsigned long long st_ld(const unsigned long long* p)
{
    unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_st(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_onepass_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                    size_t n, unsigned long long* __restrict__ status,
                                                                    const uint32_t epoch, unsigned long long* __restrict__ total,
                                                                    uint32_t* __restrict__ error_flag)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t s_excl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    const size_t base = (size_t)tile * SCAN_CHUNK;
    const unsigned long long etag = (unsigned long long)epoch << 32;
    uint32_t x[SCAN_ITEMS];
    load_items(in, base, n, x);
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) tsum += x[k];
    const uint32_t inc = warp_incl_scan(tsum, lane);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) { woff += (w < warp) ? wsum[w] : 0u; tile_total += wsum[w]; }
    if (warp == 0) {
        // publish the aggregate (tile 0: already inclusive), then walk back over the predecessors, 32 at a time
        if (lane == 0) st_st(status + tile, etag | (tile_total & ST_MASK) | (tile == 0 ? ST_INCL : ST_AGG));
        uint32_t excl = 0;
        if (tile > 0) {
            int64_t q = (int64_t)tile - 1;               // nearest predecessor not summed yet
            uint32_t spins = 0;
            for (;;) {
                const int64_t mine = q - lane;
                unsigned long long v = (mine >= 0) ? st_ld(status + mine) : (etag | ST_INCL);
                const bool ready = (uint32_t)(v >> 32) == epoch && ((uint32_t)v & (ST_AGG | ST_INCL)) != 0u;
                const unsigned not_ready = __ballot_sync(0xffffffffu, !ready);
                const unsigned incl = __ballot_sync(0xffffffffu, ready && ((uint32_t)v & ST_INCL) != 0u);
                // lanes [0, stop) are usable: up to the first inclusive entry, and only below the first entry not ready
                const int first_nr = not_ready ? (__ffs(not_ready) - 1) : 32;
                const int first_in = incl ? (__ffs(incl) - 1) : 32;
                const int stop = first_in < first_nr ? first_in + 1 : first_nr;
                uint32_t part = (lane < stop) ? ((uint32_t)v & ST_MASK) : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                excl += part;
                q -= stop;
                if (first_in < first_nr) break;          // met an inclusive prefix
                if (stop == 0) {
                    if (++spins > ST_SPIN_LIMIT) { if (lane == 0 && error_flag) atomicExch(error_flag, 1u); break; }
                    __nanosleep(20);
                }
            }
            if (lane == 0) st_st(status + tile, etag | ((excl + tile_total) & ST_MASK) | ST_INCL);
        }
        if (lane == 0) {
            s_excl = excl;
            if (total && base + SCAN_CHUNK >= n) *total = (unsigned long long)excl + tile_total;     // the last tile
        }
    }
    __syncthreads();
    uint32_t run = s_excl + woff + (inc - tsum);
    const size_t first = base + (size_t)threadIdx.x * SCAN_ITEMS;
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

// single launch; status: scan_status_bytes(n) bytes, zeroed when allocated, never cleared afterwards; epoch: a value no
// earlier scan on this table used; every partial sum must stay below 2^30
size_t scan_status_bytes(size_t n) { return (scan_chunks(n) + 1) * sizeof(unsigned long long) + 256; }

void exclusive_scan_u32_onepass(const uint32_t* in, uint32_t* out, size_t n, unsigned long long* status, uint32_t epoch,
                                unsigned long long* total_dev, uint32_t* error_flag, cudaStream_t s, int* launches)
{
    if (n == 0) {
        if (total_dev) cudaMemsetAsync(total_dev, 0, sizeof(unsigned long long), s);
        return;
    }
    scan_onepass_kernel<<<(unsigned)scan_chunks(n), SCAN_THREADS, 0, s>>>(in, out, n, status, epoch, total_dev, error_flag);
    if (launches) *launches += 1;
}

}  // namespace gsb

// radix_sort.cu — stable LSD radix sort of (uint32 key, uint32 value) pairs, sm_100a, onesweep style.
//
// Replaces the reference's CPU argsort (tbb::parallel_sort through an index indirection,
// /root/reference/gsplat_plugin/src/GSplatRenderer.C:206-208) for the depth order, and performs the
// stable tile partition of the instance list.  Keys are the raw fp32 bits of d^2
// (non-negative floats order as uint32).  LSD + stable ranking => ties keep ascending input order,
// which is the tie rule the spec fixes.
//
// Structure (after Adinets & Merrill's Onesweep): ONE upfront kernel builds the global digit histogram
// of every pass from a single read of the keys; then one kernel per <= 8-bit pass ranks a 4096-key tile,
// obtains the tile's per-digit global offset by decoupled look-back over the preceding tiles, reorders the
// tile in shared memory and scatters it coalesced.  Per pass each element is read once and written once:
// 16 B/element/pass + 4 B for the histogram (68 B per element for 4 passes).
//
// Safety: a CTA's first tile is its block index (the grid never exceeds the device's resident capacity, so all of those
// run at once), every further tile comes from an atomic ticket, so a tile only ever waits on tiles that already
// started; every spin is bounded and raises an error flag instead of hanging the GPU.
//
// r02: the frame pipeline never learns element counts on the host in time, so (i) every kernel reads n from DEVICE memory
// (n_dev; the host only supplies an upper bound for the scratch size), (ii) the pass kernel is PERSISTENT: three CTAs per SM
// (40 registers, 60 KB of shared memory) loop over ticketed tiles, so a loose upper bound costs no empty CTAs, (iii) look-back entries are 64-bit, tagged with a
// per-sort EPOCH in the high word: entries of earlier sorts read as "not published" and the table is never cleared,
// (iv) the exclusive scan of the digit histogram happens inside every pass CTA (one block scan of <= 512 bins, behind the
// first tile's key loads) instead of in a kernel of its own, and (v) the histogram itself may be supplied by the kernel that produced the keys
// (hist_ready), which removes the upfront pass over them.
//
// Ranking inside a warp needs, per key, the mask of lanes holding the same digit.  match.any does that in one
// instruction but runs on the ADU pipe at ~2 cycles per DISTINCT value (ncu r01: ADU 97 % busy, 62 cycles per
// warp for random digits).  r01 built the mask from one vote.ballot per digit bit (<= 9 ballots at ~6 SASS each: 49 % of the
// pass's instructions, r02 ncu source counters); r02 uses ONE shared-memory atomicOr per key on a [warp][digit] mask table
// aliased onto the reorder staging area (ATOMIC_MATCH; the ballot form stays selectable with GSB_RS_MATCH=ballot).
#include "common.cuh"
#include <cstdlib>

namespace gsb {

namespace {
// tile shape: compile-time knobs so tools/build_variants.sh can build experiment libraries (r01 sweep in DESIGN.md)
#ifndef GSB_RS_THREADS
#define GSB_RS_THREADS 512
#endif
#ifndef GSB_RS_ITEMS
#define GSB_RS_ITEMS 8
#endif
#ifndef GSB_RS_MINB
#define GSB_RS_MINB 3
#endif
constexpr int RS_THREADS = GSB_RS_THREADS;          // the wide shape: 512 threads x 8 keys (9-bit digits need 512 threads)
constexpr int RS_ITEMS   = GSB_RS_ITEMS;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;   // 4096: the granularity the scratch sizes are computed for
// (r02 measured a 256-thread x 8-key shape for the <= 8-bit passes, four CTAs per SM instead of two: 3 % slower frames — the
// per-tile fixed costs (counter clearing, barriers, look-back rows) double with the tile count; the template parameter stays)
constexpr int RS_MAXBITS = 9;
constexpr int RS_RADIX   = 1 << RS_MAXBITS;         // up to 512 bins per pass
constexpr int RS_MAX_PASSES = 4;
constexpr int LB_WIN     = 16;                      // entries inspected per look-back step (independent loads)
constexpr int LB_BLOCK   = 16;                      // tiles per look-back block (two-level look-back)

constexpr uint32_t LB_AGG  = 0x40000000u;   // tile aggregate available
constexpr uint32_t LB_INCL = 0x80000000u;   // inclusive prefix available
constexpr uint32_t LB_MASK = 0x3FFFFFFFu;   // counts < 2^30
constexpr uint32_t SPIN_LIMIT = 1u << 22;

struct PassPlan { int shift[RS_MAX_PASSES]; int bits[RS_MAX_PASSES]; int passes; uint32_t key_min, key_span; };
static_assert(RS_MAX_PASSES == SORT_MAX_PASSES && RS_RADIX == SORT_RADIX, "common.cuh mirrors the sort's table shape");

// dynamic shared memory of os_pass_kernel
template <int THREADS>
struct PassSmemT {
    static constexpr int RS_WARPS = THREADS / 32;
    static constexpr int RS_TILE = THREADS * RS_ITEMS;
    uint16_t cnt[RS_WARPS][RS_RADIX];   // per-warp digit counts -> warp-exclusive offsets (<= 4096)
    uint32_t local_base[RS_RADIX];      // exclusive scan of the tile's digit totals
    uint32_t global_delta[RS_RADIX];    // global offset of digit d for this tile - local_base[d]
    uint32_t tot[RS_RADIX];             // tile total per digit
    uint32_t excl[RS_RADIX];            // look-back result per digit
    uint32_t digit_base[RS_RADIX];      // exclusive scan of the pass's global digit histogram (once per CTA)
    uint32_t wtot[RS_WARPS];
    uint32_t tile;
    uint32_t sk[RS_TILE];
    uint32_t sv[RS_TILE];
};
using PassSmem = PassSmemT<RS_THREADS>;

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p)
{
    unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_volatile(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Order-preserving key compression: keys below key_min + key_span keep their order after subtracting key_min,
// everything at or above (the culled sentinel 0xFFFFFFFF) collapses onto key_span, the largest value.
// key_min = 0, key_span = 0xFFFFFFFF is the identity.
__device__ __forceinline__ uint32_t squeeze(uint32_t key, uint32_t key_min, uint32_t key_span)
{
    return min(key - key_min, key_span);
}

// digit of a key in one pass: a bit field of the squeezed key
struct BitsDigit {
    int shift; uint32_t mask, key_min, key_span;
    __device__ __forceinline__ uint32_t operator()(uint32_t key) const { return (squeeze(key, key_min, key_span) >> shift) & mask; }
};

// lanes of the warp whose digit equals this lane's digit (0 for invalid lanes); nbits = digit width of the pass
template <int NBITS>
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool valid)
{
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < NBITS; ++b) {
        const bool bit = (d & (1u << b)) != 0u;                  // one LOP3 with a predicate result, no shift
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return valid ? peers : 0u;
}

// element index of item j of this thread: warp-striped inside the warp's contiguous 256-element chunk
__device__ __forceinline__ size_t item_index(size_t base, int warp, int lane, int j)
{
    return base + (size_t)warp * (32 * RS_ITEMS) + (size_t)j * 32 + lane;
}
constexpr int RS_WARPS = RS_THREADS / 32;

// ---- upfront: global digit histograms of all passes, one read of the keys --------------------------------
// hist layout: [pass][RS_RADIX].  Persistent CTAs, shared-memory REDs, one global RED per non-zero bin per CTA.
__global__ void __launch_bounds__(RS_THREADS)
os_hist_kernel(const uint32_t* __restrict__ keys, size_t n_max, const unsigned long long* __restrict__ n_dev, PassPlan plan,
               uint32_t* __restrict__ hist)
{
    __shared__ uint32_t h[RS_MAX_PASSES][RS_RADIX];
    const size_t n = n_dev ? (size_t)min((unsigned long long)n_max, *n_dev) : n_max;
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += RS_THREADS) (&h[0][0])[i] = 0u;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * RS_THREADS;
    for (size_t i = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        const uint32_t k = squeeze(__ldg(keys + i), plan.key_min, plan.key_span);
#pragma unroll
        for (int p = 0; p < RS_MAX_PASSES; ++p)
            if (p < plan.passes) atomicAdd(&h[p][(k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * RS_RADIX; i += RS_THREADS) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// ---- one pass ------------------------------------------------------------------------------------------------
// lookback layout: [tile][nbins] 64-bit entries (epoch << 32 | flags | count): a tile publishes one coalesced row; a
// look-back step reads LB_WIN rows.  Persistent: every CTA loops over ticketed tiles until the tickets run out.
// ATOMIC_MATCH: the lanes of a warp that share a digit find each other through one shared-memory atomicOr per key on a
// [warp][digit] mask table (aliased onto the reorder staging area, which is idle during the ranking) instead of NBITS
// ballots with ~6 instructions each: the ranking loop was 49 % of the pass's instructions (r02 ncu source counters).
template <int NBITS, int THREADS, int MINB, bool ATOMIC_MATCH, class DigitFn>
__global__ void __launch_bounds__(THREADS, MINB)
os_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, const size_t n_max,
               const unsigned long long* __restrict__ n_dev,
               const DigitFn dig, const uint32_t* __restrict__ hist,
               unsigned long long* __restrict__ lookback, const unsigned tiles_max, uint32_t* __restrict__ ticket,
               uint32_t* __restrict__ error_flag, const uint32_t epoch, const int static_first)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = PassSmemT<THREADS>;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    constexpr int RS_WARPS = THREADS / 32, RS_THREADS = THREADS, RS_TILE = THREADS * RS_ITEMS;     // (shadow the wide shape's)
    static_assert((1 << NBITS) <= THREADS, "one thread per digit");

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int nbins = 1 << NBITS;
    const unsigned lt = lanemask_lt();
    const size_t n = n_dev ? (size_t)min((unsigned long long)n_max, *n_dev) : n_max;
    const unsigned num_tiles = (unsigned)((n + RS_TILE - 1) / RS_TILE);
    const unsigned long long etag = (unsigned long long)epoch << 32;

    // static_first: the first tile of a CTA is its block index and later ones come from the ticket counter (which then
    // starts at gridDim.x), so the first key loads wait neither for the ticket's round trip nor for the histogram: the three
    // global latencies of a short pass (ticket, histogram, keys) overlap.  The grid never exceeds the resident capacity, so
    // every statically assigned tile is running and the look-back cannot starve.
    uint32_t tile = blockIdx.x;
    bool first_tile = true;
    const uint32_t ticket_base = static_first ? gridDim.x : 0u;
    for (;;) {
        const bool draw = !first_tile || !static_first;                 // CTA-uniform
        if (draw && threadIdx.x == 0) sm.tile = ticket_base + atomicAdd(ticket, 1u);
        for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX / 2; i += RS_THREADS) reinterpret_cast<uint32_t*>(&sm.cnt[0][0])[i] = 0u;
        uint32_t* const masks = sm.sk;                                  // [warp][nbins]: <= 16 x 512 words = sk + sv
        if (ATOMIC_MATCH) {
            static_assert(sizeof(sm.sk) + sizeof(sm.sv) >= sizeof(uint32_t) * RS_WARPS * nbins, "mask table fits the staging area");
            for (int i = threadIdx.x; i < RS_WARPS * nbins; i += RS_THREADS) masks[i] = 0u;
        }
        if (draw) {
            __syncthreads();
            tile = sm.tile;
        }
        if (tile >= num_tiles) break;                                   // CTA-uniform

        const size_t base = (size_t)tile * RS_TILE;
        const uint32_t tile_count = (uint32_t)((n - base < (size_t)RS_TILE) ? (n - base) : (size_t)RS_TILE);
        // item j of this thread is local element li0 + 32 j of the tile (warp-striped inside the warp's 256-element chunk): one
        // pointer per array and immediate offsets, validity as a 32-bit compare against the elements left from li0 on
        const uint32_t li0 = (uint32_t)warp * (32u * RS_ITEMS) + (uint32_t)lane;
        const int rem = (int)tile_count - (int)li0;                   // item j is valid iff 32 j < rem
        const uint32_t* __restrict__ kp = keys_in + base + li0;
        const uint32_t* __restrict__ vp = vals_in + base + li0;
        uint32_t k[RS_ITEMS];
        uint32_t rd[RS_ITEMS];                                        // rank | digit << 16 (rank <= 4096, digit < 512)
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) k[j] = (32 * j < rem) ? __ldg(kp + 32 * j) : 0u;
        if (first_tile) {
            // global digit offsets of this pass: exclusive scan of the histogram, once per CTA, behind the first key loads
            // (its barriers also order the clearing of cnt / masks above when no ticket was drawn)
            const uint32_t hv = ((int)threadIdx.x < nbins) ? __ldg(hist + threadIdx.x) : 0u;
            uint32_t inc = hv;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += t; }
            if (lane == 31) sm.wtot[warp] = inc;
            __syncthreads();
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) woff += (w < warp) ? sm.wtot[w] : 0u;
            if ((int)threadIdx.x < nbins) sm.digit_base[threadIdx.x] = woff + inc - hv;
            __syncthreads();
            first_tile = false;
        }

        // stable rank of every item among equal digits of its warp, items visited in (j, lane) order
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const bool valid = 32 * j < rem;
            const uint32_t d = dig(k[j]);
            if (ATOMIC_MATCH) {
                uint32_t* const mp = masks + warp * nbins + d;
                if (valid) atomicOr(mp, 1u << lane);
                __syncwarp();
                const unsigned peers = valid ? *reinterpret_cast<volatile uint32_t*>(mp) : 0u;
                const uint32_t pre = valid ? sm.cnt[warp][d] : 0u;
                __syncwarp();                                           // everyone has read the mask and the count
                if (valid && lane == (__ffs(peers) - 1)) { sm.cnt[warp][d] = (uint16_t)(pre + __popc(peers)); *mp = 0u; }
                __syncwarp();
                rd[j] = (pre + __popc(peers & lt)) | (d << 16);
            } else {
                const unsigned peers = match_digit<NBITS>(d, valid);
                const uint32_t pre = valid ? sm.cnt[warp][d] : 0u;
                __syncwarp();
                if (valid && lane == (__ffs(peers) - 1)) sm.cnt[warp][d] = (uint16_t)(pre + __popc(peers));
                __syncwarp();
                rd[j] = (pre + __popc(peers & lt)) | (d << 16);
            }
        }
        __syncthreads();

        // the values are only needed for the reorder below: loading them here (not with the keys) keeps 8 registers free
        // during the ranking (the kernel runs at 40 registers for 3 CTAs per SM) and their latency hides behind the scans
        uint32_t v[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) v[j] = (32 * j < rem) ? __ldg(vp + 32 * j) : 0u;

        // thread d: exclusive prefix over warps for digit d, tile total for d; publish the aggregate
        uint32_t total = 0;
        if ((int)threadIdx.x < nbins) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = sm.cnt[w][threadIdx.x]; sm.cnt[w][threadIdx.x] = (uint16_t)total; total += c; }
            sm.tot[threadIdx.x] = total;
            st_volatile(lookback + (size_t)tile * nbins + threadIdx.x, etag | total | LB_AGG);
        }
        // exclusive scan of the digit totals across the block
        uint32_t inc = total;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += t; }
        if (lane == 31) sm.wtot[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) woff += (w < warp) ? sm.wtot[w] : 0u;
        if ((int)threadIdx.x < nbins) sm.local_base[threadIdx.x] = woff + inc - total;
        __syncthreads();

        // reorder through shared memory so each digit's run is contiguous (needs no global information, and lets the
        // key/value/rank registers die before the look-back)
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            if (32 * j < rem) {
                const uint32_t d = rd[j] >> 16;
                const uint32_t lp = sm.local_base[d] + sm.cnt[warp][d] + (rd[j] & 0xffffu);
                sm.sk[lp] = k[j]; sm.sv[lp] = v[j];
            }
        }

        // Decoupled look-back in TWO LEVELS (r02), one thread per digit.  A single-level walk over tile aggregates needs
        // tile / LB_WIN dependent L2 round trips when no predecessor has an inclusive prefix yet — exactly the one-wave case
        // of the depth sorts (342 tiles in flight at once: 19 round trips = 13 us of a 22 us pass).  Here tiles form blocks
        // of LB_BLOCK; the last tile of a block also publishes the block's aggregate (as soon as its 15 predecessors in the
        // block have published theirs) and later the inclusive prefix at the block's end.  A tile sums (1) the aggregates of
        // the tiles before it in its own block: one batch of independent loads, and (2) block entries backwards in windows of
        // LB_WIN until it meets an inclusive one: ceil(tile / 256) more round trips.  Rows of both tables are read coalesced
        // across the warp; entries carry the sort's epoch (anything else = not published yet); every spin is bounded.
        if ((int)threadIdx.x < nbins) {
            const unsigned long long* agg_col = lookback + threadIdx.x;                 // [tile][nbins]: tile aggregates
            unsigned long long* blk_col = lookback + (size_t)tiles_max * nbins + threadIdx.x;   // [block][nbins]
            const int blk = (int)tile / LB_BLOCK, first = blk * LB_BLOCK, cnt_in = (int)tile - first;
            const bool block_last = ((int)tile % LB_BLOCK) == LB_BLOCK - 1;
            uint32_t excl_in = 0, excl_blk = 0, spins = 0;
            bool failed = false;
            // (1) tiles [first, tile) of this block
            if (cnt_in > 0) {
                for (;;) {
                    unsigned long long val[LB_BLOCK - 1];
#pragma unroll
                    for (int j = 0; j < LB_BLOCK - 1; ++j)
                        val[j] = (j < cnt_in) ? ld_volatile(agg_col + (size_t)(first + j) * nbins) : (etag | LB_AGG);
                    bool all = true; uint32_t sum = 0;
#pragma unroll
                    for (int j = 0; j < LB_BLOCK - 1; ++j) {
                        const uint32_t lo = (uint32_t)val[j];
                        all = all && (uint32_t)(val[j] >> 32) == epoch && (lo & LB_AGG) != 0u;
                        sum += lo & LB_MASK;
                    }
                    if (all) { excl_in = sum; break; }
                    if (++spins > SPIN_LIMIT) { atomicExch(error_flag, 1u); failed = true; break; }
                    __nanosleep(20);
                }
            }
            if (block_last && !failed) st_volatile(blk_col + (size_t)blk * nbins, etag | ((excl_in + total) & LB_MASK) | LB_AGG);
            // (2) the blocks before this one, nearest first, until an inclusive entry
            int bq = blk - 1;
            bool done = bq < 0 || failed;
            while (!done) {
                unsigned long long val[LB_WIN];
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j)
                    val[j] = (bq - j >= 0) ? ld_volatile(blk_col + (size_t)(bq - j) * nbins) : (etag | LB_INCL);
                int used = 0;
                bool blocked = false;
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j) {
                    if (!done && !blocked) {
                        const uint32_t lo = (uint32_t)val[j];
                        if ((uint32_t)(val[j] >> 32) != epoch || (lo & (LB_AGG | LB_INCL)) == 0u) blocked = true;   // not published yet: retry from here
                        else { excl_blk += lo & LB_MASK; ++used; done = (lo & LB_INCL) != 0u; }
                    }
                }
                bq -= used;
                if (blocked) {
                    if (++spins > SPIN_LIMIT) { atomicExch(error_flag, 1u); break; }
                    __nanosleep(20);
                }
            }
            const uint32_t excl = excl_in + excl_blk;
            sm.excl[threadIdx.x] = excl;
            if (block_last) st_volatile(blk_col + (size_t)blk * nbins, etag | ((excl + total) & LB_MASK) | LB_INCL);
        }
        __syncthreads();
        if ((int)threadIdx.x < nbins)
            sm.global_delta[threadIdx.x] = sm.digit_base[threadIdx.x] + sm.excl[threadIdx.x] - sm.local_base[threadIdx.x];
        __syncthreads();

#pragma unroll
        for (int t = 0; t < RS_ITEMS; ++t) {
            const uint32_t i = (uint32_t)t * RS_THREADS + threadIdx.x;
            if (i < tile_count) {
                const uint32_t key = sm.sk[i];
                const uint32_t o = sm.global_delta[dig(key)] + i;          // < 2^30 elements: 32-bit offsets
                keys_out[o] = key; vals_out[o] = sm.sv[i];
            }
        }
        __syncthreads();                                                // smem is reused by the next ticket
    }
}
}  // namespace

static inline size_t rs_blocks(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// header of one sort invocation: [hist passes x 512 u32][tickets 4 u32][spare 4 u32], zero before its first use
size_t sort_header_bytes() { return ((RS_MAX_PASSES * RS_RADIX + 8) * sizeof(uint32_t) + 255) & ~size_t(255); }

// look-back table for sorts of up to n_max elements: passes x tiles x 512 bins x 8 bytes (never cleared: epoch tagged)
size_t sort_lookback_bytes(size_t n_max)
{
    const size_t nb = rs_blocks(n_max);
    return (size_t)RS_MAX_PASSES * (nb + nb / LB_BLOCK + 2) * RS_RADIX * sizeof(unsigned long long) + 256;
}

int sort_key_bits(uint32_t key_span)
{
    int b = 0; while (b < 32 && (key_span >> b) != 0u) ++b; return b < 1 ? 1 : b;
}

SortPlan sort_plan(int begin_bit, int end_bit)
{
    SortPlan p{};
    int bits_left = end_bit - begin_bit;
    if (bits_left <= 0) return p;
    const int p8 = (bits_left + 7) / 8, p9 = (bits_left + 8) / 9;
    p.passes = p9 < p8 ? p9 : p8;          // 9-bit digits only when they save a whole pass
    int shift = begin_bit;
    for (int i = 0; i < p.passes; ++i) {
        int b = (bits_left + (p.passes - i) - 1) / (p.passes - i);
        p.shift[i] = shift; p.bits[i] = b; shift += b; bits_left -= b;
    }
    return p;
}

int radix_sort_pairs(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, size_t n_max,
                     const unsigned long long* n_dev, int begin_bit, int end_bit,
                     uint32_t* header, bool header_is_zero, bool hist_ready, unsigned long long* lookback, uint32_t epoch,
                     uint32_t* error_flag, cudaStream_t s, int* launches, uint32_t key_min, uint32_t key_span)
{
    if (n_max == 0 || end_bit <= begin_bit) return 0;
    static bool attr_set = false;
    static int pass_ctas_per_sm = GSB_RS_MINB;
    if (!attr_set) {
#define GSB_SET_ATTR(B) cudaFuncSetAttribute(os_pass_kernel<B, RS_THREADS, GSB_RS_MINB, true, BitsDigit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)); \
                        cudaFuncSetAttribute(os_pass_kernel<B, RS_THREADS, GSB_RS_MINB, false, BitsDigit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem))
        GSB_SET_ATTR(1); GSB_SET_ATTR(2); GSB_SET_ATTR(3); GSB_SET_ATTR(4); GSB_SET_ATTR(5); GSB_SET_ATTR(6); GSB_SET_ATTR(7);
        GSB_SET_ATTR(8); GSB_SET_ATTR(9);
#undef GSB_SET_ATTR
        int per = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, os_pass_kernel<8, RS_THREADS, GSB_RS_MINB, true, BitsDigit>, RS_THREADS, sizeof(PassSmem)) == cudaSuccess && per >= 1)
            pass_ctas_per_sm = per;
        attr_set = true;
    }
    const unsigned nb = (unsigned)rs_blocks(n_max);
    const SortPlan sp = sort_plan(begin_bit, end_bit);
    PassPlan plan{};
    plan.key_min = key_min; plan.key_span = key_span; plan.passes = sp.passes;
    for (int p = 0; p < sp.passes; ++p) { plan.shift[p] = sp.shift[p]; plan.bits[p] = sp.bits[p]; }
    uint32_t* hist = header;
    uint32_t* tickets = hist + RS_MAX_PASSES * RS_RADIX;
    if (!error_flag) error_flag = tickets + RS_MAX_PASSES;      // nobody looks: still a valid sink
    size_t lb_off[RS_MAX_PASSES + 1]; lb_off[0] = 0;
    for (int p = 0; p < plan.passes; ++p) lb_off[p + 1] = lb_off[p] + ((size_t)(nb + nb / LB_BLOCK + 2) << plan.bits[p]);   // tile rows + block rows
    if (!header_is_zero) cudaMemsetAsync(header, 0, sort_header_bytes(), s);
    if (!hist_ready) {
        const unsigned hist_grid = nb < (unsigned)(NUM_SMS * 4) ? nb : (unsigned)(NUM_SMS * 4);
        os_hist_kernel<<<hist_grid, RS_THREADS, 0, s>>>(k0, n_max, n_dev, plan, hist);
        if (launches) *launches += 1;
    }
    // GSB_RS_MATCH=ballot in the environment selects the ballot ranking (A/B runs); default: shared-memory atomicOr
    static const bool atomic_match = [] { const char* e = getenv("GSB_RS_MATCH"); return !(e && e[0] == 'b'); }();
    static const int static_first = [] { const char* e = getenv("GSB_RS_STATIC"); return (e && atoi(e) == 0) ? 0 : 1; }();
    // resident capacity from the SMs this device really has (the static first tile relies on every CTA of the grid being
    // resident at once; NUM_SMS = 148 is the full B200, a partitioned or cut-down part reports fewer)
    static const int sms = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = NUM_SMS;
        return v;
    }();
    const unsigned cap = (unsigned)(sms * pass_ctas_per_sm);
    const unsigned grid = nb < cap ? nb : cap;
    uint32_t* kin = k0; uint32_t* vin = v0; uint32_t* kout = k1; uint32_t* vout = v1;
    int cur = 0;
    for (int p = 0; p < plan.passes; ++p) {
#define GSB_PASS_ARGS(B) kin, vin, kout, vout, n_max, n_dev, BitsDigit{ plan.shift[p], (1u << B) - 1u, key_min, key_span }, \
                        hist + p * RS_RADIX, lookback + lb_off[p], nb, tickets + p, error_flag, epoch, static_first
#define GSB_PASS(B) case B: \
            if (atomic_match) os_pass_kernel<B, RS_THREADS, GSB_RS_MINB, true, BitsDigit><<<grid, RS_THREADS, sizeof(PassSmem), s>>>(GSB_PASS_ARGS(B)); \
            else os_pass_kernel<B, RS_THREADS, GSB_RS_MINB, false, BitsDigit><<<grid, RS_THREADS, sizeof(PassSmem), s>>>(GSB_PASS_ARGS(B)); \
            break
        switch (plan.bits[p]) {
            GSB_PASS(1); GSB_PASS(2); GSB_PASS(3); GSB_PASS(4); GSB_PASS(5); GSB_PASS(6); GSB_PASS(7); GSB_PASS(8); GSB_PASS(9);
        }
#undef GSB_PASS
#undef GSB_PASS_ARGS
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        cur ^= 1;
    }
    if (launches) *launches += plan.passes;
    return cur;
}

}  // namespace gsb

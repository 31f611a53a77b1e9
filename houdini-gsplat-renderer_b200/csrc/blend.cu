// blend.cu — K5 per-tile front-to-back alpha compositing, sm_100a.
//
// Replaces the fixed-function rasteriser + fragment shader + ROP "under" blend of the reference:
//   falloff / discard / premultiply   /root/reference/gsplat_plugin/shaders/GSplatShaderSource.h:304-312
//   quad support (+-2 in eigen space) /root/reference/gsplat_plugin/shaders/GSplatShaderSource.h:168-188,277-282
//   blend (ONE_MINUS_DST_ALPHA, ONE)  /root/reference/gsplat_plugin/src/GSplatRenderer.C:613-621
//
// One CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block (one pixel per lane).  The
// tile's depth-ordered instance list is staged through shared memory in batches of 256 records with
// cp.async (LDGSTS) double buffering: thread t reads instance t's splat index and gathers its 48-byte
// record as three 16-byte async copies.  Each warp then culls the batch against its own 8x4 block
// 32 instances at a time (one instance per lane, conservative AABB test, ballot) and walks only the
// surviving bits in order, so a small splat costs one warp pass instead of eight.  Transmittance
// is kept per pixel in registers; a warp stops (checked once per 32 instances) when all its pixels are
// saturated (T < eps) and the CTA stops when all warps have (early-out: the reference has none).
//
// Per-pixel arithmetic is the spec of DESIGN.md §3 and oracle/gsplat_oracle.cpp shade(): explicit
// fmaf where the spec says fmaf, nothing else contracted (TU built with -fmad=false), so coverage
// decisions are bit-exact; only exp() differs from libm (MUFU.EX2), ~1e-7 relative.
//
// Algorithmic bytes: D_c * (4 + 48) + W*H*16 — HBM/L2-gather bound by design,
// FP32-issue bound in practice.
#include "common.cuh"
#include <cstdlib>

namespace gsb {

namespace {

constexpr int BL_THREADS = 256;
constexpr int BL_BATCH   = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// 2^t, t <= 0 and far above the denormal range here (t >= -ln(255) * log2(e) > -8): the bare MUFU.EX2 that __expf's
// t >= -126 path executes, without the range test and the two predicated scalings around it (same bits)
__device__ __forceinline__ float ex2_mufu(float t)
{
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t)); return r;
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// DEPTH: scene-depth occlusion (R.C:608-610 depth test on / writes off, SRC.h:278-282 one depth per quad):
// the window depth of every staged instance rides along in shared memory, each pixel keeps the scene depth at its
// position in a register, and a fragment that fails F.depth_func is dropped (no colour, no transmittance change).
template <bool OBB_CULL, bool DEPTH, bool QUEUE>
__global__ void __launch_bounds__(BL_THREADS)
blend_kernel(const Record* __restrict__ recs, const uint32_t* __restrict__ inst,
             const uint2* __restrict__ ranges, float4* __restrict__ fb, float4* __restrict__ fb_final,
             const __grid_constant__ FrameConsts F, const int first, const int last,
             uint32_t* __restrict__ tile_done,
             uint32_t* __restrict__ tile_consumed, unsigned long long* __restrict__ consumed_total,
             unsigned long long* __restrict__ done_tiles,
             const float* __restrict__ zdepth, const float* __restrict__ scene_depth)
{
    __shared__ __align__(16) Record srec[2][BL_BATCH];
    __shared__ float sz[DEPTH ? 2 : 1][DEPTH ? BL_BATCH : 1];
    __shared__ uint32_t s_consumed;
    __shared__ uint32_t swq[QUEUE ? BL_THREADS / 32 : 1][QUEUE ? 32 : 1];

    const int tile = blockIdx.x;
    const int ty = tile / F.tiles_x, tx = tile - ty * F.tiles_x;
    if (!owns_row(ty, F.row_rank, F.row_world, F.row_group)) return;      // CTA-uniform
    uint32_t* const done_word = tile_done + ty * done_words_per_row(F.tiles_x) + (tx >> 5);
    if (!first && ((*done_word >> (tx & 31)) & 1u) != 0u) return;         // saturated and finalised in an earlier chunk

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx = tx * TILE + (warp & 1) * 8, by = ty * TILE + (warp >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < F.width && py < F.height;
    float fpx = (float)px + 0.5f, fpy = (float)py + 0.5f;
    asm volatile("" : "+f"(fpx), "+f"(fpy));       // keep the pixel centre in registers (ptxas re-materialised fpx per visit)
    // pixel-centre box of this warp's 8x4 block
    const float wx_lo = (float)bx + 0.5f, wx_hi = (float)bx + 7.5f;
    const float wy_lo = (float)by + 0.5f, wy_hi = (float)by + 3.5f;
    const float wx_mid = (float)bx + 4.0f, wy_mid = (float)by + 2.0f;
    const float eps = F.eps_t;

    const uint2 range = ranges[tile];
    const uint32_t start = range.x, len = range.y - range.x;
    if (!first && !last && len == 0u) return;                             // nothing to add in this chunk
    if (tid == 0) s_consumed = 0u;
    __syncthreads();

    // A pixel is live while T >= eps; pixels outside the image carry T = -1 (never live, never stored), so the
    // transmittance itself is the "done" state and the inner loop needs no separate flag.
    float Cr = 0.0f, Cg = 0.0f, Cb = 0.0f, T = inside ? 1.0f : -1.0f;
    if (!first && inside) {                                               // between chunks fb holds (C, T)
        const float4 st = fb[(size_t)py * F.width + px];
        Cr = st.x; Cg = st.y; Cb = st.z; T = st.w;
    }
    // scene depth at this pixel; the farthest one of the warp's block culls whole instances
    float sd = 0.0f, sd_max = 0.0f;
    const bool lequal = F.depth_func == 2;
    if (DEPTH) {
        sd = inside ? __ldg(scene_depth + (size_t)py * F.width + px) : -1.0e30f;
        sd_max = sd;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sd_max = fmaxf(sd_max, __shfl_xor_sync(0xffffffffu, sd_max, o));
    }
    bool warp_done = __all_sync(0xffffffffu, T < eps);
    uint32_t done_pos = 0;                       // instances traversed when this pixel saturated (0: it started saturated)
    const uint32_t nb = (len + BL_BATCH - 1) / BL_BATCH;

    auto stage = [&](uint32_t b) {
        const uint32_t k = b * BL_BATCH + tid;
        if (k < len) {
            const uint32_t ref = __ldg(inst + start + k);
            const char* src = reinterpret_cast<const char*>(recs + ref);
            char* dst = reinterpret_cast<char*>(&srec[b & 1][tid]);
            cp_async16(dst, src); cp_async16(dst + 16, src + 16); cp_async16(dst + 32, src + 32);
            if (DEPTH) cp_async4(&sz[b & 1][tid], zdepth + ref);
        }
        cp_async_commit();
    };

    if (nb > 0) stage(0);
    for (uint32_t b = 0; b < nb; ++b) {
        if (b + 1 < nb) stage(b + 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const Record* buf = srec[b & 1];
        const uint32_t count = min((uint32_t)BL_BATCH, len - b * BL_BATCH);
        if (!warp_done) {
            for (uint32_t c = 0; c < count && !warp_done; c += 32) {
                const uint32_t my = c + lane;
                bool ov = false;
                if (my < count) {
                    const float2 cc = *reinterpret_cast<const float2*>(&buf[my].cx);
                    const uint32_t hp = buf[my].hpack;
                    const float hx = __half2float(__ushort_as_half((unsigned short)(hp & 0xffffu)));
                    const float hy = __half2float(__ushort_as_half((unsigned short)(hp >> 16)));
                    ov = (cc.x - hx <= wx_hi) && (cc.x + hx >= wx_lo) && (cc.y - hy <= wy_hi) && (cc.y + hy >= wy_lo);
                    if (DEPTH) { const float zw = sz[b & 1][my]; ov = ov && (lequal ? (zw <= sd_max) : (zw < sd_max)); }
                    if (OBB_CULL && ov) {
                        // second separating-axis test, in the splat's eigen space: the block's pixel centres map into the box
                        // qc +- (rx, ry); if that box misses |qx| <= 2, |qy| <= 2 or |q|^2 <= pmax no pixel of the block is
                        // covered.  Margins (1e-5 of the operand magnitudes, 1e-4 on the thresholds) are orders of magnitude
                        // above fp32 rounding of the per-pixel test below, so the cull is conservative: the frame is unchanged.
                        const float2 m0 = *reinterpret_cast<const float2*>(&buf[my].m00);
                        const float2 m1 = *reinterpret_cast<const float2*>(&buf[my].m10);
                        const float4 ma = make_float4(m0.x, m0.y, m1.x, m1.y);                     // m00 m01 m10 m11
                        const float pm = buf[my].pmax;
                        const float ddx = wx_mid - cc.x, ddy = wy_mid - cc.y;
                        const float qcx = fmaf(ddy, ma.y, ddx * ma.x), qcy = fmaf(ddy, ma.w, ddx * ma.z);
                        const float a00 = fabsf(ma.x), a01 = fabsf(ma.y), a10 = fabsf(ma.z), a11 = fabsf(ma.w);
                        const float adx = fabsf(ddx), ady = fabsf(ddy);
                        const float rx = fmaf(1.5f, a01, 3.5f * a00), ry = fmaf(1.5f, a11, 3.5f * a10);
                        const float ex = 1e-5f * fmaf(a01, ady + 1.5f, a00 * (adx + 3.5f));
                        const float ey = 1e-5f * fmaf(a11, ady + 1.5f, a10 * (adx + 3.5f));
                        const float gx = fmaxf(fabsf(qcx) - rx - ex, 0.0f), gy = fmaxf(fabsf(qcy) - ry - ey, 0.0f);
                        ov = (gx <= 2.0001f) && (gy <= 2.0001f) && (fmaf(gy, gy, gx * gx) <= fmaf(pm, 1.0001f, 1e-4f));
                    }
                }
                unsigned mask = __ballot_sync(0xffffffffu, ov);
                if (QUEUE) {
                    // the survivors' shared-memory byte offsets are compacted into a small per-warp queue, so the visit loop
                    // reads its record address with one broadcast load instead of recomputing it from the ballot mask
                    // (bit reverse + find-leading-one + two integer multiply-adds per visit in the SASS of the mask walk)
                    const int nq = __popc(mask);
                    if (ov) swq[warp][__popc(mask & ((1u << lane) - 1u))] = (uint32_t)((c + lane) * (uint32_t)sizeof(Record));
                    __syncwarp();
                    const char* bufc = reinterpret_cast<const char*>(buf);
                    uint32_t done_off = 0xffffffffu;          // queue entry that saturated this pixel (converted after the loop)
                    for (int i = 0; i < nq; ++i) {
                        const uint32_t off = swq[warp][i];
                        const float4* rp = reinterpret_cast<const float4*>(bufc + off);
                        const float4 r0 = rp[0], r1 = rp[1];
                        bool zpass = true;
                        if (DEPTH) { const float zw = sz[b & 1][off / (uint32_t)sizeof(Record)]; zpass = lequal ? (zw <= sd) : (zw < sd); }
                        if (T >= eps && zpass) {
                            const float dx = fpx - r0.x, dy = fpy - r0.y;
                            const float qx = fmaf(dy, r0.w, dx * r0.z);
                            const float qy = fmaf(dy, r1.y, dx * r1.x);
                            const float pw = fmaf(qy, qy, qx * qx);
                            if (fabsf(qx) <= 2.0f && fabsf(qy) <= 2.0f && pw <= r1.w) {
                                const float4 r2 = rp[2];
                                const float A = fminf(r1.z * ex2_mufu(pw * -1.4426950408889634f), 1.0f);   // alpha * exp(-pw)
                                const float w = T * A;
                                Cr = fmaf(w, r2.x, Cr); Cg = fmaf(w, r2.y, Cg); Cb = fmaf(w, r2.z, Cb);
                                T = T - w;
                                if (T < eps) done_off = off;
                            }
                        }
                    }
                    if (done_off != 0xffffffffu) done_pos = b * BL_BATCH + done_off / (uint32_t)sizeof(Record) + 1u;
                    __syncwarp();
                } else
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float4* rp = reinterpret_cast<const float4*>(&buf[c + j]);
                    const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
                    bool zpass = true;
                    if (DEPTH) { const float zw = sz[b & 1][c + j]; zpass = lequal ? (zw <= sd) : (zw < sd); }
                    if (T >= eps && zpass) {
                        const float dx = fpx - r0.x, dy = fpy - r0.y;
                        const float qx = fmaf(dy, r0.w, dx * r0.z);
                        const float qy = fmaf(dy, r1.y, dx * r1.x);
                        const float pw = fmaf(qy, qy, qx * qx);
                        if (fabsf(qx) <= 2.0f && fabsf(qy) <= 2.0f && pw <= r1.w) {
                            const float A = fminf(r1.z * ex2_mufu(pw * -1.4426950408889634f), 1.0f);   // alpha * exp(-pw)
                            const float w = T * A;
                            Cr = fmaf(w, r2.x, Cr); Cg = fmaf(w, r2.y, Cg); Cb = fmaf(w, r2.z, Cb);
                            T = T - w;
                            if (T < eps) done_pos = b * BL_BATCH + c + (uint32_t)j + 1u;
                        }
                    }
                }
                // early-out vote once per 32 instances, not per visit: a saturated warp may walk the rest of its group
                // with every lane predicated off (no effect on the frame); done_pos keeps the exact position
                warp_done = __all_sync(0xffffffffu, T < eps);
            }
        }
        // barrier: everyone is finished with buf before it is refilled; also the CTA-wide early-out vote
        if (__syncthreads_and(warp_done ? 1 : 0)) break;
    }
    cp_async_wait<0>();

    // the warp saturated where its last pixel did
    const uint32_t warp_pos = __reduce_max_sync(0xffffffffu, done_pos);
    if (lane == 0) atomicMax(&s_consumed, warp_done ? warp_pos : len);
    const bool tile_saturated = __syncthreads_and(warp_done ? 1 : 0) != 0;   // also orders the atomicMax
    if (inside) {
        // finished tiles go to the final frame (possibly peer memory over NVLink), unfinished ones keep (C, T) locally
        if (tile_saturated || last) fb_final[(size_t)py * F.width + px] = make_float4(Cr, Cg, Cb, 1.0f - T);
        else fb[(size_t)py * F.width + px] = make_float4(Cr, Cg, Cb, T);
    }
    if (tid == 0) {
        if (tile_saturated) { atomicOr(done_word, 1u << (tx & 31)); atomicAdd(done_tiles, 1ull); }
        if (tile_consumed) tile_consumed[tile] += s_consumed;
        if (consumed_total && s_consumed) atomicAdd(consumed_total, (unsigned long long)s_consumed);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// blend2 (r02): the same tile traversal, per-pixel arithmetic and early-out, with the visit loop split in two phases so
// that no issue slot is spent on per-visit address arithmetic, record re-loads or loop control.
//   The r01 kernel walks the instances that survive a warp's cull one at a time: every visit re-reads the 48-byte record
//   (3 LDS.128 per lane for a value all lanes share), recomputes its address, tests T and branches: 34-40 SASS per visit
//   for ~14 useful lanes.  Here the survivors of a warp are queued (compacted copies in shared memory) and processed 32 at
//   a time:
//     phase A  lane = INSTANCE: the lane keeps its record in registers and evaluates coverage and alpha for all 32 pixels
//              of the warp's 8x4 block (dx*m00, dx*m10 hoisted per column, dy per row: 4 FMA-class + 3 compares + the
//              exp per pixel, every lane busy), writing alpha (0 = not covered) to a 32 x 32 matrix in shared memory;
//     phase B  lane = PIXEL: the lane walks its row of the matrix in depth order (one LDS.128 per four instances, one
//              broadcast LDS.128 for colour) and composites: 10 instructions per instance.
//   The per-pixel operations and their order are exactly the r01 kernel's, so frames are bit-identical.
constexpr int B2_BATCH  = 128;                     // staged records per batch (double buffered)
constexpr int B2_APITCH = 36;                      // floats per pixel row of the alpha matrix (16-byte rows, conflict-free)
constexpr int B2_QUEUE  = 64;                      // queue slots per warp (<= 31 left over + 32 new)

template <bool DEPTH>
struct __align__(16) Blend2Smem {
    Record srec[2][B2_BATCH];
    Record wq[BL_THREADS / 32][B2_QUEUE];          // survivors of the warp's cull; .hpack slot = list position + 1
    float  sA[BL_THREADS / 32][32 * B2_APITCH];
    float  sz[2][DEPTH ? B2_BATCH : 1];            // window depths of the staged / queued instances (DEPTH only)
    float  wz[BL_THREADS / 32][DEPTH ? B2_QUEUE : 1];
    uint32_t consumed;
};

template <bool DEPTH>
__global__ void __launch_bounds__(BL_THREADS)
blend2_kernel(const Record* __restrict__ recs, const uint32_t* __restrict__ inst,
              const uint2* __restrict__ ranges, float4* __restrict__ fb, float4* __restrict__ fb_final,
              const __grid_constant__ FrameConsts F, const int first, const int last,
              uint32_t* __restrict__ tile_done,
              uint32_t* __restrict__ tile_consumed, unsigned long long* __restrict__ consumed_total,
              unsigned long long* __restrict__ done_tiles,
              const float* __restrict__ zdepth, const float* __restrict__ scene_depth)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Blend2Smem<DEPTH>& sm = *reinterpret_cast<Blend2Smem<DEPTH>*>(smem_raw);

    const int tile = blockIdx.x;
    const int ty = tile / F.tiles_x, tx = tile - ty * F.tiles_x;
    if (!owns_row(ty, F.row_rank, F.row_world, F.row_group)) return;      // CTA-uniform
    uint32_t* const done_word = tile_done + ty * done_words_per_row(F.tiles_x) + (tx >> 5);
    if (!first && ((*done_word >> (tx & 31)) & 1u) != 0u) return;         // saturated and finalised in an earlier chunk

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx = tx * TILE + (warp & 1) * 8, by = ty * TILE + (warp >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < F.width && py < F.height;
    const float wx_lo = (float)bx + 0.5f, wx_hi = (float)bx + 7.5f;
    const float wy_lo = (float)by + 0.5f, wy_hi = (float)by + 3.5f;
    const float wx_mid = (float)bx + 4.0f, wy_mid = (float)by + 2.0f;
    const float eps = F.eps_t;

    const uint2 range = ranges[tile];
    const uint32_t start = range.x, len = range.y - range.x;
    if (!first && !last && len == 0u) return;                             // nothing to add in this chunk
    if (tid == 0) sm.consumed = 0u;
    __syncthreads();

    float Cr = 0.0f, Cg = 0.0f, Cb = 0.0f, T = inside ? 1.0f : -1.0f;
    if (!first && inside) {                                               // between chunks fb holds (C, T)
        const float4 st = fb[(size_t)py * F.width + px];
        Cr = st.x; Cg = st.y; Cb = st.z; T = st.w;
    }
    float sd = 0.0f, sd_max = 0.0f;
    const bool lequal = F.depth_func == 2;
    if (DEPTH) {
        sd = inside ? __ldg(scene_depth + (size_t)py * F.width + px) : -1.0e30f;
        sd_max = sd;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sd_max = fmaxf(sd_max, __shfl_xor_sync(0xffffffffu, sd_max, o));
    }
    bool warp_done = __all_sync(0xffffffffu, T < eps);
    uint32_t done_pos = 0;                       // instances traversed when this pixel saturated (0: it started saturated)
    const uint32_t nb = (len + B2_BATCH - 1) / B2_BATCH;
    Record* const wq = sm.wq[warp];
    float* const sA = sm.sA[warp];
    int qn = 0;                                  // queued survivors of this warp (warp-uniform)

    auto stage = [&](uint32_t b) {
        if (tid < B2_BATCH) {
            const uint32_t k = b * B2_BATCH + tid;
            if (k < len) {
                const uint32_t ref = __ldg(inst + start + k);
                const char* src = reinterpret_cast<const char*>(recs + ref);
                char* dst = reinterpret_cast<char*>(&sm.srec[b & 1][tid]);
                cp_async16(dst, src); cp_async16(dst + 16, src + 16); cp_async16(dst + 32, src + 32);
                if (DEPTH) cp_async4(&sm.sz[b & 1][tid], zdepth + ref);
            }
        }
        cp_async_commit();
    };

    // 32 (or, at the end of the list, n < 32) queued survivors against the warp's 8x4 block
    auto process_group = [&](const int n) {
        // ---- phase A: lane = instance
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, -1.0f);      // pmax < 0: covers nothing
        if (lane < n) { const float4* rp = reinterpret_cast<const float4*>(&wq[lane]); r0 = rp[0]; r1 = rp[1]; }
        float a[8], bq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float dx = ((float)(bx + k) + 0.5f) - r0.x;
            a[k] = dx * r0.z; bq[k] = dx * r1.x;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float dy = ((float)(by + r) + 0.5f) - r0.y;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float qx = fmaf(dy, r0.w, a[k]);
                const float qy = fmaf(dy, r1.y, bq[k]);
                const float pw = fmaf(qy, qy, qx * qx);
                // |qx| <= 2 and |qy| <= 2 as one max (a NaN q makes pw NaN and fails the second test, as in the r01 kernel)
                const bool cov = fmaxf(fabsf(qx), fabsf(qy)) <= 2.0f && pw <= r1.w;
                const float A = fminf(r1.z * ex2_mufu(pw * -1.4426950408889634f), 1.0f);         // alpha * exp(-pw)
                sA[(r * 8 + k) * B2_APITCH + lane] = cov ? A : 0.0f;
            }
        }
        __syncwarp();
        // ---- phase B: lane = pixel (r = lane >> 3, k = lane & 7), instances in depth order
        const float* myA = sA + lane * B2_APITCH;
        for (int j4 = 0; j4 < n; j4 += 4) {
            const float4 A4 = *reinterpret_cast<const float4*>(myA + j4);
            const float Av[4] = { A4.x, A4.y, A4.z, A4.w };
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 c = *reinterpret_cast<const float4*>(&wq[j4 + u].r);              // r g b position (broadcast)
                bool hit = Av[u] > 0.0f && T >= eps;
                if (DEPTH) { const float zw = sm.wz[warp][j4 + u]; hit = hit && (lequal ? (zw <= sd) : (zw < sd)); }
                const float w = T * Av[u];
                const float Tn = T - w;
                // short enough to be predicated: no branch per instance
                Cr = hit ? fmaf(w, c.x, Cr) : Cr; Cg = hit ? fmaf(w, c.y, Cg) : Cg; Cb = hit ? fmaf(w, c.z, Cb) : Cb;
                T = hit ? Tn : T;
                done_pos = (hit && Tn < eps) ? __float_as_uint(c.w) : done_pos;
            }
        }
        __syncwarp();
    };

    if (nb > 0) stage(0);
    for (uint32_t b = 0; b < nb; ++b) {
        if (b + 1 < nb) stage(b + 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const Record* buf = sm.srec[b & 1];
        const uint32_t count = min((uint32_t)B2_BATCH, len - b * B2_BATCH);
        if (!warp_done) {
            for (uint32_t c = 0; c < count && !warp_done; c += 32) {
                const uint32_t my = c + lane;
                bool ov = false;
                if (my < count) {
                    const float2 cc = *reinterpret_cast<const float2*>(&buf[my].cx);
                    const uint32_t hp = buf[my].hpack;
                    const float hx = __half2float(__ushort_as_half((unsigned short)(hp & 0xffffu)));
                    const float hy = __half2float(__ushort_as_half((unsigned short)(hp >> 16)));
                    ov = (cc.x - hx <= wx_hi) && (cc.x + hx >= wx_lo) && (cc.y - hy <= wy_hi) && (cc.y + hy >= wy_lo);
                    if (DEPTH) { const float zw = sm.sz[b & 1][my]; ov = ov && (lequal ? (zw <= sd_max) : (zw < sd_max)); }
                    if (ov) {
                        // second separating-axis test in the splat's eigen space (see blend_kernel): conservative
                        const float2 m0 = *reinterpret_cast<const float2*>(&buf[my].m00);
                        const float2 m1 = *reinterpret_cast<const float2*>(&buf[my].m10);
                        const float pm = buf[my].pmax;
                        const float ddx = wx_mid - cc.x, ddy = wy_mid - cc.y;
                        const float qcx = fmaf(ddy, m0.y, ddx * m0.x), qcy = fmaf(ddy, m1.y, ddx * m1.x);
                        const float a00 = fabsf(m0.x), a01 = fabsf(m0.y), a10 = fabsf(m1.x), a11 = fabsf(m1.y);
                        const float adx = fabsf(ddx), ady = fabsf(ddy);
                        const float rx = fmaf(1.5f, a01, 3.5f * a00), ry = fmaf(1.5f, a11, 3.5f * a10);
                        const float ex = 1e-5f * fmaf(a01, ady + 1.5f, a00 * (adx + 3.5f));
                        const float ey = 1e-5f * fmaf(a11, ady + 1.5f, a10 * (adx + 3.5f));
                        const float gx = fmaxf(fabsf(qcx) - rx - ex, 0.0f), gy = fmaxf(fabsf(qcy) - ry - ey, 0.0f);
                        ov = (gx <= 2.0001f) && (gy <= 2.0001f) && (fmaf(gy, gy, gx * gx) <= fmaf(pm, 1.0001f, 1e-4f));
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, ov);
                if (mask) {
                    if (ov) {                    // queue a compacted copy of the record; its hpack slot becomes the list position
                        const int slot = qn + __popc(mask & ((1u << lane) - 1u));
                        const float4* src = reinterpret_cast<const float4*>(&buf[my]);
                        float4* dst = reinterpret_cast<float4*>(&wq[slot]);
                        float4 c2 = src[2];
                        c2.w = __uint_as_float(b * B2_BATCH + my + 1u);
                        dst[0] = src[0]; dst[1] = src[1]; dst[2] = c2;
                        if (DEPTH) sm.wz[warp][slot] = sm.sz[b & 1][my];
                    }
                    qn += __popc(mask);
                    __syncwarp();
                    if (qn >= 32) {
                        process_group(32);
                        const int rest = qn - 32;        // move the left-over entries to the front of the queue
                        float4 t0, t1, t2; float tz = 0.0f;
                        if (lane < rest) {
                            const float4* src = reinterpret_cast<const float4*>(&wq[32 + lane]);
                            t0 = src[0]; t1 = src[1]; t2 = src[2];
                            if (DEPTH) tz = sm.wz[warp][32 + lane];
                        }
                        __syncwarp();
                        if (lane < rest) {
                            float4* dst = reinterpret_cast<float4*>(&wq[lane]);
                            dst[0] = t0; dst[1] = t1; dst[2] = t2;
                            if (DEPTH) sm.wz[warp][lane] = tz;
                        }
                        __syncwarp();
                        qn = rest;
                        warp_done = __all_sync(0xffffffffu, T < eps);
                    }
                }
            }
            // the end of the tile's list: whatever is still queued
            if (b + 1 == nb && !warp_done && qn > 0) {
                process_group(qn);
                qn = 0;
                warp_done = __all_sync(0xffffffffu, T < eps);
            }
        }
        // barrier: everyone is finished with buf before it is refilled; also the CTA-wide early-out vote
        if (__syncthreads_and(warp_done ? 1 : 0)) break;
    }
    cp_async_wait<0>();

    const uint32_t warp_pos = __reduce_max_sync(0xffffffffu, done_pos);
    if (lane == 0) atomicMax(&sm.consumed, warp_done ? warp_pos : len);
    const bool tile_saturated = __syncthreads_and(warp_done ? 1 : 0) != 0;   // also orders the atomicMax
    if (inside) {
        if (tile_saturated || last) fb_final[(size_t)py * F.width + px] = make_float4(Cr, Cg, Cb, 1.0f - T);
        else fb[(size_t)py * F.width + px] = make_float4(Cr, Cg, Cb, T);
    }
    if (tid == 0) {
        if (tile_saturated) { atomicOr(done_word, 1u << (tx & 31)); atomicAdd(done_tiles, 1ull); }
        if (tile_consumed) tile_consumed[tile] += sm.consumed;
        if (consumed_total && sm.consumed) atomicAdd(consumed_total, (unsigned long long)sm.consumed);
    }
}

}  // namespace

void launch_blend(const Record* recs, const uint32_t* inst_vals, const uint2* ranges, float4* fb, float4* fb_final,
                  FrameConsts fc, int first, int last, uint32_t* tile_done, uint32_t* tile_consumed,
                  unsigned long long* consumed_total, unsigned long long* done_tiles,
                  const float* zdepth, const float* scene_depth, cudaStream_t s)
{
    const int tiles = fc.tiles_x * fc.tiles_y;
    if (tiles <= 0) return;
    const char* e = getenv("GSB_BLEND_OBB");   // GSB_BLEND_OBB=0 turns the eigen-space cull off (experiments)
    const int obb = (e && atoi(e) == 0) ? 0 : 1;
    const char* eq = getenv("GSB_BLEND_QUEUE"); // GSB_BLEND_QUEUE=0: walk the ballot mask instead of the survivor queue
    const int queue = (eq && atoi(eq) == 0) ? 0 : 1;
    const bool depth = fc.depth_func != 0 && zdepth && scene_depth;
    static const int two_phase = [] { const char* e2 = getenv("GSB_BLEND2"); return (e2 && atoi(e2) != 0) ? 1 : 0; }();   // GSB_BLEND2=1: blend2_kernel
    if (two_phase) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(blend2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Blend2Smem<false>));
            cudaFuncSetAttribute(blend2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Blend2Smem<true>));
            attr = true;
        }
        if (depth) blend2_kernel<true><<<tiles, BL_THREADS, sizeof(Blend2Smem<true>), s>>>(recs, inst_vals, ranges, fb, fb_final ? fb_final : fb, fc, first,
                       last, tile_done, tile_consumed, consumed_total, done_tiles, zdepth, scene_depth);
        else blend2_kernel<false><<<tiles, BL_THREADS, sizeof(Blend2Smem<false>), s>>>(recs, inst_vals, ranges, fb, fb_final ? fb_final : fb, fc, first,
                       last, tile_done, tile_consumed, consumed_total, done_tiles, zdepth, scene_depth);
        return;
    }
#define GSB_BLEND(O, D, Q) blend_kernel<O, D, Q><<<tiles, BL_THREADS, 0, s>>>(recs, inst_vals, ranges, fb, fb_final ? fb_final : fb, fc, first, \
                            last, tile_done, tile_consumed, consumed_total, done_tiles, zdepth, scene_depth)
#define GSB_BLEND_Q(O, D) do { if (queue) GSB_BLEND(O, D, true); else GSB_BLEND(O, D, false); } while (0)
    if (depth) { if (obb) GSB_BLEND_Q(true, true); else GSB_BLEND_Q(false, true); }
    else       { if (obb) GSB_BLEND_Q(true, false); else GSB_BLEND_Q(false, false); }
#undef GSB_BLEND_Q
#undef GSB_BLEND
}

}  // namespace gsb

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
// 32 instances at a time (one instance per lane, conservative AABB + eigen-space test, ballot), queues the
// survivors' shared-memory offsets per warp and walks that queue in order, so a small splat costs one warp pass instead
// of eight and a visit fetches its record address with one broadcast load (r02: 40 -> 34 SASS per visit).  Transmittance
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
template <bool OBB_CULL, bool DEPTH>
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
    __shared__ uint32_t swq[BL_THREADS / 32][32];

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
                {
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
    const bool depth = fc.depth_func != 0 && zdepth && scene_depth;
#define GSB_BLEND(O, D) blend_kernel<O, D><<<tiles, BL_THREADS, 0, s>>>(recs, inst_vals, ranges, fb, fb_final ? fb_final : fb, fc, first, \
                            last, tile_done, tile_consumed, consumed_total, done_tiles, zdepth, scene_depth)
    if (depth) { if (obb) GSB_BLEND(true, true); else GSB_BLEND(false, true); }
    else       { if (obb) GSB_BLEND(true, false); else GSB_BLEND(false, false); }
#undef GSB_BLEND
}

}  // namespace gsb

// project.cu — K0 pack (active-set change), K1 cull + depth key + tile rectangle (every splat, every frame) and
// K2 record + SH colour (only splats that reach a live tile), sm_100a.
//
// K0 replaces the CPU texture packing of GSplatRenderer::generateRenderGeometry
//    (/root/reference/gsplat_plugin/src/GSplatRenderer.C:448-505): the per-prim SoA arrays that
//    registerUpdate received are re-laid into 16-byte planes so K1's loads are coalesced LDG.128.
// K1 replaces, once per splat instead of once per vertex (the reference repeats it 6x):
//    depth key                        src/GSplatRenderer.C:196-202
//    centre + cull                    shaders/GSplatShaderSource.h:198-214, 277-282
//    covariance chain + eigen axes    shaders/GSplatShaderCoreLib.h:10-93
//    SH -> RGB                        shaders/GSplatShaderCoreLib.h:103-179, GSplatShaderSource.h:244-275
//
// The fp32 expression order below IS the spec and matches oracle/gsplat_oracle.cpp
// operation for operation: this TU is compiled with -fmad=false, IEEE division and sqrt
// (-prec-div=true -prec-sqrt=true, no fast-math), so keys, records and rectangles are bit-exact.
// K1 streams 40 B per submitted splat (position + discard radius, cached world-space covariance) and writes 8 B (key,
// packed tile rectangle); it is FP32-issue bound (IEEE divisions and square roots of the spec).  K2 gathers one
// 128-byte line per live splat and writes its 48-byte record, so colour / SH / record traffic is proportional to the
// splats that can still change a pixel (3 M of 20 M at 20 M / 1080p), not to the cloud.
#include "common.cuh"
#include <cstdlib>

namespace gsb {

namespace {

#define MAT(M, r, c) ((M)[(c) * 4 + (r)])

__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ float lo_h(uint32_t w) { return h2f((uint16_t)(w & 0xffffu)); }
__device__ __forceinline__ float hi_h(uint32_t w) { return h2f((uint16_t)(w >> 16)); }

// ln(x), x > 0, in double with + - * / only: the same sequence as oracle det_log() => same bits.
__device__ __forceinline__ double det_log(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    int k = (int)((u >> 52) & 0x7ffull) - 1023;
    u = (u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m = __longlong_as_double((long long)u);
    if (m > 1.4142135623730951) { m = m * 0.5; k = k + 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double p = 1.0 / 19.0;
    p = p * s2 + 1.0 / 17.0;
    p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0;
    p = p * s2 + 1.0 / 11.0;
    p = p * s2 + 1.0 / 9.0;
    p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0;
    p = p * s2 + 1.0 / 3.0;
    p = p * s2 + 1.0;
    return (2.0 * s) * p + (double)k * 0.6931471805599453;
}

// Discard radius of a splat: pmax = ln(255 alpha) as fp32 (SRC.h:308: the fragment survives iff exp(-|q|^2) alpha >= 1/255),
// or -1 if the splat can never pass the discard (alpha < 1/255, NaN).  View independent, so pack computes it ONCE per
// geometry change and K1 streams it instead of alpha: the double-precision logarithm leaves the per-frame path.
__device__ __forceinline__ float pmax_of_alpha(const float alpha)
{
    if (!(alpha >= 1.0f / 255.0f)) return -1.0f;
    const float pmax = (float)det_log((double)alpha * 255.0);
    return (pmax >= 0.0f) ? pmax : -1.0f;
}

// rr = sqrt(pmax), the discard radius in eigen space (what K1 needs for the pixel rectangle), or -1
__device__ __forceinline__ float rr_of_alpha(const float alpha)
{
    const float pmax = pmax_of_alpha(alpha);
    return (pmax >= 0.0f) ? sqrtf(pmax) : -1.0f;
}

// ------------------------------------------------------------------------------------ K0 pack
// One thread per splat of one registered prim.  Source layout = what registerUpdate receives
// (R.h:34-47): pos f32x3, Cd h3, alpha f32, scale h3, orient h4 (x,y,z,w), SH 3 x half[16].
__global__ void __launch_bounds__(256)
pack_kernel(const float* __restrict__ pos, const uint16_t* __restrict__ cd, const float* __restrict__ alpha,
            const uint16_t* __restrict__ scale, const uint16_t* __restrict__ orient,
            const uint16_t* __restrict__ shx, const uint16_t* __restrict__ shy, const uint16_t* __restrict__ shz,
            int64_t count, int64_t dst, float4* __restrict__ geomA, uint4* __restrict__ geomB,
            uint4* __restrict__ rows, int has_sh)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t o = dst + i;
    const float4 ga = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], alpha[i]);
    uint32_t s0 = scale[3 * i], s1 = scale[3 * i + 1], s2 = scale[3 * i + 2];
    uint32_t q0 = orient[4 * i], q1 = orient[4 * i + 1], q2 = orient[4 * i + 2], q3 = orient[4 * i + 3];
    const uint4 gb = make_uint4(s0 | (s1 << 16), s2 | (q0 << 16), q1 | (q2 << 16), q3);
    geomA[o] = make_float4(ga.x, ga.y, ga.z, rr_of_alpha(ga.w)); geomB[o] = gb;
    uint4* row = rows + o * ROW_U4;
    row[0] = make_uint4(__float_as_uint(ga.x), __float_as_uint(ga.y), __float_as_uint(ga.z), __float_as_uint(ga.w));
    row[1] = gb;
    uint16_t h[48];
    h[0] = cd[3 * i]; h[1] = cd[3 * i + 1]; h[2] = cd[3 * i + 2];
    if (has_sh) {
#pragma unroll
        for (int j = 0; j < 15; ++j) {
            h[3 + 3 * j] = shx[16 * i + j]; h[4 + 3 * j] = shy[16 * i + j]; h[5 + 3 * j] = shz[16 * i + j];
        }
    } else {
#pragma unroll
        for (int j = 3; j < 48; ++j) h[j] = 0;
    }
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        uint4 w;
        w.x = (uint32_t)h[8 * p + 0] | ((uint32_t)h[8 * p + 1] << 16);
        w.y = (uint32_t)h[8 * p + 2] | ((uint32_t)h[8 * p + 3] << 16);
        w.z = (uint32_t)h[8 * p + 4] | ((uint32_t)h[8 * p + 5] << 16);
        w.w = (uint32_t)h[8 * p + 6] | ((uint32_t)h[8 * p + 7] << 16);
        row[2 + p] = w;
    }
}

// ------------------------------------------------------------------------------------ K1 project
constexpr float SH_C1   = 0.4886025f;
constexpr float SH_C2_0 = 1.0925484f, SH_C2_1 = -1.0925484f, SH_C2_2 = 0.3153916f,
                SH_C2_3 = -1.0925484f, SH_C2_4 = 0.5462742f;
constexpr float SH_C3_0 = -0.5900436f, SH_C3_1 = 2.8906114f, SH_C3_2 = -0.4570458f,
                SH_C3_3 = 0.3731763f, SH_C3_4 = -0.4570458f, SH_C3_5 = 1.4453057f,
                SH_C3_6 = -0.5900436f;

// Geometry of one projected splat: everything the record needs except the colour.
struct Geom {
    float cx, cy, m00, m01, m10, m11, hx, hy;
    int   x0, x1, y0, y1;
    float psx[3];            // shader-side position (P - origin) + origin
    float clipz, clipw;      // centre clip z and w: every fragment of the quad has depth clipz / clipw (SRC.h:278-282)
};

// Centre, cull, covariance chain, eigen axes, discard radius, pixel rectangle.  Returns false if culled.
// The expression order is the spec; oracle/gsplat_oracle.cpp project_one() is its twin.
// World-space covariance Sigma = O3 R S^2 R^T O3^T of one splat (LIB.h:10-35), upper triangle S00 S01 S02 S11 S12 S22.
// Depends on scale, orient and the object matrix only, so it is cached per splat (sigma planes) and rebuilt only when
// the object matrix changes; K2 recomputes it from the splat's line.  Same operations either way => same bits.
struct Sigma { float s[6]; };
__device__ __forceinline__ Sigma sigma_of(const float* __restrict__ object, const uint4 gb)
{
    const float sx = lo_h(gb.x), sy = hi_h(gb.x), sz = lo_h(gb.y);
    const float qx = hi_h(gb.y), qy = lo_h(gb.z), qz = hi_h(gb.z), qr = lo_h(gb.w);

    float Rt[3][3];
    Rt[0][0] = 1.0f - 2.0f * (qy * qy + qz * qz); Rt[0][1] = 2.0f * (qx * qy + qr * qz); Rt[0][2] = 2.0f * (qx * qz - qr * qy);
    Rt[1][0] = 2.0f * (qx * qy - qr * qz); Rt[1][1] = 1.0f - 2.0f * (qx * qx + qz * qz); Rt[1][2] = 2.0f * (qy * qz + qr * qx);
    Rt[2][0] = 2.0f * (qx * qz + qr * qy); Rt[2][1] = 2.0f * (qy * qz - qr * qx); Rt[2][2] = 1.0f - 2.0f * (qx * qx + qy * qy);
    const float sc[3] = { sx, sy, sz };
    float Mm[3][3], M2[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Mm[a][b] = sc[a] * Rt[a][b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
            M2[a][b] = (Mm[a][0] * MAT(object, b, 0) + Mm[a][1] * MAT(object, b, 1)) + Mm[a][2] * MAT(object, b, 2);
    Sigma out;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b)
            out.s[k++] = (M2[0][a] * M2[0][b] + M2[1][a] * M2[1][b]) + M2[2][a] * M2[2][b];
    return out;
}

// rr = sqrt(pmax) (or negative: never passes the discard, see rr_of_alpha)
// owned_rows (row-partitioned frames only, else NULL): owned_rows[y] = tile rows < y this rank owns, so "does the
// rectangle touch an owned row" is two look-ups instead of a loop of integer divisions
__device__ __forceinline__ bool project_geom(const FrameConsts& F, const float p[3], const float rr,
                                             const Sigma& sig, Geom& g, const uint32_t* __restrict__ owned_rows)
{
    if (!(rr >= 0.0f)) return false;             // alpha < 1/255
#pragma unroll
    for (int k = 0; k < 3; ++k) { float t = p[k] - F.origin[k]; g.psx[k] = t + F.origin[k]; }
    const float* psx = g.psx;
    float vc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        vc[r] = ((MAT(F.obj_view, r, 0) * psx[0] + MAT(F.obj_view, r, 1) * psx[1]) + MAT(F.obj_view, r, 2) * psx[2]) + MAT(F.obj_view, r, 3);
    const float fy = -vc[1];
    float clip[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        clip[r] = ((MAT(F.proj, r, 0) * vc[0] + MAT(F.proj, r, 1) * fy) + MAT(F.proj, r, 2) * vc[2]) + MAT(F.proj, r, 3);
    const float cw = clip[3];
    if (!(cw > 0.0f)) return false;
    if (!(clip[2] >= -cw && clip[2] <= cw)) return false;
    const float ndcx = clip[0] / cw;
    const float ndcy = (-clip[1]) / cw;
    const float cx = ((ndcx + 1.0f) * 0.5f) * F.W;
    const float cy = ((ndcy + 1.0f) * 0.5f) * F.H;

    const float (&S)[6] = sig.s;        // S00 S01 S02 S11 S12 S22
#define SG(a, b) S[(a) <= (b) ? ((a) == 0 ? (b) : (a) + (b) + 1) : ((b) == 0 ? (a) : (a) + (b) + 1)]

    float t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        t[r] = ((MAT(F.view, r, 0) * psx[0] + MAT(F.view, r, 1) * psx[1]) + MAT(F.view, r, 2) * psx[2]) + MAT(F.view, r, 3);
    const float aspect = MAT(F.proj, 0, 0) / MAT(F.proj, 1, 1);
    const float tanFovX = 1.0f / MAT(F.proj, 0, 0);
    const float tanFovY = 1.0f / (MAT(F.proj, 1, 1) * aspect);
    const float limX = 1.3f * tanFovX, limY = 1.3f * tanFovY;
    const float tz = t[2];
    float rx = t[0] / tz; rx = fminf(fmaxf(rx, -limX), limX);
    float ry = t[1] / tz; ry = fminf(fmaxf(ry, -limY), limY);
    const float tx = rx * tz, ty = ry * tz;
    const float focal = (F.W * MAT(F.proj, 0, 0)) / 2.0f;
    const float j0 = focal / tz;
    const float tz2 = tz * tz;
    const float j2x = -((focal * tx) / tz2);
    const float j2y = -((focal * ty) / tz2);
    float A0[3], A1[3], B0[3], B1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        A0[k] = j0 * MAT(F.view, 0, k) + j2x * MAT(F.view, 2, k);
        A1[k] = j0 * MAT(F.view, 1, k) + j2y * MAT(F.view, 2, k);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        B0[k] = (A0[0] * SG(0, k) + A0[1] * SG(1, k)) + A0[2] * SG(2, k);
        B1[k] = (A1[0] * SG(0, k) + A1[1] * SG(1, k)) + A1[2] * SG(2, k);
    }
    const float c00 = (B0[0] * A0[0] + B0[1] * A0[1]) + B0[2] * A0[2];
    const float c01 = (B0[0] * A1[0] + B0[1] * A1[1]) + B0[2] * A1[2];
    const float c11 = (B1[0] * A1[0] + B1[1] * A1[1]) + B1[2] * A1[2];
    const float a = c00 + 0.3f, b = c01, c = c11 + 0.3f;

    const float mid = 0.5f * (a + c);
    const float hd = (a - c) / 2.0f;
    const float radius = sqrtf(hd * hd + b * b);
    const float l1 = mid + radius;
    const float l2 = fmaxf(mid - radius, 0.1f);
    const float dvx = b, dvy = l1 - a;
    const float len = sqrtf(dvx * dvx + dvy * dvy);
    if (!(len > 0.0f) || !(len <= 3.0e38f)) return false;
    const float ex = dvx / len, ey = dvy / len;
    const float s1 = fminf(sqrtf(2.0f * l1), 4096.0f);
    const float s2 = fminf(sqrtf(2.0f * l2), 4096.0f);
    if (!(s1 > 0.0f) || !(s2 > 0.0f)) return false;
    const float u1x = s1 * ex, u1y = s1 * ey;
    const float u2x = -(s2 * ey), u2y = s2 * ex;

    const float bxh = 2.0f * (fabsf(u1x) + fabsf(u2x));
    const float byh = 2.0f * (fabsf(u1y) + fabsf(u2y));
    const float exh = rr * sqrtf(u1x * u1x + u2x * u2x);
    const float eyh = rr * sqrtf(u1y * u1y + u2y * u2y);
    float hx = fminf(bxh, exh); hx = hx + (hx * 0.0001f + 0.01f);
    float hy = fminf(byh, eyh); hy = hy + (hy * 0.0001f + 0.01f);
    const float x0f = fmaxf(ceilf((cx - hx) - 0.5f), 0.0f);
    const float x1f = fminf(floorf((cx + hx) - 0.5f), F.W - 1.0f);
    const float y0f = fmaxf(ceilf((cy - hy) - 0.5f), 0.0f);
    const float y1f = fminf(floorf((cy + hy) - 0.5f), F.H - 1.0f);
    if (!(x0f <= x1f) || !(y0f <= y1f)) return false;
    g.x0 = (int)x0f; g.x1 = (int)x1f; g.y0 = (int)y0f; g.y1 = (int)y1f;
    if (owned_rows) {
        if (__ldg(owned_rows + g.y1 / TILE + 1) == __ldg(owned_rows + g.y0 / TILE)) return false;
    }
    g.cx = cx; g.cy = cy; g.clipz = clip[2]; g.clipw = cw;
    g.m00 = ex / s1; g.m01 = ey / s1;
    g.m10 = (-ey) / s2; g.m11 = ex / s2;
    g.hx = hx; g.hy = hy;
    return true;
#undef SG
}

// SH -> RGB for splat i (SRC.h:224,244-275; LIB.h:117-179), term order of LIB.h:148-174
// crow: the splat's six 16-byte colour chunks (a staged copy in shared memory)
template <int ORDER>
__device__ __forceinline__ void shade_colour(const FrameConsts& F, const uint4* crow, const float psx[3], float rgb[3])
{
    const uint4* row = crow;
    const uint4 c0 = row[0];
    rgb[0] = lo_h(c0.x); rgb[1] = hi_h(c0.x); rgb[2] = lo_h(c0.y);
    if (ORDER > 0) {
        // 48 halfs: Cd(3) then coefficient j channel ch at 3 + 3j + ch
        uint32_t w[24];
        w[0] = c0.x; w[1] = c0.y; w[2] = c0.z; w[3] = c0.w;
        constexpr int PLANES = ORDER == 1 ? 2 : (ORDER == 2 ? 4 : 6);
#pragma unroll
        for (int pl = 1; pl < PLANES; ++pl) {
            const uint4 cc = row[pl];
            w[4 * pl] = cc.x; w[4 * pl + 1] = cc.y; w[4 * pl + 2] = cc.z; w[4 * pl + 3] = cc.w;
        }
        const float wv[3] = { psx[0] - F.cam[0], psx[1] - F.cam[1], psx[2] - F.cam[2] };
        float ov[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            ov[r] = (MAT(F.inv_object, r, 0) * wv[0] + MAT(F.inv_object, r, 1) * wv[1]) + MAT(F.inv_object, r, 2) * wv[2];
        const float dl = sqrtf((ov[0] * ov[0] + ov[1] * ov[1]) + ov[2] * ov[2]);
        const float x = ov[0] / dl, y = ov[1] / dl, z = ov[2] / dl;
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            auto SH = [&](int j) -> float {      // coefficient j (0-based: sh1 = j 0)
                const int hidx = 3 + 3 * j + ch;
                const uint32_t word = w[hidx >> 1];
                return (hidx & 1) ? hi_h(word) : lo_h(word);
            };
            float res = rgb[ch];
            res = res + SH_C1 * (((-SH(0)) * y + SH(1) * z) - SH(2) * x);
            if (ORDER >= 2) {
                float t2 = (SH_C2_0 * xy) * SH(3);
                t2 = t2 + (SH_C2_1 * yz) * SH(4);
                t2 = t2 + (SH_C2_2 * ((2.0f * zz - xx) - yy)) * SH(5);
                t2 = t2 + (SH_C2_3 * xz) * SH(6);
                t2 = t2 + (SH_C2_4 * (xx - yy)) * SH(7);
                res = res + t2;
                if (ORDER >= 3) {
                    float t3 = ((SH_C3_0 * y) * (3.0f * xx - yy)) * SH(8);
                    t3 = t3 + ((SH_C3_1 * xy) * z) * SH(9);
                    t3 = t3 + ((SH_C3_2 * y) * ((4.0f * zz - xx) - yy)) * SH(10);
                    t3 = t3 + ((SH_C3_3 * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy)) * SH(11);
                    t3 = t3 + ((SH_C3_4 * x) * ((4.0f * zz - xx) - yy)) * SH(12);
                    t3 = t3 + ((SH_C3_5 * z) * (xx - yy)) * SH(13);
                    t3 = t3 + ((SH_C3_6 * x) * (xx - 3.0f * yy)) * SH(14);
                    res = res + t3;
                }
            }
            rgb[ch] = fmaxf(res, 0.0f);
        }
    }
}

// ---- K1: persistent CTAs, grid-stride over the submitted splats; everything is a coalesced stream.
// Per splat: cull + projection (project_geom), depth key on the UNMODIFIED position (R.C:196-202, 454, 584), packed tile
// rectangle.  The colour, the SH evaluation and the 48-byte record are NOT produced here: with per-pixel early-out only
// a small share of the cloud is ever blended, so they are built later (K2) for the splats that reach a live tile.
// The CTA also histograms depth_bucket(key) in shared memory (one flush per CTA) for the chunk plan.
constexpr int K1_THREADS = 256;
template <int MIN_CTAS>
__global__ void __launch_bounds__(K1_THREADS, MIN_CTAS)
project_kernel(const __grid_constant__ FrameConsts F, const float4* __restrict__ geomA,
               const float4* __restrict__ sigA, const float2* __restrict__ sigB, int64_t n, uint32_t* __restrict__ keys, uint2* __restrict__ rects,
               const int rects_all, uint32_t* __restrict__ trects, unsigned long long* __restrict__ n_visible,
               const DepthBuckets db, uint32_t* __restrict__ bucket_hist, const uint32_t* __restrict__ owned_rows)
{
    __shared__ uint32_t sh_hist[DEPTH_BUCKETS];
    if (bucket_hist) {
        for (int b = threadIdx.x; b < DEPTH_BUCKETS; b += K1_THREADS) sh_hist[b] = 0u;
        __syncthreads();
    }
    uint32_t nvis = 0;
    const int64_t stride = (int64_t)gridDim.x * K1_THREADS;
    int64_t i = (int64_t)blockIdx.x * K1_THREADS + threadIdx.x;
    float4 ga_next = make_float4(0.f, 0.f, 0.f, 0.f), sa_next = ga_next; float2 sb_next = make_float2(0.f, 0.f);
    if (i < n) { ga_next = __ldg(geomA + i); sa_next = __ldg(sigA + i); sb_next = __ldg(sigB + i); }
    for (; i < n; i += stride) {
        const float4 ga = ga_next, sa = sa_next;
        const float2 sb = sb_next;
        if (i + stride < n) {                                // prefetch
            ga_next = __ldg(geomA + i + stride); sa_next = __ldg(sigA + i + stride); sb_next = __ldg(sigB + i + stride);
        }
        const float p[3] = { ga.x, ga.y, ga.z };
        const Sigma sig{ { sa.x, sa.y, sa.z, sa.w, sb.x, sb.y } };
        Geom g;
        const bool vis = project_geom(F, p, ga.w, sig, g, owned_rows);
        uint32_t key = KEY_CULLED, tr = TRECT_CULLED;
        uint2 rect = make_uint2(1u, 1u);                     // x0=1,x1=0,y0=1,y1=0 : empty
        bool wide = false;
        if (vis) {
            const float dx = p[0] - F.cam[0], dy = p[1] - F.cam[1], dz = p[2] - F.cam[2];
            key = __float_as_uint(dx * dx + dy * dy + dz * dz);
            const int tx0 = g.x0 / TILE, tx1 = g.x1 / TILE, ty0 = g.y0 / TILE, ty1 = g.y1 / TILE;
            tr = pack_trect(tx0, tx1, ty0, ty1);
            wide = (tx1 - tx0 >= 127) || (ty1 - ty0 >= 127);
            rect = make_uint2((uint32_t)g.x0 | ((uint32_t)g.x1 << 16), (uint32_t)g.y0 | ((uint32_t)g.y1 << 16));
            ++nvis;
        }
        keys[i] = key;
        if (trects) trects[i] = tr;
        if (rects_all || wide) rects[i] = rect;
        if (bucket_hist) atomicAdd(&sh_hist[depth_bucket(key, db)], 1u);
    }
    nvis = __reduce_add_sync(0xffffffffu, nvis);             // one atomic per warp for V
    if ((threadIdx.x & 31) == 0 && nvis) atomicAdd(n_visible, (unsigned long long)nvis);
    if (bucket_hist) {
        __syncthreads();
        for (int b = threadIdx.x; b < DEPTH_BUCKETS; b += K1_THREADS) {
            const uint32_t c = sh_hist[b];
            if (c) atomicAdd(bucket_hist + b, c);
        }
    }
}

// ---- bounded K1 (GSB_OPT_LAZY_PROJECT): 20 B per evaluated splat.
// Exact (the spec's operations): the alpha / clip.w / clip.z culls and the depth key.  Bounded: the pixel rectangle.
// With T = J W (LIB.h:38-76), lambda_1 = lambda_max(T Sigma T^T) + 0.3 <= |J|_2^2 |W|_2^2 lambda_max(Sigma) + 0.3 and
// |J|_2^2 = j0^2 (1 + rx^2 + ry^2) (J J^T = j0^2 I + j2 j2^T, j2 = -j0 (rx, ry) with the clamped ratios rx, ry), so
// s1 = sqrt(2 lambda_1) is bounded; the exact half extents are hx = min(2(|u1x| + |u2x|), rr |(u1x, u2x)|)(1 + 1e-4) + 0.01
// <= min(2 sqrt 2, rr) s1 (1 + 1e-4) + 0.01 because |(u1x, u2x)| <= s1 and |u1x| + |u2x| <= sqrt 2 s1.  Divisions
// and the square root here are the approximate MUFU forms; every approximation and every rounding of the exact
// chain is covered by the explicit margins (relative 2e-3 on the extent, 0.06 px + 2e-6 |c| on the position).
// A splat the bound keeps and the exact projection culls (degenerate axes, empty exact rectangle) gets a zero live-tile
// count in K2 and no instances.  NaN positions behave like the exact kernel's fmaxf/fminf: the rectangle opens up.
struct Bound { bool vis; uint32_t key; int tx0, tx1, ty0, ty1; };
__device__ __forceinline__ Bound bound_one(const FrameConsts& F, const float4 ga, const float lm)
{
    Bound o; o.key = KEY_CULLED; o.tx0 = 1; o.tx1 = 0; o.ty0 = 1; o.ty1 = 0;
    const float p[3] = { ga.x, ga.y, ga.z };
    const float rr = ga.w;
    bool vis = rr >= 0.0f;
    float psx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { float t = p[k] - F.origin[k]; psx[k] = t + F.origin[k]; }
    float vc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        vc[r] = ((MAT(F.obj_view, r, 0) * psx[0] + MAT(F.obj_view, r, 1) * psx[1]) + MAT(F.obj_view, r, 2) * psx[2]) + MAT(F.obj_view, r, 3);
    const float fy = -vc[1];
    float clip[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        clip[r] = ((MAT(F.proj, r, 0) * vc[0] + MAT(F.proj, r, 1) * fy) + MAT(F.proj, r, 2) * vc[2]) + MAT(F.proj, r, 3);
    const float cw = clip[3];
    vis = vis && (cw > 0.0f) && (clip[2] >= -cw && clip[2] <= cw);
    if (vis) {
        // (from here on nothing has to match the exact chain bit for bit: fused multiply-adds, MUFU approximations)
        const float iw = __fdividef(1.0f, cw);
        const float hw = 0.5f * F.W, hh = 0.5f * F.H;
        const float cx = fmaf(clip[0] * iw, hw, hw);
        const float cy = fmaf((-clip[1]) * iw, hh, hh);
        float t[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            t[r] = fmaf(MAT(F.view, r, 0), psx[0], fmaf(MAT(F.view, r, 1), psx[1], fmaf(MAT(F.view, r, 2), psx[2], MAT(F.view, r, 3))));
        const float itz = __fdividef(1.0f, t[2]);
        const float rx = fminf(fmaxf(t[0] * itz, -F.lim_x), F.lim_x) * 1.0001f;
        const float ry = fminf(fmaxf(t[1] * itz, -F.lim_y), F.lim_y) * 1.0001f;
        const float j0 = F.focal * itz;
        const float nj2 = (j0 * j0) * fmaf(ry, ry, fmaf(rx, rx, 1.0f));
        const float l1 = fmaf((nj2 * F.wnorm2) * 1.002f, lm, 0.3006f);       // >= lambda_1 of the exact chain
        float s1; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(2.0f * l1));
        s1 = fminf(s1 * 1.0005f, 4096.0f);
        float hb = (fminf(rr, 2.8284272f) * s1) * 1.001f + 0.06f;
        if (!(hb <= 1.0e9f)) hb = 1.0e9f;                                   // inf / NaN bound: the whole screen
        const float hbx = fmaf(2.0e-6f, fabsf(cx), hb), hby = fmaf(2.0e-6f, fabsf(cy), hb);
        const float x0f = fmaxf(ceilf((cx - hbx) - 0.5f), 0.0f);
        const float x1f = fminf(floorf((cx + hbx) - 0.5f), F.W - 1.0f);
        const float y0f = fmaxf(ceilf((cy - hby) - 0.5f), 0.0f);
        const float y1f = fminf(floorf((cy + hby) - 0.5f), F.H - 1.0f);
        vis = (x0f <= x1f) && (y0f <= y1f);
        if (vis) {
            static_assert(TILE == 16, "tile shift");
            o.tx0 = (int)((unsigned)x0f >> 4); o.tx1 = (int)((unsigned)x1f >> 4);       // 0 <= x <= 65535
            o.ty0 = (int)((unsigned)y0f >> 4); o.ty1 = (int)((unsigned)y1f >> 4);
            const float dx = p[0] - F.cam[0], dy = p[1] - F.cam[1], dz = p[2] - F.cam[2];
            o.key = __float_as_uint(dx * dx + dy * dy + dz * dz);
        }
    }
    o.vis = vis;
    return o;
}

// Debug view of the bound (gsb_debug_fetch GSB_DBG_TRECTS / GSB_DBG_KEYS_UNSORTED with the bounded K1): the bound of EVERY
// packed splat, written at its original index.  The frame path never runs this; splat_select_kernel evaluates the same
// bound_one() only for the splats of the cells a depth chunk selects.
__global__ void __launch_bounds__(K1_THREADS)
project_bound_kernel(const __grid_constant__ FrameConsts F, const float4* __restrict__ geomA_p, const float* __restrict__ lam_p,
                     const uint32_t* __restrict__ orig, int64_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ trects)
{
    const int64_t p = (int64_t)blockIdx.x * K1_THREADS + threadIdx.x;
    if (p >= n) return;
    const Bound b = bound_one(F, __ldg(geomA_p + p), __ldg(lam_p + p));
    const uint32_t i = __ldg(orig + p);
    keys[i] = b.key;
    trects[i] = b.vis ? pack_trect(b.tx0, b.tx1, b.ty0, b.ty1) : TRECT_CULLED;
}

// ------------------------------------------------------------------------------------ spatial cells (r02)
// At pack time the active set is ordered along a Morton curve and cut into cells of CELL consecutive splats; only the
// 20-byte K1 stream (position + discard radius, eigenvalue bound) and the original index are stored in that order, the
// 128-byte lines K2 gathers stay where they were.  A cell keeps the bounding box of its positions, its largest discard
// radius and its largest eigenvalue bound.  Per frame ONE thread per cell projects the box: a conservative tile
// rectangle (the corner projections grown by the largest extent any member can have: bound_one's formula with the cell's
// worst-case operands) and a conservative depth-key interval.  A depth chunk then evaluates bound_one() only for the
// members of cells whose interval meets the chunk's and whose rectangle still holds a live tile — the two 8-byte-per-
// splat streams per chunk and the 28-byte-per-splat K1 stream of r01 shrink to the cells that can matter (and, for a
// row-partitioned frame, to the cells that touch the rank's rows: the N-proportional work no longer repeats on every GPU).
// Selection order no longer follows the splat index, so ties of the depth sort are put in index order afterwards
// (tie_fix_kernel): the spec's order (ascending key, ties ascending index) is unchanged.

__device__ __forceinline__ uint32_t spread10(uint32_t v)      // 10 bits -> every third bit
{
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

struct MortonBox { float lo[3]; float inv[3]; };             // quantisation: q = (p - lo) * inv, 10 bits per axis

__global__ void __launch_bounds__(256)
morton_kernel(const float4* __restrict__ geomA, int64_t n, const MortonBox mb, uint32_t* __restrict__ mkeys, uint32_t* __restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 g = __ldg(geomA + i);
    const float q[3] = { (g.x - mb.lo[0]) * mb.inv[0], (g.y - mb.lo[1]) * mb.inv[1], (g.z - mb.lo[2]) * mb.inv[2] };
    uint32_t key = 0x3FFFFFFFu;                               // non-finite positions go last (their cells are flagged)
    if (q[0] == q[0] && q[1] == q[1] && q[2] == q[2]) {
        const uint32_t a = (uint32_t)fminf(fmaxf(q[0], 0.0f), 1023.0f), b = (uint32_t)fminf(fmaxf(q[1], 0.0f), 1023.0f),
                       c = (uint32_t)fminf(fmaxf(q[2], 0.0f), 1023.0f);
        key = spread10(a) | (spread10(b) << 1) | (spread10(c) << 2);
    }
    mkeys[i] = key; idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
gather_geom_kernel(const uint32_t* __restrict__ orig, const float4* __restrict__ geomA, int64_t n, float4* __restrict__ geomA_p)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) geomA_p[p] = __ldg(geomA + __ldg(orig + p));
}

// one CTA per cell: lam in cell order, and the cell's box / largest discard radius / largest eigenvalue bound
__global__ void __launch_bounds__(CELL)
cell_build_kernel(const float4* __restrict__ geomA_p, const uint32_t* __restrict__ orig, const float* __restrict__ lam,
                  int64_t n, float* __restrict__ lam_p, CellBox* __restrict__ cells)
{
    __shared__ float red[CELL / 32][8];
    __shared__ uint32_t bad[CELL / 32];
    const int64_t p = (int64_t)blockIdx.x * CELL + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float lo[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, hi[3] = { -3.0e38f, -3.0e38f, -3.0e38f }, rr = -1.0f, lm = 0.0f;
    uint32_t nonfinite = 0u;
    if (p < n) {
        const float4 g = __ldg(geomA_p + p);
        lm = __ldg(lam + __ldg(orig + p));
        lam_p[p] = lm;
        const float q[3] = { g.x, g.y, g.z };
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (fabsf(q[k]) <= 3.0e38f) { lo[k] = q[k]; hi[k] = q[k]; } else nonfinite = 1u;
        }
        rr = g.w;
        if (!(lm >= 0.0f && lm <= 3.0e38f)) { lm = __int_as_float(0x7f800000); }      // inf / NaN eigenvalue bound: whole screen
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        rr = fmaxf(rr, __shfl_xor_sync(0xffffffffu, rr, o));
        lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));
        nonfinite |= __shfl_xor_sync(0xffffffffu, nonfinite, o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { red[warp][k] = lo[k]; red[warp][3 + k] = hi[k]; }
        red[warp][6] = rr; red[warp][7] = lm; bad[warp] = nonfinite;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        CellBox c;
#pragma unroll
        for (int k = 0; k < 3; ++k) { c.lo[k] = red[0][k]; c.hi[k] = red[0][3 + k]; }
        c.rr_max = red[0][6]; c.lam_max = red[0][7]; uint32_t nf = bad[0];
        for (int w = 1; w < CELL / 32; ++w) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { c.lo[k] = fminf(c.lo[k], red[w][k]); c.hi[k] = fmaxf(c.hi[k], red[w][3 + k]); }
            c.rr_max = fmaxf(c.rr_max, red[w][6]); c.lam_max = fmaxf(c.lam_max, red[w][7]); nf |= bad[w];
        }
        if (c.lo[0] > c.hi[0] || c.lo[1] > c.hi[1] || c.lo[2] > c.hi[2]) nf = 1u;       // no finite position at all
        if (nf) c.lam_max = __int_as_float(0x7f800000);                                   // flagged: whole screen, every key
        cells[blockIdx.x] = c;
    }
}

// per frame, one thread per cell: conservative tile rectangle + depth-key interval (CellView), the chunk plan's histogram
// (cell population at the bucket of the interval's middle) and the visible-splat estimate
__global__ void __launch_bounds__(256)
cell_project_kernel(const __grid_constant__ FrameConsts F, const CellBox* __restrict__ cells, const int64_t ncells, const int64_t n,
                    const DepthBuckets db, uint32_t* __restrict__ bucket_hist, uint4* __restrict__ views,
                    unsigned long long* __restrict__ n_visible)
{
    __shared__ uint32_t sh_hist[DEPTH_BUCKETS];
    if (bucket_hist) {
        for (int b = threadIdx.x; b < DEPTH_BUCKETS; b += 256) sh_hist[b] = 0u;
        __syncthreads();
    }
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t pop = 0;
    if (c < ncells) {
        const CellBox cb = cells[c];
        uint4 view = make_uint4(1u, 0u, 1u, 1u);                 // klo > khi: never selected; rectangle empty
        const int64_t left = n - c * CELL;
        const uint32_t members = (uint32_t)(left < CELL ? left : CELL);
        if (cb.rr_max >= 0.0f) {
            const bool flagged = !(cb.lam_max <= 3.0e38f) && !(cb.lo[0] <= cb.hi[0]);      // non-finite positions inside
            bool whole = !(cb.lam_max <= 3.0e38f);                                         // unbounded extent: the whole screen
            float lo[3], hi[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {                            // room for the (P - origin) + origin round trip and fp32 slop
                const float m = 4.0e-7f * (fabsf(cb.lo[k]) + fabsf(cb.hi[k]) + fabsf(F.origin[k])) + 1.0e-30f;
                lo[k] = cb.lo[k] - m; hi[k] = cb.hi[k] + m;
            }
            // depth-key interval: squared distance from the camera to the box, nearest and farthest point
            float dmin2 = 0.0f, dmax2 = 0.0f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float a = lo[k] - F.cam[k], b = F.cam[k] - hi[k];
                const float nr = fmaxf(fmaxf(a, b), 0.0f), fr = fmaxf(fabsf(F.cam[k] - lo[k]), fabsf(F.cam[k] - hi[k]));
                dmin2 = fmaf(nr, nr, dmin2); dmax2 = fmaf(fr, fr, dmax2);
            }
            uint32_t klo = __float_as_uint(fmaxf(dmin2 * (1.0f - 2.0e-5f), 0.0f));
            uint32_t khi = (dmax2 <= 3.0e38f) ? __float_as_uint(dmax2 * (1.0f + 2.0e-5f)) : 0x7F800000u;
            if (!(dmin2 == dmin2) || !(dmax2 == dmax2) || flagged || !(cb.lo[0] <= cb.hi[0])) { klo = 0u; khi = KEY_CULLED - 1u; whole = true; }
            // the eight corners: centre range on screen, nearest view-space depth
            float cxmin = 3.0e38f, cxmax = -3.0e38f, cymin = 3.0e38f, cymax = -3.0e38f, tzmin = 3.0e38f, tzmax = -3.0e38f;
            const float hw = 0.5f * F.W, hh = 0.5f * F.H;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float px = (q & 1) ? hi[0] : lo[0], py = (q & 2) ? hi[1] : lo[1], pz = (q & 4) ? hi[2] : lo[2];
                float vc[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    vc[r] = fmaf(MAT(F.obj_view, r, 0), px, fmaf(MAT(F.obj_view, r, 1), py, fmaf(MAT(F.obj_view, r, 2), pz, MAT(F.obj_view, r, 3))));
                const float fy = -vc[1];
                const float c0 = fmaf(MAT(F.proj, 0, 0), vc[0], fmaf(MAT(F.proj, 0, 1), fy, fmaf(MAT(F.proj, 0, 2), vc[2], MAT(F.proj, 0, 3))));
                const float c1 = fmaf(MAT(F.proj, 1, 0), vc[0], fmaf(MAT(F.proj, 1, 1), fy, fmaf(MAT(F.proj, 1, 2), vc[2], MAT(F.proj, 1, 3))));
                const float c3 = fmaf(MAT(F.proj, 3, 0), vc[0], fmaf(MAT(F.proj, 3, 1), fy, fmaf(MAT(F.proj, 3, 2), vc[2], MAT(F.proj, 3, 3))));
                const float tz = fmaf(MAT(F.view, 2, 0), px, fmaf(MAT(F.view, 2, 1), py, fmaf(MAT(F.view, 2, 2), pz, MAT(F.view, 2, 3))));
                if (!(c3 > 1.0e-20f)) whole = true;                 // the box reaches the camera plane (or NaN)
                const float iw = 1.0f / c3;
                const float cx = fmaf(c0 * iw, hw, hw), cy = fmaf((-c1) * iw, hh, hh);
                cxmin = fminf(cxmin, cx); cxmax = fmaxf(cxmax, cx); cymin = fminf(cymin, cy); cymax = fmaxf(cymax, cy);
                tzmin = fminf(tzmin, tz); tzmax = fmaxf(tzmax, tz);
                if (!(cx == cx) || !(cy == cy) || !(tz == tz)) whole = true;
            }
            if (!(tzmin > 0.0f) && !(tzmax < 0.0f)) whole = true;   // view-space depth changes sign inside the box: |J| unbounded
            int tx0 = 0, tx1 = F.tiles_x - 1, ty0 = 0, ty1 = F.tiles_y - 1;
            bool empty = false;
            if (!whole) {
                const float az = fminf(fabsf(tzmin), fabsf(tzmax));
                const float itz = (1.0f / az) * 1.00001f;
                const float j0 = F.focal * itz;
                const float nj2 = (j0 * j0) * (1.0f + 1.0003f * (F.lim_x * F.lim_x + F.lim_y * F.lim_y));
                const float l1 = (nj2 * F.wnorm2) * 1.003f * cb.lam_max + 0.3006f;
                const float s1 = fminf(sqrtf(2.0f * l1) * 1.001f, 4096.0f);
                float hb = (fminf(cb.rr_max, 2.8284272f) * s1) * 1.002f + 0.1f;
                if (!(hb <= 1.0e9f)) hb = 1.0e9f;
                const float hbx = hb + 4.0e-6f * fmaxf(fabsf(cxmin), fabsf(cxmax)) + 0.1f;
                const float hby = hb + 4.0e-6f * fmaxf(fabsf(cymin), fabsf(cymax)) + 0.1f;
                const float x0f = fmaxf(ceilf((cxmin - hbx) - 0.5f), 0.0f), x1f = fminf(floorf((cxmax + hbx) - 0.5f), F.W - 1.0f);
                const float y0f = fmaxf(ceilf((cymin - hby) - 0.5f), 0.0f), y1f = fminf(floorf((cymax + hby) - 0.5f), F.H - 1.0f);
                if (!(x0f <= x1f) || !(y0f <= y1f)) empty = true;
                else {
                    tx0 = (int)((unsigned)x0f >> 4); tx1 = (int)((unsigned)x1f >> 4);
                    ty0 = (int)((unsigned)y0f >> 4); ty1 = (int)((unsigned)y1f >> 4);
                }
            }
            if (!empty) {
                view = make_uint4(klo, khi, (uint32_t)tx0 | ((uint32_t)tx1 << 16), (uint32_t)ty0 | ((uint32_t)ty1 << 16));
                pop = members;
                if (bucket_hist) atomicAdd(&sh_hist[depth_bucket((klo >> 1) + (khi >> 1), db)], members);
            }
        }
        views[c] = view;
    }
    pop = __reduce_add_sync(0xffffffffu, pop);
    if ((threadIdx.x & 31) == 0 && pop) atomicAdd(n_visible, (unsigned long long)pop);
    if (bucket_hist) {
        __syncthreads();
        for (int b = threadIdx.x; b < DEPTH_BUCKETS; b += 256) {
            const uint32_t v = sh_hist[b];
            if (v) atomicAdd(bucket_hist + b, v);
        }
    }
}

// per depth chunk, one thread per cell: cells whose key interval meets the chunk's and whose rectangle holds a live tile
__global__ void __launch_bounds__(256)
cell_select_kernel(const uint4* __restrict__ views, const int64_t ncells, const ChunkPlan* __restrict__ plan, const int chunk,
                   const int tiles_x, const uint32_t* __restrict__ sat, uint32_t* __restrict__ sel_cells, uint32_t* __restrict__ n_sel)
{
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const uint32_t key_lo = plan ? __ldg(&plan->key_lo[chunk]) : 0u;
    const uint32_t key_hi = plan ? __ldg(&plan->key_lo[chunk + 1]) : KEY_CULLED;
    bool sel = false;
    if (c < ncells) {
        const uint4 v = __ldg(views + c);
        if (v.x <= v.y && v.y >= key_lo && v.x < key_hi)
            sel = live_tiles((int)(v.z & 0xffffu), (int)(v.z >> 16), (int)(v.w & 0xffffu), (int)(v.w >> 16), tiles_x, sat) != 0u;
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if (m) {
        const int lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(n_sel, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (sel) sel_cells[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)c;
    }
}

// per depth chunk: bound_one() for the members of the selected cells; the splats whose key lies in the chunk's interval
// and whose bound holds a live tile are appended (key, ORIGINAL index) to the live list, their number and instance bound
// are accumulated, and the digit histograms of the depth sort that follows are built on the way (no pass over the keys).
// Persistent CTAs; one CTA handles one cell at a time (thread t = member t).
struct SelectSort { int shift[SORT_MAX_PASSES]; int bits[SORT_MAX_PASSES]; int passes; uint32_t key_min, key_span; };
__global__ void __launch_bounds__(CELL)
splat_select_kernel(const __grid_constant__ FrameConsts F, const float4* __restrict__ geomA_p, const float* __restrict__ lam_p,
                    const uint32_t* __restrict__ orig, const int64_t n, const uint32_t* __restrict__ sel_cells,
                    const uint32_t* __restrict__ n_sel, const ChunkPlan* __restrict__ plan, const int chunk,
                    const uint32_t* __restrict__ sat, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                    unsigned long long* __restrict__ l_total, unsigned long long* __restrict__ d_total,
                    const SelectSort ss, uint32_t* __restrict__ sort_hist)
{
    __shared__ uint32_t sh_hist[SORT_MAX_PASSES][SORT_RADIX];
    __shared__ uint32_t s_wl[CELL / 32], s_wd[CELL / 32];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_MAX_PASSES * SORT_RADIX; i += CELL) (&sh_hist[0][0])[i] = 0u;
    const uint32_t key_lo = plan ? __ldg(&plan->key_lo[chunk]) : 0u;
    const uint32_t key_hi = plan ? __ldg(&plan->key_lo[chunk + 1]) : KEY_CULLED;
    const uint32_t nsel = *n_sel;
    __syncthreads();
    // software pipeline: the next cell's 20 bytes per member are in flight while this cell is evaluated (ncu r02: the
    // kernel stalled mostly on these loads, one dependent chain cell id -> position -> bound per iteration)
    int64_t p_next = -1;
    float4 ga_next = make_float4(0.f, 0.f, 0.f, -1.f); float lm_next = 0.f;
    if (blockIdx.x < nsel) {
        p_next = (int64_t)__ldg(sel_cells + blockIdx.x) * CELL + threadIdx.x;
        if (p_next < n) { ga_next = __ldg(geomA_p + p_next); lm_next = __ldg(lam_p + p_next); }
    }
    for (uint32_t s = blockIdx.x; s < nsel; s += gridDim.x) {
        const int64_t p = p_next;
        const float4 ga = ga_next; const float lm = lm_next;
        if (s + gridDim.x < nsel) {
            p_next = (int64_t)__ldg(sel_cells + s + gridDim.x) * CELL + threadIdx.x;
            if (p_next < n) { ga_next = __ldg(geomA_p + p_next); lm_next = __ldg(lam_p + p_next); }
        }
        uint32_t cnt = 0, key = KEY_CULLED;
        if (p < n) {
            const Bound b = bound_one(F, ga, lm);
            key = b.key;
            if (b.vis && key >= key_lo && key < key_hi) cnt = live_tiles(b.tx0, b.tx1, b.ty0, b.ty1, F.tiles_x, sat);
        }
        const unsigned m = __ballot_sync(0xffffffffu, cnt != 0u);
        const uint32_t dsum = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) { s_wl[warp] = (uint32_t)__popc(m); s_wd[warp] = dsum; }
        __syncthreads();
        uint32_t wbase = 0, l = 0, d = 0;
#pragma unroll
        for (int w = 0; w < CELL / 32; ++w) { wbase += (w < warp) ? s_wl[w] : 0u; l += s_wl[w]; d += s_wd[w]; }
        if (threadIdx.x == 0 && l) {
            s_base = atomicAdd(l_total, (unsigned long long)l);
            atomicAdd(d_total, (unsigned long long)d);
        }
        __syncthreads();
        if (cnt != 0u) {
            const size_t o = (size_t)s_base + wbase + (uint32_t)__popc(m & ((1u << lane) - 1u));
            keys_out[o] = key; vals_out[o] = __ldg(orig + p);
            const uint32_t q = sort_squeeze(key, ss.key_min, ss.key_span);
#pragma unroll
            for (int ps = 0; ps < SORT_MAX_PASSES; ++ps)
                if (ps < ss.passes) atomicAdd(&sh_hist[ps][(q >> ss.shift[ps]) & ((1u << ss.bits[ps]) - 1u)], 1u);
        }
        // (s_wl / s_wd / s_base are rewritten only after the next iteration's first barrier... no: s_wl is written before
        // it; this barrier keeps a fast warp of the next iteration from overwriting what a slow warp still reads)
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ss.passes * SORT_RADIX; i += CELL) {
        const uint32_t v = (&sh_hist[0][0])[i];
        if (v) atomicAdd(sort_hist + i, v);
    }
}

// Upper bound of the largest eigenvalue of the symmetric 3x3 matrix in sg (the fp32 values the exact kernels use), in
// double: the trigonometric closed form lambda_max = q + 2 p cos(acos(det((A - qI)/p) / 2) / 3), falling back to
// q + 2p (cos <= 1) when the argument degenerates; rounded up by 1e-6 relative before the conversion to fp32.
// Not finite -> +inf (the bounded K1 then keeps the splat on the whole screen; the exact projection decides).
__device__ __forceinline__ float lambda_max_upper(const Sigma& sg)
{
    const double a = sg.s[0], d = sg.s[1], e = sg.s[2], b = sg.s[3], f = sg.s[4], c = sg.s[5];
    const double q = (a + b + c) / 3.0;
    const double p1 = d * d + e * e + f * f;
    const double p2 = (a - q) * (a - q) + (b - q) * (b - q) + (c - q) * (c - q) + 2.0 * p1;
    double lmax = q;
    if (p2 > 0.0) {
        const double p = sqrt(p2 / 6.0);
        const double ip = 1.0 / p;
        const double b00 = (a - q) * ip, b11 = (b - q) * ip, b22 = (c - q) * ip, b01 = d * ip, b02 = e * ip, b12 = f * ip;
        const double det = b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02);
        double r = 0.5 * det;
        double cs = 1.0;                                   // cos <= 1: always a valid bound
        if (r > -1.0 && r < 1.0) cs = cos(acos(r) / 3.0);
        else if (r <= -1.0) cs = 0.5;                      // phi = pi/3
        lmax = q + 2.0 * p * cs;
    }
    lmax = fabs(lmax) * (1.0 + 1e-6) + 1e-300;
    const float out = (float)lmax;
    return (out >= 0.0f && out <= 3.0e38f) ? __fmul_ru(out, 1.0000002f) : __int_as_float(0x7f800000);
}

// ---- sigma planes: world-space covariance per splat, rebuilt when the object matrix changes
struct ObjMat { float m[16]; };
__global__ void __launch_bounds__(256)
sigma_kernel(const __grid_constant__ ObjMat O, const uint4* __restrict__ geomB, int64_t n,
             float4* __restrict__ sigA, float2* __restrict__ sigB, float* __restrict__ lam)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Sigma sg = sigma_of(O.m, __ldg(geomB + i));
    if (sigA) {                                             // the covariance planes are only read by the exact K1
        sigA[i] = make_float4(sg.s[0], sg.s[1], sg.s[2], sg.s[3]);
        sigB[i] = make_float2(sg.s[4], sg.s[5]);
    }
    lam[i] = lambda_max_upper(sg);
}

// ---- chunk plan: one CTA of DEPTH_BUCKETS threads
__global__ void __launch_bounds__(DEPTH_BUCKETS)
choose_chunks_kernel(const uint32_t* __restrict__ hist, const int nchunks, const int shift, const DepthBuckets db,
                     ChunkPlan* __restrict__ plan)
{
    __shared__ uint32_t wtot[DEPTH_BUCKETS / 32];
    __shared__ uint32_t csize[MAX_CHUNKS + 1];
    __shared__ uint8_t  s_lut[DEPTH_BUCKETS];
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    if (b <= MAX_CHUNKS) csize[b] = 0u;
    const uint32_t mine = (b < DEPTH_BUCKETS - 1) ? hist[b] : 0u;       // the last bucket holds the culled splats
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, V = 0;
#pragma unroll
    for (int w = 0; w < DEPTH_BUCKETS / 32; ++w) { woff += (w < warp) ? wtot[w] : 0u; V += wtot[w]; }
    const uint32_t excl = woff + inc - mine;                            // visible splats in nearer buckets
    int chunk = 0;
    for (int c = 1; c < nchunks; ++c) {
        unsigned long long tq = ((unsigned long long)V * ((1ull << c) - 1ull)) >> shift;
        const uint32_t target = (uint32_t)(tq < (unsigned long long)V ? tq : (unsigned long long)V);
        if (excl >= target && target > 0u) ++chunk;                     // thresholds ascend with c: monotone in b
    }
    if (b == DEPTH_BUCKETS - 1) chunk = nchunks;
    plan->lut[b] = (uint8_t)chunk; s_lut[b] = (uint8_t)chunk;
    const uint32_t cnt = (b == DEPTH_BUCKETS - 1) ? hist[b] : mine;
    if (cnt) atomicAdd(&csize[chunk], cnt);
    __syncthreads();
    if (b <= MAX_CHUNKS) plan->size[b] = (b <= nchunks) ? csize[b] : 0u;
    // key boundaries: key -> bucket -> chunk is monotone, so chunk c starts at the smallest key whose chunk is >= c
    // (bisection over the non-negative fp32 bit patterns up to +inf); chunks beyond the last non-empty one start at
    // KEY_CULLED, i.e. are empty
    if (b <= nchunks) {
        uint32_t klo = 0u;
        if (b == nchunks) klo = KEY_CULLED;
        else if (b > 0) {
            uint32_t lo = 0u, hi = 0x7F800000u;
            if ((int)s_lut[depth_bucket(hi, db)] < b) klo = KEY_CULLED;
            else {
                while (lo < hi) {
                    const uint32_t mid = lo + ((hi - lo) >> 1);
                    if ((int)s_lut[depth_bucket(mid, db)] >= b) hi = mid; else lo = mid + 1u;
                }
                klo = lo;
            }
        }
        plan->key_lo[b] = klo;
    }
    if (b == nchunks + 1 && b <= MAX_CHUNKS + 1) plan->key_lo[b] = KEY_CULLED;
}

// ---- K2: one thread per live splat, in depth order.  The splat's whole 128-byte line (geometry + colour + SH) is
// one DRAM burst pair.  Lines are gathered cooperatively: 8 lanes fetch the 8 16-byte chunks of one line, so every load
// instruction of a warp covers 4 complete lines and all loads of the warp's 32 lines are in flight together (one
// latency, not three dependent ones); the lines are staged in shared memory (144-byte pitch: conflict-free both ways)
// and each thread then reads its own.  The projection is redone with K1's code (same bits); the record goes to recs[j],
// and the splat's tile rectangle and live-tile count (the binning inputs) to tile_rects[j] / counts[j].
constexpr int K2_THREADS = 256;
constexpr int K2_PITCH   = ROW_U4 + 1;             // uint4 per staged line

// the K2 work of one thread: gather (cooperatively, per warp) the lines of 32 consecutive live ranks, redo the projection
// with K1's code, write the record of rank j, return its tile rectangle and live-tile count (0: the exact projection culls it)
template <int ORDER>
__device__ __forceinline__ uint32_t record_one(const FrameConsts& F, const uint4* __restrict__ rows, uint4* sw, const int lane,
                                               const uint32_t my, const bool valid, const int64_t j, const uint32_t* __restrict__ sat,
                                               Record* __restrict__ recs, float* __restrict__ zdepth,
                                               const uint32_t* __restrict__ owned_rows, uint2& trect)
{
    constexpr int NCH = 2 + (ORDER == 0 ? 1 : (ORDER == 1 ? 2 : (ORDER == 2 ? 4 : 6)));     // chunks of the line this order reads
    const int c = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3);
        const uint32_t idx = __shfl_sync(0xffffffffu, my, r);
        if (c < NCH) sw[r * K2_PITCH + c] = __ldg(rows + (size_t)idx * ROW_U4 + c);
    }
    __syncwarp();
    trect = make_uint2(1u, 1u);
    if (!valid) return 0u;
    const uint4* row = sw + lane * K2_PITCH;
    const uint4 ra = row[0];
    const uint4 gb = row[1];
    const float p[3] = { __uint_as_float(ra.x), __uint_as_float(ra.y), __uint_as_float(ra.z) };
    const float alpha = __uint_as_float(ra.w);
    Geom g;
    float4* out = reinterpret_cast<float4*>(recs + j);
    const float pmax = pmax_of_alpha(alpha);
    if (!project_geom(F, p, (pmax >= 0.0f) ? sqrtf(pmax) : -1.0f, sigma_of(F.object, gb), g, owned_rows)) {
        // the bound kept it, the exact projection culls it (degenerate axes, empty exact rectangle, no owned row): no instances
        out[0] = make_float4(-1.0e9f, -1.0e9f, 0.f, 0.f); out[1] = make_float4(0.f, 0.f, 0.f, -1.0f);
        out[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (zdepth) zdepth[j] = 0.0f;
        return 0u;
    }
    if (zdepth) zdepth[j] = ((g.clipz / g.clipw) * F.depth_hr) + F.depth_hm;     // window depth of the whole quad
    const int tx0 = g.x0 / TILE, tx1 = g.x1 / TILE, ty0 = g.y0 / TILE, ty1 = g.y1 / TILE;
    trect = make_uint2((uint32_t)tx0 | ((uint32_t)tx1 << 16), (uint32_t)ty0 | ((uint32_t)ty1 << 16));
    const uint32_t count = live_tiles(tx0, tx1, ty0, ty1, F.tiles_x, sat);
    float rgb[3];
    shade_colour<ORDER>(F, row + 2, g.psx, rgb);
    const uint32_t hpack = (uint32_t)__half_as_ushort(__float2half_ru(g.hx)) |
                           ((uint32_t)__half_as_ushort(__float2half_ru(g.hy)) << 16);
    out[0] = make_float4(g.cx, g.cy, g.m00, g.m01);
    out[1] = make_float4(g.m10, g.m11, alpha, pmax);
    out[2] = make_float4(rgb[0], rgb[1], rgb[2], __uint_as_float(hpack));
    return count;
}

template <int ORDER>
__global__ void __launch_bounds__(K2_THREADS)
records_kernel(const __grid_constant__ FrameConsts F, const uint4* __restrict__ rows,
               const uint32_t* __restrict__ live_splats, const int64_t n_live, const uint32_t* __restrict__ sat,
               Record* __restrict__ recs, uint2* __restrict__ tile_rects, uint32_t* __restrict__ counts,
               float* __restrict__ zdepth, const uint32_t* __restrict__ owned_rows)
{
    __shared__ uint4 srow[K2_THREADS / 32][32 * K2_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t j = (int64_t)blockIdx.x * K2_THREADS + threadIdx.x;
    const bool valid = j < n_live;
    const uint32_t my = valid ? __ldg(live_splats + j) : 0u;
    uint2 tr;
    const uint32_t cnt = record_one<ORDER>(F, rows, srow[warp], lane, my, valid, j, sat, recs, zdepth, owned_rows, tr);
    if (valid) { tile_rects[j] = tr; counts[j] = cnt; }
}

}  // namespace

void launch_pack(const float* pos, const uint16_t* cd_h, const float* alpha, const uint16_t* scale_h,
                 const uint16_t* orient_h, const uint16_t* shx, const uint16_t* shy, const uint16_t* shz,
                 int64_t count, int64_t dst_offset, float4* geomA, uint4* geomB, uint4* rows, int has_sh, cudaStream_t s)
{
    if (count <= 0) return;
    unsigned grid = (unsigned)((count + 255) / 256);
    pack_kernel<<<grid, 256, 0, s>>>(pos, cd_h, alpha, scale_h, orient_h, shx, shy, shz, count, dst_offset,
                                     geomA, geomB, rows, has_sh);
}

void launch_sigma(const float object[16], const uint4* geomB, int64_t n, float4* sigA, float2* sigB, float* lam, cudaStream_t s)
{
    if (n <= 0) return;
    ObjMat O; for (int k = 0; k < 16; ++k) O.m[k] = object[k];
    sigma_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(O, geomB, n, sigA, sigB, lam);
}

void launch_project_bound_debug(const FrameConsts& fc, const float4* geomA_p, const float* lam_p, const uint32_t* orig, int64_t n,
                                uint32_t* keys, uint32_t* trects, cudaStream_t s)
{
    if (n <= 0) return;
    project_bound_kernel<<<(unsigned)((n + K1_THREADS - 1) / K1_THREADS), K1_THREADS, 0, s>>>(fc, geomA_p, lam_p, orig, n, keys, trects);
}

void launch_morton(const float4* geomA, int64_t n, const float bbox[6], uint32_t* mkeys, uint32_t* idx, cudaStream_t s)
{
    if (n <= 0) return;
    MortonBox mb;
    for (int k = 0; k < 3; ++k) {
        const float ext = bbox[3 + k] - bbox[k];
        mb.lo[k] = bbox[k];
        mb.inv[k] = (ext > 0.0f && ext <= 3.0e38f) ? 1024.0f / ext : 0.0f;
    }
    morton_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(geomA, n, mb, mkeys, idx);
}

void launch_gather_geom(const uint32_t* orig, const float4* geomA, int64_t n, float4* geomA_p, cudaStream_t s)
{
    if (n <= 0) return;
    gather_geom_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(orig, geomA, n, geomA_p);
}

void launch_cell_build(const float4* geomA_p, const uint32_t* orig, const float* lam, int64_t n, float* lam_p, CellBox* cells,
                       cudaStream_t s)
{
    if (n <= 0) return;
    cell_build_kernel<<<(unsigned)((n + CELL - 1) / CELL), CELL, 0, s>>>(geomA_p, orig, lam, n, lam_p, cells);
}

void launch_cell_project(const FrameConsts& fc, const CellBox* cells, int64_t n, DepthBuckets db, uint32_t* bucket_hist,
                         uint4* views, unsigned long long* n_visible, cudaStream_t s)
{
    if (n <= 0) return;
    const int64_t ncells = (n + CELL - 1) / CELL;
    cell_project_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, s>>>(fc, cells, ncells, n, db, bucket_hist, views, n_visible);
}

void launch_cell_select(const uint4* views, int64_t n, const ChunkPlan* plan, int chunk, const FrameConsts& fc,
                        const uint32_t* sat, uint32_t* sel_cells, uint32_t* n_sel, cudaStream_t s)
{
    if (n <= 0) return;
    const int64_t ncells = (n + CELL - 1) / CELL;
    cell_select_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, s>>>(views, ncells, plan, chunk, fc.tiles_x, sat, sel_cells, n_sel);
}

void launch_splat_select(const FrameConsts& fc, const float4* geomA_p, const float* lam_p, const uint32_t* orig, int64_t n,
                         const uint32_t* sel_cells, const uint32_t* n_sel, const ChunkPlan* plan, int chunk,
                         const uint32_t* sat, uint32_t* keys_out, uint32_t* vals_out,
                         unsigned long long* l_total, unsigned long long* d_total,
                         const SortPlan& sp, uint32_t key_min, uint32_t key_span, uint32_t* sort_hist, cudaStream_t s)
{
    if (n <= 0) return;
    static int per_sm = 0;
    if (!per_sm) {
        cudaError_t rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, splat_select_kernel, CELL, 0);
        if (rc != cudaSuccess || per_sm < 1) per_sm = 4;
    }
    const int64_t ncells = (n + CELL - 1) / CELL, cap = (int64_t)NUM_SMS * per_sm;
    SelectSort ss{};
    ss.passes = sp.passes; ss.key_min = key_min; ss.key_span = key_span;
    for (int p = 0; p < sp.passes; ++p) { ss.shift[p] = sp.shift[p]; ss.bits[p] = sp.bits[p]; }
    splat_select_kernel<<<(unsigned)(ncells < cap ? ncells : cap), CELL, 0, s>>>(fc, geomA_p, lam_p, orig, n, sel_cells, n_sel, plan, chunk,
                                                                               sat, keys_out, vals_out, l_total, d_total, ss, sort_hist);
}

void launch_project(const FrameConsts& fc, const PackedSplats& ps, int64_t n,
                    uint32_t* keys, uint2* rects, int rects_all, uint32_t* trects,
                    unsigned long long* n_visible, DepthBuckets db, uint32_t* bucket_hist, const uint32_t* owned_rows,
                    cudaStream_t s)
{
    if (n <= 0) return;
    // resident CTAs per SM: 4 (54 registers), 5 (48) or 6 (40, a few spill bytes); GSB_K1_OCC overrides for experiments
    static int per_sm_of[3] = { 0, 0, 0 };
    const char* e = getenv("GSB_K1_OCC");
    int occ = e ? atoi(e) : 5;
    if (occ < 4 || occ > 6) occ = 5;
    int& per_sm = per_sm_of[occ - 4];
    if (!per_sm) {
        cudaError_t rc = occ == 4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, project_kernel<4>, K1_THREADS, 0)
                       : occ == 5 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, project_kernel<5>, K1_THREADS, 0)
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, project_kernel<6>, K1_THREADS, 0);
        if (rc != cudaSuccess || per_sm < 1) per_sm = occ;
    }
    const int64_t want = (n + K1_THREADS - 1) / K1_THREADS, cap = (int64_t)NUM_SMS * per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
#define GSB_K1(M) project_kernel<M><<<grid, K1_THREADS, 0, s>>>(fc, ps.geomA, ps.sigA, ps.sigB, n, keys, rects, rects_all, trects, \
                                                                n_visible, db, bucket_hist, owned_rows)
    if (occ == 4) GSB_K1(4); else if (occ == 5) GSB_K1(5); else GSB_K1(6);
#undef GSB_K1
}

void launch_choose_chunks(const uint32_t* bucket_hist, int nchunks, int shift, DepthBuckets db, ChunkPlan* plan, cudaStream_t s)
{
    choose_chunks_kernel<<<1, DEPTH_BUCKETS, 0, s>>>(bucket_hist, nchunks, shift, db, plan);
}

void launch_records(const FrameConsts& fc, const PackedSplats& ps, const uint32_t* live_splats, int64_t n_live,
                    const uint32_t* sat, Record* recs, uint2* tile_rects, uint32_t* counts, float* zdepth,
                    const uint32_t* owned_rows, cudaStream_t s)
{
    if (n_live <= 0) return;
    const unsigned grid = (unsigned)((n_live + K2_THREADS - 1) / K2_THREADS);
    switch (fc.sh_order) {
    case 0:  records_kernel<0><<<grid, K2_THREADS, 0, s>>>(fc, ps.rows, live_splats, n_live, sat, recs, tile_rects, counts, zdepth, owned_rows); break;
    case 1:  records_kernel<1><<<grid, K2_THREADS, 0, s>>>(fc, ps.rows, live_splats, n_live, sat, recs, tile_rects, counts, zdepth, owned_rows); break;
    case 2:  records_kernel<2><<<grid, K2_THREADS, 0, s>>>(fc, ps.rows, live_splats, n_live, sat, recs, tile_rects, counts, zdepth, owned_rows); break;
    default: records_kernel<3><<<grid, K2_THREADS, 0, s>>>(fc, ps.rows, live_splats, n_live, sat, recs, tile_rects, counts, zdepth, owned_rows); break;
    }
}

}  // namespace gsb

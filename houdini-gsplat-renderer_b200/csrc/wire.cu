// wire.cu — wireframe / selection overlay of a GSplat primitive, sm_100a.
//
// Replaces the reference's wire vertex shader (/root/reference/gsplat_plugin/shaders/GSplatShaderSource.h:22-90), which the
// reference runs 8 times per splat over a VBO holding 8 copies of every attribute (src/GR_GSplat.C:376-421, drawn as
// RE_PRIM_LINES at GR_GSplat.C:474-483): the outline of the splat's +-2 sigma quad, vertices 0..7 = the four edges
// (-2,-2)-(2,-2), (2,-2)-(2,2), (2,2)-(-2,2), (-2,2)-(-2,-2), colour = Cd.  Here ONE thread per splat evaluates the
// covariance chain once and writes the 8 clip-space positions (16 B each) and the colour; the shim draws them as lines
// from an interop buffer with a pass-through shader, or — without OpenGL — wire_overlay_kernel rasterises the outlines
// into an RGBA32F frame (nearest splat wins per pixel, opaque; a fixed DDA rule the oracle restates).
//
// The wire shader differs from the main one on purpose (reproduced): the position is the raw P (no origin round trip), the
// covariance ignores the object matrix (SRC.h:69), nothing is culled (a splat behind the camera still emits vertices; GL
// clips the lines).  fp32 expression order = the spec shared with oracle/gsplat_oracle.cpp orc_wire_vertices (bit-exact).
#include "common.cuh"

namespace gsb {

namespace {

#define MAT(M, r, c) ((M)[(c) * 4 + (r)])

__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

__global__ void __launch_bounds__(256)
wire_vertices_kernel(const __grid_constant__ FrameConsts F, const float* __restrict__ pos, const uint16_t* __restrict__ cd,
                     const uint16_t* __restrict__ scale, const uint16_t* __restrict__ orient, const int64_t n,
                     float4* __restrict__ verts, float* __restrict__ colors)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p[3] = { pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] };
    // centre (SRC.h:64-66)
    float vc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        vc[r] = ((MAT(F.obj_view, r, 0) * p[0] + MAT(F.obj_view, r, 1) * p[1]) + MAT(F.obj_view, r, 2) * p[2]) + MAT(F.obj_view, r, 3);
    const float fy = -vc[1];
    float clip[4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
        clip[r] = ((MAT(F.proj, r, 0) * vc[0] + MAT(F.proj, r, 1) * fy) + MAT(F.proj, r, 2) * vc[2]) + MAT(F.proj, r, 3);
    // covariance: M = S R^T, Sigma = M^T M (LIB.h:10-35), no object matrix (SRC.h:74)
    const float sx = h2f(scale[3 * i]), sy = h2f(scale[3 * i + 1]), sz = h2f(scale[3 * i + 2]);
    const float qx = h2f(orient[4 * i]), qy = h2f(orient[4 * i + 1]), qz = h2f(orient[4 * i + 2]), qr = h2f(orient[4 * i + 3]);
    float Rt[3][3];
    Rt[0][0] = 1.0f - 2.0f * (qy * qy + qz * qz); Rt[0][1] = 2.0f * (qx * qy + qr * qz); Rt[0][2] = 2.0f * (qx * qz - qr * qy);
    Rt[1][0] = 2.0f * (qx * qy - qr * qz); Rt[1][1] = 1.0f - 2.0f * (qx * qx + qz * qz); Rt[1][2] = 2.0f * (qy * qz + qr * qx);
    Rt[2][0] = 2.0f * (qx * qz + qr * qy); Rt[2][1] = 2.0f * (qy * qz - qr * qx); Rt[2][2] = 1.0f - 2.0f * (qx * qx + qy * qy);
    const float sc[3] = { sx, sy, sz };
    float Mm[3][3], S[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Mm[a][b] = sc[a] * Rt[a][b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) S[a][b] = (Mm[0][a] * Mm[0][b] + Mm[1][a] * Mm[1][b]) + Mm[2][a] * Mm[2][b];
    // EWA projection (LIB.h:38-76) with the view matrix and the raw position
    float t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        t[r] = ((MAT(F.view, r, 0) * p[0] + MAT(F.view, r, 1) * p[1]) + MAT(F.view, r, 2) * p[2]) + MAT(F.view, r, 3);
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
        B0[k] = (A0[0] * S[0][k] + A0[1] * S[1][k]) + A0[2] * S[2][k];
        B1[k] = (A1[0] * S[0][k] + A1[1] * S[1][k]) + A1[2] * S[2][k];
    }
    const float c00 = (B0[0] * A0[0] + B0[1] * A0[1]) + B0[2] * A0[2];
    const float c01 = (B0[0] * A1[0] + B0[1] * A1[1]) + B0[2] * A1[2];
    const float c11 = (B1[0] * A1[0] + B1[1] * A1[1]) + B1[2] * A1[2];
    const float a = c00 + 0.3f, b = c01, c = c11 + 0.3f;
    // eigen axes (LIB.h:79-93): v1 = s1 (ex, -ey), v2 = s2 (-ey, -ex) before the final y flip
    const float mid = 0.5f * (a + c);
    const float hd = (a - c) / 2.0f;
    const float radius = sqrtf(hd * hd + b * b);
    const float l1 = mid + radius;
    const float l2 = fmaxf(mid - radius, 0.1f);
    const float dvx = b, dvy = l1 - a;
    const float len = sqrtf(dvx * dvx + dvy * dvy);
    const float ex = dvx / len, ey = dvy / len;                 // NaN for a degenerate splat, as in the GLSL
    const float s1 = fminf(sqrtf(2.0f * l1), 4096.0f);
    const float s2 = fminf(sqrtf(2.0f * l2), 4096.0f);
    const float v1x = s1 * ex, v1y = s1 * (-ey);
    const float v2x = s2 * (-ey), v2y = s2 * (-ex);
    // the 8 vertices (SRC.h:35-56, 82-88): quad corner per vertex index
    const float qcx[8] = { -2.f, 2.f, 2.f, 2.f, 2.f, -2.f, -2.f, -2.f };
    const float qcy[8] = { -2.f, -2.f, -2.f, 2.f, 2.f, 2.f, 2.f, -2.f };
    const float r0 = h2f(cd[3 * i]), g0 = h2f(cd[3 * i + 1]), b0 = h2f(cd[3 * i + 2]);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float dx = ((qcx[v] * v1x + qcy[v] * v2x) * 2.0f) / F.W;
        const float dy = ((qcx[v] * v1y + qcy[v] * v2y) * 2.0f) / F.H;
        const float ox = clip[0] + dx * clip[3];
        const float oy = -(clip[1] + dy * clip[3]);
        verts[8 * i + v] = make_float4(ox, oy, clip[2], clip[3]);
        if (colors) { colors[(8 * i + v) * 3] = r0; colors[(8 * i + v) * 3 + 1] = g0; colors[(8 * i + v) * 3 + 2] = b0; }
    }
}

// ---- overlay: the four edges of every splat into a per-pixel (depth, splat) key; the nearest splat wins.
// Segment rule (the oracle restates it): window coordinates x = (ndc.x + 1) W / 2, y = (ndc.y + 1) H / 2 in fp32;
// n = ceil(max(|dx|, |dy|)) steps (at least 1, at most 65536); sample i = 0..n at a + (b - a) * (i / n); pixel = floor.
// Splats whose centre fails 0 < w, -w <= z <= w (GL's clip volume) are skipped, like the main path.
__global__ void __launch_bounds__(256)
wire_overlay_kernel(const float4* __restrict__ verts, const int64_t n, const int width, const int height,
                    unsigned long long* __restrict__ owner)
{
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // 4 segments per splat
    if (seg >= n * 4) return;
    const int64_t i = seg >> 2;
    const float4 va = verts[8 * i + 2 * (seg & 3)], vb = verts[8 * i + 2 * (seg & 3) + 1];
    const float w = va.w, z = va.z;
    if (!(w > 0.0f) || !(z >= -w && z <= w)) return;
    const float W = (float)width, H = (float)height;
    const float ax = ((va.x / w + 1.0f) * 0.5f) * W, ay = ((va.y / w + 1.0f) * 0.5f) * H;
    const float bx = ((vb.x / w + 1.0f) * 0.5f) * W, by = ((vb.y / w + 1.0f) * 0.5f) * H;
    if (!(ax == ax) || !(ay == ay) || !(bx == bx) || !(by == by)) return;
    const float ddx = bx - ax, ddy = by - ay;
    const float m = fmaxf(fabsf(ddx), fabsf(ddy));
    if (!(m <= 1.0e9f)) return;
    int steps = (int)ceilf(m);
    steps = steps < 1 ? 1 : (steps > 65536 ? 65536 : steps);
    const float depth = z / w;                                  // one depth per splat (SRC.h:85: z, w of the centre)
    const unsigned long long key = ((unsigned long long)__float_as_uint(depth * 0.5f + 0.5f) << 32) | (unsigned long long)(uint32_t)i;
    const float fn = (float)steps;
    for (int k = 0; k <= steps; ++k) {
        const float tpar = (float)k / fn;
        const float x = ax + ddx * tpar, y = ay + ddy * tpar;
        const float fx = floorf(x), fyy = floorf(y);
        if (fx >= 0.0f && fyy >= 0.0f && fx < W && fyy < H)
            atomicMin(owner + (size_t)fyy * width + (size_t)fx, key);
    }
}

__global__ void __launch_bounds__(256)
wire_resolve_kernel(const unsigned long long* __restrict__ owner, const uint16_t* __restrict__ cd, const int64_t px,
                    float4* __restrict__ rgba)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= px) return;
    const unsigned long long o = owner[k];
    if (o == 0xFFFFFFFFFFFFFFFFull) return;
    const uint32_t i = (uint32_t)o;
    rgba[k] = make_float4(h2f(cd[3 * (size_t)i]), h2f(cd[3 * (size_t)i + 1]), h2f(cd[3 * (size_t)i + 2]), 1.0f);   // GSplatWireFragmentShader: (color, 1)
}

}  // namespace

void launch_wire_vertices(const FrameConsts& fc, const float* pos, const uint16_t* cd, const uint16_t* scale,
                          const uint16_t* orient, int64_t n, float4* verts, float* colors, cudaStream_t s)
{
    if (n <= 0) return;
    wire_vertices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(fc, pos, cd, scale, orient, n, verts, colors);
}

void launch_wire_overlay(const float4* verts, const uint16_t* cd, int64_t n, int width, int height,
                         unsigned long long* owner, float4* rgba, cudaStream_t s)
{
    if (n <= 0 || width <= 0 || height <= 0) return;
    const int64_t px = (int64_t)width * height;
    cudaMemsetAsync(owner, 0xFF, (size_t)px * 8, s);
    wire_overlay_kernel<<<(unsigned)((n * 4 + 255) / 256), 256, 0, s>>>(verts, n, width, height, owner);
    wire_resolve_kernel<<<(unsigned)((px + 255) / 256), 256, 0, s>>>(owner, cd, px, rgba);
}

}  // namespace gsb

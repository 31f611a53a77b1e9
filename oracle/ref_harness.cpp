/* ref_harness.cpp — oracle/_ref/libgsplat_ref.so: the REFERENCE'S OWN SHADER TEXT, compiled for the host.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under houdini-gsplat-renderer_b200/ may include, link, import or execute
 * this.  Built by oracle/build_ref.py only where /root/reference exists; the GPU box uses the prebuilt .so.
 *
 * What is the reference's and what is restated here:
 *   REFERENCE, unmodified text, read from /root/reference at build time and compiled as C++ against oracle/glsl_cxx.h:
 *       GSplatCoreLib, GSplatSphericalHarmonicsLib          shaders/GSplatShaderCoreLib.h:8-95, 101-181
 *       _GSplatMainVertexShader, _GSplatMainFragmentShader   shaders/GSplatShaderSource.h:117-288, 291-314
 *       _GSplatWireVertexShader                              shaders/GSplatShaderSource.h:22-90
 *     (build_ref.py applies four token-level rewrites GLSL needs to be C++ — `out T x` parameters become `T& x`, the
 *      `in/out parms {..} name;` interface blocks and `out vec4 color_out;` become thread-local structs/variables, the wire
 *      shader's `in vecN attr;` become thread-local variables — and records how often each fired in _ref/manifest.json.)
 *   RESTATED here, because the reference leaves it to Houdini's RE_* wrapper and the OpenGL driver:
 *       texture packing as virtual textures     src/GSplatRenderer.C:448-505 (same texel values, no 2^k x 2^k copies)
 *       texture dimensions                      src/GSplatRenderer.C:106-139, 155-163
 *       uniforms                                src/GSplatRenderer.C:625-645, shaders/GSplatShaderSource.h:153-159
 *       draw: 6 vertices x N instances, 2 triangles per instance   src/GSplatRenderer.C:29, 647
 *       clipping (z against +-w), an ideal rasteriser (pixel centres, exact edge functions in double, one owner per shared
 *       edge), perspective-correct varying interpolation, depth test on / depth writes off (R.C:608-610),
 *       blend ADD with src = 1 - dst.a, dst = 1 on an RGBA32F target (R.C:613-621)
 *       argsortByDistance                       src/GSplatRenderer.C:176-216 (ties: ascending index)
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#include <parallel/algorithm>
#include <omp.h>

#include "glsl_cxx.h"

namespace refglsl {
using namespace glsl;

/* GL built-ins, per shader invocation */
thread_local int  gl_InstanceID = 0, gl_VertexID = 0;
thread_local vec4 gl_Position;
thread_local bool glsl_discarded = false;

#define float flt
#define uniform
#define discard do { glsl_discarded = true; return; } while (0)

namespace main_vs {
#define main vs_main
#include "core_lib.inc"
#include "sh_lib.inc"
#include "main_vs.inc"
#undef main
}  // namespace main_vs

namespace main_fs {
#define main fs_main
#include "main_fs.inc"
#undef main
}  // namespace main_fs

namespace wire_vs {
#define main wire_main
#include "core_lib.inc"
#include "wire_vs.inc"
#undef main
}  // namespace wire_vs

#undef float
#undef uniform
#undef discard
}  // namespace refglsl

using glsl::flt;
using glsl::vec2;
using glsl::vec3;
using glsl::vec4;
using glsl::mat4;

namespace {

inline float h2f(uint16_t h)
{
    uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
    if (e == 0) {
        if (m == 0) u = s << 31;
        else { int k = 0; while (!(m & 0x400u)) { m <<= 1; ++k; } m &= 0x3ffu; u = (s << 31) | ((uint32_t)(113 - k) << 23) | (m << 13); }
    } else if (e == 31) u = (s << 31) | 0x7f800000u | (m << 13);
    else u = (s << 31) | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}

/* closestSqrtPowerOf2, R.C:155-163 */
int closest_sqrt_pow2(int n)
{
    if (n <= 1) return 2;
    float sqrtVal = (float)std::sqrt((double)n);
    unsigned int power = (unsigned int)std::ceil(std::log2(sqrtVal));
    return (int)std::pow(2, power);
}

/* the arrays registerUpdate receives (R.h:34-47) + what generateRenderGeometry derives from them */
struct Scene {
    int64_t n = 0;
    const float* pos = nullptr; const uint16_t* cd = nullptr; const float* alpha = nullptr;
    const uint16_t* scale = nullptr; const uint16_t* orient = nullptr;
    const uint16_t* shx = nullptr; const uint16_t* shy = nullptr; const uint16_t* shz = nullptr;
    float origin[3] = { 0, 0, 0 };
    const int32_t* zorder = nullptr;
} g_scene;

/* virtual textures: texel values exactly as R.C:456-502 writes them */
vec4 fetch_tex0(const void*, int linear)
{
    const Scene& s = g_scene;
    const int64_t i = linear >> 2; const int k = linear & 3;
    if (i >= s.n) return vec4(0, 0, 0, 0);
    switch (k) {
    case 0: return vec4(flt(s.pos[3 * i] - s.origin[0]), flt(s.pos[3 * i + 1] - s.origin[1]), flt(s.pos[3 * i + 2] - s.origin[2]), flt(0.0f));
    case 1: return vec4(flt(h2f(s.cd[3 * i])), flt(h2f(s.cd[3 * i + 1])), flt(h2f(s.cd[3 * i + 2])), flt(s.alpha[i]));
    case 2: return vec4(flt(h2f(s.scale[3 * i])), flt(h2f(s.scale[3 * i + 1])), flt(h2f(s.scale[3 * i + 2])), flt(0.0f));
    default: return vec4(flt(h2f(s.orient[4 * i])), flt(h2f(s.orient[4 * i + 1])), flt(h2f(s.orient[4 * i + 2])), flt(h2f(s.orient[4 * i + 3])));
    }
}
vec4 fetch_sh(const void* which, int linear)        /* RGB16F: alpha reads as 1 */
{
    const Scene& s = g_scene;
    const int64_t i = linear >> 3; int j = linear & 7;
    if (which) { if (j == 7) return vec4(0, 0, 0, 1); j += 8; }          /* degree-3 texture: coefficients 8..14, texel 7 = padding */
    if (i >= s.n || !s.shx) return vec4(0, 0, 0, 1);
    return vec4(flt(h2f(s.shx[16 * i + j])), flt(h2f(s.shy[16 * i + j])), flt(h2f(s.shz[16 * i + j])), flt(1.0f));
}
int fetch_z(const void*, int linear)
{
    return (linear < g_scene.n && g_scene.zorder) ? g_scene.zorder[linear] : 0;
}

mat4 to_mat4(const float* m)          /* 16 floats, column-major */
{
    return mat4(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
}

int g_width = 0, g_height = 0;
float g_depth_range[2] = { 0.0f, 1.0f };

struct VSOut { float p[4]; float pos[4]; float color[3]; float opacity; };

inline void run_vs(int instance, int vertex, VSOut& o)
{
    namespace V = refglsl::main_vs;
    refglsl::gl_InstanceID = instance; refglsl::gl_VertexID = vertex;
    /* a varying the shader does not write keeps the previous invocation's value on real hardware (undefined); start clean */
    V::vsOut.pos = vec4(0, 0, 0, 0); V::vsOut.color = vec3(0, 0, 0); V::vsOut.opacity = flt(0.0f);
    V::vs_main();
    const vec4& g = refglsl::gl_Position;
    o.p[0] = g.d[0].v; o.p[1] = g.d[1].v; o.p[2] = g.d[2].v; o.p[3] = g.d[3].v;
    o.pos[0] = V::vsOut.pos.d[0].v; o.pos[1] = V::vsOut.pos.d[1].v; o.pos[2] = V::vsOut.pos.d[2].v; o.pos[3] = V::vsOut.pos.d[3].v;
    o.color[0] = V::vsOut.color.d[0].v; o.color[1] = V::vsOut.color.d[1].v; o.color[2] = V::vsOut.color.d[2].v;
    o.opacity = V::vsOut.opacity.v;
}

inline bool run_fs(const float pos[4], const float color[3], float opacity, float out[4])
{
    namespace F = refglsl::main_fs;
    F::fsIn.pos = vec4(flt(pos[0]), flt(pos[1]), flt(pos[2]), flt(pos[3]));
    F::fsIn.color = vec3(flt(color[0]), flt(color[1]), flt(color[2]));
    F::fsIn.opacity = flt(opacity);
    refglsl::glsl_discarded = false;
    F::fs_main();
    if (refglsl::glsl_discarded) return false;
    out[0] = F::color_out.d[0].v; out[1] = F::color_out.d[1].v; out[2] = F::color_out.d[2].v; out[3] = F::color_out.d[3].v;
    return true;
}

/* one owner per shared edge: a pixel centre exactly on an edge belongs to the triangle for which the edge (a -> b, interior
 * on the left) runs upwards, or leftwards when horizontal; the reversed edge of the neighbour then refuses it */
inline bool edge_owns(double dx, double dy) { return dy > 0.0 || (dy == 0.0 && dx < 0.0); }

}  // namespace

extern "C" {

typedef struct {
    float view[16], proj[16], object[16], inv_object[16], obj_view[16];
    float cam[3];
    float origin[3];
    int32_t width, height;
    int32_t sh_order;               /* GSplatShOrder as bound at R.C:623,628 (0 when the data has no SH) */
    float depth_range[2];
} ref_frame;

/* Binds one scene + frame: the uniforms of R.C:625-645 and the glH_* built-ins, the four samplers. */
void ref_bind(int64_t n, const float* pos, const uint16_t* cd, const float* alpha, const uint16_t* scale,
              const uint16_t* orient, const uint16_t* shx, const uint16_t* shy, const uint16_t* shz,
              const int32_t* zorder, const ref_frame* f)
{
    namespace V = refglsl::main_vs;
    namespace W = refglsl::wire_vs;
    Scene& s = g_scene;
    s.n = n; s.pos = pos; s.cd = cd; s.alpha = alpha; s.scale = scale; s.orient = orient;
    s.shx = shx; s.shy = shy; s.shz = shz; s.zorder = zorder;
    memcpy(s.origin, f->origin, 12);
    g_width = f->width; g_height = f->height;
    g_depth_range[0] = f->depth_range[0]; g_depth_range[1] = f->depth_range[1];
    const bool has_sh = shx && shy && shz;
    V::WorldSpaceCameraPos = vec3(flt(f->cam[0]), flt(f->cam[1]), flt(f->cam[2]));
    V::GSplatCount = (int)n;
    V::GSplatVertexCount = 6;
    V::GSplatZOrderTexDim = closest_sqrt_pow2((int)n);
    V::GSplatPosColorAlphaScaleOrientTexDim = closest_sqrt_pow2((int)n * 4);
    V::GSplatShDeg1And2TexDim = has_sh ? closest_sqrt_pow2((int)n * 8) : 0;
    V::GSplatShDeg3TexDim = V::GSplatShDeg1And2TexDim;
    V::GSplatShOrder = has_sh ? f->sh_order : 0;
    V::GSplatOrigin = vec3(flt(f->origin[0]), flt(f->origin[1]), flt(f->origin[2]));
    V::GSplatZOrderIntegerTexSampler.fetch = fetch_z; V::GSplatZOrderIntegerTexSampler.dim = V::GSplatZOrderTexDim;
    V::GSplatPosColorAlphaScaleOrientTexSampler.fetch = fetch_tex0;
    V::GSplatPosColorAlphaScaleOrientTexSampler.dim = V::GSplatPosColorAlphaScaleOrientTexDim;
    V::GSplatShDeg1And2TexSampler.fetch = fetch_sh; V::GSplatShDeg1And2TexSampler.ctx = nullptr;
    V::GSplatShDeg1And2TexSampler.dim = V::GSplatShDeg1And2TexDim;
    V::GSplatShDeg3TexSampler.fetch = fetch_sh; V::GSplatShDeg3TexSampler.ctx = (const void*)1;
    V::GSplatShDeg3TexSampler.dim = V::GSplatShDeg3TexDim;
    V::glH_ObjViewMatrix = to_mat4(f->obj_view); V::glH_ObjectMatrix = to_mat4(f->object);
    V::glH_InvObjectMatrix = to_mat4(f->inv_object); V::glH_ViewMatrix = to_mat4(f->view);
    V::glH_ProjectMatrix = to_mat4(f->proj);
    V::glH_DepthRange = vec2(flt(f->depth_range[0]), flt(f->depth_range[1]));
    V::glH_ScreenSize = vec2(flt((float)f->width), flt((float)f->height));
    W::glH_ObjViewMatrix = V::glH_ObjViewMatrix; W::glH_ViewMatrix = V::glH_ViewMatrix;
    W::glH_ProjectMatrix = V::glH_ProjectMatrix; W::glH_ScreenSize = V::glH_ScreenSize;
}

/* one vertex-shader invocation: out = gl_Position[4], vsOut.pos[4], vsOut.color[3], vsOut.opacity */
void ref_vs_main(int instance, int vertex, float out[12])
{
    VSOut o; run_vs(instance, vertex, o);
    memcpy(out, o.p, 16); memcpy(out + 4, o.pos, 16); memcpy(out + 8, o.color, 12); out[11] = o.opacity;
}

/* the vertex shader over instances [i0, i1) x 6 vertices, OpenMP: out[(i - i0) * 6 + v][12] */
void ref_vs_batch(int64_t i0, int64_t i1, float* out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i)
        for (int v = 0; v < 6; ++v) ref_vs_main((int)i, v, out + ((i - i0) * 6 + v) * 12);
}

/* one fragment-shader invocation; returns 1 if the fragment was discarded */
int ref_fs_main(const float pos[4], const float color[3], float opacity, float rgba[4])
{
    return run_fs(pos, color, opacity, rgba) ? 0 : 1;
}

/* wireframe vertex shader (SRC.h:22-90): one invocation; attributes as the VBO holds them (GR.C:393-405) */
void ref_wire_vs(int vertex_id, const float P[3], const float Cd[3], const float scale[3], const float orient[4],
                 float gl_pos[4], float color[3])
{
    namespace W = refglsl::wire_vs;
    W::P = vec3(flt(P[0]), flt(P[1]), flt(P[2])); W::Cd = vec3(flt(Cd[0]), flt(Cd[1]), flt(Cd[2]));
    W::scale = vec3(flt(scale[0]), flt(scale[1]), flt(scale[2]));
    W::orient = vec4(flt(orient[0]), flt(orient[1]), flt(orient[2]), flt(orient[3]));
    refglsl::gl_VertexID = vertex_id;
    W::wire_main();
    const vec4& g = refglsl::gl_Position;
    gl_pos[0] = g.d[0].v; gl_pos[1] = g.d[1].v; gl_pos[2] = g.d[2].v; gl_pos[3] = g.d[3].v;
    color[0] = W::vsOut.color.d[0].v; color[1] = W::vsOut.color.d[1].v; color[2] = W::vsOut.color.d[2].v;
}

/* argsortByDistance, R.C:188-208 (ties by ascending index; the reference's sort leaves them unspecified) */
void ref_argsort_by_distance(const float* pos, int64_t n, const float cam[3], int32_t* order)
{
    std::vector<float> dist((size_t)n);
    for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* p = pos + 3 * i;
        float dx = p[0] - cam[0], dy = p[1] - cam[1], dz = p[2] - cam[2];
        dist[(size_t)i] = dx * dx + dy * dy + dz * dz;
    }
    const float* d = dist.data();
    __gnu_parallel::sort(order, order + n, [d](int32_t a, int32_t b) { return d[a] < d[b] || (d[a] == d[b] && a < b); });
}

/* The draw of R.C:647 on the bound scene: instances 0..n-1 in z-order-texture order, 6 vertices each = triangles (0,1,2)
 * and (3,4,5); fragments blended into rgba (H x W x 4, row 0 = bottom scanline, caller-initialised = the incoming frame
 * buffer).  scene_depth (H x W window depths) + depth_func 1 = LESS, 2 = LEQUAL: the depth test of R.C:608-610; NULL / 0:
 * no test.  stats[0] += fragments shaded, stats[1] += fragments that survived the discard.  No tiles, no termination. */
void ref_draw_ex(float* rgba, const float* scene_depth, int depth_func, int64_t* stats, uint8_t* unsafe, float tol);
void ref_draw(float* rgba, const float* scene_depth, int depth_func, int64_t* stats)
{
    ref_draw_ex(rgba, scene_depth, depth_func, stats, nullptr, 0.0f);
}

/* unsafe (H x W bytes, may be NULL): set to 1 where some splat's support edge (|q.x| = 2, |q.y| = 2) or discard threshold
 * (alpha = 1/255) passes within tol_q of the pixel centre (tol_q = tol + the splat's own fp32 position uncertainty in q
 * units, see below) — there a last-bit difference in the evaluation order of the
 * reference's formulas decides coverage, so two valid evaluations of the same GLSL may differ by up to alpha there
 * (SURVEY.md §7 "discontinuous support") — provided the fragment could still change the pixel by more than 1e-4 (its
 * opacity times the transmittance left at that point).  Diagnostic only; the frame itself does not depend on it. */
void ref_draw_ex(float* rgba, const float* scene_depth, int depth_func, int64_t* stats, uint8_t* unsafe, float tol)
{
    const int W = g_width, H = g_height;
    const int64_t n = g_scene.n;
    const int BATCH = 1 << 15, BAND = 8;
    const int nbands = (H + BAND - 1) / BAND;
    std::vector<VSOut> vs((size_t)BATCH * 6);
    std::vector<int> ylo((size_t)BATCH), yhi((size_t)BATCH);
    int64_t shaded = 0, kept = 0;
    const float hr = (g_depth_range[1] - g_depth_range[0]) * 0.5f, hm = (g_depth_range[1] + g_depth_range[0]) * 0.5f;
    for (int64_t b0 = 0; b0 < n; b0 += BATCH) {
        const int64_t b1 = std::min<int64_t>(n, b0 + BATCH);
        /* vertex stage: 6 invocations per instance, exactly as drawInstanced issues them */
#pragma omp parallel for schedule(static)
        for (int64_t i = b0; i < b1; ++i) {
            VSOut* o = &vs[(size_t)(i - b0) * 6];
            for (int v = 0; v < 6; ++v) run_vs((int)i, v, o[v]);
            /* clip: every vertex carries the centre's z and w (SRC.h:278-282), so the quad is inside or outside as a whole */
            int lo = 1, hi = 0;
            const float w = o[0].p[3], z = o[0].p[2];
            if (w > 0.0f && z >= -w && z <= w) {
                double mn = 1e300, mx = -1e300;
                for (int v = 0; v < 6; ++v) {
                    const double yw = ((double)o[v].p[1] / (double)o[v].p[3] * 0.5 + 0.5) * H;
                    mn = std::min(mn, yw); mx = std::max(mx, yw);
                }
                if (mx >= 0.0 && mn <= (double)H && mn == mn && mx == mx) {
                    lo = (int)std::max(0.0, std::floor(mn - 0.5)); hi = (int)std::min((double)(H - 1), std::ceil(mx - 0.5));
                }
            }
            ylo[(size_t)(i - b0)] = lo; yhi[(size_t)(i - b0)] = hi;
        }
        /* raster + fragment + blend: bands of scanlines in parallel, instances in order inside a band */
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : shaded, kept)
        for (int band = 0; band < nbands; ++band) {
            const int by0 = band * BAND, by1 = std::min(H, by0 + BAND) - 1;
            for (int64_t i = b0; i < b1; ++i) {
                const size_t k = (size_t)(i - b0);
                if (ylo[k] > by1 || yhi[k] < by0 || ylo[k] > yhi[k]) continue;
                const VSOut* o = &vs[k * 6];
                const float ndc_z = o[0].p[2] / o[0].p[3];
                const float zw = ndc_z * hr + hm;
                if (unsafe) {
                    /* the varying `pos` is one affine function of the window position over the whole quad: take it from the
                     * first triangle and evaluate it on the quad's bounding box grown by one pixel */
                    double X[3], Y[3];
                    for (int c = 0; c < 3; ++c) {
                        X[c] = ((double)o[c].p[0] / (double)o[c].p[3] * 0.5 + 0.5) * W;
                        Y[c] = ((double)o[c].p[1] / (double)o[c].p[3] * 0.5 + 0.5) * H;
                    }
                    const double area = (X[1] - X[0]) * (Y[2] - Y[0]) - (Y[1] - Y[0]) * (X[2] - X[0]);
                    if (area != 0.0 && area == area) {
                        double xmn = 1e300, xmx = -1e300, ymn = 1e300, ymx = -1e300;
                        for (int v = 0; v < 6; ++v) {
                            const double xw = ((double)o[v].p[0] / (double)o[v].p[3] * 0.5 + 0.5) * W;
                            const double yw = ((double)o[v].p[1] / (double)o[v].p[3] * 0.5 + 0.5) * H;
                            xmn = std::min(xmn, xw); xmx = std::max(xmx, xw); ymn = std::min(ymn, yw); ymx = std::max(ymx, yw);
                        }
                        /* tolerance of THIS splat in q units.  Any fp32 evaluation of the reference's formulas carries the
                         * window position to one ulp of the screen extent (delta_px: the reference's NDC vertices are fp32, one
                         * ulp of 1.0 is W / 2 * 2^-23 px; a pixel-space evaluation subtracts fp32 centres with ulp 2^-23 * W),
                         * and q changes by |grad q|_1 per pixel — so a 2 px splat at x = 1900 has a q uncertainty of ~1e-4,
                         * five times the floor `tol`.  tol_q = tol + 2 * delta_px * |grad q|_1. */
                        const double gx0 = ((Y[2] - Y[0]) * (o[1].pos[0] - o[0].pos[0]) - (Y[1] - Y[0]) * (o[2].pos[0] - o[0].pos[0])) / area;
                        const double gy0 = (-(X[2] - X[0]) * (o[1].pos[0] - o[0].pos[0]) + (X[1] - X[0]) * (o[2].pos[0] - o[0].pos[0])) / area;
                        const double gx1 = ((Y[2] - Y[0]) * (o[1].pos[1] - o[0].pos[1]) - (Y[1] - Y[0]) * (o[2].pos[1] - o[0].pos[1])) / area;
                        const double gy1 = (-(X[2] - X[0]) * (o[1].pos[1] - o[0].pos[1]) + (X[1] - X[0]) * (o[2].pos[1] - o[0].pos[1])) / area;
                        const double grad = std::max(std::fabs(gx0) + std::fabs(gy0), std::fabs(gx1) + std::fabs(gy1));
                        const double delta_px = std::ldexp((double)std::max(W, H), -23);
                        const double tq = (double)tol + 2.0 * delta_px * grad;
                        const int ux0 = (int)std::max(0.0, std::floor(xmn - 1.5)), ux1 = (int)std::min((double)(W - 1), std::ceil(xmx + 0.5));
                        const int uy0 = std::max(by0, (int)std::max(0.0, std::floor(ymn - 1.5)));
                        const int uy1 = std::min(by1, (int)std::min((double)(H - 1), std::ceil(ymx + 0.5)));
                        for (int py = uy0; py <= uy1; ++py)
                            for (int px = ux0; px <= ux1; ++px) {
                                const double cx = px + 0.5, cy = py + 0.5;
                                const double l1 = ((cx - X[0]) * (Y[2] - Y[0]) - (cy - Y[0]) * (X[2] - X[0])) / area;
                                const double l2 = ((X[1] - X[0]) * (cy - Y[0]) - (Y[1] - Y[0]) * (cx - X[0])) / area;
                                const double l0 = 1.0 - l1 - l2;
                                const double qx = l0 * o[0].pos[0] + l1 * o[1].pos[0] + l2 * o[2].pos[0];
                                const double qy = l0 * o[0].pos[1] + l1 * o[1].pos[1] + l2 * o[2].pos[1];
                                const double ax = std::fabs(qx), ay = std::fabs(qy);
                                bool u = (std::fabs(ax - 2.0) < tq && ay < 2.0 + tq) || (std::fabs(ay - 2.0) < tq && ax < 2.0 + tq);
                                if (!u && ax <= 2.0 + tq && ay <= 2.0 + tq) {
                                    const double a = std::exp(-(qx * qx + qy * qy)) * (double)o[0].opacity;
                                    u = std::fabs(a - 1.0 / 255.0) < tq * 0.05;      /* |da/dq| = 2 |q| a <= 0.02 on the ring */
                                }
                                /* a flip of this fragment moves the pixel by at most (1 - dst.a) * opacity: deep layers behind an
                                 * (almost) opaque pixel cannot matter and are not flagged */
                                if (u && (1.0f - rgba[((size_t)py * W + px) * 4 + 3]) * o[0].opacity > 1e-4f) unsafe[(size_t)py * W + px] = 1;
                            }
                    }
                }
                for (int t = 0; t < 2; ++t) {
                    const VSOut* v = o + 3 * t;
                    double X[3], Y[3], iw[3];
                    for (int c = 0; c < 3; ++c) {
                        X[c] = ((double)v[c].p[0] / (double)v[c].p[3] * 0.5 + 0.5) * W;
                        Y[c] = ((double)v[c].p[1] / (double)v[c].p[3] * 0.5 + 0.5) * H;
                        iw[c] = 1.0 / (double)v[c].p[3];
                    }
                    double area = (X[1] - X[0]) * (Y[2] - Y[0]) - (Y[1] - Y[0]) * (X[2] - X[0]);
                    if (!(area != 0.0) || area != area) continue;
                    int a = 0, b = 1, c = 2;
                    if (area < 0.0) { b = 2; c = 1; area = -area; }          /* no face culling: orient counter-clockwise */
                    const double xmn = std::min({ X[0], X[1], X[2] }), xmx = std::max({ X[0], X[1], X[2] });
                    const double ymn = std::min({ Y[0], Y[1], Y[2] }), ymx = std::max({ Y[0], Y[1], Y[2] });
                    const int px0 = (int)std::max(0.0, std::ceil(xmn - 0.5)), px1 = (int)std::min((double)(W - 1), std::floor(xmx - 0.5));
                    const int py0 = std::max(by0, (int)std::max(0.0, std::ceil(ymn - 0.5)));
                    const int py1 = std::min(by1, (int)std::min((double)(H - 1), std::floor(ymx - 0.5)));
                    const int idx[3] = { a, b, c };
                    const bool same_w = v[0].p[3] == v[1].p[3] && v[1].p[3] == v[2].p[3];
                    const bool flat = same_w && memcmp(v[0].color, v[1].color, 16) == 0 && memcmp(v[0].color, v[2].color, 16) == 0;
                    for (int py = py0; py <= py1; ++py)
                        for (int px = px0; px <= px1; ++px) {
                            const double cx = px + 0.5, cy = py + 0.5;
                            double lam[3]; bool in = true;
                            for (int e = 0; e < 3 && in; ++e) {
                                const int p = idx[(e + 1) % 3], q = idx[(e + 2) % 3];      /* edge opposite vertex idx[e] */
                                const double dx = X[q] - X[p], dy = Y[q] - Y[p];
                                const double E = dx * (cy - Y[p]) - dy * (cx - X[p]);
                                if (E < 0.0 || (E == 0.0 && !edge_owns(dx, dy))) in = false;
                                lam[e] = E / area;
                            }
                            if (!in) continue;
                            if (scene_depth && depth_func) {
                                const float sd = scene_depth[(size_t)py * W + px];
                                if (!(depth_func == 1 ? (zw < sd) : (zw <= sd))) continue;
                            }
                            /* perspective-correct interpolation of the varyings (lam_i / w_i, normalised); every vertex of a
                             * quad carries the same w, colour and opacity, for which this reduces to the plain weights */
                            float fpos[4], fcol[3], src[4], fop;
                            if (same_w) {
                                for (int d = 0; d < 4; ++d)
                                    fpos[d] = (float)(lam[0] * (double)v[idx[0]].pos[d] + lam[1] * (double)v[idx[1]].pos[d] + lam[2] * (double)v[idx[2]].pos[d]);
                            } else {
                                double wsum = 0.0, pp[4] = { 0, 0, 0, 0 };
                                for (int e = 0; e < 3; ++e) {
                                    const double l = lam[e] * iw[idx[e]];
                                    wsum += l;
                                    for (int d = 0; d < 4; ++d) pp[d] += l * (double)v[idx[e]].pos[d];
                                }
                                for (int d = 0; d < 4; ++d) fpos[d] = (float)(pp[d] / wsum);
                            }
                            if (flat) { fcol[0] = v[0].color[0]; fcol[1] = v[0].color[1]; fcol[2] = v[0].color[2]; fop = v[0].opacity; }
                            else {
                                double wsum = 0.0, cc[3] = { 0, 0, 0 }, op = 0.0;
                                for (int e = 0; e < 3; ++e) {
                                    const double l = lam[e] * iw[idx[e]];
                                    wsum += l;
                                    for (int d = 0; d < 3; ++d) cc[d] += l * (double)v[idx[e]].color[d];
                                    op += l * (double)v[idx[e]].opacity;
                                }
                                for (int d = 0; d < 3; ++d) fcol[d] = (float)(cc[d] / wsum);
                                fop = (float)(op / wsum);
                            }
                            ++shaded;
                            if (!run_fs(fpos, fcol, fop, src)) continue;
                            ++kept;
                            /* blend equation ADD, srcRGB = srcA = ONE_MINUS_DST_ALPHA, dstRGB = dstA = ONE (R.C:613-621), fp32 target */
                            float* dst = rgba + ((size_t)py * W + px) * 4;
                            const float f = 1.0f - dst[3];
                            dst[0] = src[0] * f + dst[0]; dst[1] = src[1] * f + dst[1];
                            dst[2] = src[2] * f + dst[2]; dst[3] = src[3] * f + dst[3];
                        }
                }
            }
        }
    }
    if (stats) { stats[0] += shaded; stats[1] += kept; }
}

int  ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int  ref_closest_sqrt_pow2(int n) { return closest_sqrt_pow2(n); }

}  /* extern "C" */

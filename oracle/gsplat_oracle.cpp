/* gsplat_oracle.cpp — CPU ORACLE for the splat -> framebuffer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under houdini-gsplat-renderer_b200/ may include, link,
 * import or execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / reported baseline.
 *
 * PARITY STATUS: pinned to the reference's own shader text.  The reference (rubendhz/houdini-gsplat-renderer @ 833a2124)
 * ships no tests, fixtures or golden vectors and its plugin cannot be built here (Houdini HDK +
 * OpenGL), but its arithmetic lives in GLSL strings that CAN be compiled: oracle/_ref/libgsplat_ref.so is that text,
 * unmodified, built as C++ (oracle/build_ref.py, glsl_cxx.h, ref_harness.cpp).  Pins:
 *   (1) tests/test_ref_pin.py — per-splat vertex-shader outputs, whole frames, depth-tested frames and fragment-shader
 *       known answers of the compiled reference text vs this oracle; tests/test_ref_golden.py — frames rendered by the
 *       reference text, committed as fixtures (tests/golden/ref_*.npz + the script that made them);
 *   (2) analytic known-answer tests (tests/test_oracle_kat.py);
 *   (3) an independent literal emulation of the GLSL text in numpy (oracle/glsl_literal.py);
 *   (4) golden vectors this oracle generated itself (tests/golden/g*.npz, regression pin only).
 *
 * What it restates (paths relative to /root/reference/gsplat_plugin):
 *   keys + order         src/GSplatRenderer.C:176-216   (argsortByDistance)
 *   camera position      src/GSplatRenderer.C:551-563
 *   shader position      src/GSplatRenderer.C:459-461 + shaders/GSplatShaderSource.h:201-202
 *   centre + cull        shaders/GSplatShaderSource.h:204-214, 277-282
 *   covariance chain     shaders/GSplatShaderCoreLib.h:10-93 via GSplatShaderSource.h:230-242
 *   SH colour            shaders/GSplatShaderCoreLib.h:103-179 via GSplatShaderSource.h:244-275
 *   quad / falloff       shaders/GSplatShaderSource.h:168-188, 304-312
 *   blend equation       src/GSplatRenderer.C:613-621
 *   fp16 inputs          src/GR_GSplat.C:314-318,337-366 (the boundary already receives halfs)
 *
 * The arithmetic below is the SPEC: every fp32 expression has a fixed
 * evaluation order, no FMA contraction (build with -ffp-contract=off) except where fmaf()
 * is written explicitly, IEEE division and sqrt.  The CUDA kernels follow the same order,
 * so keys, records, pixel rectangles, tile lists and per-pixel coverage are bit-exact; only
 * exp() differs (libm here, MUFU.EX2 on the GPU), which bounds RGBA error to ~1e-6.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <numeric>
#include <parallel/algorithm>
#include <omp.h>

extern "C" {

typedef struct {
    float view[16];        /* glH_ViewMatrix,    column-major (GL) */
    float proj[16];        /* glH_ProjectMatrix                    */
    float object[16];      /* glH_ObjectMatrix                     */
    float inv_object[16];  /* glH_InvObjectMatrix                  */
    float obj_view[16];    /* glH_ObjViewMatrix                    */
    float cam[3];          /* WorldSpaceCameraPos (R.C:551-563)    */
    float origin[3];       /* GSplatOrigin        (R.C:403-418)    */
    int32_t width, height; /* glH_ScreenSize                       */
    int32_t sh_order;      /* GSplatShOrder, already 0 if no SH data (R.C:623,628) */
    int32_t row_rank, row_world; /* tile-row ownership: row ty owned iff (ty / row_group) % world == rank */
    int32_t row_group;           /* tile rows per band, >= 1 */
    float eps_t;           /* transmittance early-out threshold; 0 = never stop (reference) */
} orc_frame;

typedef struct {
    float cx, cy, m00, m01;
    float m10, m11, alpha, pmax;
    float r, g, b;
    uint32_t hpack;        /* half(hx) | half(hy)<<16, rounded toward +inf */
} orc_record;              /* 48 bytes */

typedef struct { uint16_t x0, x1, y0, y1; } orc_rect;   /* inclusive pixel rectangle */

typedef struct {
    int64_t n_submitted, n_visible, n_instances, n_consumed;
    double  ms_sort_reference, ms_project, ms_sort, ms_bin, ms_blend, ms_total;
} orc_stats;

#define ORC_TILE 16
#define ORC_KEY_CULLED 0xFFFFFFFFu

static inline int orc_owns_row(const orc_frame* F, int ty)
{
    int g = F->row_group > 1 ? F->row_group : 1;
    return F->row_world <= 1 || ((ty / g) % F->row_world) == F->row_rank;
}

/* ---------------------------------------------------------------- half <-> float */
static inline float h2f(uint16_t h)
{
    uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
    if (e == 0) {
        if (m == 0) u = s << 31;
        else { int k = 0; while (!(m & 0x400u)) { m <<= 1; ++k; } m &= 0x3ffu;
               u = (s << 31) | ((uint32_t)(113 - k) << 23) | (m << 13); }
    } else if (e == 31) u = (s << 31) | 0x7f800000u | (m << 13);
    else u = (s << 31) | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}
/* smallest half >= f, for f >= 0 (NaN -> +inf) */
static inline uint16_t f2h_ru_pos(float f)
{
    if (!(f == f)) return 0x7c00;
    if (f > 65504.0f) return 0x7c00;
    uint32_t u; memcpy(&u, &f, 4);
    int e = (int)((u >> 23) & 0xffu); uint32_t m = u & 0x7fffffu;
    if (e < 103) return f > 0.0f ? 1 : 0;
    if (e < 113) { uint32_t full = m | 0x800000u; int sh = 113 - e + 13;
                   uint32_t mh = full >> sh; if (full & ((1u << sh) - 1u)) ++mh; return (uint16_t)mh; }
    uint32_t mh = m >> 13; if (m & 0x1fffu) ++mh;
    return (uint16_t)((((uint32_t)(e - 112)) << 10) + mh);
}
float    orc_half_to_float(uint16_t h) { return h2f(h); }
uint16_t orc_float_to_half_ru(float f) { return f2h_ru_pos(f); }

/* ---------------------------------------------------------------- deterministic ln()
 * ln(x) for x > 0 in double using only + - * / (IEEE, no contraction) so the same
 * sequence gives the same bits on the GPU.  x = m * 2^k, m in [sqrt(1/2), sqrt(2)). */
static inline double det_log(double x)
{
    uint64_t u; memcpy(&u, &x, 8);
    int k = (int)((u >> 52) & 0x7ffu) - 1023;
    u = (u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m; memcpy(&m, &u, 8);                      /* m in [1,2) */
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
double orc_det_log(double x) { return det_log(x); }

/* ---------------------------------------------------------------- camera (R.C:551-563)
 * cam = (0,0,0,1) * inverse(view) in double, rounded to f32: the translation of the
 * inverse, i.e. x = -adj-based solve.  Spec: solve R^T-free general 4x4 via cofactors of
 * the upper-left 3x3 and translation column; bottom row assumed (0,0,0,1) for affine views,
 * otherwise the full 4x4 adjugate is used. */
static void inv4_d(const double m[16], double inv[16])   /* column-major, full adjugate */
{
    double a[16];
    a[0] = m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
    a[4] = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
    a[8] = m[4]*m[9]*m[15] - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
    a[12] = -m[4]*m[9]*m[14] + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
    a[1] = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10];
    a[5] = m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10];
    a[9] = -m[0]*m[9]*m[15] + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9];
    a[13] = m[0]*m[9]*m[14] - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
    a[2] = m[1]*m[6]*m[15] - m[1]*m[7]*m[14] - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7] - m[13]*m[3]*m[6];
    a[6] = -m[0]*m[6]*m[15] + m[0]*m[7]*m[14] + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7] + m[12]*m[3]*m[6];
    a[10] = m[0]*m[5]*m[15] - m[0]*m[7]*m[13] - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7] - m[12]*m[3]*m[5];
    a[14] = -m[0]*m[5]*m[14] + m[0]*m[6]*m[13] + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6] + m[12]*m[2]*m[5];
    a[3] = -m[1]*m[6]*m[11] + m[1]*m[7]*m[10] + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7] + m[9]*m[3]*m[6];
    a[7] = m[0]*m[6]*m[11] - m[0]*m[7]*m[10] - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7] - m[8]*m[3]*m[6];
    a[11] = -m[0]*m[5]*m[11] + m[0]*m[7]*m[9] + m[4]*m[1]*m[11] - m[4]*m[3]*m[9] - m[8]*m[1]*m[7] + m[8]*m[3]*m[5];
    a[15] = m[0]*m[5]*m[10] - m[0]*m[6]*m[9] - m[4]*m[1]*m[10] + m[4]*m[2]*m[9] + m[8]*m[1]*m[6] - m[8]*m[2]*m[5];
    double det = m[0]*a[0] + m[1]*a[4] + m[2]*a[8] + m[3]*a[12];
    double r = 1.0 / det;
    for (int i = 0; i < 16; ++i) inv[i] = a[i] * r;
}
void orc_camera_from_view(const float view[16], float cam[3])
{
    double m[16], inv[16];
    for (int i = 0; i < 16; ++i) m[i] = (double)view[i];
    inv4_d(m, inv);
    /* column-vector convention: camera = inv * (0,0,0,1) = 4th column (w assumed 1) */
    cam[0] = (float)inv[12]; cam[1] = (float)inv[13]; cam[2] = (float)inv[14];
}

/* ---------------------------------------------------------------- keys (R.C:196-202) */
static inline uint32_t key_of(const float* p, const float* cam)
{
    float dx = p[0] - cam[0], dy = p[1] - cam[1], dz = p[2] - cam[2];
    float d2 = dx * dx + dy * dy + dz * dz;     /* ((dx*dx + dy*dy) + dz*dz), no FMA */
    uint32_t u; memcpy(&u, &d2, 4); return u;
}
void orc_keys(const float* pos, int64_t n, const float cam[3], uint32_t* keys)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) keys[i] = key_of(pos + 3 * i, cam);
}

/* stable argsort by uint32 key: ascending key, ties ascending index */
void orc_sort(const uint32_t* keys, int64_t n, int32_t* order)
{
    for (int64_t i = 0; i < n; ++i) order[i] = (int32_t)i;
    __gnu_parallel::stable_sort(order, order + n,
        [keys](int32_t a, int32_t b) { return keys[a] < keys[b]; });
}

/* the reference's own per-camera-move CPU stage, restated (R.C:188-208): iota, fp32 squared distances of EVERY splat,
 * parallel comparison sort through the index (tbb::parallel_sort there, __gnu_parallel::sort here).  The reference's
 * comparator is dist[a] < dist[b] alone, which leaves ties to the (unstable) sort; here ties are broken by ascending
 * index so the result is the spec's order and the frame built from it is the oracle's frame. */
void orc_sort_reference_style(const float* pos, int64_t n, const float cam[3], int32_t* order)
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

/* ---------------------------------------------------------------- per-splat projection */
/* M is column-major: element (row r, col c) = M[c*4 + r] */
#define MAT(M, r, c) ((M)[(c) * 4 + (r)])

static const float SH_C1   = 0.4886025f;
static const float SH_C2_0 = 1.0925484f, SH_C2_1 = -1.0925484f, SH_C2_2 = 0.3153916f,
                   SH_C2_3 = -1.0925484f, SH_C2_4 = 0.5462742f;
static const float SH_C3_0 = -0.5900436f, SH_C3_1 = 2.8906114f, SH_C3_2 = -0.4570458f,
                   SH_C3_3 = 0.3731763f, SH_C3_4 = -0.4570458f, SH_C3_5 = 1.4453057f,
                   SH_C3_6 = -0.5900436f;

/* returns 1 if the splat survives (record/rect valid), 0 if culled */
static int project_one(const orc_frame* F,
                       const float* p, const uint16_t* cd_h, float alpha,
                       const uint16_t* scale_h, const uint16_t* orient_h,
                       const uint16_t* shx, const uint16_t* shy, const uint16_t* shz,
                       orc_record* rec, orc_rect* rect)
{
    const float W = (float)F->width, H = (float)F->height;
    /* alpha can never pass the 1/255 discard (SRC.h:308-310) */
    if (!(alpha >= 1.0f / 255.0f)) return 0;

    /* shader-side position: texel holds P - origin, shader adds origin back (R.C:459-461, SRC.h:201-202) */
    float ps[3];
    for (int k = 0; k < 3; ++k) { float t = p[k] - F->origin[k]; ps[k] = t + F->origin[k]; }

    /* centre (SRC.h:204-214) */
    const float* OV = F->obj_view; const float* P = F->proj;
    float vc[3];
    for (int r = 0; r < 3; ++r)
        vc[r] = ((MAT(OV, r, 0) * ps[0] + MAT(OV, r, 1) * ps[1]) + MAT(OV, r, 2) * ps[2]) + MAT(OV, r, 3);
    float fy = -vc[1];
    float clip[4];
    for (int r = 0; r < 4; ++r)
        clip[r] = ((MAT(P, r, 0) * vc[0] + MAT(P, r, 1) * fy) + MAT(P, r, 2) * vc[2]) + MAT(P, r, 3);
    float cw = clip[3];
    if (!(cw > 0.0f)) return 0;                          /* clip.w <= 0 -> degenerate */
    if (!(clip[2] >= -cw && clip[2] <= cw)) return 0;    /* GL clip volume on the constant-z quad */
    float ndcx = clip[0] / cw;
    float ndcy = (-clip[1]) / cw;                        /* final y = -y (SRC.h:281) */
    float cx = ((ndcx + 1.0f) * 0.5f) * W;
    float cy = ((ndcy + 1.0f) * 0.5f) * H;

    /* covariance (LIB.h:10-35; SRC.h:230-236): M = S * R^T * O3^T, Sigma = M^T M */
    float sx = h2f(scale_h[0]), sy = h2f(scale_h[1]), sz = h2f(scale_h[2]);
    float qx = h2f(orient_h[0]), qy = h2f(orient_h[1]), qz = h2f(orient_h[2]), qr = h2f(orient_h[3]);
    /* R^T rows (GLSL mat3 columns at LIB.h:21-25 with rot=(r,x,y,z)) */
    float Rt[3][3];
    Rt[0][0] = 1.0f - 2.0f * (qy * qy + qz * qz); Rt[0][1] = 2.0f * (qx * qy + qr * qz); Rt[0][2] = 2.0f * (qx * qz - qr * qy);
    Rt[1][0] = 2.0f * (qx * qy - qr * qz); Rt[1][1] = 1.0f - 2.0f * (qx * qx + qz * qz); Rt[1][2] = 2.0f * (qy * qz + qr * qx);
    Rt[2][0] = 2.0f * (qx * qz + qr * qy); Rt[2][1] = 2.0f * (qy * qz - qr * qx); Rt[2][2] = 1.0f - 2.0f * (qx * qx + qy * qy);
    float Mm[3][3];
    const float sc[3] = { sx, sy, sz };
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Mm[i][j] = sc[i] * Rt[i][j];
    const float* O = F->object;
    float M2[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        M2[i][j] = (Mm[i][0] * MAT(O, j, 0) + Mm[i][1] * MAT(O, j, 1)) + Mm[i][2] * MAT(O, j, 2);
    float S[3][3];
    for (int i = 0; i < 3; ++i) for (int j = i; j < 3; ++j) {
        S[i][j] = (M2[0][i] * M2[0][j] + M2[1][i] * M2[1][j]) + M2[2][i] * M2[2][j];
        S[j][i] = S[i][j];
    }

    /* EWA projection (LIB.h:38-76) */
    const float* V = F->view;
    float t[3];
    for (int r = 0; r < 3; ++r)
        t[r] = ((MAT(V, r, 0) * ps[0] + MAT(V, r, 1) * ps[1]) + MAT(V, r, 2) * ps[2]) + MAT(V, r, 3);
    float aspect = MAT(P, 0, 0) / MAT(P, 1, 1);
    float tanFovX = 1.0f / MAT(P, 0, 0);
    float tanFovY = 1.0f / (MAT(P, 1, 1) * aspect);
    float limX = 1.3f * tanFovX, limY = 1.3f * tanFovY;
    float tz = t[2];
    float rx = t[0] / tz; rx = fminf(fmaxf(rx, -limX), limX);
    float ry = t[1] / tz; ry = fminf(fmaxf(ry, -limY), limY);
    float tx = rx * tz, ty = ry * tz;
    float focal = (W * MAT(P, 0, 0)) / 2.0f;
    float j0 = focal / tz;
    float tz2 = tz * tz;
    float j2x = -((focal * tx) / tz2);
    float j2y = -((focal * ty) / tz2);
    float A0[3], A1[3];
    for (int k = 0; k < 3; ++k) {
        A0[k] = j0 * MAT(V, 0, k) + j2x * MAT(V, 2, k);
        A1[k] = j0 * MAT(V, 1, k) + j2y * MAT(V, 2, k);
    }
    float B0[3], B1[3];
    for (int k = 0; k < 3; ++k) {
        B0[k] = (A0[0] * S[0][k] + A0[1] * S[1][k]) + A0[2] * S[2][k];
        B1[k] = (A1[0] * S[0][k] + A1[1] * S[1][k]) + A1[2] * S[2][k];
    }
    float c00 = (B0[0] * A0[0] + B0[1] * A0[1]) + B0[2] * A0[2];
    float c01 = (B0[0] * A1[0] + B0[1] * A1[1]) + B0[2] * A1[2];
    float c11 = (B1[0] * A1[0] + B1[1] * A1[1]) + B1[2] * A1[2];
    float a = c00 + 0.3f, b = c01, c = c11 + 0.3f;

    /* eigen-decomposition (LIB.h:79-93) */
    float mid = 0.5f * (a + c);
    float hd = (a - c) / 2.0f;
    float radius = sqrtf(hd * hd + b * b);
    float l1 = mid + radius;
    float l2 = fmaxf(mid - radius, 0.1f);
    float dvx = b, dvy = l1 - a;
    float len = sqrtf(dvx * dvx + dvy * dvy);
    if (!(len > 0.0f) || !(len <= 3.0e38f)) return 0;     /* normalize(0,0) = NaN in GLSL -> vanishes */
    float ex = dvx / len, ey = dvy / len;
    float s1 = fminf(sqrtf(2.0f * l1), 4096.0f);
    float s2 = fminf(sqrtf(2.0f * l2), 4096.0f);
    if (!(s1 > 0.0f) || !(s2 > 0.0f)) return 0;           /* NaN covariance */
    /* on-screen axes, y-up pixel frame: u1 = s1*(ex,ey), u2 = s2*(-ey,ex) */
    float u1x = s1 * ex, u1y = s1 * ey;
    float u2x = -(s2 * ey), u2y = s2 * ex;

    /* discard radius: A < 1/255  <=>  |q|^2 > ln(255*alpha) (SRC.h:306-310) */
    float pmax = (float)det_log((double)alpha * 255.0);
    if (!(pmax >= 0.0f)) return 0;

    /* pixel rectangle of the visible support (tightened): the +-2 box AABB and the
     * |q|^2 <= pmax ellipse AABB, whichever is smaller, plus a rounding-safety margin */
    float bxh = 2.0f * (fabsf(u1x) + fabsf(u2x));
    float byh = 2.0f * (fabsf(u1y) + fabsf(u2y));
    float rr = sqrtf(pmax);
    float exh = rr * sqrtf(u1x * u1x + u2x * u2x);
    float eyh = rr * sqrtf(u1y * u1y + u2y * u2y);
    float hx = fminf(bxh, exh); hx = hx + (hx * 0.0001f + 0.01f);
    float hy = fminf(byh, eyh); hy = hy + (hy * 0.0001f + 0.01f);
    float x0f = fmaxf(ceilf((cx - hx) - 0.5f), 0.0f);
    float x1f = fminf(floorf((cx + hx) - 0.5f), W - 1.0f);
    float y0f = fmaxf(ceilf((cy - hy) - 0.5f), 0.0f);
    float y1f = fminf(floorf((cy + hy) - 0.5f), H - 1.0f);
    if (!(x0f <= x1f) || !(y0f <= y1f)) return 0;
    int x0 = (int)x0f, x1 = (int)x1f, y0 = (int)y0f, y1 = (int)y1f;
    /* tile-row ownership (multi-GPU): survive iff some owned tile row is touched */
    if (F->row_world > 1) {
        int ty0 = y0 / ORC_TILE, ty1 = y1 / ORC_TILE, any = 0;
        for (int ty = ty0; ty <= ty1 && !any; ++ty) any = orc_owns_row(F, ty);
        if (!any) return 0;
    }
    rect->x0 = (uint16_t)x0; rect->x1 = (uint16_t)x1; rect->y0 = (uint16_t)y0; rect->y1 = (uint16_t)y1;

    /* colour (SRC.h:224, 244-275; LIB.h:117-179) */
    float rgb[3] = { h2f(cd_h[0]), h2f(cd_h[1]), h2f(cd_h[2]) };
    int order = F->sh_order;
    if (order > 0 && shx) {
        const float* OI = F->inv_object;
        float wv[3] = { ps[0] - F->cam[0], ps[1] - F->cam[1], ps[2] - F->cam[2] };
        float ov[3];
        for (int r = 0; r < 3; ++r)
            ov[r] = (MAT(OI, r, 0) * wv[0] + MAT(OI, r, 1) * wv[1]) + MAT(OI, r, 2) * wv[2];
        float dl = sqrtf((ov[0] * ov[0] + ov[1] * ov[1]) + ov[2] * ov[2]);
        float x = ov[0] / dl, y = ov[1] / dl, z = ov[2] / dl;
        const uint16_t* shc[3] = { shx, shy, shz };
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        for (int ch = 0; ch < 3; ++ch) {
            float sh[15];
            for (int j = 0; j < 15; ++j) sh[j] = h2f(shc[ch][j]);   /* element j at (j/4, j%4) row-major */
            float res = rgb[ch];
            res = res + SH_C1 * (((-sh[0]) * y + sh[1] * z) - sh[2] * x);
            if (order >= 2) {
                float t2 = (SH_C2_0 * xy) * sh[3];
                t2 = t2 + (SH_C2_1 * yz) * sh[4];
                t2 = t2 + (SH_C2_2 * ((2.0f * zz - xx) - yy)) * sh[5];
                t2 = t2 + (SH_C2_3 * xz) * sh[6];
                t2 = t2 + (SH_C2_4 * (xx - yy)) * sh[7];
                res = res + t2;
                if (order >= 3) {
                    float t3 = ((SH_C3_0 * y) * (3.0f * xx - yy)) * sh[8];
                    t3 = t3 + ((SH_C3_1 * xy) * z) * sh[9];
                    t3 = t3 + ((SH_C3_2 * y) * ((4.0f * zz - xx) - yy)) * sh[10];
                    t3 = t3 + ((SH_C3_3 * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy)) * sh[11];
                    t3 = t3 + ((SH_C3_4 * x) * ((4.0f * zz - xx) - yy)) * sh[12];
                    t3 = t3 + ((SH_C3_5 * z) * (xx - yy)) * sh[13];
                    t3 = t3 + ((SH_C3_6 * x) * (xx - 3.0f * yy)) * sh[14];
                    res = res + t3;
                }
            }
            rgb[ch] = fmaxf(res, 0.0f);
        }
    }

    rec->cx = cx; rec->cy = cy;
    rec->m00 = ex / s1;      rec->m01 = ey / s1;
    rec->m10 = (-ey) / s2;   rec->m11 = ex / s2;
    rec->alpha = alpha; rec->pmax = pmax;
    rec->r = rgb[0]; rec->g = rgb[1]; rec->b = rgb[2];
    rec->hpack = (uint32_t)f2h_ru_pos(hx) | ((uint32_t)f2h_ru_pos(hy) << 16);
    return 1;
}

/* keys for ALL n (culled -> 0xFFFFFFFF), records/rects valid where vis[i] != 0 */
int64_t orc_project(const orc_frame* F, int64_t n,
                    const float* pos, const uint16_t* cd_h, const float* alpha,
                    const uint16_t* scale_h, const uint16_t* orient_h,
                    const uint16_t* shx, const uint16_t* shy, const uint16_t* shz,
                    uint32_t* keys, orc_record* recs, orc_rect* rects, uint8_t* vis)
{
    int64_t nvis = 0;
#pragma omp parallel for schedule(static) reduction(+ : nvis)
    for (int64_t i = 0; i < n; ++i) {
        orc_record r; orc_rect q; memset(&r, 0, sizeof r); memset(&q, 0, sizeof q);
        int ok = project_one(F, pos + 3 * i, cd_h + 3 * i, alpha[i], scale_h + 3 * i, orient_h + 4 * i,
                             shx ? shx + 16 * i : nullptr, shy ? shy + 16 * i : nullptr,
                             shz ? shz + 16 * i : nullptr, &r, &q);
        vis[i] = (uint8_t)ok;
        keys[i] = ok ? key_of(pos + 3 * i, F->cam) : ORC_KEY_CULLED;
        recs[i] = r; rects[i] = q;
        nvis += ok;
    }
    return nvis;
}

/* ---------------------------------------------------------------- scene-depth occlusion
 * The reference draws with the depth test ON and depth writes OFF (R.C:608-610), and every vertex of a splat's quad
 * carries the CENTRE's clip z and w (SRC.h:278-282: out_vertex = centerClipPos, only xy displaced), so all fragments of
 * a splat have one window depth: zw = ndc_z * (far - near)/2 + (far + near)/2 with ndc_z = clip.z / clip.w and
 * (near, far) = glDepthRange (glH_DepthRange, SRC.h:158).  A fragment survives iff zw passes the depth function against
 * the scene depth already in the buffer at its pixel.  Spec: hr = (far - near) * 0.5f, hm = (far + near) * 0.5f,
 * zw = (ndc_z * hr) + hm, fp32, no contraction.  zw[i] is defined for every splat whose clip.w > 0 (else 0). */
void orc_window_depth(const orc_frame* F, int64_t n, const float* pos, const float depth_range[2], float* zw)
{
    const float hr = (depth_range[1] - depth_range[0]) * 0.5f, hm = (depth_range[1] + depth_range[0]) * 0.5f;
    const float* OV = F->obj_view; const float* P = F->proj;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* p = pos + 3 * i;
        float ps[3];
        for (int k = 0; k < 3; ++k) { float t = p[k] - F->origin[k]; ps[k] = t + F->origin[k]; }
        float vc[3];
        for (int r = 0; r < 3; ++r)
            vc[r] = ((MAT(OV, r, 0) * ps[0] + MAT(OV, r, 1) * ps[1]) + MAT(OV, r, 2) * ps[2]) + MAT(OV, r, 3);
        float fy = -vc[1];
        float cz = ((MAT(P, 2, 0) * vc[0] + MAT(P, 2, 1) * fy) + MAT(P, 2, 2) * vc[2]) + MAT(P, 2, 3);
        float cw = ((MAT(P, 3, 0) * vc[0] + MAT(P, 3, 1) * fy) + MAT(P, 3, 2) * vc[2]) + MAT(P, 3, 3);
        zw[i] = (cw > 0.0f) ? ((cz / cw) * hr) + hm : 0.0f;
    }
}

/* ---------------------------------------------------------------- binning
 * Instances are emitted in global depth order and stably partitioned by tile id.
 * tile id = ty * TX + tx, origin bottom-left.  Only owned tile rows are emitted.
 * tile_start has TX*TY + 1 entries.  Pass inst = NULL to only count.  Returns D. */
int64_t orc_bin(const orc_frame* F, int64_t n, const int32_t* order, const uint8_t* vis,
                const orc_rect* rects, int64_t* tile_start, int32_t* inst)
{
    /* parallel stable counting sort: the depth order is cut into one contiguous range per thread; per-thread tile
     * counts, then for every tile the exclusive prefix over threads gives each thread its write cursor, so every
     * tile's list keeps the global depth order.  Same output as the obvious serial loop. */
    const int TX = (F->width + ORC_TILE - 1) / ORC_TILE, TY = (F->height + ORC_TILE - 1) / ORC_TILE;
    const int64_t NT = (int64_t)TX * TY;
    int T = omp_get_max_threads();
    if (T < 1) T = 1;
    if ((int64_t)T * NT > (int64_t)64 << 20) T = (int)std::max<int64_t>(1, ((int64_t)64 << 20) / NT);   /* bound the table */
    std::vector<int64_t> cnt((size_t)T * (size_t)NT, 0);
    auto range = [&](int t, int64_t& r0, int64_t& r1) { r0 = n * t / T; r1 = n * (t + 1) / T; };
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        int64_t r0, r1; range(t, r0, r1);
        int64_t* c = cnt.data() + (size_t)t * (size_t)NT;
        for (int64_t r = r0; r < r1; ++r) {
            int32_t i = order[r]; if (!vis[i]) continue;
            const orc_rect& q = rects[i];
            for (int ty = q.y0 / ORC_TILE; ty <= q.y1 / ORC_TILE; ++ty) {
                if (!orc_owns_row(F, ty)) continue;
                for (int tx = q.x0 / ORC_TILE; tx <= q.x1 / ORC_TILE; ++tx) c[(size_t)ty * TX + tx] += 1;
            }
        }
    }
    /* tile totals -> tile_start; per-thread counts -> per-thread cursors (in place) */
    int64_t acc = 0;
    for (int64_t tile = 0; tile < NT; ++tile) {
        tile_start[tile] = acc;
        for (int t = 0; t < T; ++t) {
            int64_t c = cnt[(size_t)t * (size_t)NT + (size_t)tile];
            cnt[(size_t)t * (size_t)NT + (size_t)tile] = acc;
            acc += c;
        }
    }
    tile_start[NT] = acc;
    if (!inst) return acc;
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        int64_t r0, r1; range(t, r0, r1);
        int64_t* cur = cnt.data() + (size_t)t * (size_t)NT;
        for (int64_t r = r0; r < r1; ++r) {
            int32_t i = order[r]; if (!vis[i]) continue;
            const orc_rect& q = rects[i];
            for (int ty = q.y0 / ORC_TILE; ty <= q.y1 / ORC_TILE; ++ty) {
                if (!orc_owns_row(F, ty)) continue;
                for (int tx = q.x0 / ORC_TILE; tx <= q.x1 / ORC_TILE; ++tx) inst[cur[(size_t)ty * TX + tx]++] = i;
            }
        }
    }
    return acc;
}

/* ---------------------------------------------------------------- per-pixel (SRC.h:304-312, R.C:613-621)
 * state (C, T=1-dst.a); covered iff |qx|<=2, |qy|<=2, pw<=pmax; A=min(alpha*exp(-pw),1);
 * w=T*A; C=fma(w,rgb,C); T=T-w.  Returns 1 when the pixel just became saturated. */
static inline int shade(const orc_record& s, float px, float py, float eps, float* C, float* T)
{
    float dx = px - s.cx, dy = py - s.cy;
    float qx = fmaf(dy, s.m01, dx * s.m00);
    float qy = fmaf(dy, s.m11, dx * s.m10);
    float pw = fmaf(qy, qy, qx * qx);
    if (!(fabsf(qx) <= 2.0f && fabsf(qy) <= 2.0f && pw <= s.pmax)) return 0;
    float A = fminf(s.alpha * expf(-pw), 1.0f);
    float w = (*T) * A;
    C[0] = fmaf(w, s.r, C[0]); C[1] = fmaf(w, s.g, C[1]); C[2] = fmaf(w, s.b, C[2]);
    *T = *T - w;
    return (*T < eps) ? 1 : 0;
}

/* tiled blend.  rgba: H*W*4 floats, row 0 = bottom (GL).  Un-owned tile rows are left as is.
 * consumed[t] (may be NULL) = instances traversed until every pixel of tile t saturated. */
enum { ORC_DEPTH_NONE = 0, ORC_DEPTH_LESS = 1, ORC_DEPTH_LEQUAL = 2 };
static inline int depth_pass(int func, float zw, float sd)
{
    return func == ORC_DEPTH_LESS ? (zw < sd) : (func == ORC_DEPTH_LEQUAL ? (zw <= sd) : 1);
}

/* zw: window depth per splat (orc_window_depth), scene_depth: H*W floats, row 0 = bottom; depth_func = ORC_DEPTH_*.
 * A fragment that fails the depth test is dropped (no colour, no transmittance change). */
int64_t orc_blend_depth(const orc_frame* F, const orc_record* recs, const int64_t* tile_start,
                        const int32_t* inst, float* rgba, int64_t* consumed,
                        const float* zw, const float* scene_depth, int depth_func)
{
    if (!zw || !scene_depth) depth_func = ORC_DEPTH_NONE;
    const int W = F->width, H = F->height;
    const int TX = (W + ORC_TILE - 1) / ORC_TILE, TY = (H + ORC_TILE - 1) / ORC_TILE;
    const float eps = F->eps_t;
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
    for (int t = 0; t < TX * TY; ++t) {
        int ty = t / TX, tx = t % TX;
        if (!orc_owns_row(F, ty)) { if (consumed) consumed[t] = 0; continue; }
        float C[ORC_TILE * ORC_TILE][3]; float T[ORC_TILE * ORC_TILE]; uint8_t done[ORC_TILE * ORC_TILE];
        int live = 0;
        for (int j = 0; j < ORC_TILE; ++j) for (int i = 0; i < ORC_TILE; ++i) {
            int k = j * ORC_TILE + i; C[k][0] = C[k][1] = C[k][2] = 0.0f; T[k] = 1.0f;
            int inside = (tx * ORC_TILE + i < W) && (ty * ORC_TILE + j < H);
            done[k] = (uint8_t)!inside; live += inside;
        }
        int64_t s = tile_start[t], e = tile_start[t + 1], used = e - s;
        for (int64_t q = s; q < e; ++q) {
            const orc_record& sp = recs[inst[q]];
            for (int k = 0; k < ORC_TILE * ORC_TILE; ++k) {
                if (done[k]) continue;
                float px = (float)(tx * ORC_TILE + (k % ORC_TILE)) + 0.5f;
                float py = (float)(ty * ORC_TILE + (k / ORC_TILE)) + 0.5f;
                if (depth_func != ORC_DEPTH_NONE) {
                    int x = tx * ORC_TILE + (k % ORC_TILE), y = ty * ORC_TILE + (k / ORC_TILE);
                    if (!depth_pass(depth_func, zw[inst[q]], scene_depth[(size_t)y * W + x])) continue;
                }
                if (shade(sp, px, py, eps, C[k], &T[k])) { done[k] = 1; --live; }
            }
            if (live == 0) { used = q - s + 1; break; }
        }
        if (consumed) consumed[t] = used;
        total += used;
        for (int j = 0; j < ORC_TILE; ++j) for (int i = 0; i < ORC_TILE; ++i) {
            int x = tx * ORC_TILE + i, y = ty * ORC_TILE + j; if (x >= W || y >= H) continue;
            int k = j * ORC_TILE + i; float* o = rgba + ((size_t)y * W + x) * 4;
            o[0] = C[k][0]; o[1] = C[k][1]; o[2] = C[k][2]; o[3] = 1.0f - T[k];
        }
    }
    return total;
}

int64_t orc_blend(const orc_frame* F, const orc_record* recs, const int64_t* tile_start,
                  const int32_t* inst, float* rgba, int64_t* consumed)
{
    return orc_blend_depth(F, recs, tile_start, inst, rgba, consumed, nullptr, nullptr, ORC_DEPTH_NONE);
}

/* brute force: every pixel walks every visible splat in depth order, no rectangles, no tiles.
 * Validates that rect/tile culling (and the margin) never changes a pixel.  Small scenes only. */
void orc_blend_bruteforce(const orc_frame* F, int64_t n, const int32_t* order, const uint8_t* vis,
                          const orc_record* recs, float* rgba)
{
    const int W = F->width, H = F->height; const float eps = F->eps_t;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        float C[3] = { 0, 0, 0 }, T = 1.0f;
        float px = (float)x + 0.5f, py = (float)y + 0.5f;
        for (int64_t r = 0; r < n; ++r) {
            int32_t i = order[r]; if (!vis[i]) continue;
            if (shade(recs[i], px, py, eps, C, &T)) break;
        }
        float* o = rgba + ((size_t)y * W + x) * 4; o[0] = C[0]; o[1] = C[1]; o[2] = C[2]; o[3] = 1.0f - T;
    }
}

/* ---------------------------------------------------------------- whole frame (CPU baseline)
 * "reference semantics on N host cores (restatement, not llvmpipe)", BASELINE.md §3.
 * Times the reference's own CPU stage (distance + comparison argsort, R.C:176-216) separately. */
static double now_ms() { return omp_get_wtime() * 1e3; }

int orc_render(const orc_frame* F, int64_t n,
               const float* pos, const uint16_t* cd_h, const float* alpha,
               const uint16_t* scale_h, const uint16_t* orient_h,
               const uint16_t* shx, const uint16_t* shy, const uint16_t* shz,
               float* rgba, orc_stats* st, int time_reference_sort)
{
    const int TX = (F->width + ORC_TILE - 1) / ORC_TILE, TY = (F->height + ORC_TILE - 1) / ORC_TILE;
    std::vector<uint32_t> keys((size_t)n); std::vector<orc_record> recs((size_t)n);
    std::vector<orc_rect> rects((size_t)n); std::vector<uint8_t> vis((size_t)n);
    std::vector<int32_t> order((size_t)n);
    std::vector<int64_t> tile_start((size_t)TX * TY + 1);
    orc_stats s; memset(&s, 0, sizeof s); s.n_submitted = n;
    double t0 = now_ms();
    /* time_reference_sort: the depth order comes from the reference's own CPU stage (distance of every splat + comparison
     * argsort, R.C:176-216) and is used for the frame — one sort, as in the reference.  Otherwise: the oracle's stable
     * sort of the keys (culled splats last).  Both give the same sequence of visible splats (NaN distances aside). */
    if (time_reference_sort) { orc_sort_reference_style(pos, n, F->cam, order.data()); s.ms_sort_reference = now_ms() - t0; }
    double t1 = now_ms();
    s.n_visible = orc_project(F, n, pos, cd_h, alpha, scale_h, orient_h, shx, shy, shz,
                              keys.data(), recs.data(), rects.data(), vis.data());
    double t2 = now_ms(); s.ms_project = t2 - t1;
    if (!time_reference_sort) orc_sort(keys.data(), n, order.data());
    double t3 = now_ms(); s.ms_sort = t3 - t2;
    int64_t D = orc_bin(F, n, order.data(), vis.data(), rects.data(), tile_start.data(), nullptr);
    std::vector<int32_t> inst((size_t)D);
    orc_bin(F, n, order.data(), vis.data(), rects.data(), tile_start.data(), inst.data());
    double t4 = now_ms(); s.ms_bin = t4 - t3; s.n_instances = D;
    memset(rgba, 0, (size_t)F->width * F->height * 16);
    s.n_consumed = orc_blend(F, recs.data(), tile_start.data(), inst.data(), rgba, nullptr);
    double t5 = now_ms(); s.ms_blend = t5 - t4; s.ms_total = t5 - t0;
    if (st) *st = s;
    return 0;
}

/* ---------------------------------------------------------------- wireframe overlay
 * The reference's wire vertex shader (SRC.h:22-90) for all 8 vertices of every splat: raw position (no origin round
 * trip), covariance WITHOUT the object matrix (SRC.h:74), no culling.  verts[8 n][4] = gl_Position, colors[8 n][3] = Cd.
 * Same fp32 operation order as csrc/wire.cu (bit-exact). */
void orc_wire_vertices(const orc_frame* F, int64_t n, const float* pos, const uint16_t* cd_h, const uint16_t* scale_h,
                       const uint16_t* orient_h, float* verts, float* colors)
{
    const float W = (float)F->width, H = (float)F->height;
    const float* OV = F->obj_view; const float* P = F->proj; const float* V = F->view;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* p = pos + 3 * i;
        float vc[3];
        for (int r = 0; r < 3; ++r)
            vc[r] = ((MAT(OV, r, 0) * p[0] + MAT(OV, r, 1) * p[1]) + MAT(OV, r, 2) * p[2]) + MAT(OV, r, 3);
        float fy = -vc[1];
        float clip[4];
        for (int r = 0; r < 4; ++r)
            clip[r] = ((MAT(P, r, 0) * vc[0] + MAT(P, r, 1) * fy) + MAT(P, r, 2) * vc[2]) + MAT(P, r, 3);
        float sx = h2f(scale_h[3 * i]), sy = h2f(scale_h[3 * i + 1]), sz = h2f(scale_h[3 * i + 2]);
        float qx = h2f(orient_h[4 * i]), qy = h2f(orient_h[4 * i + 1]), qz = h2f(orient_h[4 * i + 2]), qr = h2f(orient_h[4 * i + 3]);
        float Rt[3][3];
        Rt[0][0] = 1.0f - 2.0f * (qy * qy + qz * qz); Rt[0][1] = 2.0f * (qx * qy + qr * qz); Rt[0][2] = 2.0f * (qx * qz - qr * qy);
        Rt[1][0] = 2.0f * (qx * qy - qr * qz); Rt[1][1] = 1.0f - 2.0f * (qx * qx + qz * qz); Rt[1][2] = 2.0f * (qy * qz + qr * qx);
        Rt[2][0] = 2.0f * (qx * qz + qr * qy); Rt[2][1] = 2.0f * (qy * qz - qr * qx); Rt[2][2] = 1.0f - 2.0f * (qx * qx + qy * qy);
        const float sc[3] = { sx, sy, sz };
        float Mm[3][3], S[3][3];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Mm[a][b] = sc[a] * Rt[a][b];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] = (Mm[0][a] * Mm[0][b] + Mm[1][a] * Mm[1][b]) + Mm[2][a] * Mm[2][b];
        float t[3];
        for (int r = 0; r < 3; ++r)
            t[r] = ((MAT(V, r, 0) * p[0] + MAT(V, r, 1) * p[1]) + MAT(V, r, 2) * p[2]) + MAT(V, r, 3);
        float aspect = MAT(P, 0, 0) / MAT(P, 1, 1);
        float tanFovX = 1.0f / MAT(P, 0, 0);
        float tanFovY = 1.0f / (MAT(P, 1, 1) * aspect);
        float limX = 1.3f * tanFovX, limY = 1.3f * tanFovY;
        float tz = t[2];
        float rx = t[0] / tz; rx = fminf(fmaxf(rx, -limX), limX);
        float ry = t[1] / tz; ry = fminf(fmaxf(ry, -limY), limY);
        float tx = rx * tz, ty = ry * tz;
        float focal = (W * MAT(P, 0, 0)) / 2.0f;
        float j0 = focal / tz;
        float tz2 = tz * tz;
        float j2x = -((focal * tx) / tz2);
        float j2y = -((focal * ty) / tz2);
        float A0[3], A1[3], B0[3], B1[3];
        for (int k = 0; k < 3; ++k) {
            A0[k] = j0 * MAT(V, 0, k) + j2x * MAT(V, 2, k);
            A1[k] = j0 * MAT(V, 1, k) + j2y * MAT(V, 2, k);
        }
        for (int k = 0; k < 3; ++k) {
            B0[k] = (A0[0] * S[0][k] + A0[1] * S[1][k]) + A0[2] * S[2][k];
            B1[k] = (A1[0] * S[0][k] + A1[1] * S[1][k]) + A1[2] * S[2][k];
        }
        float c00 = (B0[0] * A0[0] + B0[1] * A0[1]) + B0[2] * A0[2];
        float c01 = (B0[0] * A1[0] + B0[1] * A1[1]) + B0[2] * A1[2];
        float c11 = (B1[0] * A1[0] + B1[1] * A1[1]) + B1[2] * A1[2];
        float a = c00 + 0.3f, b = c01, c = c11 + 0.3f;
        float mid = 0.5f * (a + c);
        float hd = (a - c) / 2.0f;
        float radius = sqrtf(hd * hd + b * b);
        float l1 = mid + radius;
        float l2 = fmaxf(mid - radius, 0.1f);
        float dvx = b, dvy = l1 - a;
        float len = sqrtf(dvx * dvx + dvy * dvy);
        float ex = dvx / len, ey = dvy / len;
        float s1 = fminf(sqrtf(2.0f * l1), 4096.0f);
        float s2 = fminf(sqrtf(2.0f * l2), 4096.0f);
        float v1x = s1 * ex, v1y = s1 * (-ey);
        float v2x = s2 * (-ey), v2y = s2 * (-ex);
        static const float qcx[8] = { -2.f, 2.f, 2.f, 2.f, 2.f, -2.f, -2.f, -2.f };
        static const float qcy[8] = { -2.f, -2.f, -2.f, 2.f, 2.f, 2.f, 2.f, -2.f };
        for (int v = 0; v < 8; ++v) {
            float dx = ((qcx[v] * v1x + qcy[v] * v2x) * 2.0f) / W;
            float dy = ((qcx[v] * v1y + qcy[v] * v2y) * 2.0f) / H;
            float* o = verts + (8 * i + v) * 4;
            o[0] = clip[0] + dx * clip[3]; o[1] = -(clip[1] + dy * clip[3]); o[2] = clip[2]; o[3] = clip[3];
            if (colors) { float* c3 = colors + (8 * i + v) * 3; c3[0] = h2f(cd_h[3 * i]); c3[1] = h2f(cd_h[3 * i + 1]); c3[2] = h2f(cd_h[3 * i + 2]); }
        }
    }
}

/* the outlines rasterised into rgba (H x W x 4, left as is where no line passes): per pixel the nearest splat wins (ties: the
 * lower index), colour (Cd, 1).  Segment rule of csrc/wire.cu: n = ceil(max(|dx|, |dy|)) steps, samples a + (b - a)(i / n),
 * pixel = floor; splats whose centre is outside GL's clip volume are skipped. */
void orc_wire_overlay(int64_t n, const float* verts, const uint16_t* cd_h, int width, int height, float* rgba)
{
    std::vector<uint64_t> owner((size_t)width * height, ~0ull);
    const float W = (float)width, H = (float)height;
    for (int64_t seg = 0; seg < n * 4; ++seg) {
        const int64_t i = seg >> 2;
        const float* va = verts + (8 * i + 2 * (seg & 3)) * 4; const float* vb = va + 4;
        const float w = va[3], z = va[2];
        if (!(w > 0.0f) || !(z >= -w && z <= w)) continue;
        const float ax = ((va[0] / w + 1.0f) * 0.5f) * W, ay = ((va[1] / w + 1.0f) * 0.5f) * H;
        const float bx = ((vb[0] / w + 1.0f) * 0.5f) * W, by = ((vb[1] / w + 1.0f) * 0.5f) * H;
        if (!(ax == ax) || !(ay == ay) || !(bx == bx) || !(by == by)) continue;
        const float ddx = bx - ax, ddy = by - ay;
        const float m = fmaxf(fabsf(ddx), fabsf(ddy));
        if (!(m <= 1.0e9f)) continue;
        int steps = (int)ceilf(m);
        steps = steps < 1 ? 1 : (steps > 65536 ? 65536 : steps);
        const float depth = z / w;
        const float dk = depth * 0.5f + 0.5f;
        uint32_t db; memcpy(&db, &dk, 4);
        const uint64_t key = ((uint64_t)db << 32) | (uint64_t)(uint32_t)i;
        const float fn = (float)steps;
        for (int k = 0; k <= steps; ++k) {
            const float tpar = (float)k / fn;
            const float x = ax + ddx * tpar, y = ay + ddy * tpar;
            const float fx = floorf(x), fyy = floorf(y);
            if (fx >= 0.0f && fyy >= 0.0f && fx < W && fyy < H) {
                uint64_t& o = owner[(size_t)fyy * width + (size_t)fx];
                if (key < o) o = key;
            }
        }
    }
    for (size_t k = 0; k < owner.size(); ++k) {
        if (owner[k] == ~0ull) continue;
        const uint32_t i = (uint32_t)owner[k];
        rgba[4 * k] = h2f(cd_h[3 * (size_t)i]); rgba[4 * k + 1] = h2f(cd_h[3 * (size_t)i + 1]); rgba[4 * k + 2] = h2f(cd_h[3 * (size_t)i + 2]);
        rgba[4 * k + 3] = 1.0f;
    }
}

int orc_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline sets its thread count explicitly */
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_sizeof_frame(void) { return (int)sizeof(orc_frame); }
int orc_sizeof_record(void) { return (int)sizeof(orc_record); }

/* ---------------------------------------------------------------- config 1 (CPU only, no render)
 * HDK-free model of SOP cook + prim build + baryCentre (SOP_GSplat.C:93-117, GEO_GSplat.C:413-431,
 * 338-351): one prim, N vertices wired 1:1 to points 0..N-1; barycentre = sequential fp32 sum / N. */
void orc_build_prim(const float* pos, int64_t n, int32_t* vertex_to_point, float bary[3], float bbox[6])
{
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        vertex_to_point[i] = (int32_t)i;
        sx += pos[3 * i]; sy += pos[3 * i + 1]; sz += pos[3 * i + 2];
        if (i == 0) { bbox[0] = bbox[3] = pos[0]; bbox[1] = bbox[4] = pos[1]; bbox[2] = bbox[5] = pos[2]; }
        for (int k = 0; k < 3; ++k) {
            bbox[k] = fminf(bbox[k], pos[3 * i + k]); bbox[3 + k] = fmaxf(bbox[3 + k], pos[3 * i + k]);
        }
    }
    float fn = (float)n;
    bary[0] = sx / fn; bary[1] = sy / fn; bary[2] = sz / fn;
}

} /* extern "C" */

/* glsl_cxx.h — GLSL 3.30 value semantics in C++17, just enough to compile the reference's shader text on the host.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  Used by oracle/ref_harness.cpp to build oracle/_ref/libgsplat_ref.so from the
 * UNMODIFIED shader strings of /root/reference/gsplat_plugin/shaders/GSplatShaderCoreLib.h and GSplatShaderSource.h
 * (they are read from where they lie at build time, oracle/build_ref.py; nothing of them is stored in this repository).
 *
 * What is modelled, and how:
 *   float            -> class flt wrapping a C float: every operation rounds to fp32 (GLSL highp float), literals such
 *                       as 2.0 or 1 convert implicitly, and nothing is evaluated in double.  The shader text is compiled
 *                       with `#define float flt`.  Build with -ffp-contract=off: no FMA contraction.
 *   vecN / ivec2     -> structs with x y z w / r g b a members and the swizzles the text uses (.xy .xyz .rgb .rgba .xyzw
 *                       .wxyz) as proxy members aliasing the same storage (readable, and writable: `v.xy += ...`).
 *   mat3 / mat4      -> COLUMN-major: mat3(a,b,c, d,e,f, g,h,i) fills column 0 with (a,b,c); m[c][r]; mat3(mat4) takes the
 *                       upper-left 3x3; M * v and A * B as in GLSL (sums taken left to right over k = 0,1,2,...).
 *   built-ins        -> transpose, dot, length, normalize (v * inversesqrt(dot(v,v)) is one valid GLSL evaluation; here
 *                       v / sqrt(dot(v,v))), clamp, min, max, sqrt, exp (libm expf), texelFetch on sampler objects that
 *                       the harness binds to host arrays.
 *   evaluation order -> C++'s: operators of equal precedence associate left to right, as GLSL's grammar does.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

struct flt {
    float v;
    flt() = default;
    flt(float a) : v(a) {}
    flt(double a) : v((float)a) {}
    flt(int a) : v((float)a) {}
};
inline flt operator+(flt a, flt b) { return flt(a.v + b.v); }
inline flt operator-(flt a, flt b) { return flt(a.v - b.v); }
inline flt operator*(flt a, flt b) { return flt(a.v * b.v); }
inline flt operator/(flt a, flt b) { return flt(a.v / b.v); }
inline flt operator-(flt a) { return flt(-a.v); }
inline flt& operator+=(flt& a, flt b) { a.v = a.v + b.v; return a; }
inline flt& operator-=(flt& a, flt b) { a.v = a.v - b.v; return a; }
inline flt& operator*=(flt& a, flt b) { a.v = a.v * b.v; return a; }
inline flt& operator/=(flt& a, flt b) { a.v = a.v / b.v; return a; }
inline bool operator<(flt a, flt b) { return a.v < b.v; }
inline bool operator>(flt a, flt b) { return a.v > b.v; }
inline bool operator<=(flt a, flt b) { return a.v <= b.v; }
inline bool operator>=(flt a, flt b) { return a.v >= b.v; }
inline bool operator==(flt a, flt b) { return a.v == b.v; }
inline bool operator!=(flt a, flt b) { return a.v != b.v; }

inline flt sqrt(flt a) { return flt(::sqrtf(a.v)); }
inline flt exp(flt a) { return flt(::expf(a.v)); }
inline flt min(flt a, flt b) { return flt(b.v < a.v ? b.v : a.v); }          /* GLSL: y < x ? y : x */
inline flt max(flt a, flt b) { return flt(a.v < b.v ? b.v : a.v); }          /* GLSL: x < y ? y : x */
inline flt clamp(flt x, flt lo, flt hi) { return min(max(x, lo), hi); }      /* GLSL: min(max(x, minVal), maxVal) */
inline flt abs(flt a) { return flt(::fabsf(a.v)); }

struct vec2; struct vec3; struct vec4;

/* Component and swizzle proxies.  They live inside the vector's union and alias its storage d[]: `v.x`, `v.a`, `v.xy`,
 * `v.wxyz` read and write the components of v itself (a GNU anonymous struct cannot hold a class with constructors, so
 * even the single components are proxies; each converts to flt and assigns from flt). */
template <int I> struct comp {
    flt d[4];
    operator flt() const { return d[I]; }
    comp& operator=(flt o) { d[I] = o; return *this; }
    comp& operator=(const comp& o) { d[I] = o.d[I]; return *this; }
    comp& operator+=(flt o) { d[I] = d[I] + o; return *this; }
    comp& operator-=(flt o) { d[I] = d[I] - o; return *this; }
    comp& operator*=(flt o) { d[I] = d[I] * o; return *this; }
    comp& operator/=(flt o) { d[I] = d[I] / o; return *this; }
};
template <int A, int B> struct swz2 {
    flt d[4];
    inline operator vec2() const;
    inline swz2& operator=(const vec2& o);
    inline swz2& operator+=(const vec2& o);
};
template <int A, int B, int C> struct swz3 {
    flt d[4];
    inline operator vec3() const;
};
template <int A, int B, int C, int D> struct swz4 {
    flt d[4];
    inline operator vec4() const;
};

struct vec2 {
    union {
        flt d[4];                      /* two used */
        comp<0> x; comp<1> y;
        comp<0> r; comp<1> g;
    };
    vec2() = default;
    vec2(const vec2& o) { d[0] = o.d[0]; d[1] = o.d[1]; }
    vec2& operator=(const vec2& o) { d[0] = o.d[0]; d[1] = o.d[1]; return *this; }
    vec2(flt a, flt b) { d[0] = a; d[1] = b; }
    explicit vec2(flt a) { d[0] = a; d[1] = a; }
    flt& operator[](int i) { return d[i]; }
    const flt& operator[](int i) const { return d[i]; }
};
struct vec3 {
    union {
        flt d[4];                      /* three used */
        comp<0> x; comp<1> y; comp<2> z;
        comp<0> r; comp<1> g; comp<2> b;
    };
    vec3() = default;
    vec3(const vec3& o) { d[0] = o.d[0]; d[1] = o.d[1]; d[2] = o.d[2]; }
    vec3& operator=(const vec3& o) { d[0] = o.d[0]; d[1] = o.d[1]; d[2] = o.d[2]; return *this; }
    vec3(flt a, flt b, flt c) { d[0] = a; d[1] = b; d[2] = c; }
    explicit vec3(flt a) { d[0] = a; d[1] = a; d[2] = a; }
    flt& operator[](int i) { return d[i]; }
    const flt& operator[](int i) const { return d[i]; }
};
struct vec4 {
    union {
        flt d[4];
        comp<0> x; comp<1> y; comp<2> z; comp<3> w;
        comp<0> r; comp<1> g; comp<2> b; comp<3> a;
        swz2<0, 1> xy;
        swz3<0, 1, 2> xyz;
        swz3<0, 1, 2> rgb;
        swz4<0, 1, 2, 3> xyzw;
        swz4<0, 1, 2, 3> rgba;
        swz4<3, 0, 1, 2> wxyz;
    };
    vec4() = default;
    vec4(const vec4& o) { for (int i = 0; i < 4; ++i) d[i] = o.d[i]; }
    vec4& operator=(const vec4& o) { for (int i = 0; i < 4; ++i) d[i] = o.d[i]; return *this; }
    vec4(flt p, flt q, flt s, flt t) { d[0] = p; d[1] = q; d[2] = s; d[3] = t; }
    vec4(const vec3& v, flt t) { d[0] = v.d[0]; d[1] = v.d[1]; d[2] = v.d[2]; d[3] = t; }
    vec4(const vec2& v, flt s, flt t) { d[0] = v.d[0]; d[1] = v.d[1]; d[2] = s; d[3] = t; }
    flt& operator[](int i) { return d[i]; }
    const flt& operator[](int i) const { return d[i]; }
};
template <int A, int B> inline swz2<A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int A, int B> inline swz2<A, B>& swz2<A, B>::operator=(const vec2& o) { d[A] = o.d[0]; d[B] = o.d[1]; return *this; }
template <int A, int B> inline swz2<A, B>& swz2<A, B>::operator+=(const vec2& o) { d[A] = d[A] + o.d[0]; d[B] = d[B] + o.d[1]; return *this; }
template <int A, int B, int C> inline swz3<A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int A, int B, int C, int D> inline swz4<A, B, C, D>::operator vec4() const { return vec4(d[A], d[B], d[C], d[D]); }

struct ivec2 {
    int x, y;
    ivec2() = default;
    ivec2(int a, int b) : x(a), y(b) {}
};
inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
struct ivec4 {
    union { struct { int x, y, z, w; }; struct { int r, g, b, a; }; };
};

/* ---- vec2 */
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, const vec2& b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(const vec2& a, flt s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(flt s, const vec2& a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(const vec2& a, flt s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator+(const vec2& a, flt s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(const vec2& a, flt s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec2& operator*=(vec2& a, flt s) { a.x = a.x * s; a.y = a.y * s; return a; }
inline vec2& operator+=(vec2& a, const vec2& b) { a.x = a.x + b.x; a.y = a.y + b.y; return a; }
inline flt dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline flt length(const vec2& a) { return sqrt(dot(a, a)); }
inline vec2 normalize(const vec2& a) { flt l = length(a); return vec2(a.x / l, a.y / l); }

/* ---- vec3 */
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, flt s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(flt s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, flt s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, const vec3& b) { a.x = a.x + b.x; a.y = a.y + b.y; a.z = a.z + b.z; return a; }
inline vec3& operator*=(vec3& a, flt s) { a.x = a.x * s; a.y = a.y * s; a.z = a.z * s; return a; }
inline flt dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline flt length(const vec3& a) { return sqrt(dot(a, a)); }
inline vec3 normalize(const vec3& a) { flt l = length(a); return vec3(a.x / l, a.y / l, a.z / l); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }

/* ---- vec4 */
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const vec4& a, flt s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }

/* ---- matrices, column-major: c[col][row] */
struct mat4;
struct mat3 {
    vec3 c[3];
    mat3() = default;
    mat3(flt a0, flt a1, flt a2, flt b0, flt b1, flt b2, flt c0, flt c1, flt c2)
    { c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2); }
    explicit inline mat3(const mat4& m);
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() = default;
    mat4(flt a0, flt a1, flt a2, flt a3, flt b0, flt b1, flt b2, flt b3, flt c0, flt c1, flt c2, flt c3,
         flt d0, flt d1, flt d2, flt d3)
    { c[0] = vec4(a0, a1, a2, a3); c[1] = vec4(b0, b1, b2, b3); c[2] = vec4(c0, c1, c2, c3); c[3] = vec4(d0, d1, d2, d3); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m)
{
    for (int j = 0; j < 3; ++j) c[j] = vec3(m.c[j].d[0], m.c[j].d[1], m.c[j].d[2]);
}
inline mat3 transpose(const mat3& m)
{
    mat3 t;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) t.c[j][i] = m.c[i][j];
    return t;
}
inline vec3 operator*(const mat3& m, const vec3& v)
{
    vec3 r;
    for (int i = 0; i < 3; ++i) r[i] = m.c[0][i] * v.x + m.c[1][i] * v.y + m.c[2][i] * v.z;
    return r;
}
inline mat3 operator*(const mat3& a, const mat3& b)
{
    mat3 r;
    for (int j = 0; j < 3; ++j) r.c[j] = a * b.c[j];
    return r;
}
inline vec4 operator*(const mat4& m, const vec4& v)
{
    vec4 r;
    for (int i = 0; i < 4; ++i) r[i] = m.c[0][i] * v.x + m.c[1][i] * v.y + m.c[2][i] * v.z + m.c[3][i] * v.w;
    return r;
}
inline mat4 operator*(const mat4& a, const mat4& b)
{
    mat4 r;
    for (int j = 0; j < 4; ++j) r.c[j] = a * b.c[j];
    return r;
}

/* ---- samplers: bound by the harness to host arrays laid out exactly as the reference's textures
 * (src/GSplatRenderer.C:448-505): texel (x, y) of a dim x dim texture is element y * dim + x. */
struct sampler2D {
    const void* ctx = nullptr;
    vec4 (*fetch)(const void* ctx, int linear) = nullptr;
    int dim = 0;
};
struct isampler2D {
    const void* ctx = nullptr;
    int (*fetch)(const void* ctx, int linear) = nullptr;
    int dim = 0;
};
inline vec4 texelFetch(const sampler2D& s, ivec2 p, int /*lod*/) { return s.fetch(s.ctx, p.y * s.dim + p.x); }
inline ivec4 texelFetch(const isampler2D& s, ivec2 p, int /*lod*/)
{
    ivec4 r; r.x = s.fetch(s.ctx, p.y * s.dim + p.x); r.y = 0; r.z = 0; r.w = 1; return r;
}

}  // namespace glsl

// ingest.cu — attribute ingestion on the GPU: what GR_PrimGsplat::update does with TBB on the
// CPU (/root/reference/gsplat_plugin/src/GR_GSplat.C:302-372) — read the raw fp32 point attributes, apply the defaults
// (GR.C:309-312), quantise Cd / scale / orient / SH to IEEE half (UT_Vector3H(...), round-to-nearest-even) and lay the
// SH coefficients out as three 4x4 half matrices (coefficient j at (j/4, j%4), [3][3] = 0) from any of the three
// encodings the reference accepts (GR.C:145-189, 326-368).  Output = exactly the arrays registerUpdate receives.
// Pure streaming byte work, HBM-bound: <= 236 B read and 132 B written per point.
#include "common.cuh"

namespace gsb {

namespace {

__device__ __forceinline__ uint16_t f2h(float v) { return __half_as_ushort(__float2half_rn(v)); }

// exp(x) in double with + - * and one rint only (no libm, no FMA contraction: this TU is built with -fmad=false), so the
// numpy restatement (oracle/ingest.py det_exp) produces the same bits: x = k ln2 + r, |r| <= 0.35, Taylor to r^13 / 13!
// (truncation 5e-18), scaled by 2^k in two exact steps.  Used by the INRIA activation below.
__device__ __forceinline__ double det_exp(double x)
{
    if (!(x == x)) return x;
    if (x > 709.0) return __longlong_as_double(0x7ff0000000000000ll);
    if (x < -745.0) return 0.0;
    const double k = rint(x * 1.4426950408889634);
    const double r = (x - k * 6.93147180369123816490e-01) - k * 1.90821492927058770002e-10;
    double p = 1.0 / 6227020800.0;                 // 1/13!
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    const int ki = (int)k, k1 = ki / 2, k2 = ki - k1;        // 2^k = 2^k1 * 2^k2, both normal doubles
    const double s1 = __longlong_as_double((long long)(k1 + 1023) << 52), s2 = __longlong_as_double((long long)(k2 + 1023) << 52);
    return (p * s1) * s2;
}

// pos / alpha copy, Cd / scale / orient quantisation with defaults.
// activation = GSB_ACT_INRIA: the sources are the raw columns of an INRIA 3D-Gaussian-splatting PLY and the
// conversion of the example scene's point wrangles (SURVEY 8a note N1; hip/GSplatPlugin_simpleScene_v001.hip) runs here:
//   Cd = 0.28209479177387814 f_dc + 0.5;  alpha = 1 / (1 + exp(-opacity));  scale = exp(scale_raw);
//   orient = normalize(rot_1, rot_2, rot_3, rot_0)   (INRIA stores w first, Houdini quaternions are (x, y, z, w))
// fp32 results (exp in double, deterministic, then rounded), then the same half quantisation as without activation.
__global__ void __launch_bounds__(256)
ingest_core_kernel(const float* __restrict__ P, const float* __restrict__ Cd, const float* __restrict__ alpha,
                   const float* __restrict__ scale, const float* __restrict__ orient, int64_t n, const int activation,
                   float* __restrict__ pos_out, uint16_t* __restrict__ cd_out, float* __restrict__ alpha_out,
                   uint16_t* __restrict__ scale_out, uint16_t* __restrict__ orient_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool inria = activation == 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pos_out[3 * i + k] = P[3 * i + k];
        float c = Cd ? Cd[3 * i + k] : 0.0f;                             // default colour (0,0,0)
        if (inria && Cd) c = 0.28209479177387814f * c + 0.5f;
        cd_out[3 * i + k] = f2h(c);
        float sc = scale ? scale[3 * i + k] : 1.0f;                      // default scale (1,1,1)
        if (inria && scale) sc = (float)det_exp((double)sc);
        scale_out[3 * i + k] = f2h(sc);
    }
    float a = alpha ? alpha[i] : 1.0f;                                   // default alpha 1
    if (inria && alpha) a = (float)(1.0 / (1.0 + det_exp(-(double)a)));
    alpha_out[i] = a;
    float q[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
    if (orient) {
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = orient[4 * i + k];
        if (inria) {
            const float w = q[0], x = q[1], y = q[2], z = q[3];
            const float nn = sqrtf(((x * x + y * y) + z * z) + w * w);
            const float d = nn > 0.0f ? nn : 1.0f;
            q[0] = x / d; q[1] = y / d; q[2] = z / d; q[3] = w / d;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) orient_out[4 * i + k] = f2h(q[k]);
}

// SH from `sh_coefficients` ([n][len][3]) or from 15 vec3 attributes laid out as [15][n][3]
__global__ void __launch_bounds__(256)
ingest_sh_vec3_kernel(const float* __restrict__ src, int64_t n, int len, int planar,
                      uint16_t* __restrict__ shx, uint16_t* __restrict__ shy, uint16_t* __restrict__ shz)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float x = 0.0f, y = 0.0f, z = 0.0f;                              // matrices start as UT_Matrix4F(0.0), GR.C:322-324
        if (j < len && j < 15) {
            const float* p = planar ? src + ((size_t)j * n + i) * 3 : src + ((size_t)i * len + j) * 3;
            x = p[0]; y = p[1]; z = p[2];
        }
        shx[16 * i + j] = f2h(x); shy[16 * i + j] = f2h(y); shz[16 * i + j] = f2h(z);
    }
}

// SH from f_rest_0..44 laid out as [45][n]: coefficient j has R = f_rest_j, G = f_rest_{j+15}, B = f_rest_{j+30} (GR.C:357-366)
__global__ void __launch_bounds__(256)
ingest_sh_rest_kernel(const float* __restrict__ rest, int64_t n,
                      uint16_t* __restrict__ shx, uint16_t* __restrict__ shy, uint16_t* __restrict__ shz)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float x = 0.0f, y = 0.0f, z = 0.0f;
        if (j < 15) { x = rest[(size_t)j * n + i]; y = rest[(size_t)(j + 15) * n + i]; z = rest[(size_t)(j + 30) * n + i]; }
        shx[16 * i + j] = f2h(x); shy[16 * i + j] = f2h(y); shz[16 * i + j] = f2h(z);
    }
}
}  // namespace

void launch_ingest_core(const float* P, const float* Cd, const float* alpha, const float* scale, const float* orient,
                        int64_t n, int activation, float* pos_out, uint16_t* cd_out, float* alpha_out, uint16_t* scale_out,
                        uint16_t* orient_out, cudaStream_t s)
{
    if (n <= 0) return;
    ingest_core_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(P, Cd, alpha, scale, orient, n, activation, pos_out, cd_out,
                                                                  alpha_out, scale_out, orient_out);
}

void launch_ingest_sh_vec3(const float* src, int64_t n, int len, int planar, uint16_t* shx, uint16_t* shy, uint16_t* shz,
                           cudaStream_t s)
{
    if (n <= 0) return;
    ingest_sh_vec3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, n, len, planar, shx, shy, shz);
}

void launch_ingest_sh_rest(const float* rest, int64_t n, uint16_t* shx, uint16_t* shy, uint16_t* shz, cudaStream_t s)
{
    if (n <= 0) return;
    ingest_sh_rest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(rest, n, shx, shy, shz);
}

}  // namespace gsb

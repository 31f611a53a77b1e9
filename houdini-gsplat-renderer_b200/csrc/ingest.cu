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

// pos / alpha copy, Cd / scale / orient quantisation with defaults
__global__ void __launch_bounds__(256)
ingest_core_kernel(const float* __restrict__ P, const float* __restrict__ Cd, const float* __restrict__ alpha,
                   const float* __restrict__ scale, const float* __restrict__ orient, int64_t n,
                   float* __restrict__ pos_out, uint16_t* __restrict__ cd_out, float* __restrict__ alpha_out,
                   uint16_t* __restrict__ scale_out, uint16_t* __restrict__ orient_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pos_out[3 * i + k] = P[3 * i + k];
        cd_out[3 * i + k] = f2h(Cd ? Cd[3 * i + k] : 0.0f);              // default colour (0,0,0)
        scale_out[3 * i + k] = f2h(scale ? scale[3 * i + k] : 1.0f);     // default scale (1,1,1)
    }
    alpha_out[i] = alpha ? alpha[i] : 1.0f;                              // default alpha 1
#pragma unroll
    for (int k = 0; k < 4; ++k) orient_out[4 * i + k] = f2h(orient ? orient[4 * i + k] : (k == 3 ? 1.0f : 0.0f));
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
                        int64_t n, float* pos_out, uint16_t* cd_out, float* alpha_out, uint16_t* scale_out,
                        uint16_t* orient_out, cudaStream_t s)
{
    if (n <= 0) return;
    ingest_core_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(P, Cd, alpha, scale, orient, n, pos_out, cd_out, alpha_out,
                                                                  scale_out, orient_out);
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

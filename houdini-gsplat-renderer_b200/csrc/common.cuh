// common.cuh — shared device/host declarations for libgsplat_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstddef>

namespace gsb {

constexpr int      TILE       = 16;            // screen tile edge
constexpr uint32_t KEY_CULLED = 0xFFFFFFFFu;   // depth key of a culled splat: sorts last
constexpr int      NUM_SMS    = 148;           // B200

// Per-frame constants handed to the kernels by value (fits the 4 KB kernel-parameter space).
// Matrices are column-major: element (row r, col c) = m[c*4 + r].
struct FrameConsts {
    float view[16], proj[16], object[16], inv_object[16], obj_view[16];
    float cam[3];
    float origin[3];
    float W, H;                 // glH_ScreenSize as floats
    int   width, height;
    int   tiles_x, tiles_y;
    int   sh_order;             // effective order (0 when the packed set has no SH)
    int   row_rank, row_world;  // tile-row ownership: row ty is owned iff (ty / row_group) % row_world == row_rank
    int   row_group;
    float eps_t;                // transmittance early-out threshold
};

// 2-D record written by project and gathered by blend: 48 bytes, three 16-byte chunks.
struct __align__(16) Record {
    float cx, cy, m00, m01;
    float m10, m11, alpha, pmax;
    float r, g, b;
    uint32_t hpack;             // half(hx) | half(hy) << 16, rounded toward +inf (conservative cull data)
};
static_assert(sizeof(Record) == 48, "Record must be 48 bytes");

// Packed, render-layout splat attributes (built by pack on active-set change):
//   geomA[i] = (p.x, p.y, p.z, alpha)                         16 B
//   geomB[i] = (scale h3 | orient h4 (x,y,z,w) | pad h1)      16 B
//   col[k][i], k = 0..5: 48 halfs = Cd(3) then SH coefficient j channel c at 3+3j+c   6 x 16 B
// Degree d needs halfs [0, 3 + 3*{0,3,8,15}) -> planes {1,2,4,6}.
struct PackedSplats {
    const float4* geomA;
    const uint4*  geomB;
    const uint4*  col[6];
};

// radix sort / scan primitives (radix_sort.cu, scan.cu)
size_t   scan_scratch_bytes(size_t n);
// exclusive scan of n uint32; out may alias in.  If total_dev != nullptr the 64-bit grand total is stored there.
void     exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void* scratch,
                            unsigned long long* total_dev, cudaStream_t s, int* launches);
size_t   sort_scratch_bytes(size_t n);
// Stable LSD radix sort of (key,val) pairs on key bits [begin_bit, end_bit).  Ping-pongs between
// (k0,v0) and (k1,v1); returns 0 if the result is in (k0,v0), 1 if in (k1,v1).  n must be < 2^30.
// *error_flag (device, may be NULL) is set to 1 if a bounded look-back spin ever times out.
// If gather_src != NULL the last pass also writes gather_dst[sorted position] = gather_src[value] (8-byte items).
int      radix_sort_pairs(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, size_t n,
                          int begin_bit, int end_bit, void* scratch, uint32_t* error_flag, cudaStream_t s, int* launches,
                          const uint2* gather_src = nullptr, uint2* gather_dst = nullptr,
                          uint32_t key_min = 0u, uint32_t key_span = 0xFFFFFFFFu);
// The sort orders by min(key - key_min, key_span) (order-preserving for keys in [key_min, key_min + key_span),
// everything above collapses onto key_span); pass end_bit = sort_key_bits(key_span) to sort only the bits that vary.
int      sort_key_bits(uint32_t key_span);

// project.cu
void launch_pack(const float* pos, const uint16_t* cd_h, const float* alpha, const uint16_t* scale_h,
                 const uint16_t* orient_h, const uint16_t* shx, const uint16_t* shy, const uint16_t* shz,
                 int64_t count, int64_t dst_offset, float4* geomA, uint4* geomB, uint4* const col[6],
                 int planes, cudaStream_t s);
// K1: keys (culled -> KEY_CULLED), vals = splat index, rects (x0>x1 = culled), records, *n_visible += V
// vis_flags (may be NULL): 1/0 per splat, the input of launch_compact's scan
void launch_project(const FrameConsts& fc, const PackedSplats& ps, int64_t n,
                    uint32_t* keys, uint32_t* vals, Record* recs, uint2* rects,
                    unsigned long long* n_visible, uint32_t* vis_flags, cudaStream_t s);
// order-preserving compaction of the surviving (key, index) pairs; positions = exclusive scan of vis_flags
void launch_compact(const uint32_t* keys, const uint32_t* positions, int64_t n, uint32_t* keys_out, uint32_t* vals_out,
                    cudaStream_t s);

// binning.cu
// ranks [r0, r0+n) of the depth order; rects_sorted in depth order; tile_done (may be NULL) = saturation flags
void launch_tile_counts(const uint2* rects_sorted, int64_t r0, int64_t n, FrameConsts fc,
                        const uint32_t* tile_done, uint32_t* counts, cudaStream_t s);
// offsets = exclusive scan of the tile counts, *total = its grand total (device)
void launch_emit(const uint32_t* order, const uint2* rects_sorted, const uint32_t* offsets,
                 const unsigned long long* total, int64_t r0, int64_t n, FrameConsts fc, const uint32_t* tile_done,
                 uint32_t* inst_keys, uint32_t* inst_vals, cudaStream_t s);
void launch_tile_ranges(const uint32_t* sorted_tile_ids, uint64_t d, uint2* ranges, int num_tiles,
                        cudaStream_t s);

// ingest.cu: raw fp32 point attributes -> the arrays registerUpdate receives (NULL source = default)
void launch_ingest_core(const float* P, const float* Cd, const float* alpha, const float* scale, const float* orient,
                        int64_t n, float* pos_out, uint16_t* cd_out, float* alpha_out, uint16_t* scale_out,
                        uint16_t* orient_out, cudaStream_t s);
// src = [n][len][3] (planar = 0, `sh_coefficients`) or [15][n][3] (planar = 1, `sh1`..`sh15`)
void launch_ingest_sh_vec3(const float* src, int64_t n, int len, int planar, uint16_t* shx, uint16_t* shy, uint16_t* shz,
                           cudaStream_t s);
// rest = [45][n] (`f_rest_0`..`f_rest_44`)
void launch_ingest_sh_rest(const float* rest, int64_t n, uint16_t* shx, uint16_t* shy, uint16_t* shz, cudaStream_t s);

// blend.cu
// One depth chunk.  first: pixel state starts at (0,0,0,T=1), otherwise it is reloaded from fb, which between
// chunks holds (C, T).  A tile whose pixels are all saturated is finalised to (C, 1-T) and flagged in tile_done;
// last: every remaining tile is finalised.  Finalised tiles are stored to fb_final (NULL = fb; may be peer memory).  *done_tiles counts the tiles flagged so far (early termination).
void launch_blend(const Record* recs, const uint32_t* inst_vals, const uint2* ranges, float4* fb, float4* fb_final,
                  FrameConsts fc, int first, int last, uint32_t* tile_done, uint32_t* tile_consumed,
                  unsigned long long* consumed_total, unsigned long long* done_tiles, cudaStream_t s);

// tile-row ownership rule shared by every kernel
__host__ __device__ __forceinline__ bool owns_row(int ty, int rank, int world, int group)
{
    return world <= 1 || ((ty / group) % world) == rank;
}

}  // namespace gsb

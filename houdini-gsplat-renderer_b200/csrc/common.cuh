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
    // bounded K1 (project_bound_kernel): constants of the covariance chain evaluated on the host with the spec's fp32
    // operations (focal = W*P00/2, limX/Y = 1.3*tanFov) and the squared spectral norm of mat3(view), rounded up
    float focal, lim_x, lim_y, wnorm2;
    int   depth_func;           // scene-depth occlusion: 0 none, 1 LESS, 2 LEQUAL (gsb_depth_func)
    float depth_hr, depth_hm;   // window depth = (clip.z / clip.w) * depth_hr + depth_hm  ((far-near)/2, (far+near)/2)
};

// 2-D record written by project and gathered by blend: 48 bytes, three 16-byte chunks.
struct __align__(16) Record {
    float cx, cy, m00, m01;
    float m10, m11, alpha, pmax;
    float r, g, b;
    uint32_t hpack;             // half(hx) | half(hy) << 16, rounded toward +inf (conservative cull data)
};
static_assert(sizeof(Record) == 48, "Record must be 48 bytes");

// Packed, render-layout splat attributes (built by pack on active-set change).  Two copies of the geometry:
//   streamed by K1 (every splat, every frame), SoA planes:
//     geomA[i] = (p.x, p.y, p.z, rr)                            16 B   rr = sqrt(ln(255 alpha)) or -1 (never passes the discard)
//     geomB[i] = (scale h3 | orient h4 (x,y,z,w) | pad h1)      16 B
//   gathered by K2 (only the splats that reach a live tile), one 128-byte line per splat:
//     rows[8*i + 0] = geomA bits, rows[8*i + 1] = geomB,
//     rows[8*i + 2 + k], k = 0..5: 48 halfs = Cd(3) then SH coefficient j channel c at 3+3j+c
//   Degree d needs halfs [0, 3 + 3*{0,3,8,15}) -> colour chunks {1,2,4,6} of the line.
constexpr int ROW_U4 = 8;                      // uint4 per splat line
//   cached per object matrix (launch_sigma): world-space covariance, upper triangle
//     sigA[i] = (S00, S01, S02, S11), sigB[i] = (S12, S22)      24 B
//     lam[i]  = upper bound of the largest eigenvalue of that covariance                       4 B   (bounded K1)
struct PackedSplats {
    const float4* geomA;
    const uint4*  geomB;
    const uint4*  rows;
    const float4* sigA;
    const float2* sigB;
    const float*  lam;
};

// Depth buckets: a monotone (non-decreasing) map from the depth key to [0, DEPTH_BUCKETS-2]: the key's offset from the
// smallest possible key of the packed set, shifted so the whole key range spans the buckets; culled splats (KEY_CULLED)
// take the last bucket.  Used to cut the depth order into chunks without sorting or moving the cloud: the chunk plan
// only needs counts per bucket (quantiles), so any monotone map works, and this one is two integer instructions.
constexpr int DEPTH_BUCKETS = 512;
constexpr int MAX_CHUNKS    = 16;
struct DepthBuckets { uint32_t key_min; int shift; };
__device__ __forceinline__ uint32_t depth_bucket(uint32_t key, DepthBuckets db)
{
    if (key == KEY_CULLED) return (uint32_t)(DEPTH_BUCKETS - 1);
    const uint32_t o = key > db.key_min ? key - db.key_min : 0u;
    return min(o >> db.shift, (uint32_t)(DEPTH_BUCKETS - 2));
}
// Chunk plan chosen on the device from the bucket histogram (choose_chunks).
struct ChunkPlan {
    uint32_t size[MAX_CHUNKS + 1];             // visible splats per chunk; entry nchunks = culled
    uint32_t key_lo[MAX_CHUNKS + 2];           // chunk c holds the keys in [key_lo[c], key_lo[c + 1]); key_lo[nchunks] = KEY_CULLED
    uint8_t  lut[DEPTH_BUCKETS];               // bucket -> chunk
};

// radix sort / scan primitives (radix_sort.cu, scan.cu)
size_t   scan_scratch_bytes(size_t n);
// exclusive scan of n uint32; out may alias in.  If total_dev != nullptr the 64-bit grand total is stored there.
void     exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void* scratch,
                            unsigned long long* total_dev, cudaStream_t s, int* launches);
size_t   scan_status_bytes(size_t n);
void     exclusive_scan_u32_onepass(const uint32_t* in, uint32_t* out, size_t n, unsigned long long* status, uint32_t epoch,
                                    unsigned long long* total_dev, uint32_t* error_flag, cudaStream_t s, int* launches);
// Stable LSD radix sort of (key,val) pairs on key bits [begin_bit, end_bit) of the squeezed key (see below).
// Ping-pongs between (k0,v0) and (k1,v1); returns 0 if the result is in (k0,v0), 1 if in (k1,v1).
//   n_max      host-side upper bound of the element count (< 2^30): sizes the grid cap and the look-back table
//   n_dev      device pointer to the real count (clamped to n_max), or NULL: n_max is the count
//   header     sort_header_bytes() of device memory for this invocation: digit histograms + tile tickets.  Must be
//              zero when the first kernel that touches it runs (header_is_zero = false: this call clears it itself)
//   hist_ready the histograms in the header were already filled by the kernel that produced the keys (sort_plan gives
//              the digit layout: pass p counts ((squeeze(key) >> shift[p]) & (2^bits[p] - 1)) at header[p * SORT_RADIX + d])
//   lookback   sort_lookback_bytes(n_max) bytes, shared by all sorts of a stream; zero it ONCE after allocation, then
//              never again: entries carry the epoch of the sort that wrote them.  epoch must differ from every epoch
//              the table has seen (a counter)
//   error_flag device word (may be NULL), set to 1 if a bounded look-back spin ever times out
// The sort orders by min(key - key_min, key_span) (order-preserving for keys in [key_min, key_min + key_span),
// everything above collapses onto key_span); pass end_bit = sort_key_bits(key_span) to sort only the bits that vary.
constexpr int SORT_MAX_PASSES = 4;
constexpr int SORT_RADIX      = 512;
struct SortPlan { int shift[SORT_MAX_PASSES]; int bits[SORT_MAX_PASSES]; int passes; };
SortPlan sort_plan(int begin_bit, int end_bit);
size_t   sort_header_bytes();
size_t   sort_lookback_bytes(size_t n_max);
int      radix_sort_pairs(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, size_t n_max,
                          const unsigned long long* n_dev, int begin_bit, int end_bit,
                          uint32_t* header, bool header_is_zero, bool hist_ready, unsigned long long* lookback, uint32_t epoch,
                          uint32_t* error_flag, cudaStream_t s, int* launches,
                          uint32_t key_min = 0u, uint32_t key_span = 0xFFFFFFFFu);
int      sort_key_bits(uint32_t key_span);
__host__ __device__ __forceinline__ uint32_t sort_squeeze(uint32_t key, uint32_t key_min, uint32_t key_span)
{
    const uint32_t d = key - key_min;
    return d < key_span ? d : key_span;
}

// project.cu
void launch_pack(const float* pos, const uint16_t* cd_h, const float* alpha, const uint16_t* scale_h,
                 const uint16_t* orient_h, const uint16_t* shx, const uint16_t* shy, const uint16_t* shz,
                 int64_t count, int64_t dst_offset, float4* geomA, uint4* geomB, uint4* rows, int has_sh, cudaStream_t s);
// world-space covariance planes (sigA / sigB, may be NULL: only the exact K1 reads them) and the eigenvalue bound lam from
// geomB and the object matrix (run when the object matrix or the packed set changes)
void launch_sigma(const float object[16], const uint4* geomB, int64_t n, float4* sigA, float2* sigB, float* lam, cudaStream_t s);
// K1 (every submitted splat): cull, depth key (culled -> KEY_CULLED), packed tile rectangle (trects, may be NULL), exact
// pixel rectangle (rects: every splat if rects_all, else only the "wide" ones the packed form cannot hold),
// *n_visible += V, and (bucket_hist != NULL) the DEPTH_BUCKETS-bin histogram of depth_bucket(key).
void launch_project(const FrameConsts& fc, const PackedSplats& ps, int64_t n,
                    uint32_t* keys, uint2* rects, int rects_all, uint32_t* trects,
                    unsigned long long* n_visible, DepthBuckets db, uint32_t* bucket_hist, const uint32_t* owned_rows,
                    cudaStream_t s);
// owned_rows (exact K1 / K2 launchers): NULL for a whole frame; for a row-partitioned frame owned_rows[y] = number of tile
// rows < y this rank owns (tiles_y + 1 entries), so the ownership cull is two look-ups.

// ---- bounded K1 over spatial cells (GSB_OPT_LAZY_PROJECT, the production path; see project.cu) ----------------------
constexpr int CELL = 256;                       // splats per cell = threads of the per-cell selection CTA
struct CellBox { float lo[3], hi[3]; float rr_max, lam_max; };   // 32 B: box of the positions, largest discard radius
                                                                 // (-1: nothing visible), largest eigenvalue bound (inf: unbounded)
// pack time: Morton keys of the packed positions (bbox = min xyz, max xyz of the set); after sorting (mkeys, idx) the
// values are the original indices in cell order (orig); geomA_p = geomA in that order
void launch_morton(const float4* geomA, int64_t n, const float bbox[6], uint32_t* mkeys, uint32_t* idx, cudaStream_t s);
void launch_gather_geom(const uint32_t* orig, const float4* geomA, int64_t n, float4* geomA_p, cudaStream_t s);
// whenever lam changes (object matrix): lam in cell order + the cell boxes
void launch_cell_build(const float4* geomA_p, const uint32_t* orig, const float* lam, int64_t n, float* lam_p, CellBox* cells,
                       cudaStream_t s);
// per frame: views[c] = (key_lo, key_hi, tx0 | tx1 << 16, ty0 | ty1 << 16), key_lo > key_hi = invisible; the chunk plan's
// bucket histogram (may be NULL); *n_visible += members of the visible cells (an upper bound of V)
void launch_cell_project(const FrameConsts& fc, const CellBox* cells, int64_t n, DepthBuckets db, uint32_t* bucket_hist,
                         uint4* views, unsigned long long* n_visible, cudaStream_t s);
// per depth chunk: the cells that can hold a live splat of the chunk -> sel_cells[0 .. *n_sel) (*n_sel zero before)
void launch_cell_select(const uint4* views, int64_t n, const ChunkPlan* plan, int chunk, const FrameConsts& fc,
                        const uint32_t* sat, uint32_t* sel_cells, uint32_t* n_sel, cudaStream_t s);
// per depth chunk: the live splats of the selected cells -> (keys_out, vals_out = ORIGINAL index), unordered;
// *l_total += their number, *d_total += an upper bound of their tile instances; sort_hist (zero before) receives the digit
// histograms of the depth sort described by sp / key_min / key_span
void launch_splat_select(const FrameConsts& fc, const float4* geomA_p, const float* lam_p, const uint32_t* orig, int64_t n,
                         const uint32_t* sel_cells, const uint32_t* n_sel, const ChunkPlan* plan, int chunk,
                         const uint32_t* sat, uint32_t* keys_out, uint32_t* vals_out,
                         unsigned long long* l_total, unsigned long long* d_total,
                         const SortPlan& sp, uint32_t key_min, uint32_t key_span, uint32_t* sort_hist, cudaStream_t s);
// debug view: the bound of every packed splat, by original index (same bound_one() the selection evaluates)
void launch_project_bound_debug(const FrameConsts& fc, const float4* geomA_p, const float* lam_p, const uint32_t* orig, int64_t n,
                                uint32_t* keys, uint32_t* trects, cudaStream_t s);
// chunk plan from the bucket histogram: chunk c (< nchunks - 1) ends at the first bucket whose exclusive count reaches
// V * (2^(c+1) - 1) / 2^shift (the first chunk holds V / 2^shift splats, every further one doubles; the last takes the rest)
// and the bucket boundaries are turned into key boundaries (plan->key_lo: the smallest key whose bucket belongs to the
// chunk, found by bisection over the monotone map key -> bucket -> chunk), so membership is two integer compares
void launch_choose_chunks(const uint32_t* bucket_hist, int nchunks, int shift, DepthBuckets db, ChunkPlan* plan, cudaStream_t s);
// K2 (only splats that reach a live tile, in depth order): gather the splat's 128-byte line, redo the projection,
// evaluate SH, write the 48-byte record of live rank j to recs[j], its tile rectangle (tx0 | tx1 << 16, ty0 | ty1 << 16)
// to tile_rects[j] and the number of live tiles it touches to counts[j] (sat: launch_live_sat, NULL = every tile live)
void launch_records(const FrameConsts& fc, const PackedSplats& ps, const uint32_t* live_splats, int64_t n_live,
                    const uint32_t* sat, Record* recs, uint2* tile_rects, uint32_t* counts, float* zdepth,
                    const uint32_t* owned_rows, cudaStream_t s);
// zdepth (may be NULL): window depth of live rank j (scene-depth occlusion)

// binning.cu
// tile_done: saturation flags, one bit per tile, tile (tx, ty) at bit tx & 31 of word ty * done_words_per_row(tiles_x) +
// tx / 32 (NULL = none).  A tile is live if this rank owns its row and it is not saturated.  sat: (tiles_y + 1) x
// (tiles_x + 1) uint32, sat[y][x] = live tiles in rows < y and columns < x, so the live tiles of any tile rectangle are
// four look-ups.  One small CTA per depth chunk.
void launch_live_sat(FrameConsts fc, const uint32_t* tile_done, uint32_t* sat, cudaStream_t s);
// live selection of one depth chunk over the submitted splats: elements whose depth key belongs to the chunk
// (plan->key_lo[chunk] <= key < plan->key_lo[chunk + 1]; plan NULL = every visible splat) and that touch a live tile are
// compacted, order preserving, into (keys_out, vals_out = splat index); *l_total = their number, *d_total = the
// instances they will emit.  trects = K1's packed tile rectangles (NULL: exact rectangles in rects); stage_k/v:
// select_stage_elems(n) uint32 each (CTA-local runs before the gather; may be the second halves of the sort's ping-pong
// buffers); scratch: select_scratch_bytes(n).  Three launches, no spin-waits.
size_t select_scratch_bytes(int64_t n);
size_t select_stage_elems(int64_t n);
void launch_select_live(const uint32_t* keys, const uint32_t* trects, const uint2* rects, int64_t n,
                        const ChunkPlan* plan, int chunk,
                        FrameConsts fc, const uint32_t* sat, uint32_t* keys_out, uint32_t* vals_out,
                        uint32_t* stage_k, uint32_t* stage_v,
                        void* scratch, unsigned long long* l_total, unsigned long long* d_total,
                        cudaStream_t s);
// instance (tile id, live rank) pairs at offsets[k] .., rows ascending then columns ascending, live tiles only;
// tile_rects = K2's rectangles by live rank; offsets = exclusive scan of K2's counts, *total = its grand total (device)
// tile_hist (may be NULL; zero before): receives the digit histograms of the tile partition described by tile_plan
void launch_emit(const uint2* tile_rects, const uint32_t* offsets,
                 const unsigned long long* total, int64_t n, FrameConsts fc, const uint32_t* tile_done,
                 uint32_t* inst_keys, uint32_t* inst_vals, const SortPlan& tile_plan, uint32_t* tile_hist, cudaStream_t s);
// d_max: host-side upper bound (grid size), d_dev: the exact count on the device (NULL: d_max is exact)
void launch_tile_ranges(const uint32_t* sorted_tile_ids, uint64_t d_max, const unsigned long long* d_dev, uint2* ranges,
                        int num_tiles, cudaStream_t s);
// after the depth sort: runs of equal keys are put in ascending value (= splat index) order; keys stay where they are
void launch_tie_fix(const uint32_t* keys_sorted, const uint32_t* vals_sorted, uint64_t l_max, const unsigned long long* l_dev,
                    uint32_t* vals_out, cudaStream_t s);
// debug views (GSB_OPT_KEEP_INTERMEDIATES): records by splat index, instances as splat indices
void launch_debug_views(const Record* recs, const uint32_t* live_splats, int64_t n_live, Record* recs_by_splat,
                        const uint32_t* inst_refs, uint64_t d, uint32_t* inst_splats, cudaStream_t s);

// ingest.cu: raw fp32 point attributes -> the arrays registerUpdate receives (NULL source = default)
// activation: 0 = the attributes are Houdini's (the reference's contract), 1 = raw INRIA PLY columns, activated here (8f-2)
void launch_ingest_core(const float* P, const float* Cd, const float* alpha, const float* scale, const float* orient,
                        int64_t n, int activation, float* pos_out, uint16_t* cd_out, float* alpha_out, uint16_t* scale_out,
                        uint16_t* orient_out, cudaStream_t s);
// src = [n][len][3] (planar = 0, `sh_coefficients`) or [15][n][3] (planar = 1, `sh1`..`sh15`)
void launch_ingest_sh_vec3(const float* src, int64_t n, int len, int planar, uint16_t* shx, uint16_t* shy, uint16_t* shz,
                           cudaStream_t s);
// rest = [45][n] (`f_rest_0`..`f_rest_44`)
void launch_ingest_sh_rest(const float* rest, int64_t n, uint16_t* shx, uint16_t* shy, uint16_t* shz, cudaStream_t s);

// wire.cu: the reference's wireframe vertex shader, once per splat: verts[8 n] = gl_Position of the outline's
// 8 line vertices, colors[8 n][3] = Cd (may be NULL); and an overlay of the outlines into an RGBA32F frame (owner: width x
// height u64 scratch; the nearest splat wins per pixel)
void launch_wire_vertices(const FrameConsts& fc, const float* pos, const uint16_t* cd, const uint16_t* scale,
                          const uint16_t* orient, int64_t n, float4* verts, float* colors, cudaStream_t s);
void launch_wire_overlay(const float4* verts, const uint16_t* cd, int64_t n, int width, int height,
                         unsigned long long* owner, float4* rgba, cudaStream_t s);

// blend.cu
// One depth chunk.  first: pixel state starts at (0,0,0,T=1), otherwise it is reloaded from fb, which between
// chunks holds (C, T).  A tile whose pixels are all saturated is finalised to (C, 1-T) and flagged in tile_done (bit map);
// last: every remaining tile is finalised.  Finalised tiles are stored to fb_final (NULL = fb; may be peer memory).  *done_tiles counts the tiles flagged so far (early termination).
// zdepth / scene_depth (both NULL unless fc.depth_func != 0): window depth per live rank and the scene's depth buffer;
// a fragment whose depth fails fc.depth_func against the scene depth at its pixel is dropped.
void launch_blend(const Record* recs, const uint32_t* inst_vals, const uint2* ranges, float4* fb, float4* fb_final,
                  FrameConsts fc, int first, int last, uint32_t* tile_done, uint32_t* tile_consumed,
                  unsigned long long* consumed_total, unsigned long long* done_tiles,
                  const float* zdepth, const float* scene_depth, cudaStream_t s);

// Packed tile rectangle carried through the depth sort: tx0:9 | ty0:9 | (tx1-tx0):7 | (ty1-ty0):7.  Extents of 127 tiles
// or more saturate to 127 = "wide: read the exact pixel rectangle by splat index".  0xFFFFFFFF = culled.
// Valid while the screen has at most 512 x 512 tiles (8192 px); larger screens use the exact rectangles only.
constexpr uint32_t TRECT_CULLED = 0xFFFFFFFFu;
__host__ __device__ __forceinline__ uint32_t pack_trect(int tx0, int tx1, int ty0, int ty1)
{
    const int w = tx1 - tx0, h = ty1 - ty0;
    return (uint32_t)tx0 | ((uint32_t)ty0 << 9) | ((uint32_t)(w > 127 ? 127 : w) << 18) | ((uint32_t)(h > 127 ? 127 : h) << 25);
}

// number of live tiles of the tile rectangle [tx0, tx1] x [ty0, ty1] (live = row owned by this rank, tile not yet
// saturated): four look-ups in the summed-area table of the live map (sat[y * (tiles_x + 1) + x] = live tiles in rows
// < y, columns < x), no loop and no divergence; sat == NULL means every tile is live (single rank, first depth chunk)
__device__ __forceinline__ uint32_t live_tiles(int tx0, int tx1, int ty0, int ty1, int tiles_x, const uint32_t* __restrict__ sat)
{
    if (!sat) return (uint32_t)((tx1 - tx0 + 1) * (ty1 - ty0 + 1));
    const int st = tiles_x + 1;
    const uint32_t* r0 = sat + ty0 * st;
    const uint32_t* r1 = sat + (ty1 + 1) * st;
    return (__ldg(r1 + tx1 + 1) - __ldg(r0 + tx1 + 1)) - (__ldg(r1 + tx0) - __ldg(r0 + tx0));
}

// words per tile row of the saturation bit map (tile_done)
__host__ __device__ __forceinline__ int done_words_per_row(int tiles_x) { return (tiles_x + 31) >> 5; }

// Saturation flags (tile_done): one BIT per tile, rows padded to whole 32-bit words (done_words_per_row), set by the
// blend of an earlier depth chunk.  The live tiles of a splat's row segment [tx0, tx1] are then a mask and a popcount
// per word (one word for almost every splat) instead of a load per tile.
__device__ __forceinline__ uint32_t live_word(const uint32_t* __restrict__ done, int wpr, int ty, int w, int tx0, int tx1)
{
    uint32_t m = 0xffffffffu;
    if (w == (tx0 >> 5)) m &= 0xffffffffu << (tx0 & 31);
    if (w == (tx1 >> 5)) m &= 0xffffffffu >> (31 - (tx1 & 31));
    return m & ~__ldg(done + ty * wpr + w);
}

// tile-row ownership rule shared by every kernel
__host__ __device__ __forceinline__ bool owns_row(int ty, int rank, int world, int group)
{
    return world <= 1 || ((ty / group) % world) == rank;
}

}  // namespace gsb

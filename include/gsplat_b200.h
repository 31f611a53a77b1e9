/* gsplat_b200.h — C ABI of libgsplat_b200.so: the B200-native replacement for the
 * splat -> framebuffer hot path of rubendhz/houdini-gsplat-renderer (plugin v1.4.1).
 *
 * Every entry point below replaces one member of the reference's `GSplatRenderer` singleton
 * (reference paths relative to /root/reference/gsplat_plugin); the HDK shim keeps the reference's
 * class names and forwards 1:1 (INTEGRATION.md shows the binding).  Plain pointers and sizes only:
 * no C++/torch types cross this boundary, nothing throws, every call returns a status code and
 * leaves a message for gsb_last_error().  One caller thread per context (the reference is not
 * thread-safe either, include/GSplatRenderer.h:29-32); one context per GPU / process.
 *
 * Conventions: matrices are 16 floats, column-major, column-vector convention (OpenGL), i.e.
 * exactly the glH_* builtins the reference's GLSL reads (shaders/GSplatShaderSource.h:153-159).
 * Framebuffers are RGBA32F, premultiplied alpha, row 0 = bottom scanline (GL texture order).
 * Half-precision inputs are IEEE binary16 bit patterns in uint16_t (UT_Vector3H / fpreal16).
 */
#ifndef GSPLAT_B200_H
#define GSPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GSB_API __attribute__((visibility("default")))
#else
#define GSB_API
#endif

#define GSB_ABI_VERSION 3
#define GSB_TILE 16                       /* screen tile edge in pixels */
#define GSB_REFERENCE_SPLAT_CAP 8388607   /* GSPLAT_COUNT_MAX - 1, include/GSplatRenderer.h:26, src/GSplatRenderer.C:336 */
#define GSB_ID_MAX 128                    /* bytes for a registry id string incl. NUL */

typedef struct gsb_context gsb_context;

enum gsb_status {
    GSB_OK = 0,
    GSB_ERR_INVALID = 1,     /* bad argument */
    GSB_ERR_CUDA = 2,        /* CUDA runtime error (message in gsb_last_error) */
    GSB_ERR_NOMEM = 3,
    GSB_ERR_NOT_FOUND = 4,   /* unknown registry id */
    GSB_ERR_LIMIT = 5        /* more than 2^31 tile instances, screen > 65535 px, ... */
};

/* Identity of one GSplat primitive, the fields the reference hashes into its registry id
 * "<gdp ptr hex>__<vtx0>__<v0>_<v1>_<v2>_<v3>" (src/GSplatRenderer.C:241-243). */
typedef struct gsb_prim_key {
    uint64_t gdp;            /* GU_Detail pointer value */
    int64_t  vtx0;           /* GA_Offset of the prim's first vertex */
    int64_t  version[4];     /* RE_CacheVersion elements 0..3 */
} gsb_prim_key;

/* Per-redraw inputs the reference reads from RE_Render and the glH_* uniforms
 * (src/GSplatRenderer.C:558-562; shaders/GSplatShaderSource.h:153-159). */
typedef struct gsb_frame {
    float   view[16];        /* glH_ViewMatrix      (also r->getMatrix(), R.C:558) */
    float   proj[16];        /* glH_ProjectMatrix   */
    float   object[16];      /* glH_ObjectMatrix    */
    float   inv_object[16];  /* glH_InvObjectMatrix */
    float   obj_view[16];    /* glH_ObjViewMatrix   */
    int32_t width, height;   /* glH_ScreenSize      */
    int32_t is_object_level; /* DM_SceneHookData::disp_options->isObjectLevel() (DM_GSplatHook.C:34) */
    int32_t row_rank;        /* multi-GPU: this context blends tile rows ty with (ty / row_group) % row_world == row_rank */
    int32_t row_world;       /* 1 = whole frame */
    int32_t row_group;       /* tile rows per interleaved band; 0 or 1 = single rows.  Wider bands duplicate fewer splats */
    /* Scene-depth occlusion.  The reference draws with the depth test on and depth writes off
     * (src/GSplatRenderer.C:608-610) and gives every vertex of a splat's quad the CENTRE's clip z and w
     * (shaders/GSplatShaderSource.h:278-282), so a splat's fragments all carry one window depth
     * zw = clip.z/clip.w * (far-near)/2 + (far+near)/2; a fragment is kept iff zw passes depth_func against the scene
     * depth at its pixel.  Nothing is written to the depth buffer. */
    int32_t depth_func;        /* enum gsb_depth_func; GSB_DEPTH_NONE (0) = no occlusion */
    uint32_t gl_depth_texture; /* optional: an R32F / DEPTH_COMPONENT32F GL_TEXTURE_2D of the frame size holding the scene's
                                  window depth; mapped read-only through CUDA<->GL interop (needs a current GL context).
                                  Used when scene_depth is NULL */
    float   depth_range[2];    /* glDepthRange near, far (glH_DepthRange, GSplatShaderSource.h:158); {0,0} is read as {0,1} */
    const void* scene_depth;   /* device pointer, width*height floats, window depth in [0,1], row 0 = bottom scanline */
} gsb_frame;

enum gsb_depth_func {
    GSB_DEPTH_NONE = 0,
    GSB_DEPTH_LESS = 1,        /* GL_LESS   (OpenGL's default) */
    GSB_DEPTH_LEQUAL = 2       /* GL_LEQUAL (what Houdini's viewport passes use) */
};

/* Where the finished frame goes.  All optional; with everything NULL the frame stays in the
 * library-owned device buffer (gsb_device_framebuffer). */
typedef struct gsb_target {
    void*    device_rgba;    /* caller-owned device buffer, width*height*16 bytes; NULL = library buffer */
    void*    host_rgba;      /* if non-NULL the frame is delivered here (D2H inside the call, call returns when done); see
                                GSB_OPT_HOST_DIRECT for pinned memory.  Row-partitioned frames (row_world > 1): a pinned target
                                receives only the tiles this rank owns (the rest is left untouched, so every rank can write
                                into one shared host frame); a pageable target receives the whole local frame (zeros elsewhere) */
    uint32_t gl_texture;     /* CUDA<->GL interop target: an RGBA32F GL_TEXTURE_2D of the frame size; the frame is copied into it
                                device->device (cudaGraphicsGLRegisterImage).  Needs a current GL context on the calling thread
                                (Houdini's main thread); without one the call fails with GSB_ERR_CUDA.  0 = none */
    uint32_t flags;          /* reserved, 0 */
    void*    final_rgba;     /* optional: finished tiles are stored HERE instead of device_rgba (which then only holds the
                                per-chunk blend state).  May be PEER memory of another GPU (gsb_ipc_open): the blend kernel
                                writes this rank's tile rows straight into the display GPU's frame over NVLink, so a
                                row-partitioned frame needs no gather pass.  Not readable until every rank has synchronised. */
} gsb_target;

typedef struct gsb_stats {
    int64_t n_submitted;     /* N: splats in the packed active set */
    int64_t n_visible;       /* V: survive cull; an upper bound with GSB_OPT_LAZY_PROJECT (see there) */
    int64_t n_instances;     /* D: tile instances emitted (summed over depth chunks; saturated tiles receive none) */
    int64_t n_consumed;      /* D_c: instances traversed before every pixel of their tile saturated */
    int32_t rendered;        /* 1 if the last gsb_render drew, 0 if it early-returned like R.C:536-549 */
    int32_t repacked;        /* 1 if the last gsb_generate_render_geometry rebuilt the packed set */
    int32_t sh_order_used;   /* GSplatShOrder actually bound (0 if no SH data, R.C:623,628) */
    int32_t width, height;
    int32_t tiles_x, tiles_y;
    int32_t launches;        /* kernels launched by the last gsb_render */
    int32_t depth_chunks;    /* depth chunks the last frame was binned/blended in */
    int32_t warnings;        /* enum gsb_warning bits raised by the last gsb_render (the reference logs these, R.C:565-581) */
    float   camera[3];       /* WorldSpaceCameraPos used for keys and SH */
    float   origin[3];       /* GSplatOrigin (mean of barycentres, R.C:403-418) */
    /* device time per stage of the last gsb_render, CUDA events on the library stream;
     * valid when GSB_OPT_STAGE_TIMING is on (gsb_get_stats synchronises) */
    float   ms_project, ms_sort, ms_bin, ms_blend, ms_copy, ms_total;
    float   ms_records;      /* K2 (records + SH of the live splats); ms_sort = chunk partition + live selection + live depth sort */
    float   reserved1;
    int64_t n_live;          /* L: splats that reached a live (owned, un-saturated) tile, summed over depth chunks: only these
                                are depth-sorted, given a 2-D record (SH evaluated) and binned */
} gsb_stats;

enum gsb_warning {
    GSB_WARN_OBJECT_LEVEL = 1    /* gsb_frame.is_object_level: OBJ-level transforms other than identity are not supported by the
                                    reference's formulas (R.C:565-577), which this library reproduces; raised on the
                                    first OBJ-level frame after a SOP-level one, like the reference's one-time log line */
};

enum gsb_option {
    GSB_OPT_SPLAT_CAP = 1,       /* max splats packed; default GSB_REFERENCE_SPLAT_CAP; 0 = unlimited */
    GSB_OPT_EPS_T = 2,           /* transmittance early-out threshold; default 1e-5; 0 = never stop (reference) */
    GSB_OPT_STAGE_TIMING = 3,    /* record per-stage CUDA events (default 0) */
    GSB_OPT_KEEP_INTERMEDIATES = 4, /* keep unsorted keys etc. for gsb_debug_fetch (default 0) */
    GSB_OPT_COMPACT = 6,         /* accepted for ABI compatibility, no effect: the live splats of every depth chunk are always
                                    compacted before their depth sort */
    GSB_OPT_CHUNK_SHIFT = 7,     /* the first depth chunk holds V / 2^shift visible splats, each further chunk doubles, the last
                                    takes the rest; 0 = auto.  Any value gives the same frame */
    GSB_OPT_HOST_DIRECT = 8,     /* 1 (default): when gsb_target.host_rgba is pinned host memory the device can address
                                    (cudaHostAlloc / cudaHostRegister), the blend kernel stores finished tiles straight into it
                                    over PCIe while later depth chunks are still being binned (no separate D2H pass);
                                    pageable memory always takes the staged copy.  0: always cudaMemcpyAsync.  Same bytes */
    GSB_OPT_LAZY_PROJECT = 9,    /* 1 (default): K1 computes, for every submitted splat, the exact cheap culls, the exact depth key and a
                                    conservative bound of its tile rectangle (1/5 of the exact kernel's instructions); the exact
                                    projection runs only in K2 for the splats a depth chunk selects.  Same frame.  gsb_stats then
                                    reports upper bounds: n_visible = splats that pass the cheap culls with a non-empty bound,
                                    n_live = splats selected by the bound.  0: exact projection of every splat in K1 (exact
                                    n_visible / n_live; always used with GSB_OPT_KEEP_INTERMEDIATES and on screens wider than
                                    8192 px) */
    GSB_OPT_DEPTH_CHUNKS = 5     /* bin+blend in this many front-to-back depth chunks, skipping saturated tiles in later
                                    chunks; 1 = single pass (full tile lists, what the parity tests fetch); 0 = auto */
};

enum gsb_debug_buffer {
    GSB_DBG_KEYS_UNSORTED = 0,   /* uint32[N]  depth keys in submission order */
    GSB_DBG_ORDER = 1,           /* uint32[L]  splat index by depth rank: the live splats of the last depth chunk (with
                                    GSB_OPT_DEPTH_CHUNKS = 1: every visible splat, i.e. the global depth order) */
    GSB_DBG_RECORDS = 2,         /* 48 B x N   2-D records by splat index (valid for the live splats of the last chunk;
                                    needs KEEP_INTERMEDIATES) */
    GSB_DBG_RECTS = 3,           /* uint16[4] x N  inclusive pixel rectangle x0,x1,y0,y1 (x0>x1 = culled; needs KEEP_INTERMEDIATES) */
    GSB_DBG_TILE_RANGES = 4,     /* uint32[2] x tiles  [start,end) into the instance list (last depth chunk) */
    GSB_DBG_INSTANCES = 5,       /* uint32[D]  splat index per tile instance, tile-major, depth order inside (last chunk;
                                    needs KEEP_INTERMEDIATES) */
    GSB_DBG_FRAMEBUFFER = 6,     /* float[4] x W x H */
    GSB_DBG_KEYS_SORTED = 7,     /* uint32[L]  keys in the order of GSB_DBG_ORDER */
    GSB_DBG_TILE_CONSUMED = 8,   /* uint32 x tiles  instances traversed per tile */
    GSB_DBG_TRECTS = 9           /* uint32[N]  K1's packed tile rectangle per splat, tx0:9 | ty0:9 | (tx1-tx0):7 | (ty1-ty0):7
                                    (extent 127 = "127 or more"), 0xFFFFFFFF = culled: exact, or with GSB_OPT_LAZY_PROJECT the
                                    conservative bound (a superset of the exact rectangle); screens up to 8192 px */
};

/* ---- lifetime -------------------------------------------------------------------------- */
GSB_API int  gsb_abi_version(void);
GSB_API int  gsb_create(int cuda_device, gsb_context** out);   /* replaces GSplatRenderer::getInstance(), R.h:29-32 */
GSB_API int  gsb_destroy(gsb_context* ctx);
GSB_API const char* gsb_last_error(void);                      /* thread-local, never NULL */

/* ---- the reference's public surface, include/GSplatRenderer.h:34-56 ---------------------- */

/* GSplatRenderer::registerUpdate (R.h:34-47, R.C:218-291).  Copies the arrays to the device
 * immediately (the reference keeps raw pointers, R.C:277-284; this ABI never borrows host memory).
 * Evicts entries of the same gdp with a different version (R.C:246-265).  shx/shy/shz may be NULL
 * together (no SH data).  id_out receives the registry id string (GSB_ID_MAX bytes). */
GSB_API int gsb_register_update(gsb_context* ctx, const gsb_prim_key* key, int64_t splat_count,
                        const float origin[3],
                        const float*    pos,        /* [count][3]  UT_Vector3Array  */
                        const uint16_t* cd_h,       /* [count][3]  UT_Vector3HArray */
                        const float*    alpha,      /* [count]     UT_FloatArray    */
                        const uint16_t* scale_h,    /* [count][3]  UT_Vector3HArray */
                        const uint16_t* orient_h,   /* [count][4]  UT_Vector4HArray (x,y,z,w) */
                        const uint16_t* shx_h,      /* [count][16] MyUT_Matrix4HArray, coeff j at (j/4, j%4) */
                        const uint16_t* shy_h,
                        const uint16_t* shz_h,
                        char* id_out);

GSB_API int gsb_include_in_render_pass(gsb_context* ctx, const char* id);              /* R.C:313-320 */
GSB_API int gsb_flush_entries_for_matching_detail(gsb_context* ctx, const char* id);   /* R.C:293-311 */
GSB_API int gsb_generate_render_geometry(gsb_context* ctx);                            /* R.C:322-532 */
GSB_API int gsb_render(gsb_context* ctx, const gsb_frame* frame, const gsb_target* target); /* R.C:534-658 */
GSB_API int gsb_post_render(gsb_context* ctx);                                         /* R.C:660-678 */
GSB_API int gsb_set_rendering_enabled(gsb_context* ctx, int enabled);                  /* R.C:680-683 */
GSB_API int gsb_set_explicit_camera_pos(gsb_context* ctx, const float pos[3]);         /* R.C:685-689 */
GSB_API int gsb_set_spherical_harmonics_order(gsb_context* ctx, int sh_order);         /* R.C:691-694 */

/* ---- caller side of the boundary: GR_PrimGsplat::update (src/GR_GSplat.C:191-458), SURVEY.md §8 f-1 ------- */

/* Raw fp32 point attributes of one GSplat primitive, as Houdini stores them.  NULL = attribute absent. */
typedef struct gsb_raw_attributes {
    int64_t      count;                /* points of the prim (vertex i <-> point i, GEO_GSplat.C:413-431) */
    const float* P;                    /* [count][3]  required */
    const float* Cd;                   /* [count][3]  absent -> (0,0,0)            GR.C:309 */
    const float* opacity;              /* [count]                                   GR.C:240 */
    const float* Alpha;                /* [count]     preferred over opacity when both exist; neither -> 1   GR.C:241-257,310 */
    const float* scale;                /* [count][3]  absent -> (1,1,1)            GR.C:311 */
    const float* orient;               /* [count][4]  (x,y,z,w), absent -> (0,0,0,1)  GR.C:312 */
    const float* sh_coefficients;      /* [count][sh_coefficients_len][3]  vec3-array attribute, tried first  GR.C:93-113 */
    int32_t      sh_coefficients_len;  /* vec3 entries per point; entries >= 15 are ignored */
    int32_t      activation;           /* enum gsb_activation.  GSB_ACT_INRIA: the arrays are the RAW columns of an INRIA 3DGS .ply
                                           — Cd = f_dc_0..2, opacity = the logit, scale = log scales, orient =
                                          (rot_0, rot_1, rot_2, rot_3) = (w, x, y, z), f_rest as is — and the conversion of the
                                          example scene's wrangles (Cd = SH_C0 f_dc + 0.5, sigmoid, exp, quaternion reorder +
                                          normalise) runs on the GPU inside the ingestion kernel */
    const float* sh[15];               /* sh1..sh15, each [count][3]; used if sh_coefficients is absent and all 15 exist  GR.C:115-128,160-171 */
    const float* f_rest[45];           /* f_rest_0..44, each [count]; used last, all 45 must exist; coefficient j = (f_rest_j, f_rest_j+15, f_rest_j+30)  GR.C:130-143,173-184,357-366 */
    int32_t      has_sh_order;         /* detail attribute gsplat__sh_order present  GR.C:284-289 */
    int32_t      sh_order;
    int32_t      has_explicit_camera;  /* detail attribute gsplat__explicit_camera_pos present  GR.C:277-282 */
    float        explicit_camera[3];
} gsb_raw_attributes;

enum gsb_activation { GSB_ACT_NONE = 0, GSB_ACT_INRIA = 1 };

/* What update() leaves in the GR primitive for its render() to push every pass (GR.C:438-457, 485-492). */
typedef struct gsb_update_result {
    char    id[GSB_ID_MAX];            /* registry id (same text as registerUpdate's) */
    int32_t sh_order;                  /* 3 by default; attribute value if in 0..3; 0 (and an error logged) otherwise */
    int32_t sh_order_invalid;          /* 1 if gsplat__sh_order was outside 0..3 */
    int32_t sh_data_found;             /* 1 if one of the three SH encodings was complete */
    int32_t set_explicit_camera;
    float   explicit_camera[3];
    float   barycentre[3];             /* GEO_PrimGsplat::baryCenter: sequential fp32 sum / count (GEO_GSplat.C:338-351) */
} gsb_update_result;

/* GR_PrimGsplat::update: extracts + quantises on the GPU and registers the prim (as registerUpdate would). */
GSB_API int gsb_update_from_attributes(gsb_context* ctx, const gsb_prim_key* key, const gsb_raw_attributes* attrs,
                                       gsb_update_result* out);

/* ---- wireframe / selection overlay: GR_PrimGsplat::render's wire pass (src/GR_GSplat.C:474-483), SURVEY.md §8 f-4 ------- */

/* The reference keeps a VBO with 8 copies of every splat attribute (GR_GSplat.C:376-421) and runs its wire vertex shader
 * (shaders/GSplatShaderSource.h:22-90) on all of them, drawn as RE_PRIM_LINES: the outline of each splat's +-2 sigma quad,
 * colour = Cd.  Here the vertex shader runs once per splat on the GPU: vertices[8 count][4] = gl_Position of line vertices
 * 0..7 (edges (0,1) (2,3) (4,5) (6,7)), colors[8 count][3] = Cd.  Everything optional; with all pointers NULL the vertices
 * stay in a library buffer (gsb_wire_device_vertices).  overlay_rgba: the outlines are also rasterised into that RGBA32F
 * device frame (frame.width x frame.height, row 0 = bottom): per pixel the nearest splat wins, colour (Cd, 1), pixels no
 * line touches are left as they are. */
typedef struct gsb_wire_target {
    void* device_vertices;   /* caller-owned device buffer, 8 * count * 16 bytes (e.g. a mapped GL buffer); NULL = library buffer */
    void* device_colors;     /* caller-owned device buffer, 8 * count * 12 bytes; NULL = not produced unless host_colors is set */
    void* host_vertices;     /* optional host copies (the call returns when they are there) */
    void* host_colors;
    void* overlay_rgba;      /* optional device frame the outlines are drawn into */
    void* overlay_host_rgba; /* optional host copy of overlay_rgba after the overlay */
} gsb_wire_target;
GSB_API int   gsb_render_wireframe(gsb_context* ctx, const char* id, const gsb_frame* frame, const gsb_wire_target* target);
GSB_API void* gsb_wire_device_vertices(gsb_context* ctx);

/* ---- additions the reference has no equivalent for -------------------------------------- */
GSB_API int   gsb_set_option(gsb_context* ctx, int option, double value);
GSB_API int   gsb_get_stats(gsb_context* ctx, gsb_stats* out);
GSB_API int   gsb_set_stream(gsb_context* ctx, void* cuda_stream);   /* run on a caller stream (NULL = library stream) */
GSB_API int   gsb_synchronize(gsb_context* ctx);
GSB_API void* gsb_device_framebuffer(gsb_context* ctx);              /* device pointer of the last library-owned frame */
GSB_API int   gsb_registry_size(gsb_context* ctx);                   /* live registry entries */

/* Multi-GPU frame sharing (one process per GPU).  The display rank allocates a frame and exports a CUDA IPC handle; the
 * other ranks open it and pass the mapped pointer as gsb_target.final_rgba. */
#define GSB_IPC_HANDLE_BYTES 64
GSB_API int gsb_ipc_export_frame(gsb_context* ctx, int32_t width, int32_t height, unsigned char handle_out[GSB_IPC_HANDLE_BYTES],
                                 void** local_ptr_out);
GSB_API int gsb_ipc_open(gsb_context* ctx, const unsigned char handle[GSB_IPC_HANDLE_BYTES], void** peer_ptr_out);
GSB_API int gsb_ipc_close(gsb_context* ctx, void* peer_ptr);
GSB_API int gsb_copy_to_host(gsb_context* ctx, const void* device_ptr, void* host_ptr, uint64_t bytes);  /* stream-ordered, synchronous */
/* Page-lock a host range the caller owns (e.g. a frame in POSIX shared memory mapped by every rank of a row-partitioned
 * job) and map it into this context's device: cudaHostRegister(Portable | Mapped).  Passed as gsb_target.host_rgba it
 * then receives this rank's finished tiles straight from the blend kernel over the rank's own PCIe link
 * (GSB_OPT_HOST_DIRECT); tile rows the rank does not own are left untouched.  ptr and bytes must be page aligned. */
GSB_API int gsb_host_register(gsb_context* ctx, void* host_ptr, uint64_t bytes);
GSB_API int gsb_host_unregister(gsb_context* ctx, void* host_ptr);

/* Test hook: the arrays a registered prim holds (what registerUpdate received / update quantised). */
enum gsb_entry_array { GSB_ENT_POS = 0, GSB_ENT_CD = 1, GSB_ENT_ALPHA = 2, GSB_ENT_SCALE = 3, GSB_ENT_ORIENT = 4,
                       GSB_ENT_SHX = 5, GSB_ENT_SHY = 6, GSB_ENT_SHZ = 7 };
GSB_API int gsb_debug_fetch_entry(gsb_context* ctx, const char* id, int which, void* dst, uint64_t dst_bytes, uint64_t* bytes_needed);

/* Test hooks: copy an intermediate device buffer to the host.  *bytes_needed is always set;
 * the copy happens only if dst != NULL and dst_bytes >= needed. */
GSB_API int gsb_debug_fetch(gsb_context* ctx, int which, void* dst, uint64_t dst_bytes, uint64_t* bytes_needed);
/* Stand-alone exercisers for the device primitives (stable LSD radix sort on [begin_bit,end_bit),
 * exclusive scan), host buffers in and out. */
GSB_API int gsb_debug_sort_pairs(gsb_context* ctx, const uint32_t* keys, const uint32_t* vals, uint64_t n,
                         int begin_bit, int end_bit, uint32_t* keys_out, uint32_t* vals_out);
GSB_API int gsb_debug_exclusive_scan(gsb_context* ctx, const uint32_t* in, uint64_t n, uint32_t* out, uint64_t* total);

#ifdef __cplusplus
}
#endif
#endif /* GSPLAT_B200_H */

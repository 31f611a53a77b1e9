// renderer.cu — host side of libgsplat_b200.so: the C ABI of include/gsplat_b200.h and a C++ mirror
// of the reference's GSplatRenderer registry / state machine
// (/root/reference/gsplat_plugin/include/GSplatRenderer.h:29-131, src/GSplatRenderer.C:141-153,
// 218-320, 322-378, 403-418, 534-563, 660-694).  What the reference does with GL textures, a CPU
// argsort and one instanced draw is done here with CUDA launches on one stream:
//   pack (on active-set change) | K1 cull+key+tile rect for every splat -> chunk plan | per depth chunk: live map SAT ->
//   live selection -> depth radix sort -> K2 records+SH+tile rects+counts -> scan -> emit -> tile radix partition ->
//   tile ranges -> blend (finished tiles straight to the device / pinned host / peer-GPU frame) | optional D2H, GL interop.
// No CPU fallback exists: every entry point fails with GSB_ERR_CUDA if the device is unusable.
#include "common.cuh"
#include "host_pool.h"
#include "../../include/gsplat_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

using namespace gsb;

// CUDA<->OpenGL interop entry point of the CUDA runtime (cuda_gl_interop.h needs <GL/gl.h>, which this image does not
// have; the symbol itself lives in libcudart and only needs a current GL context at run time, i.e. inside Houdini).
extern "C" cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource** resource, unsigned int image,
                                                   unsigned int target, unsigned int flags);

namespace {

thread_local std::string g_err = "";

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            std::ostringstream o_;                                                                \
            o_ << #call << " failed: " << cudaGetErrorString(e_) << " (" << __FILE__ << ":" << __LINE__ << ")"; \
            return fail(e_ == cudaErrorMemoryAllocation ? GSB_ERR_NOMEM : GSB_ERR_CUDA, o_.str()); \
        }                                                                                         \
    } while (0)

struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    // per-frame buffers whose size follows the camera (instances, records): half as much again, so a camera that orbits
    // towards a denser view re-allocates (a device-wide sync) a handful of times in total instead of every few frames
    cudaError_t ensure_grow(size_t bytes) { return ensure(bytes, 1); }
    cudaError_t ensure(size_t bytes, int slack_shift = 3)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + (bytes >> slack_shift) + 256;       // headroom so D jitter does not realloc every frame
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// One registered primitive (GSplatRegisterEntry, R.h:59-76).  Arrays are device copies.
struct Entry {
    uint64_t gdp = 0;
    int64_t  version[4] = { 0, 0, 0, 0 };
    int64_t  vtx0 = 0;
    int64_t  count = 0;
    float    origin[3] = { 0, 0, 0 };
    bool     active = false;
    int      age = -1, age_since_last_active = -1;
    bool     has_sh = false;
    bool     bbox_valid = false;           // all positions finite
    float    bbox[6] = { 0, 0, 0, 0, 0, 0 }; // min xyz, max xyz of the prim's points
    DevBuf   pos, cd, alpha, scale, orient, shx, shy, shz;
};

// ---- cold path (geometry change): host -> device through pinned, double-buffered staging.  A cudaMemcpyAsync from
// pageable memory is staged by the driver through one small bounce buffer (r01: 2.64 GB in 0.64 s, 4 GB/s); here a few
// host threads (a persistent pool, host_pool.h) fill one pinned slot while the DMA engine drains the others, so the rate is
// min(host memcpy, PCIe).
// GSB_TRACE_COLD=1 in the environment: wall-clock phases of the cold path on stderr (a diagnostic, not an interface)
struct ColdTrace {
    const char* what; bool on; std::chrono::steady_clock::time_point t0, tl; std::string line;
    explicit ColdTrace(const char* w) : what(w), on(getenv("GSB_TRACE_COLD") != nullptr), t0(std::chrono::steady_clock::now()), tl(t0) {}
    void mark(const char* phase)
    {
        if (!on) return;
        const auto t = std::chrono::steady_clock::now();
        char b[96]; snprintf(b, sizeof b, " %s=%.1f", phase, std::chrono::duration<double, std::milli>(t - tl).count());
        line += b; tl = t;
    }
    ~ColdTrace()
    {
        if (!on) return;
        fprintf(stderr, "[gsb cold] %s total=%.1f ms:%s\n", what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), line.c_str());
    }
};
template <class F> void parallel_ranges(size_t n, size_t grain, F&& f) { HostPool::get().ranges(n, grain, std::forward<F>(f)); }

struct Stager {
    static constexpr size_t SLOT = (size_t)64 << 20;
    static constexpr int NSLOT = 3;
    char* slot[NSLOT] = { nullptr, nullptr, nullptr };
    cudaEvent_t done[NSLOT] = { nullptr, nullptr, nullptr };
    bool busy[NSLOT] = { false, false, false };
    int next = 0;
    ~Stager()
    {
        for (int i = 0; i < NSLOT; ++i) { if (slot[i]) cudaFreeHost(slot[i]); if (done[i]) cudaEventDestroy(done[i]); }
    }
    cudaError_t init()
    {
        for (int i = 0; i < NSLOT; ++i) {
            if (!slot[i]) { cudaError_t e = cudaMallocHost(&slot[i], SLOT); if (e != cudaSuccess) return e; }
            if (!done[i]) { cudaError_t e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
        }
        return cudaSuccess;
    }
    // stream-ordered copy of a pageable host range; returns before the last DMA finishes (the slots are ours)
    cudaError_t copy(void* dst, const void* src, size_t bytes, cudaStream_t s)
    {
        cudaError_t e = init();
        if (e != cudaSuccess) return e;
        const char* sp = static_cast<const char*>(src);
        char* dp = static_cast<char*>(dst);
        for (size_t off = 0; off < bytes; off += SLOT) {
            const size_t len = std::min(SLOT, bytes - off);
            const int k = next; next = (next + 1) % NSLOT;
            if (busy[k]) { e = cudaEventSynchronize(done[k]); if (e != cudaSuccess) return e; busy[k] = false; }
            char* buf = slot[k];
            parallel_ranges(len, (size_t)1 << 20, [&](size_t a, size_t b) { memcpy(buf + a, sp + off + a, b - a); });
            e = cudaMemcpyAsync(dp + off, buf, len, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return e;
            e = cudaEventRecord(done[k], s);
            if (e != cudaSuccess) return e;
            busy[k] = true;
        }
        return cudaSuccess;
    }
};

enum { EV_START = 0, EV_PROJECT, EV_SORT, EV_BIN, EV_BLEND, EV_COPY, EV_COUNT };

}  // namespace

struct gsb_context {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;

    // registry + state machine (names follow the reference's members)
    std::map<std::string, std::unique_ptr<Entry>> registry;   // myRenderStateRegistry (ordered: deterministic iteration)
    std::set<std::string> active_set;                         // myActiveRegistries
    bool    pack_dirty = false;                               // a packed prim was re-registered / the cap changed: re-pack even if
                                                              // the requested set equals active_set (which stays the packed set)
    bool    render_enabled = true;                            // myIsRenderEnabled
    bool    can_render = false;                               // myCanRender
    bool    sh_present = false;                               // myIsShDataPresent
    bool    explicit_cam_set = false;                         // myIsExplicitCameraPosSet
    float   explicit_cam[3] = { 0, 0, 0 };
    int     sh_order = 0;                                     // myShOrder
    int64_t splat_count = 0;                                  // myGSplatCount
    float   origin[3] = { 0, 0, 0 };                          // mySplatOrigin
    bool    bbox_valid = false;                               // union of the packed prims' bounding boxes
    float   bbox[6] = { 0, 0, 0, 0, 0, 0 };

    // options
    int64_t cap = GSB_REFERENCE_SPLAT_CAP;
    float   eps_t = 1e-5f;
    bool    stage_timing = false, keep_intermediates = false;
    int     depth_chunks = 0;                                 // 0 = auto
    int     chunk_shift = 0;                                  // first chunk = V / 2^shift; 0 = auto (= depth_chunks)
    bool    lazy_project = true;                              // bounded K1 (GSB_OPT_LAZY_PROJECT); keep_intermediates forces the exact K1
    bool    host_direct = true;                               // finished tiles go straight to pinned host targets
    int     compact_mode = 0;                                 // 0 = auto (when row-partitioned), 1 = always, 2 = never

    // packed render-layout attributes
    DevBuf geomA, geomB, rows, sigA, sigB, lam;
    bool   sigma_valid = false;                      // lam (and the cell boxes) match the packed set and sigma_object
    bool   sigma_planes_valid = false;               // ... and so do sigA / sigB (only the exact K1 reads them)
    float  sigma_object[16] = {};
    // spatial cells of the bounded K1 (built by gsb_generate_render_geometry): the K1 stream in Morton order
    DevBuf geomA_p, lam_p, orig, cells, cell_views, sel_cells;
    DevBuf arena;                                    // per-frame counters, histograms, sort headers, tile flags: ONE memset per frame
    DevBuf lookback;                                 // radix-sort look-back table (epoch tagged, cleared on allocation only)
    uint32_t sort_epoch = 0;
    cudaEvent_t ev_sel = nullptr;                    // "the chunk's selection counters are in pinned memory"
    std::vector<uint32_t> owned_rows_h; int owned_key[4] = { -1, -1, -1, -1 };
    FrameConsts last_fc{}; bool last_lazy = false;   // for the on-demand debug view of the bound
    bool obj_level_warned = false;                   // _justPrintedOBJLevelRenderingWarning (R.h:127)
    int auto_shift = 4; long long auto_key[4] = { -1, -1, -1, -1 };   // adaptive first-chunk size (gsb_render)
    uint32_t* last_tile_consumed = nullptr;

    // per-frame device buffers.  keys/trects: K1 output in submission order (never moved).
    // lkeys/lvals: the live splats of the current chunk (ping-pong of their depth sort); ltiles: their tile rectangles (K2).
    DevBuf keys, trects, lkeys[2], lvals[2], ltiles, recs, rects, counts, ikeys[2], ivals[2],
           ranges, live_sat, owned_rows, fb, plan;
    DevBuf zdepth, scene_depth_buf;                  // scene-depth occlusion: window depth per live rank; GL depth copy
    struct cudaGraphicsResource* gl_depth_res = nullptr; uint32_t gl_depth_tex = 0; int gl_depth_w = 0, gl_depth_h = 0;
    DevBuf dbg_recs, dbg_inst;                       // GSB_OPT_KEEP_INTERMEDIATES views (by splat index)
    struct cudaGraphicsResource* gl_res = nullptr;  // registered viewport texture (CUDA<->GL interop hand-back)
    uint32_t gl_tex = 0; int gl_w = 0, gl_h = 0;
    Stager stager;                                   // pinned double-buffered staging of the cold-path uploads
    DevBuf wire_verts, wire_cols, wire_owner;         // wireframe overlay (gsb_render_wireframe)
    DevBuf shared_frame;                             // exported through CUDA IPC to the other ranks (display rank only)
    DevBuf scan_scratch;
    DevBuf scan_status;                               // single-pass count scan: epoch-tagged tile states (zeroed when allocated)
    unsigned long long* counters_h = nullptr;        // pinned mirror: [0..8) frame counters, [8..16) the current chunk's, [16..) every chunk's at frame end
    int order_buf = 0, order_vals_buf = 0, inst_buf = 0;
    int chunks_last = 0;
    int64_t last_live = 0;                           // live splats of the last depth chunk
    int64_t last_n = 0, last_sorted = 0; uint64_t last_d = 0; int last_tiles = 0; int last_w = 0, last_h = 0;
    float4* last_fb = nullptr;

    cudaEvent_t ev[EV_COUNT] = {};
    cudaEvent_t evc[16][5] = {};                     // per depth chunk: start, after the live sort, records, binning, blend
    int evc_chunks = 0;
    bool ev_valid = false;
    gsb_stats stats{};
};

namespace {

std::string make_id(const gsb_prim_key& k)
{
    // same text the reference builds at R.C:241-243
    std::ostringstream oss;
    oss << std::hex << std::showbase << (uintptr_t)k.gdp << "__" << std::dec << k.vtx0 << "__"
        << k.version[0] << "_" << k.version[1] << "_" << k.version[2] << "_" << k.version[3];
    return oss.str();
}

int upload(Stager& st, DevBuf& b, const void* src, size_t bytes, cudaStream_t s)
{
    CU(b.ensure(bytes ? bytes : 16));
    if (!bytes) return GSB_OK;
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, src) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    if (pinned || bytes < ((size_t)4 << 20)) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
    else CU(st.copy(b.p, src, bytes, s));
    return GSB_OK;
}

struct DevBufView { char* p; };
int upload_into(Stager& st, void* dst, const void* src, size_t bytes, cudaStream_t s)
{
    if (!bytes) return GSB_OK;
    if (bytes < ((size_t)4 << 20)) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    else CU(st.copy(dst, src, bytes, s));
    return GSB_OK;
}

// bounding box + finiteness of n xyz triples, over the host threads
void bbox_of(const float* pos, size_t n, float lo[3], float hi[3], bool& finite)
{
    const int T = host_threads();
    std::vector<float> part((size_t)T * 6, 0.0f);
    std::vector<char> fin((size_t)T, 1), used((size_t)T, 0);
    std::vector<std::thread> th;
    const int t = (int)std::min<size_t>((size_t)T, (n + 65535) / 65536);
    auto work = [&](int k, size_t a, size_t b) {
        float l[3] = { pos[3 * a], pos[3 * a + 1], pos[3 * a + 2] }, h[3] = { l[0], l[1], l[2] };
        bool f = true;
        for (size_t i = a; i < b; ++i)
            for (int c = 0; c < 3; ++c) {
                const float v = pos[3 * i + c];
                f = f && std::isfinite(v);
                l[c] = v < l[c] ? v : l[c]; h[c] = v > h[c] ? v : h[c];
            }
        for (int c = 0; c < 3; ++c) { part[(size_t)k * 6 + c] = l[c]; part[(size_t)k * 6 + 3 + c] = h[c]; }
        fin[(size_t)k] = f ? 1 : 0; used[(size_t)k] = 1;
    };
    if (t <= 1) work(0, 0, n);
    else {
        for (int k = 0; k < t; ++k) {
            const size_t a = n * (size_t)k / (size_t)t, b = n * (size_t)(k + 1) / (size_t)t;
            if (b > a) th.emplace_back(work, k, a, b);
        }
        for (auto& x : th) x.join();
    }
    finite = true; bool first = true;
    for (int k = 0; k < T; ++k) {
        if (!used[(size_t)k]) continue;
        finite = finite && fin[(size_t)k];
        for (int c = 0; c < 3; ++c) {
            const float l = part[(size_t)k * 6 + c], h = part[(size_t)k * 6 + 3 + c];
            lo[c] = first ? l : (l < lo[c] ? l : lo[c]); hi[c] = first ? h : (h > hi[c] ? h : hi[c]);
        }
        first = false;
    }
}

// camera = (0,0,0,1) * inverse(view) in double, rounded to f32 (R.C:558-562): 4th column of the
// inverse in column-vector convention.  Full adjugate, same formula as the oracle's.
void camera_from_view(const float view[16], float cam[3])
{
    double m[16], a[16];
    for (int i = 0; i < 16; ++i) m[i] = (double)view[i];
    a[0] = m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
    a[4] = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
    a[8] = m[4]*m[9]*m[15] - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
    a[12] = -m[4]*m[9]*m[14] + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
    a[13] = m[0]*m[9]*m[14] - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
    a[14] = -m[0]*m[5]*m[14] + m[0]*m[6]*m[13] + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6] + m[12]*m[2]*m[5];
    double det = m[0]*a[0] + m[1]*a[4] + m[2]*a[8] + m[3]*a[12];
    double r = 1.0 / det;
    cam[0] = (float)(a[12] * r); cam[1] = (float)(a[13] * r); cam[2] = (float)(a[14] * r);
}

// largest eigenvalue of the symmetric 3x3 matrix g (double; closed form, cos <= 1 fallback): |mat3(view)|_2^2 for the
// bounded K1, rounded up
double sym3_lambda_max(const double g[3][3])
{
    const double a = g[0][0], b = g[1][1], c = g[2][2], d = g[0][1], e = g[0][2], f = g[1][2];
    const double q = (a + b + c) / 3.0;
    const double p2 = (a - q) * (a - q) + (b - q) * (b - q) + (c - q) * (c - q) + 2.0 * (d * d + e * e + f * f);
    if (!(p2 > 0.0)) return q;
    const double p = std::sqrt(p2 / 6.0), ip = 1.0 / p;
    const double b00 = (a - q) * ip, b11 = (b - q) * ip, b22 = (c - q) * ip, b01 = d * ip, b02 = e * ip, b12 = f * ip;
    const double det = b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02);
    const double r = 0.5 * det;
    double cs = 1.0;
    if (r > -1.0 && r < 1.0) cs = std::cos(std::acos(r) / 3.0); else if (r <= -1.0) cs = 0.5;
    return q + 2.0 * p * cs;
}

int ceil_log2(uint32_t v) { int b = 0; while ((1u << b) < v) ++b; return b; }

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace

// Nothing may throw across the C ABI (std::bad_alloc from the registry's containers, std::length_error, ...): every
// entry point that can allocate is a function-try-block ending in this handler.
#define GSB_CATCH_ALL                                                                                      \
    catch (const std::bad_alloc&) { return fail(GSB_ERR_NOMEM, "out of host memory"); }                   \
    catch (const std::exception& e_) { return fail(GSB_ERR_INVALID, std::string("internal error: ") + e_.what()); } \
    catch (...) { return fail(GSB_ERR_INVALID, "internal error: unknown exception"); }

// look-back table of the radix sorts: entries carry the epoch of the sort that wrote them, so the table is cleared only
// when it is (re)allocated or when the 32-bit epoch counter wraps
static int ensure_lookback(gsb_context* ctx, size_t n_max)
{
    const size_t need = sort_lookback_bytes(n_max);
    if (need > ctx->lookback.cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(ctx->lookback.ensure(need));
        CU(cudaMemsetAsync(ctx->lookback.p, 0, ctx->lookback.cap, ctx->stream));
    }
    return GSB_OK;
}
static uint32_t next_epoch(gsb_context* ctx)
{
    if (ctx->sort_epoch >= 0xFFFFFFF0u) {
        if (ctx->lookback.p) cudaMemsetAsync(ctx->lookback.p, 0, ctx->lookback.cap, ctx->stream);
        if (ctx->scan_status.p) cudaMemsetAsync(ctx->scan_status.p, 0, ctx->scan_status.cap, ctx->stream);
        ctx->sort_epoch = 0;
    }
    return ++ctx->sort_epoch;
}

// ============================================================================== C ABI
extern "C" {

int gsb_abi_version(void) { return GSB_ABI_VERSION; }

const char* gsb_last_error(void) { return g_err.c_str(); }

int gsb_create(int cuda_device, gsb_context** out)
try {
    if (!out) return fail(GSB_ERR_INVALID, "gsb_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cuda_device < 0 || cuda_device >= ndev) return fail(GSB_ERR_INVALID, "gsb_create: no such CUDA device");
    CU(cudaSetDevice(cuda_device));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, cuda_device));
    if (prop.major < 10)
        return fail(GSB_ERR_CUDA, std::string("gsb_create: device is ") + prop.name +
                                      " (sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                      "); this library contains sm_100a code only");
    std::unique_ptr<gsb_context> c(new gsb_context);
    c->device = cuda_device;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (int i = 0; i < EV_COUNT; ++i) CU(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 5; ++j) CU(cudaEventCreate(&c->evc[i][j]));
    CU(cudaEventCreateWithFlags(&c->ev_sel, cudaEventDisableTiming));
    CU(cudaMallocHost(&c->counters_h, (16 + MAX_CHUNKS * 8) * 8));
    memset(c->counters_h, 0, (16 + MAX_CHUNKS * 8) * 8);
    CU(c->plan.ensure(sizeof(ChunkPlan)));
    // the cold path's pinned slots and host threads exist from here on: the first registration does not pay for them
    CU(c->stager.init());
    (void)HostPool::get();
    *out = c.release();
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_destroy(gsb_context* ctx)
try {
    if (!ctx) return GSB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->gl_res) cudaGraphicsUnregisterResource(ctx->gl_res);
    if (ctx->gl_depth_res) cudaGraphicsUnregisterResource(ctx->gl_depth_res);
    for (int i = 0; i < EV_COUNT; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 5; ++j) if (ctx->evc[i][j]) cudaEventDestroy(ctx->evc[i][j]);
    if (ctx->counters_h) cudaFreeHost(ctx->counters_h);
    if (ctx->ev_sel) cudaEventDestroy(ctx->ev_sel);
    cudaStream_t s = ctx->own_stream;
    delete ctx;
    if (s) cudaStreamDestroy(s);
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_set_stream(gsb_context* ctx, void* cuda_stream)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_synchronize(gsb_context* ctx)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_set_option(gsb_context* ctx, int option, double value)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    switch (option) {
    case GSB_OPT_SPLAT_CAP:
        if (value < 0) return fail(GSB_ERR_INVALID, "GSB_OPT_SPLAT_CAP must be >= 0");
        if (ctx->cap != (int64_t)value) ctx->pack_dirty = true;      // force a re-pack under the new budget
        ctx->cap = (int64_t)value; return GSB_OK;
    case GSB_OPT_EPS_T:
        if (!(value >= 0.0 && value < 1.0)) return fail(GSB_ERR_INVALID, "GSB_OPT_EPS_T must be in [0,1)");
        ctx->eps_t = (float)value; return GSB_OK;
    case GSB_OPT_STAGE_TIMING: ctx->stage_timing = value != 0; return GSB_OK;
    case GSB_OPT_KEEP_INTERMEDIATES: ctx->keep_intermediates = value != 0; return GSB_OK;
    case GSB_OPT_COMPACT:
        if (value < 0 || value > 2) return fail(GSB_ERR_INVALID, "GSB_OPT_COMPACT must be 0 (auto), 1 (always) or 2 (never)");
        ctx->compact_mode = (int)value; return GSB_OK;
    case GSB_OPT_HOST_DIRECT: ctx->host_direct = value != 0; return GSB_OK;
    case GSB_OPT_LAZY_PROJECT: ctx->lazy_project = value != 0; return GSB_OK;
    case GSB_OPT_CHUNK_SHIFT:
        if (value < 0 || value > 16) return fail(GSB_ERR_INVALID, "GSB_OPT_CHUNK_SHIFT must be 0 (auto) .. 16");
        ctx->chunk_shift = (int)value; return GSB_OK;
    case GSB_OPT_DEPTH_CHUNKS:
        if (value < 0 || value > 16) return fail(GSB_ERR_INVALID, "GSB_OPT_DEPTH_CHUNKS must be 0 (auto) .. 16");
        ctx->depth_chunks = (int)value; return GSB_OK;
    default: return fail(GSB_ERR_INVALID, "unknown option");
    }
} GSB_CATCH_ALL

int gsb_registry_size(gsb_context* ctx) { return ctx ? (int)ctx->registry.size() : 0; }

// ------------------------------------------------------------------ registerUpdate, R.C:218-291
int gsb_register_update(gsb_context* ctx, const gsb_prim_key* key, int64_t splat_count, const float origin[3],
                        const float* pos, const uint16_t* cd_h, const float* alpha, const uint16_t* scale_h,
                        const uint16_t* orient_h, const uint16_t* shx_h, const uint16_t* shy_h,
                        const uint16_t* shz_h, char* id_out)
try {
    if (!ctx || !key || !origin) return fail(GSB_ERR_INVALID, "gsb_register_update: NULL argument");
    if (splat_count < 0 || splat_count > 0x3fffffffLL) return fail(GSB_ERR_INVALID, "gsb_register_update: bad splat_count");
    if (splat_count > 0 && (!pos || !cd_h || !alpha || !scale_h || !orient_h))
        return fail(GSB_ERR_INVALID, "gsb_register_update: NULL attribute array");
    const bool has_sh = shx_h && shy_h && shz_h;
    if (!has_sh && (shx_h || shy_h || shz_h))
        return fail(GSB_ERR_INVALID, "gsb_register_update: shx/shy/shz must be all NULL or all non-NULL");
    CU(cudaSetDevice(ctx->device));
    const std::string id = make_id(*key);

    // The new entry is built on the side and enters the registry only when every upload has succeeded: a failed call
    // (GSB_ERR_NOMEM at 20 M splats ...) leaves the registry exactly as it was.
    std::unique_ptr<Entry> ne(new Entry);
    Entry& e = *ne;
    e.gdp = key->gdp; e.vtx0 = key->vtx0; memcpy(e.version, key->version, sizeof e.version);
    e.count = splat_count; memcpy(e.origin, origin, sizeof e.origin);
    e.active = false; e.age = -1; e.age_since_last_active = -1;
    e.has_sh = has_sh && splat_count > 0;
    const size_t n = (size_t)splat_count;
    cudaStream_t s = ctx->stream;
    int rc;
    ColdTrace tr("gsb_register_update");
    if ((rc = upload(ctx->stager, e.pos, pos, n * 12, s))) return rc;
    tr.mark("pos");
    // bounding box of the prim (bounds the depth keys, see gsb_render): host threads, while the first DMA runs
    e.bbox_valid = n > 0;
    if (n > 0) {
        float lo[3], hi[3]; bool finite = true;
        bbox_of(pos, n, lo, hi, finite);
        e.bbox_valid = finite;
        for (int k = 0; k < 3; ++k) { e.bbox[k] = lo[k]; e.bbox[3 + k] = hi[k]; }
    }
    tr.mark("bbox");
    if ((rc = upload(ctx->stager, e.cd, cd_h, n * 6, s))) return rc;
    if ((rc = upload(ctx->stager, e.alpha, alpha, n * 4, s))) return rc;
    if ((rc = upload(ctx->stager, e.scale, scale_h, n * 6, s))) return rc;
    if ((rc = upload(ctx->stager, e.orient, orient_h, n * 8, s))) return rc;
    tr.mark("cd_alpha_scale_orient");
    if (e.has_sh) {
        if ((rc = upload(ctx->stager, e.shx, shx_h, n * 32, s))) return rc;
        if ((rc = upload(ctx->stager, e.shy, shy_h, n * 32, s))) return rc;
        if ((rc = upload(ctx->stager, e.shz, shz_h, n * 32, s))) return rc;
        tr.mark("sh");
    }
    CU(cudaStreamSynchronize(s));      // the caller's arrays are not borrowed past this call
    tr.mark("sync");

    // same gdp, different version -> erase (R.C:246-265); then the entry replaces any older one under the same id
    for (auto it = ctx->registry.begin(); it != ctx->registry.end();) {
        Entry& o = *it->second;
        if (o.gdp == key->gdp && memcmp(o.version, key->version, sizeof o.version) != 0) it = ctx->registry.erase(it);
        else ++it;
    }
    ctx->registry[id] = std::move(ne);
    // a re-registered id carries new data: re-pack even if the active set looks unchanged (active_set itself stays the
    // packed set, so the comparison of R.C:141-153 keeps its meaning)
    if (ctx->active_set.count(id)) ctx->pack_dirty = true;
    if (id_out) { strncpy(id_out, id.c_str(), GSB_ID_MAX - 1); id_out[GSB_ID_MAX - 1] = 0; }
    return GSB_OK;
} GSB_CATCH_ALL

// ------------------------------------------------------------------ GR_PrimGsplat::update, GR.C:191-458 (SURVEY f-1)
int gsb_update_from_attributes(gsb_context* ctx, const gsb_prim_key* key, const gsb_raw_attributes* a, gsb_update_result* out)
try {
    if (!ctx || !key || !a || !out) return fail(GSB_ERR_INVALID, "gsb_update_from_attributes: NULL argument");
    if (a->count < 0 || a->count > 0x3fffffffLL) return fail(GSB_ERR_INVALID, "gsb_update_from_attributes: bad count");
    if (a->count > 0 && !a->P) return fail(GSB_ERR_INVALID, "gsb_update_from_attributes: P is required");
    if (a->activation != GSB_ACT_NONE && a->activation != GSB_ACT_INRIA)
        return fail(GSB_ERR_INVALID, "gsb_update_from_attributes: activation must be GSB_ACT_NONE or GSB_ACT_INRIA");
    CU(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof *out);
    const size_t n = (size_t)a->count;
    cudaStream_t s = ctx->stream;

    // which SH encoding: sh_coefficients, else sh1..sh15, else f_rest_0..44 (GR.C:145-189); an encoding counts only if complete
    int sh_kind = 0;
    if (a->sh_coefficients && a->sh_coefficients_len > 0) sh_kind = 1;
    if (!sh_kind) { bool all = true; for (int j = 0; j < 15; ++j) all = all && a->sh[j]; if (all) sh_kind = 2; }
    if (!sh_kind) { bool all = true; for (int j = 0; j < 45; ++j) all = all && a->f_rest[j]; if (all) sh_kind = 3; }
    const bool sh_found = sh_kind != 0 && n > 0;

    // barycentre: sequential fp32 sum / count (GEO_GSplat.C:338-351 — the order of the additions is the spec, so this
    // one pass stays serial; it runs on its own thread beside the uploads), bounding box over the host threads
    float sum[3] = { 0, 0, 0 }, lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
    bool finite = n > 0;
    std::thread bary_thread([&] {
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        const float* P = a->P;
        for (size_t i = 0; i < n; ++i) { sx += P[3 * i]; sy += P[3 * i + 1]; sz += P[3 * i + 2]; }
        sum[0] = sx; sum[1] = sy; sum[2] = sz;
    });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{ bary_thread };

    const std::string id = make_id(*key);
    std::unique_ptr<Entry> ne(new Entry);          // enters the registry only after every upload and kernel was issued
    Entry& e = *ne;
    e.gdp = key->gdp; e.vtx0 = key->vtx0; memcpy(e.version, key->version, sizeof e.version);
    e.count = a->count;
    e.active = false; e.age = -1; e.age_since_last_active = -1;
    e.has_sh = sh_found;

    // raw fp32 -> device staging, then quantise on the GPU
    DevBuf dP, dCd, dA, dS, dO, dSH;
    int rc;
    Stager& sg = ctx->stager;
    const float* alpha_src = a->Alpha ? a->Alpha : a->opacity;                     // Alpha wins when both exist (GR.C:246-257)
    if ((rc = upload(sg, dP, a->P, n * 12, s))) return rc;
    if (n > 0) bbox_of(a->P, n, lo, hi, finite);
    if (a->Cd && (rc = upload(sg, dCd, a->Cd, n * 12, s))) return rc;
    if (alpha_src && (rc = upload(sg, dA, alpha_src, n * 4, s))) return rc;
    if (a->scale && (rc = upload(sg, dS, a->scale, n * 12, s))) return rc;
    if (a->orient && (rc = upload(sg, dO, a->orient, n * 16, s))) return rc;
    CU(e.pos.ensure(n * 12 + 16)); CU(e.cd.ensure(n * 6 + 16)); CU(e.alpha.ensure(n * 4 + 16));
    CU(e.scale.ensure(n * 6 + 16)); CU(e.orient.ensure(n * 8 + 16));
    launch_ingest_core(dP.as<float>(), a->Cd ? dCd.as<float>() : nullptr, alpha_src ? dA.as<float>() : nullptr,
                       a->scale ? dS.as<float>() : nullptr, a->orient ? dO.as<float>() : nullptr, a->count, a->activation,
                       e.pos.as<float>(), e.cd.as<uint16_t>(), e.alpha.as<float>(), e.scale.as<uint16_t>(), e.orient.as<uint16_t>(), s);
    if (sh_found) {
        CU(e.shx.ensure(n * 32)); CU(e.shy.ensure(n * 32)); CU(e.shz.ensure(n * 32));
        if (sh_kind == 1) {
            const size_t len = (size_t)a->sh_coefficients_len;
            if ((rc = upload(sg, dSH, a->sh_coefficients, n * len * 12, s))) return rc;
            launch_ingest_sh_vec3(dSH.as<float>(), a->count, (int)len, 0, e.shx.as<uint16_t>(), e.shy.as<uint16_t>(), e.shz.as<uint16_t>(), s);
        } else if (sh_kind == 2) {
            CU(dSH.ensure(n * 15 * 12));
            for (int j = 0; j < 15; ++j) {
                DevBufView v{ dSH.as<char>() + (size_t)j * n * 12 };
                if ((rc = upload_into(sg, v.p, a->sh[j], n * 12, s))) return rc;
            }
            launch_ingest_sh_vec3(dSH.as<float>(), a->count, 15, 1, e.shx.as<uint16_t>(), e.shy.as<uint16_t>(), e.shz.as<uint16_t>(), s);
        } else {
            CU(dSH.ensure(n * 45 * 4));
            for (int j = 0; j < 45; ++j) {
                DevBufView v{ dSH.as<char>() + (size_t)j * n * 4 };
                if ((rc = upload_into(sg, v.p, a->f_rest[j], n * 4, s))) return rc;
            }
            launch_ingest_sh_rest(dSH.as<float>(), a->count, e.shx.as<uint16_t>(), e.shy.as<uint16_t>(), e.shz.as<uint16_t>(), s);
        }
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));          // staging buffers die here; the caller's arrays are not borrowed
    bary_thread.join();
    float bary[3] = { 0, 0, 0 };
    if (n > 0) { const float fn = (float)a->count; for (int k = 0; k < 3; ++k) bary[k] = sum[k] / fn; }
    memcpy(e.origin, bary, sizeof bary);
    e.bbox_valid = finite;
    for (int k = 0; k < 3; ++k) { e.bbox[k] = lo[k]; e.bbox[3 + k] = hi[k]; }

    for (auto it = ctx->registry.begin(); it != ctx->registry.end();) {          // same eviction as registerUpdate
        Entry& o = *it->second;
        if (o.gdp == key->gdp && memcmp(o.version, key->version, sizeof o.version) != 0) it = ctx->registry.erase(it);
        else ++it;
    }
    ctx->registry[id] = std::move(ne);
    if (ctx->active_set.count(id)) ctx->pack_dirty = true;

    strncpy(out->id, id.c_str(), GSB_ID_MAX - 1);
    out->sh_data_found = sh_found ? 1 : 0;
    out->sh_order = 3;                                                             // GR.C:444
    if (a->has_sh_order) {
        out->sh_order = a->sh_order;
        if (a->sh_order < 0 || a->sh_order > 3) { out->sh_order = 0; out->sh_order_invalid = 1; }   // GR.C:447-452
    }
    out->set_explicit_camera = a->has_explicit_camera ? 1 : 0;
    memcpy(out->explicit_camera, a->explicit_camera, 12);
    memcpy(out->barycentre, bary, 12);
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_debug_fetch_entry(gsb_context* ctx, const char* id, int which, void* dst, uint64_t dst_bytes, uint64_t* bytes_needed)
try {
    if (!ctx || !id) return fail(GSB_ERR_INVALID, "NULL argument");
    auto it = ctx->registry.find(id);
    if (it == ctx->registry.end()) return fail(GSB_ERR_NOT_FOUND, "gsb_debug_fetch_entry: unknown id");
    const Entry& e = *it->second;
    const uint64_t n = (uint64_t)e.count;
    const void* src = nullptr; uint64_t need = 0;
    switch (which) {
    case GSB_ENT_POS: src = e.pos.p; need = n * 12; break;
    case GSB_ENT_CD: src = e.cd.p; need = n * 6; break;
    case GSB_ENT_ALPHA: src = e.alpha.p; need = n * 4; break;
    case GSB_ENT_SCALE: src = e.scale.p; need = n * 6; break;
    case GSB_ENT_ORIENT: src = e.orient.p; need = n * 8; break;
    case GSB_ENT_SHX: src = e.shx.p; need = e.has_sh ? n * 32 : 0; break;
    case GSB_ENT_SHY: src = e.shy.p; need = e.has_sh ? n * 32 : 0; break;
    case GSB_ENT_SHZ: src = e.shz.p; need = e.has_sh ? n * 32 : 0; break;
    default: return fail(GSB_ERR_INVALID, "gsb_debug_fetch_entry: unknown array");
    }
    if (bytes_needed) *bytes_needed = need;
    if (dst && need) {
        if (dst_bytes < need) return fail(GSB_ERR_INVALID, "gsb_debug_fetch_entry: destination too small");
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
    }
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_include_in_render_pass(gsb_context* ctx, const char* id)      // R.C:313-320
try {
    if (!ctx || !id) return fail(GSB_ERR_INVALID, "NULL argument");
    auto it = ctx->registry.find(id);
    if (it == ctx->registry.end()) return GSB_OK;      // the reference silently ignores unknown ids
    it->second->active = true;
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_flush_entries_for_matching_detail(gsb_context* ctx, const char* id)   // R.C:293-311
try {
    if (!ctx || !id) return fail(GSB_ERR_INVALID, "NULL argument");
    auto it = ctx->registry.find(id);
    if (it == ctx->registry.end()) return GSB_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const uint64_t gdp = it->second->gdp;
    for (auto j = ctx->registry.begin(); j != ctx->registry.end();) {
        if (j->second->gdp == gdp) j = ctx->registry.erase(j); else ++j;
    }
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_set_rendering_enabled(gsb_context* ctx, int enabled)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    ctx->render_enabled = enabled != 0; return GSB_OK;
} GSB_CATCH_ALL
int gsb_set_explicit_camera_pos(gsb_context* ctx, const float pos[3])
try {
    if (!ctx || !pos) return fail(GSB_ERR_INVALID, "NULL argument");
    ctx->explicit_cam_set = true; memcpy(ctx->explicit_cam, pos, 12); return GSB_OK;
} GSB_CATCH_ALL
int gsb_set_spherical_harmonics_order(gsb_context* ctx, int sh_order)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    ctx->sh_order = sh_order; return GSB_OK;       // validation (0..3) is the caller's, GR_GSplat.C:444-457
} GSB_CATCH_ALL

// ------------------------------------------------------------------ generateRenderGeometry, R.C:322-532
int gsb_generate_render_geometry(gsb_context* ctx)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    ctx->stats.repacked = 0;
    // isRenderStateRegistryCurrent (R.C:141-153)
    std::set<std::string> requested;
    for (auto& kv : ctx->registry) if (kv.second->active) requested.insert(kv.first);
    if (requested == ctx->active_set && !ctx->pack_dirty) return GSB_OK;
    ctx->pack_dirty = false;

    CU(cudaSetDevice(ctx->device));
    const int64_t cap = ctx->cap > 0 ? ctx->cap : INT64_MAX;
    ctx->active_set.clear();
    ctx->can_render = false;
    int64_t total = 0;
    bool sh_all = true, cap_hit = false;
    for (auto& kv : ctx->registry) {                       // R.C:344-358
        Entry& e = *kv.second;
        if (!cap_hit && e.active && e.count > 0) {
            ctx->active_set.insert(kv.first);
            total += e.count;
            sh_all = sh_all && e.has_sh;                   // SH present iff every active prim has it
        }
        if (total >= cap) cap_hit = true;
    }
    if (!total) return GSB_OK;                             // R.C:360-363
    ctx->splat_count = std::min(total, cap);
    if (ctx->splat_count > 0x3fffffffLL) return fail(GSB_ERR_LIMIT, "more than 2^30-1 splats in the active set");
    ctx->sh_present = sh_all;

    // origin = mean of the active prims' barycentres, fp32, iteration order = ascending id (R.C:403-418)
    float o[3] = { 0, 0, 0 }; int clusters = 0;
    for (auto& id : ctx->active_set) {
        const Entry& e = *ctx->registry[id];
        o[0] += e.origin[0]; o[1] += e.origin[1]; o[2] += e.origin[2]; ++clusters;
    }
    if (clusters > 0) { const float c = (float)clusters; o[0] /= c; o[1] /= c; o[2] /= c; }
    memcpy(ctx->origin, o, sizeof o);
    ctx->bbox_valid = true;
    bool bb_first = true;
    for (auto& id : ctx->active_set) {
        const Entry& e = *ctx->registry[id];
        ctx->bbox_valid = ctx->bbox_valid && e.bbox_valid;
        for (int k = 0; k < 3; ++k) {
            ctx->bbox[k] = bb_first ? e.bbox[k] : std::min(ctx->bbox[k], e.bbox[k]);
            ctx->bbox[3 + k] = bb_first ? e.bbox[3 + k] : std::max(ctx->bbox[3 + k], e.bbox[3 + k]);
        }
        bb_first = false;
    }

    const size_t n = (size_t)ctx->splat_count;
    ColdTrace tr("gsb_generate_render_geometry");
    CU(ctx->geomA.ensure(n * 16)); CU(ctx->geomB.ensure(n * 16)); CU(ctx->rows.ensure(n * ROW_U4 * 16));
    tr.mark("alloc_packed");
    int64_t offset = 0;
    for (auto& id : ctx->active_set) {                     // R.C:420-511
        const Entry& e = *ctx->registry[id];
        const int64_t left = ctx->splat_count - offset;
        if (left <= 0) break;
        const int64_t cnt = std::min(e.count, left);
        launch_pack(e.pos.as<float>(), e.cd.as<uint16_t>(), e.alpha.as<float>(), e.scale.as<uint16_t>(),
                    e.orient.as<uint16_t>(), e.shx.as<uint16_t>(), e.shy.as<uint16_t>(), e.shz.as<uint16_t>(),
                    cnt, offset, ctx->geomA.as<float4>(), ctx->geomB.as<uint4>(), ctx->rows.as<uint4>(), sh_all ? 1 : 0, ctx->stream);
        offset += cnt;
    }
    CU(cudaGetLastError());
    // spatial cells of the bounded K1: order the set along a Morton curve (30-bit keys over the set's bounding box), keep
    // the original index of every position of that order, and lay the K1 stream (position + discard radius) out in it.
    // The cell boxes depend on the eigenvalue bounds, i.e. on the object matrix: gsb_render builds them with lam.
    {
        const int64_t ncells = ((int64_t)n + CELL - 1) / CELL;
        CU(ctx->geomA_p.ensure(n * 16)); CU(ctx->lam_p.ensure(n * 4)); CU(ctx->orig.ensure(n * 4 + 16));
        CU(ctx->cells.ensure((size_t)ncells * sizeof(CellBox))); CU(ctx->cell_views.ensure((size_t)ncells * 16));
        CU(ctx->sel_cells.ensure((size_t)ncells * 4 + 16));
        DevBuf mk[2], mi[2], hdr;
        for (int b = 0; b < 2; ++b) { CU(mk[b].ensure(n * 4 + 16)); CU(mi[b].ensure(n * 4 + 16)); }
        CU(hdr.ensure(sort_header_bytes()));
        int rc2 = ensure_lookback(ctx, n);
        if (rc2) return rc2;
        tr.mark("alloc_cells");
        float bb[6] = { 0, 0, 0, 0, 0, 0 };
        if (ctx->bbox_valid) memcpy(bb, ctx->bbox, sizeof bb);          // invalid box: every key 0, the order stays the index order
        launch_morton(ctx->geomA.as<float4>(), (int64_t)n, bb, mk[0].as<uint32_t>(), mi[0].as<uint32_t>(), ctx->stream);
        const int cur = radix_sort_pairs(mk[0].as<uint32_t>(), mi[0].as<uint32_t>(), mk[1].as<uint32_t>(), mi[1].as<uint32_t>(), n, nullptr,
                                         0, 30, hdr.as<uint32_t>(), false, false, ctx->lookback.as<unsigned long long>(),
                                         next_epoch(ctx), nullptr, ctx->stream, nullptr);
        CU(cudaMemcpyAsync(ctx->orig.p, mi[cur].p, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        launch_gather_geom(ctx->orig.as<uint32_t>(), ctx->geomA.as<float4>(), (int64_t)n, ctx->geomA_p.as<float4>(), ctx->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));                          // the temporaries die here
        tr.mark("pack_morton_gather");
    }
    ctx->sigma_valid = false; ctx->sigma_planes_valid = false;
    ctx->can_render = true;
    ctx->stats.repacked = 1;
    return GSB_OK;
} GSB_CATCH_ALL

// ------------------------------------------------------------------ render, R.C:534-658
int gsb_render(gsb_context* ctx, const gsb_frame* fr, const gsb_target* target)
try {
    if (!ctx || !fr) return fail(GSB_ERR_INVALID, "gsb_render: NULL argument");
    gsb_stats& st = ctx->stats;
    st.rendered = 0; st.launches = 0; st.warnings = 0;
    if (!ctx->render_enabled || !ctx->can_render) return GSB_OK;          // R.C:536-539
    bool any = false;
    for (auto& kv : ctx->registry) any |= kv.second->active;
    if (!any) return GSB_OK;                                             // R.C:541-549
    // OBJ-level rendering: the reference warns once per switch to OBJ level and renders anyway (R.C:565-581)
    if (fr->is_object_level) {
        if (!ctx->obj_level_warned) { st.warnings |= GSB_WARN_OBJECT_LEVEL; ctx->obj_level_warned = true; }
    } else ctx->obj_level_warned = false;
    if (fr->width < 1 || fr->height < 1 || fr->width > 65535 || fr->height > 65535)
        return fail(GSB_ERR_LIMIT, "gsb_render: screen size must be 1..65535");
    if (fr->row_world < 1 || fr->row_rank < 0 || fr->row_rank >= fr->row_world)
        return fail(GSB_ERR_INVALID, "gsb_render: bad tile-row partition");
    if (fr->row_group < 0) return fail(GSB_ERR_INVALID, "gsb_render: row_group must be >= 0");
    if (fr->depth_func < GSB_DEPTH_NONE || fr->depth_func > GSB_DEPTH_LEQUAL)
        return fail(GSB_ERR_INVALID, "gsb_render: depth_func must be GSB_DEPTH_NONE, GSB_DEPTH_LESS or GSB_DEPTH_LEQUAL");
    if (fr->depth_func != GSB_DEPTH_NONE && !fr->scene_depth && fr->gl_depth_texture == 0)
        return fail(GSB_ERR_INVALID, "gsb_render: depth_func set but neither scene_depth nor gl_depth_texture given");
    CU(cudaSetDevice(ctx->device));
    (void)cudaGetLastError();            // a non-sticky error left by an earlier failed call (e.g. GL interop without a GL context) is not this frame's
    cudaStream_t s = ctx->stream;

    FrameConsts fc{};
    memcpy(fc.view, fr->view, 64); memcpy(fc.proj, fr->proj, 64); memcpy(fc.object, fr->object, 64);
    memcpy(fc.inv_object, fr->inv_object, 64); memcpy(fc.obj_view, fr->obj_view, 64);
    if (ctx->explicit_cam_set) memcpy(fc.cam, ctx->explicit_cam, 12);     // R.C:552-555
    else camera_from_view(fr->view, fc.cam);                              // R.C:556-563
    memcpy(fc.origin, ctx->origin, 12);
    fc.width = fr->width; fc.height = fr->height; fc.W = (float)fr->width; fc.H = (float)fr->height;
    fc.tiles_x = (fr->width + TILE - 1) / TILE; fc.tiles_y = (fr->height + TILE - 1) / TILE;
    const bool do_sh = ctx->sh_order > 0 && ctx->sh_present;              // R.C:623
    fc.sh_order = do_sh ? std::min(ctx->sh_order, 3) : 0;
    fc.row_rank = fr->row_rank; fc.row_world = fr->row_world; fc.row_group = fr->row_group > 1 ? fr->row_group : 1;
    fc.eps_t = ctx->eps_t;
    // scene-depth occlusion (R.C:608-610; SRC.h:278-282): window depth = ndc_z * (far - near)/2 + (far + near)/2
    fc.depth_func = fr->depth_func;
    {
        float dn = fr->depth_range[0], df = fr->depth_range[1];
        if (dn == 0.0f && df == 0.0f) df = 1.0f;                            // unset -> glDepthRange default
        fc.depth_hr = (df - dn) * 0.5f; fc.depth_hm = (df + dn) * 0.5f;
    }
    {   // constants of the covariance chain for the bounded K1: the spec's fp32 operations (project_geom), on the host
        const float p00 = fr->proj[0], p11 = fr->proj[5];
        const float aspect = p00 / p11;
        const float tan_x = 1.0f / p00, tan_y = 1.0f / (p11 * aspect);
        fc.lim_x = 1.3f * tan_x; fc.lim_y = 1.3f * tan_y;
        fc.focal = (fc.W * p00) / 2.0f;
        double g[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            g[i][j] = 0.0;
            for (int r = 0; r < 3; ++r) g[i][j] += (double)fr->view[i * 4 + r] * (double)fr->view[j * 4 + r];   // (W^T W)_ij
        }
        const double wn = sym3_lambda_max(g) * (1.0 + 1e-6);
        fc.wnorm2 = std::nextafterf((float)wn, INFINITY);
    }
    const int num_tiles = fc.tiles_x * fc.tiles_y;
    const int64_t n = ctx->splat_count;
    const size_t  N = (size_t)n;

    // depth chunks: the frame is binned and blended front to back in nchunks ranges of the depth order
    int nchunks = ctx->depth_chunks;
    // auto: a first chunk that saturates most tiles + the rest.  A chunk costs ~80 us of fixed work (selection, two sorts
    // of small lists, launches) and every instance it avoids ~33 ns, so a third chunk pays when the instance count is large:
    // at 4K and beyond (r02 sweeps, 20 M splats: 8K 5.57 ms with 3 chunks vs 5.78 with 2; 1080p 1.415 vs 1.403).
    // The SIZE of the first chunk is not a constant: it follows the previous frame's outcome (auto_shift, below).
    const bool auto_chunks = nchunks <= 0;
    if (auto_chunks) nchunks = (n >= (int64_t)500000) ? (((int64_t)fr->width * fr->height >= (int64_t)3840 * 2160) ? 3 : 2) : 1;
    nchunks = std::min(nchunks, (int)MAX_CHUNKS);

    // Every depth key is the fp32 bit pattern of a squared distance from the camera to a point inside the packed set's
    // bounding box, so keys lie in [bits(dmin^2), bits(dmax^2)]: sorting key - key_min needs only the bits that vary
    // (25 instead of 32 for the benchmark cloud => 3 passes instead of 4) and gives the identical order.  The same
    // bound makes the depth buckets (chunk partition) linear in the distance over [dmin, dmax].
    uint32_t key_min = 0u, key_span = 0xFFFFFFFFu;
    DepthBuckets db{ 0u, 0 };
    bool range_ok = false;
    if (ctx->bbox_valid) {
        double dmin2 = 0.0, dmax2 = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double c = (double)fc.cam[k], lo = (double)ctx->bbox[k], hi = (double)ctx->bbox[3 + k];
            const double near_d = c < lo ? lo - c : (c > hi ? c - hi : 0.0);
            const double far_d = std::max(std::fabs(c - lo), std::fabs(c - hi));
            dmin2 += near_d * near_d; dmax2 += far_d * far_d;
        }
        const float fmin = std::nextafterf((float)(dmin2 * (1.0 - 1e-5)), 0.0f);
        const float fmax = std::nextafterf((float)(dmax2 * (1.0 + 1e-5)), INFINITY);
        if (std::isfinite(fmin) && std::isfinite(fmax) && fmin >= 0.0f && fmax >= fmin) {
            uint32_t bmin, bmax; memcpy(&bmin, &fmin, 4); memcpy(&bmax, &fmax, 4);
            key_min = bmin; key_span = bmax - bmin + 1u;      // valid keys squeeze to [0, span-1], culled to span
            // buckets: (key - key_min) >> shift, the shift that maps the span onto [0, DEPTH_BUCKETS - 2]
            db.key_min = bmin; db.shift = 0;
            while (((key_span - 1u) >> db.shift) > (uint32_t)(DEPTH_BUCKETS - 2)) ++db.shift;
            range_ok = key_span > 1u;
        }
    }
    if (!range_ok) nchunks = 1;                               // no finite depth range: a single chunk needs no buckets
    const int key_bits = sort_key_bits(key_span);

    // buffers
    // exact K1 only: packed tile rectangles ride along as a payload (screens up to 512 x 512 tiles); otherwise exact rectangles by index
    const bool use_trects = fc.tiles_x <= 512 && fc.tiles_y <= 512;
    // bounded K1 over spatial cells (production): conservative tile rectangles, evaluated only for the members of the cells
    // a depth chunk selects; the exact projection runs in K2 for the splats that were selected.  The debug views (exact
    // rectangles by splat index) keep the exact K1 for every splat.
    const bool lazy = ctx->lazy_project && !ctx->keep_intermediates &&
                      std::isfinite(fc.lim_x) && std::isfinite(fc.lim_y) && std::isfinite(fc.focal) && std::isfinite(fc.wnorm2);
    if (!lazy) {
        CU(ctx->keys.ensure(N * 4));
        if (use_trects) CU(ctx->trects.ensure(N * 4));
        CU(ctx->rects.ensure(N * 8));
    }
    CU(ctx->scan_scratch.ensure(std::max(scan_scratch_bytes(N), select_scratch_bytes(n))));
    if (scan_status_bytes(N) > ctx->scan_status.cap) {
        CU(ctx->scan_status.ensure(scan_status_bytes(N)));
        CU(cudaMemsetAsync(ctx->scan_status.p, 0, ctx->scan_status.cap, s));
    }
    CU(ctx->ranges.ensure((size_t)num_tiles * 8));
    const size_t done_bytes = (size_t)done_words_per_row(fc.tiles_x) * (size_t)fc.tiles_y * 4;      // one bit per tile
    CU(ctx->live_sat.ensure((size_t)(fc.tiles_x + 1) * (size_t)(fc.tiles_y + 1) * 4));
    float4* fb = nullptr;
    const size_t fb_bytes = (size_t)fr->width * fr->height * 16;
    if (target && target->device_rgba) fb = static_cast<float4*>(target->device_rgba);
    else { CU(ctx->fb.ensure(fb_bytes)); fb = ctx->fb.as<float4>(); }
    float4* fb_final = (target && target->final_rgba) ? static_cast<float4*>(target->final_rgba) : fb;
    // A pinned, device-addressable host target receives the finished tiles directly from the blend kernel (zero-copy
    // stores over PCIe, overlapped with the binning of the deeper chunks) instead of a D2H pass after the frame.
    bool host_is_final = false;
    if (target && target->host_rgba && !target->final_rgba && ctx->host_direct) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, target->host_rgba) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
            fb_final = static_cast<float4*>(pa.devicePointer);
            host_is_final = true;
        } else (void)cudaGetLastError();
    }

    const float* scene_depth = nullptr;
    if (fc.depth_func != GSB_DEPTH_NONE) {
        scene_depth = static_cast<const float*>(fr->scene_depth);
        if (!scene_depth) {
            // Houdini's depth attachment, copied by the shim into an R32F texture: map it and copy it device->device
            if (ctx->gl_depth_res && (ctx->gl_depth_tex != fr->gl_depth_texture || ctx->gl_depth_w != fr->width || ctx->gl_depth_h != fr->height)) {
                cudaGraphicsUnregisterResource(ctx->gl_depth_res); ctx->gl_depth_res = nullptr;
            }
            if (!ctx->gl_depth_res) {
                CU(cudaGraphicsGLRegisterImage(&ctx->gl_depth_res, fr->gl_depth_texture, 0x0DE1u /* GL_TEXTURE_2D */,
                                               cudaGraphicsRegisterFlagsReadOnly));
                ctx->gl_depth_tex = fr->gl_depth_texture; ctx->gl_depth_w = fr->width; ctx->gl_depth_h = fr->height;
            }
            CU(ctx->scene_depth_buf.ensure((size_t)fr->width * fr->height * 4));
            cudaArray_t arr = nullptr;
            CU(cudaGraphicsMapResources(1, &ctx->gl_depth_res, s));
            CU(cudaGraphicsSubResourceGetMappedArray(&arr, ctx->gl_depth_res, 0, 0));
            CU(cudaMemcpy2DFromArrayAsync(ctx->scene_depth_buf.p, (size_t)fr->width * 4, arr, 0, 0, (size_t)fr->width * 4,
                                          (size_t)fr->height, cudaMemcpyDeviceToDevice, s));
            CU(cudaGraphicsUnmapResources(1, &ctx->gl_depth_res, s));
            scene_depth = ctx->scene_depth_buf.as<float>();
        }
    }

    // ---- the frame arena: every counter, histogram, sort header and tile flag of the frame, cleared by ONE memset.
    //   [0, 64)                    frame counters (u64): [0] V  [1] D_c  [2] sort error flag  [3] finished tiles
    //   [64, 64 + 64 MAX_CHUNKS)   per depth chunk (u64): [0] L  [1] bound of D (selection)  [2] selected cells (u32)  [3] D (exact)
    //   bucket histogram of the chunk plan, 2 sort headers per chunk, tile_done bit map, tile_consumed
    const size_t hdr_bytes = sort_header_bytes();
    const size_t off_chunk = 64, off_hist = off_chunk + (size_t)MAX_CHUNKS * 64, off_sort = align256(off_hist + DEPTH_BUCKETS * 4),
                 off_done = off_sort + (size_t)MAX_CHUNKS * 2 * hdr_bytes, off_cons = align256(off_done + done_bytes),
                 arena_bytes = off_cons + (size_t)num_tiles * 4;
    CU(ctx->arena.ensure(arena_bytes));
    char* const arena = ctx->arena.as<char>();
    const bool tm = ctx->stage_timing;
    if (tm) CU(cudaEventRecord(ctx->ev[EV_START], s));
    CU(cudaMemsetAsync(arena, 0, arena_bytes, s));
    unsigned long long* const cnt = reinterpret_cast<unsigned long long*>(arena);
    uint32_t* const tile_done = reinterpret_cast<uint32_t*>(arena + off_done);
    uint32_t* const tile_consumed = reinterpret_cast<uint32_t*>(arena + off_cons);
    ctx->last_tile_consumed = tile_consumed;
    {
        int rc2 = ensure_lookback(ctx, N);
        if (rc2) return rc2;
    }

    // eigenvalue bounds (and, for the exact K1, the covariance planes) follow the object matrix; so do the cell boxes
    const bool object_changed = memcmp(ctx->sigma_object, fr->object, 64) != 0;
    if (!ctx->sigma_valid || object_changed || (!lazy && !ctx->sigma_planes_valid)) {
        CU(ctx->lam.ensure(N * 4));
        if (!lazy) { CU(ctx->sigA.ensure(N * 16)); CU(ctx->sigB.ensure(N * 8)); }
        launch_sigma(fr->object, ctx->geomB.as<uint4>(), n, lazy ? nullptr : ctx->sigA.as<float4>(),
                     lazy ? nullptr : ctx->sigB.as<float2>(), ctx->lam.as<float>(), s);
        launch_cell_build(ctx->geomA_p.as<float4>(), ctx->orig.as<uint32_t>(), ctx->lam.as<float>(), n, ctx->lam_p.as<float>(),
                          ctx->cells.as<CellBox>(), s);
        memcpy(ctx->sigma_object, fr->object, 64);
        ctx->sigma_valid = true;
        ctx->sigma_planes_valid = !lazy;
        st.launches += 2;
    }
    // row-partitioned frame, exact K1 / K2: prefix count of the tile rows this rank owns (cached per partition)
    const uint32_t* owned_rows = nullptr;
    if (fr->row_world > 1) {
        const int key4[4] = { fc.tiles_y, fc.row_rank, fc.row_world, fc.row_group };
        if (memcmp(key4, ctx->owned_key, sizeof key4) != 0) {
            CU(cudaStreamSynchronize(s));                   // nothing in flight reads the old table
            ctx->owned_rows_h.assign((size_t)fc.tiles_y + 1, 0u);
            for (int ty = 0; ty < fc.tiles_y; ++ty)
                ctx->owned_rows_h[ty + 1] = ctx->owned_rows_h[ty] + (owns_row(ty, fc.row_rank, fc.row_world, fc.row_group) ? 1u : 0u);
            CU(ctx->owned_rows.ensure(ctx->owned_rows_h.size() * 4));
            CU(cudaMemcpyAsync(ctx->owned_rows.p, ctx->owned_rows_h.data(), ctx->owned_rows_h.size() * 4, cudaMemcpyHostToDevice, s));
            memcpy(ctx->owned_key, key4, sizeof key4);
        }
        owned_rows = ctx->owned_rows.as<uint32_t>();
    }
    PackedSplats ps{ ctx->geomA.as<float4>(), ctx->geomB.as<uint4>(), ctx->rows.as<uint4>(), ctx->sigA.as<float4>(),
                     ctx->sigB.as<float2>(), ctx->lam.as<float>() };
    uint32_t* bucket_hist = nchunks > 1 ? reinterpret_cast<uint32_t*>(arena + off_hist) : nullptr;
    // K1.  Bounded: one thread per CELL projects the cell's box (tile rectangle + key interval + the chunk plan's histogram);
    // the per-splat bound runs later, inside the chunks, for the selected cells only.  Exact: every submitted splat.
    if (lazy) launch_cell_project(fc, ctx->cells.as<CellBox>(), n, db, bucket_hist, ctx->cell_views.as<uint4>(), cnt + 0, s);
    else launch_project(fc, ps, n, ctx->keys.as<uint32_t>(), ctx->rects.as<uint2>(),
                        (ctx->keep_intermediates || !use_trects) ? 1 : 0, use_trects ? ctx->trects.as<uint32_t>() : nullptr,
                        cnt + 0, db, bucket_hist, owned_rows, s);
    const uint2* exact_rects = lazy ? nullptr : ctx->rects.as<uint2>();
    st.launches += 1;
    // The depth order is cut into chunks WITHOUT sorting or moving the cloud: the chunk plan maps every depth bucket to
    // a chunk and turns the bucket boundaries into key boundaries; each chunk then selects its own live splats.
    const ChunkPlan* chunk_plan = nullptr;
    if (nchunks > 1) {
        ChunkPlan* plan = ctx->plan.as<ChunkPlan>();
        // first chunk = V / 2^shift.  Explicit chunk count: geometric (shift = count).  Auto: the shift that the feedback
        // below settled on for this (cloud, screen, chunk count); the starting guess only matters for the first frames
        const long long akey[4] = { (long long)n, fr->width, fr->height, nchunks };
        if (memcmp(akey, ctx->auto_key, sizeof akey) != 0) {
            memcpy(ctx->auto_key, akey, sizeof akey);
            ctx->auto_shift = nchunks + 1 + (n >= (int64_t)10000000 ? 1 : 0);
        }
        const int shift = ctx->chunk_shift > 0 ? ctx->chunk_shift : (ctx->depth_chunks > 0 ? nchunks : ctx->auto_shift);
        launch_choose_chunks(bucket_hist, nchunks, shift, db, plan, s);
        st.launches += 1;
        chunk_plan = plan;
    }
    const uint32_t* pkeys = ctx->keys.as<uint32_t>();
    const uint32_t* ptrects = use_trects ? ctx->trects.as<uint32_t>() : nullptr;
    if (tm) CU(cudaEventRecord(ctx->ev[EV_PROJECT], s));
    if (tm) CU(cudaEventRecord(ctx->ev[EV_SORT], s));

    // Per depth chunk: live selection -> depth sort -> K2 records -> binning -> blend.  The splats of the chunk that still
    // touch a live tile (owned by this rank, not saturated by nearer chunks) are compacted, depth-sorted, binned, given their
    // records and blended; tiles whose pixels all saturated are flagged.  The per-pixel sequence of blended instances is the
    // one a single global sort would give, so the frame is bit-identical to the single-chunk result.
    // Every kernel reads its element count from the device; the host needs L and a bound of D only to size buffers, and
    // learns them from ONE event wait per chunk that it reaches after it has queued the chunk's depth sort.
    if (fr->row_world > 1 && fb_final == fb) CU(cudaMemsetAsync(fb, 0, fb_bytes, s));   // rows this rank does not own stay zero
    const int tile_bits = std::max(1, ceil_log2((uint32_t)num_tiles));
    const SortPlan depth_plan = sort_plan(0, key_bits);
    // tiles this rank owns: when all of them are saturated no deeper splat can change a pixel and the frame is done
    int n_owned_rows = 0;
    for (int ty = 0; ty < fc.tiles_y; ++ty) n_owned_rows += owns_row(ty, fc.row_rank, fc.row_world, fc.row_group) ? 1 : 0;
    const uint64_t owned_tiles = (uint64_t)n_owned_rows * (uint64_t)fc.tiles_x;
    uint64_t L_total = 0, V = 0, D = 0, L = 0;
    uint64_t L_chunk[MAX_CHUNKS] = {};
    int chunks_run = 0;
    // the live buffers are sized once for the cloud, so no size has to come back from the device before the selection
    // (exact K1: the second halves double as the selection's staging area: CTA-local runs, dead before the sort ping-pongs)
    const size_t live_bytes = select_stage_elems(n) * 4 + 16;
    for (int b = 0; b < 2; ++b) {
        CU(ctx->lkeys[b].ensure(live_bytes)); CU(ctx->lvals[b].ensure(live_bytes));
    }
    uint32_t* const err_flag = reinterpret_cast<uint32_t*>(cnt + 2);
    for (int c = 0; c < nchunks; ++c) {
        const bool first = (c == 0);
        if (tm) CU(cudaEventRecord(ctx->evc[c][0], s));
        unsigned long long* const cc = cnt + 8 + (size_t)c * 8;              // this chunk's counters
        uint32_t* const hdr_depth = reinterpret_cast<uint32_t*>(arena + off_sort + (size_t)(2 * c) * hdr_bytes);
        uint32_t* const hdr_tile = reinterpret_cast<uint32_t*>(arena + off_sort + (size_t)(2 * c + 1) * hdr_bytes);
        // live map of this chunk as a summed-area table (first chunk of a single rank: every tile is live, no table)
        const uint32_t* sat = nullptr;
        if (!first || fr->row_world > 1) {
            launch_live_sat(fc, first ? nullptr : tile_done, ctx->live_sat.as<uint32_t>(), s);
            st.launches += 1;
            sat = ctx->live_sat.as<uint32_t>();
        }
        // live selection: the splats of the chunk that still touch a live tile -> (key, index) pairs, their number L and a
        // bound of the instances D they will emit
        if (lazy) {
            launch_cell_select(ctx->cell_views.as<uint4>(), n, chunk_plan, c, fc, sat, ctx->sel_cells.as<uint32_t>(),
                               reinterpret_cast<uint32_t*>(cc + 2), s);
            launch_splat_select(fc, ctx->geomA_p.as<float4>(), ctx->lam_p.as<float>(), ctx->orig.as<uint32_t>(), n,
                                ctx->sel_cells.as<uint32_t>(), reinterpret_cast<const uint32_t*>(cc + 2), chunk_plan, c, sat,
                                ctx->lkeys[0].as<uint32_t>(), ctx->lvals[0].as<uint32_t>(), cc + 0, cc + 1,
                                depth_plan, key_min, key_span, hdr_depth, s);
            st.launches += 2;
        } else {
            launch_select_live(pkeys, ptrects, exact_rects, n, chunk_plan, c, fc,
                               sat, ctx->lkeys[0].as<uint32_t>(), ctx->lvals[0].as<uint32_t>(),
                               ctx->lkeys[1].as<uint32_t>(), ctx->lvals[1].as<uint32_t>(), ctx->scan_scratch.p, cc + 0, cc + 1, s);
            st.launches += 3;
        }
        CU(cudaMemcpyAsync(ctx->counters_h, cnt, 64, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(ctx->counters_h + 8, cc, 64, cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(ctx->ev_sel, s));
        // depth sort of the live splats, queued BEFORE the host waits: the count is read on the device, the persistent pass
        // kernels do not depend on it, and the GPU has work while the host sizes the chunk's buffers
        ctx->order_buf = radix_sort_pairs(ctx->lkeys[0].as<uint32_t>(), ctx->lvals[0].as<uint32_t>(),
                                          ctx->lkeys[1].as<uint32_t>(), ctx->lvals[1].as<uint32_t>(), N, cc + 0, 0, key_bits,
                                          hdr_depth, true, lazy, ctx->lookback.as<unsigned long long>(), next_epoch(ctx), err_flag, s, &st.launches,
                                          key_min, key_span);
        // the one host wait of the chunk: V, this chunk's L and bound of D, the tiles finished by the previous chunks
        CU(cudaEventSynchronize(ctx->ev_sel));
        V = ctx->counters_h[0]; L = ctx->counters_h[8]; D = ctx->counters_h[9];
        L_chunk[c] = L;
        const bool all_done = ctx->counters_h[3] >= owned_tiles;        // implies L == 0
        // the last chunk that has work, or the last chunk at all, finalises the un-saturated tiles
        const bool last = (c == nchunks - 1) || all_done;
        if (D > 0x3fffffffull) return fail(GSB_ERR_LIMIT, "more than 2^30-1 tile instances in one depth chunk");
        L_total += L;
        chunks_run = c + 1;
        if (all_done && !first) {                                        // every tile was finalised when it saturated
            if (tm) for (int e = 1; e < 5; ++e) CU(cudaEventRecord(ctx->evc[c][e], s));
            break;
        }
        for (int b = 0; b < 2; ++b) { CU(ctx->ikeys[b].ensure_grow((size_t)D * 4 + 16)); CU(ctx->ivals[b].ensure_grow((size_t)D * 4 + 16)); }
        CU(ctx->recs.ensure_grow((size_t)L * sizeof(Record) + 16)); CU(ctx->ltiles.ensure_grow((size_t)L * 8 + 16));
        CU(ctx->counts.ensure_grow((size_t)L * 4 + 16));
        if (scene_depth) CU(ctx->zdepth.ensure_grow((size_t)L * 4 + 16));
        float* zdepth = scene_depth ? ctx->zdepth.as<float>() : nullptr;
        {
            int rc2 = ensure_lookback(ctx, (size_t)std::max<uint64_t>(D, N));
            if (rc2) return rc2;
        }
        uint32_t* counts = ctx->counts.as<uint32_t>();
        // ties of the depth sort in index order (the live list arrives in cell order): the final order.  (r02 also measured
        // this inside K2: the divergent tie loops cost the gather-bound kernel 32 us per frame, the separate kernel 23.)
        ctx->order_vals_buf = ctx->order_buf ^ 1;
        launch_tie_fix(ctx->lkeys[ctx->order_buf].as<uint32_t>(), ctx->lvals[ctx->order_buf].as<uint32_t>(), L, nullptr,
                       ctx->lvals[ctx->order_vals_buf].as<uint32_t>(), s);
        st.launches += (L ? 1 : 0);
        const uint32_t* order = ctx->lvals[ctx->order_vals_buf].as<uint32_t>();
        if (tm) CU(cudaEventRecord(ctx->evc[c][1], s));
        // K2: records of the live splats, in depth order, plus their tile rectangles and live-tile counts
        launch_records(fc, ps, order, (int64_t)L, sat, ctx->recs.as<Record>(), ctx->ltiles.as<uint2>(), counts, zdepth, owned_rows, s);
        st.launches += (L ? 1 : 0);
        if (tm) CU(cudaEventRecord(ctx->evc[c][2], s));
        // K4: live-tile counts (K2) -> offsets (their exact total D stays on the device, cc + 3) -> instances (the emit also
        // builds the digit histograms of the tile partition; r02 measured the scan INSIDE the emit, one decoupled look-back
        // per CTA: 16 us slower per frame than a scan kernel of its own) -> stable partition by tile -> tile ranges
        const SortPlan tile_plan = sort_plan(0, tile_bits);
        // (one launch: reduce, decoupled look-back over epoch-tagged tile states and apply; GSB_SCAN=3 in the environment keeps
        // the three-launch scan for A/B runs.  Every partial sum is below the 2^30 - 1 instance limit checked above.)
        static const bool scan3 = [] { const char* e = getenv("GSB_SCAN"); return e && atoi(e) == 3; }();
        if (scan3) exclusive_scan_u32(counts, counts, (size_t)L, ctx->scan_scratch.p, cc + 3, s, &st.launches);
        else exclusive_scan_u32_onepass(counts, counts, (size_t)L, ctx->scan_status.as<unsigned long long>(), next_epoch(ctx), cc + 3,
                                        err_flag, s, &st.launches);
        launch_emit(ctx->ltiles.as<uint2>(), counts, cc + 3, (int64_t)L, fc,
                    first ? nullptr : tile_done, ctx->ikeys[0].as<uint32_t>(), ctx->ivals[0].as<uint32_t>(), tile_plan, hdr_tile, s);
        st.launches += (L ? 1 : 0);
        // stable partition of the instances by tile -> tile ranges
        ctx->inst_buf = radix_sort_pairs(ctx->ikeys[0].as<uint32_t>(), ctx->ivals[0].as<uint32_t>(),
                                         ctx->ikeys[1].as<uint32_t>(), ctx->ivals[1].as<uint32_t>(), (size_t)D, cc + 3, 0, tile_bits,
                                         hdr_tile, true, true, ctx->lookback.as<unsigned long long>(), next_epoch(ctx), err_flag, s,
                                         &st.launches);
        launch_tile_ranges(ctx->ikeys[ctx->inst_buf].as<uint32_t>(), D, cc + 3, ctx->ranges.as<uint2>(), num_tiles, s);
        st.launches += (D ? 1 : 0);
        if (tm) CU(cudaEventRecord(ctx->evc[c][3], s));
        launch_blend(ctx->recs.as<Record>(), ctx->ivals[ctx->inst_buf].as<uint32_t>(), ctx->ranges.as<uint2>(), fb, fb_final, fc,
                     first ? 1 : 0, last ? 1 : 0, tile_done, tile_consumed, cnt + 1, cnt + 3,
                     zdepth, scene_depth, s);
        st.launches += 1;
        if (tm) CU(cudaEventRecord(ctx->evc[c][4], s));
    }
    // Feedback on the first chunk's size (auto mode).  Too small a first chunk leaves most tiles unsaturated and the LAST
    // chunk then selects a large part of the cloud (20 M / 1080p: first = V/32 -> 6.2 M live splats instead of 2.8 M); too
    // large a first chunk sorts, shades and bins splats nobody sees.  Balance: the last chunk's live count should stay
    // between a quarter and twice the earlier chunks' sum; outside that band the shift moves by one for the next frame.
    // Any value gives the same frame, so this only steers cost, never pixels.
    if (auto_chunks && ctx->chunk_shift == 0 && nchunks > 1) {
        uint64_t early = 0;
        for (int c = 0; c + 1 < nchunks; ++c) early += L_chunk[c];
        const uint64_t lastL = chunks_run == nchunks ? L_chunk[nchunks - 1] : 0;     // stopped early: every tile saturated
        if (lastL > 2 * early && ctx->auto_shift > 1) --ctx->auto_shift;
        else if (early > 4 * lastL && ctx->auto_shift < 10) ++ctx->auto_shift;
    }
    nchunks = chunks_run;
    CU(cudaGetLastError());
    ctx->evc_chunks = nchunks; ctx->chunks_last = nchunks;
    if (tm) CU(cudaEventRecord(ctx->ev[EV_BLEND], s));
    // D_c, the sort error flag and every chunk's exact D: to pinned memory, read after the next synchronisation
    CU(cudaMemcpyAsync(ctx->counters_h, cnt, 64, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ctx->counters_h + 16, cnt + 8, (size_t)MAX_CHUNKS * 64, cudaMemcpyDeviceToHost, s));

    if (target && target->host_rgba) {
        if (!host_is_final) CU(cudaMemcpyAsync(target->host_rgba, fb_final, fb_bytes, cudaMemcpyDeviceToHost, s));
        if (tm) CU(cudaEventRecord(ctx->ev[EV_COPY], s));
        CU(cudaStreamSynchronize(s));                  // the call returns when the frame is in host memory
        if (ctx->counters_h[2] != 0ull)                // bounded look-back spin of a radix-sort pass timed out: the frame is not valid
            return fail(GSB_ERR_CUDA, "radix sort look-back timed out (internal error); the delivered frame is invalid");
    } else if (tm) CU(cudaEventRecord(ctx->ev[EV_COPY], s));
    if (target && target->gl_texture != 0) {
        // hand the frame to the viewport without a host round trip: device->device copy into the mapped
        // RGBA32F texture; the shim then draws it with the reference's blend state (R.C:613-621).  Needs a current GL context.
        if (ctx->gl_res && (ctx->gl_tex != target->gl_texture || ctx->gl_w != fr->width || ctx->gl_h != fr->height)) {
            cudaGraphicsUnregisterResource(ctx->gl_res); ctx->gl_res = nullptr;
        }
        if (!ctx->gl_res) {
            CU(cudaGraphicsGLRegisterImage(&ctx->gl_res, target->gl_texture, 0x0DE1u /* GL_TEXTURE_2D */,
                                           cudaGraphicsRegisterFlagsWriteDiscard));
            ctx->gl_tex = target->gl_texture; ctx->gl_w = fr->width; ctx->gl_h = fr->height;
        }
        cudaArray_t arr = nullptr;
        CU(cudaGraphicsMapResources(1, &ctx->gl_res, s));
        CU(cudaGraphicsSubResourceGetMappedArray(&arr, ctx->gl_res, 0, 0));
        CU(cudaMemcpy2DToArrayAsync(arr, 0, 0, fb_final, (size_t)fr->width * 16, (size_t)fr->width * 16, (size_t)fr->height,
                                    cudaMemcpyDeviceToDevice, s));
        CU(cudaGraphicsUnmapResources(1, &ctx->gl_res, s));
    }
    ctx->ev_valid = tm;

    uint64_t D_last = 0;                  // exact instance count of the last chunk (its buffers are the ones still around)
    if (ctx->keep_intermediates) {        // debug views of the last chunk: records by splat index, instances as splat indices
        CU(cudaStreamSynchronize(s));
        D_last = nchunks > 0 ? ctx->counters_h[16 + (size_t)(nchunks - 1) * 8 + 3] : 0;
        CU(ctx->dbg_recs.ensure(N * sizeof(Record) + 16)); CU(ctx->dbg_inst.ensure((size_t)D_last * 4 + 16));
        CU(cudaMemsetAsync(ctx->dbg_recs.p, 0, N * sizeof(Record), s));
        launch_debug_views(ctx->recs.as<Record>(), ctx->lvals[ctx->order_vals_buf].as<uint32_t>(), (int64_t)L, ctx->dbg_recs.as<Record>(),
                           ctx->ivals[ctx->inst_buf].as<uint32_t>(), D_last, ctx->dbg_inst.as<uint32_t>(), s);
        CU(cudaGetLastError());
    }
    ctx->last_live = (int64_t)L;
    ctx->last_n = n; ctx->last_sorted = (int64_t)L; ctx->last_d = D_last; ctx->last_tiles = num_tiles; ctx->last_w = fr->width; ctx->last_h = fr->height;
    ctx->last_fb = fb_final;
    ctx->last_fc = fc; ctx->last_lazy = lazy;
    st.depth_chunks = nchunks;
    st.rendered = 1; st.n_submitted = n; st.n_visible = (int64_t)V; st.n_instances = 0 /* gsb_get_stats */; st.n_live = (int64_t)L_total;
    st.sh_order_used = fc.sh_order; st.width = fr->width; st.height = fr->height;
    st.tiles_x = fc.tiles_x; st.tiles_y = fc.tiles_y;
    memcpy(st.camera, fc.cam, 12); memcpy(st.origin, fc.origin, 12);
    return GSB_OK;
} GSB_CATCH_ALL

// ------------------------------------------------------------------ wire pass of GR_PrimGsplat::render, GR.C:474-483 (SURVEY f-4)
int gsb_render_wireframe(gsb_context* ctx, const char* id, const gsb_frame* fr, const gsb_wire_target* target)
try {
    if (!ctx || !id || !fr) return fail(GSB_ERR_INVALID, "gsb_render_wireframe: NULL argument");
    if (fr->width < 1 || fr->height < 1 || fr->width > 65535 || fr->height > 65535)
        return fail(GSB_ERR_LIMIT, "gsb_render_wireframe: screen size must be 1..65535");
    auto it = ctx->registry.find(id);
    if (it == ctx->registry.end()) return fail(GSB_ERR_NOT_FOUND, "gsb_render_wireframe: unknown id");
    const Entry& e = *it->second;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t n = (size_t)e.count;
    FrameConsts fc{};
    memcpy(fc.view, fr->view, 64); memcpy(fc.proj, fr->proj, 64); memcpy(fc.object, fr->object, 64);
    memcpy(fc.inv_object, fr->inv_object, 64); memcpy(fc.obj_view, fr->obj_view, 64);
    fc.width = fr->width; fc.height = fr->height; fc.W = (float)fr->width; fc.H = (float)fr->height;
    float4* verts = target && target->device_vertices ? static_cast<float4*>(target->device_vertices) : nullptr;
    if (!verts) { CU(ctx->wire_verts.ensure(n * 8 * 16 + 16)); verts = ctx->wire_verts.as<float4>(); }
    float* cols = target && target->device_colors ? static_cast<float*>(target->device_colors) : nullptr;
    if (!cols && target && target->host_colors) { CU(ctx->wire_cols.ensure(n * 8 * 12 + 16)); cols = ctx->wire_cols.as<float>(); }
    launch_wire_vertices(fc, e.pos.as<float>(), e.cd.as<uint16_t>(), e.scale.as<uint16_t>(), e.orient.as<uint16_t>(), (int64_t)n,
                         verts, cols, s);
    if (target && target->overlay_rgba) {
        CU(ctx->wire_owner.ensure((size_t)fr->width * fr->height * 8));
        launch_wire_overlay(verts, e.cd.as<uint16_t>(), (int64_t)n, fr->width, fr->height, ctx->wire_owner.as<unsigned long long>(),
                            static_cast<float4*>(target->overlay_rgba), s);
    }
    CU(cudaGetLastError());
    bool sync = false;
    if (target && target->host_vertices && n) { CU(cudaMemcpyAsync(target->host_vertices, verts, n * 8 * 16, cudaMemcpyDeviceToHost, s)); sync = true; }
    if (target && target->host_colors && n) { CU(cudaMemcpyAsync(target->host_colors, cols, n * 8 * 12, cudaMemcpyDeviceToHost, s)); sync = true; }
    if (target && target->overlay_host_rgba && target->overlay_rgba) {
        CU(cudaMemcpyAsync(target->overlay_host_rgba, target->overlay_rgba, (size_t)fr->width * fr->height * 16, cudaMemcpyDeviceToHost, s));
        sync = true;
    }
    if (sync) CU(cudaStreamSynchronize(s));
    return GSB_OK;
} GSB_CATCH_ALL

void* gsb_wire_device_vertices(gsb_context* ctx) { return ctx ? ctx->wire_verts.p : nullptr; }

// ------------------------------------------------------------------ postRender, R.C:660-678
int gsb_post_render(gsb_context* ctx)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    for (auto& kv : ctx->registry) {
        Entry& e = *kv.second;
        if (e.active) e.age_since_last_active = 0;
        else if (e.age_since_last_active > -1) ++e.age_since_last_active;
        e.active = false;
        ++e.age;
    }
    ctx->explicit_cam_set = false;
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_get_stats(gsb_context* ctx, gsb_stats* out)
try {
    if (!ctx || !out) return fail(GSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    gsb_stats& st = ctx->stats;
    if (st.rendered) {
        st.n_consumed = (int64_t)ctx->counters_h[1];
        unsigned long long d = 0;                                      // exact instance counts of the chunks, summed
        for (int c = 0; c < ctx->chunks_last; ++c) d += ctx->counters_h[16 + (size_t)c * 8 + 3];
        st.n_instances = (int64_t)d;
    }
    if (st.rendered && ctx->counters_h[2] != 0ull)
        return fail(GSB_ERR_CUDA, "radix sort look-back timed out (internal error); the last frame is invalid");
    st.ms_project = st.ms_sort = st.ms_bin = st.ms_blend = st.ms_copy = st.ms_total = st.ms_records = 0.0f;
    if (st.rendered && ctx->ev_valid) {
        cudaEventElapsedTime(&st.ms_project, ctx->ev[EV_START], ctx->ev[EV_PROJECT]);
        cudaEventElapsedTime(&st.ms_sort, ctx->ev[EV_PROJECT], ctx->ev[EV_SORT]);
        for (int c = 0; c < ctx->evc_chunks; ++c) {                        // summed over depth chunks
            float t[4] = { 0.f, 0.f, 0.f, 0.f };
            for (int e = 0; e < 4; ++e) cudaEventElapsedTime(&t[e], ctx->evc[c][e], ctx->evc[c][e + 1]);
            st.ms_sort += t[0]; st.ms_records += t[1]; st.ms_bin += t[2]; st.ms_blend += t[3];
        }
        cudaEventElapsedTime(&st.ms_copy, ctx->ev[EV_BLEND], ctx->ev[EV_COPY]);
        cudaEventElapsedTime(&st.ms_total, ctx->ev[EV_START], ctx->ev[EV_COPY]);
    }
    *out = st;
    return GSB_OK;
} GSB_CATCH_ALL

void* gsb_device_framebuffer(gsb_context* ctx) { return ctx ? (void*)ctx->last_fb : nullptr; }

int gsb_ipc_export_frame(gsb_context* ctx, int32_t width, int32_t height, unsigned char handle_out[GSB_IPC_HANDLE_BYTES],
                         void** local_ptr_out)
try {
    if (!ctx || !handle_out || !local_ptr_out || width < 1 || height < 1) return fail(GSB_ERR_INVALID, "gsb_ipc_export_frame: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == GSB_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(ctx->device));
    CU(ctx->shared_frame.ensure((size_t)width * height * 16));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->shared_frame.p));
    memcpy(handle_out, &h, GSB_IPC_HANDLE_BYTES);
    *local_ptr_out = ctx->shared_frame.p;
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_ipc_open(gsb_context* ctx, const unsigned char handle[GSB_IPC_HANDLE_BYTES], void** peer_ptr_out)
try {
    if (!ctx || !handle || !peer_ptr_out) return fail(GSB_ERR_INVALID, "gsb_ipc_open: NULL argument");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h; memcpy(&h, handle, GSB_IPC_HANDLE_BYTES);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));     // maps the peer GPU's memory (NVLink P2P)
    *peer_ptr_out = p;
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_ipc_close(gsb_context* ctx, void* peer_ptr)
try {
    if (!ctx || !peer_ptr) return fail(GSB_ERR_INVALID, "gsb_ipc_close: NULL argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaIpcCloseMemHandle(peer_ptr));
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_host_register(gsb_context* ctx, void* host_ptr, uint64_t bytes)
try {
    if (!ctx || !host_ptr || !bytes) return fail(GSB_ERR_INVALID, "gsb_host_register: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_host_unregister(gsb_context* ctx, void* host_ptr)
try {
    if (!ctx || !host_ptr) return fail(GSB_ERR_INVALID, "gsb_host_unregister: NULL argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaHostUnregister(host_ptr));
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_copy_to_host(gsb_context* ctx, const void* device_ptr, void* host_ptr, uint64_t bytes)
try {
    if (!ctx || !device_ptr || !host_ptr) return fail(GSB_ERR_INVALID, "gsb_copy_to_host: NULL argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(host_ptr, device_ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_debug_fetch(gsb_context* ctx, int which, void* dst, uint64_t dst_bytes, uint64_t* bytes_needed)
try {
    if (!ctx) return fail(GSB_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const void* src = nullptr; uint64_t need = 0;
    const uint64_t n = (uint64_t)ctx->last_n, d = ctx->last_d, t = (uint64_t)ctx->last_tiles;
    switch (which) {
    case GSB_DBG_KEYS_UNSORTED: src = ctx->keys.p; need = n * 4; break;
    case GSB_DBG_ORDER:         src = ctx->lvals[ctx->order_vals_buf].p; need = (uint64_t)ctx->last_sorted * 4; break;
    case GSB_DBG_KEYS_SORTED:   src = ctx->lkeys[ctx->order_buf].p; need = (uint64_t)ctx->last_sorted * 4; break;
    case GSB_DBG_RECORDS:       src = ctx->dbg_recs.p; need = ctx->keep_intermediates ? n * sizeof(Record) : 0; break;
    case GSB_DBG_RECTS:         src = ctx->rects.p; need = ctx->keep_intermediates ? n * 8 : 0; break;
    case GSB_DBG_TILE_RANGES:   src = ctx->ranges.p; need = t * 8; break;
    case GSB_DBG_INSTANCES:     src = ctx->dbg_inst.p; need = ctx->keep_intermediates ? d * 4 : 0; break;
    case GSB_DBG_FRAMEBUFFER:   src = ctx->last_fb; need = (uint64_t)ctx->last_w * ctx->last_h * 16; break;
    case GSB_DBG_TILE_CONSUMED: src = ctx->last_tile_consumed; need = t * 4; break;
    case GSB_DBG_TRECTS:        src = ctx->trects.p; need = (ctx->last_w <= 8192 && ctx->last_h <= 8192) ? n * 4 : 0; break;
    default: return fail(GSB_ERR_INVALID, "gsb_debug_fetch: unknown buffer");
    }
    if (bytes_needed) *bytes_needed = need;
    if (dst && need && ctx->last_lazy && (which == GSB_DBG_KEYS_UNSORTED || which == GSB_DBG_TRECTS) && n) {
        // the bounded K1 evaluates its bound only inside the cells a chunk selects; the debug view computes it for EVERY
        // splat with the same device function (bound_one) and the last frame's constants, by original index
        CU(ctx->keys.ensure(n * 4)); CU(ctx->trects.ensure(n * 4));
        launch_project_bound_debug(ctx->last_fc, ctx->geomA_p.as<float4>(), ctx->lam_p.as<float>(), ctx->orig.as<uint32_t>(), (int64_t)n,
                                   ctx->keys.as<uint32_t>(), ctx->trects.as<uint32_t>(), ctx->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));
        src = which == GSB_DBG_KEYS_UNSORTED ? ctx->keys.p : ctx->trects.p;
    }
    if (dst && need) {
        if (dst_bytes < need) return fail(GSB_ERR_INVALID, "gsb_debug_fetch: destination too small");
        if (!src) return fail(GSB_ERR_INVALID, "gsb_debug_fetch: buffer not available");
        CU(cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
    }
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_debug_sort_pairs(gsb_context* ctx, const uint32_t* keys, const uint32_t* vals, uint64_t n,
                         int begin_bit, int end_bit, uint32_t* keys_out, uint32_t* vals_out)
try {
    if (!ctx || (n && (!keys || !vals || !keys_out || !vals_out))) return fail(GSB_ERR_INVALID, "NULL argument");
    if (begin_bit < 0 || end_bit > 32 || end_bit < begin_bit) return fail(GSB_ERR_INVALID, "bad bit range");
    CU(cudaSetDevice(ctx->device));
    DevBuf k[2], v[2], hdr;
    for (int b = 0; b < 2; ++b) { CU(k[b].ensure(n * 4 + 16)); CU(v[b].ensure(n * 4 + 16)); }
    CU(hdr.ensure(sort_header_bytes()));
    {
        int rc2 = ensure_lookback(ctx, (size_t)n);
        if (rc2) return rc2;
    }
    cudaStream_t s = ctx->stream;
    CU(cudaMemcpyAsync(k[0].p, keys, n * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(v[0].p, vals, n * 4, cudaMemcpyHostToDevice, s));
    int launches = 0;
    int r = radix_sort_pairs(k[0].as<uint32_t>(), v[0].as<uint32_t>(), k[1].as<uint32_t>(), v[1].as<uint32_t>(), n, nullptr,
                             begin_bit, end_bit, hdr.as<uint32_t>(), false, false, ctx->lookback.as<unsigned long long>(),
                             next_epoch(ctx), nullptr, s, &launches);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(keys_out, k[r].p, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(vals_out, v[r].p, n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return GSB_OK;
} GSB_CATCH_ALL

int gsb_debug_exclusive_scan(gsb_context* ctx, const uint32_t* in, uint64_t n, uint32_t* out, uint64_t* total)
try {
    if (!ctx || (n && (!in || !out))) return fail(GSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    DevBuf a, b, scr, tot, status;
    CU(a.ensure(n * 4 + 16)); CU(b.ensure(n * 4 + 16)); CU(scr.ensure(scan_scratch_bytes(n))); CU(tot.ensure(16));
    CU(status.ensure(scan_status_bytes(n)));
    cudaStream_t s = ctx->stream;
    CU(cudaMemcpyAsync(a.p, in, n * 4, cudaMemcpyHostToDevice, s));
    int launches = 0;
    // inputs whose total stays below 2^30 go through the single-pass scan the frame uses, anything larger through the
    // three-launch scan (64-bit prefixes)
    unsigned long long sum = 0;
    for (uint64_t i = 0; i < n; ++i) sum += in[i];
    if (sum < (1ull << 30)) {
        CU(cudaMemsetAsync(status.p, 0, status.cap, s));
        exclusive_scan_u32_onepass(a.as<uint32_t>(), b.as<uint32_t>(), n, status.as<unsigned long long>(), 1u, tot.as<unsigned long long>(),
                                   nullptr, s, &launches);
    } else exclusive_scan_u32(a.as<uint32_t>(), b.as<uint32_t>(), n, scr.p, tot.as<unsigned long long>(), s, &launches);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, b.p, n * 4, cudaMemcpyDeviceToHost, s));
    unsigned long long t = 0;
    CU(cudaMemcpyAsync(&t, tot.p, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (total) *total = t;
    return GSB_OK;
} GSB_CATCH_ALL

}  // extern "C"

// rr_api.cu — C ABI (include/rr.h) over the sm_100a kernels in rr_kernels.cuh.
// Replaces the reference's OpenCL layer for the draw path: ocl.h / ocl.cpp / clstate.* (context, queue, program),
// engine.cpp's render_tris / generate_realtime_shadowing launch code, and cl_gl_interop_texture.hpp (colour targets).
// There is no CPU fallback anywhere in this file: without a usable CUDA device rr_create() fails.
#include "rr_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace rr;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return fail(e__ == cudaErrorMemoryAllocation ? RR_ERR_OOM : RR_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

RotSC make_rotsc(float rx, float ry, float rz) {      // native_sin / native_cos pinned: double on the host, rounded to float
    RotSC r;
    r.sx = (float)sin((double)rx); r.sy = (float)sin((double)ry); r.sz = (float)sin((double)rz);
    r.cx = (float)cos((double)rx); r.cy = (float)cos((double)ry); r.cz = (float)cos((double)rz);
    return r;
}

FaceTable make_face_table() {                         // r_struct, cl2.cl:4487-4511 / 2538-2558 (float arithmetic on M_PI = 3.1415927f)
    const float PI = RR_PI_F;
    const float e[6][3] = {{0, 0, 0}, {PI / 2.0f, 0, 0}, {0, PI, 0}, {3.0f * PI / 2.0f, 0, 0}, {0, 3.0f * PI / 2.0f, 0}, {0, PI / 2.0f, 0}};
    FaceTable t;
    for (int k = 0; k < 6; k++) t.r[k] = make_rotsc(e[k][0], e[k][1], e[k][2]);
    return t;
}

// largest float t with sqrtf(t) <= depth_far (IEEE sqrt is monotone, so sqrtf(d2) > depth_far <=> d2 > t)
float far2_threshold() {
    float t = RR_DEPTH_FAR * RR_DEPTH_FAR;
    while (sqrtf(t) > RR_DEPTH_FAR) t = nextafterf(t, 0.f);
    while (sqrtf(nextafterf(t, INFINITY)) <= RR_DEPTH_FAR) t = nextafterf(t, INFINITY);
    return t;
}

enum { EV_SH0, EV_SH1, EV_F0, EV_SETUP, EV_DEPTH, EV_IDS, EV_SHADE, EV_COUNT };

}  // namespace

// The device scene that is NOT being rendered: object_context::build(async) fills it on the upload stream while frames keep
// coming from the front scene (the flat fields of rr_ctx), and rr_scene_build_commit swaps the two (flip_buffers,
// object_context.cpp:520-590). Same fields as the front scene; buffers are reused when they are large enough.
struct SceneBuf {
    uint32_t n_tris = 0, n_objs = 0, tri_cap = 0, obj_cap = 0;
    rr_triangle* d_tris = nullptr;
    float4 *d_pa = nullptr, *d_pb = nullptr;
    float2* d_pc = nullptr;
    rr_obj_desc* d_objs = nullptr;
    ObjLite* d_objlite = nullptr;
    uint32_t* d_obj_r2 = nullptr;
    int2* d_obj_rows = nullptr;
    uint32_t n_clusters = 0;
    ClusterBox* d_clusters = nullptr;
    uint8_t* d_cluster_vis = nullptr;
    uint4* d_cluster_faces = nullptr;
    uint32_t *d_active = nullptr, *d_skipped = nullptr;
    float4 *d_cutdown = nullptr, *d_scutdown = nullptr;
    uint32_t cap_cut = 0;
    unsigned long long* d_lookback = nullptr;
    uint32_t lookback_blocks = 0;
    rr_obj_desc* h_objs_pinned = nullptr;
};

struct rr_ctx {
    rr_config cfg;
    float fov = 0;
    int W = 0, H = 0, L = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    bool have_shadow_ev = false, have_frame_ev = false;
    // scene
    uint32_t n_tris = 0, n_objs = 0;
    uint32_t tri_cap = 0, obj_cap = 0;           // what the scene buffers were sized for (they swap with the back scene, rr_scene_build_*)
    rr_triangle* d_tris = nullptr;
    float4 *d_pa = nullptr, *d_pb = nullptr;
    float2* d_pc = nullptr;
    rr_obj_desc* d_objs = nullptr;
    ObjLite* d_objlite = nullptr;
    uint32_t* d_obj_r2 = nullptr;                // per-object bounding radius^2 (float bits), object space
    int2* d_obj_rows = nullptr;                  // per-object screen rows of this frame (band mode)
    bool objlite_dirty = true;
    // atlas
    uchar4* d_atlas = nullptr;
    cudaTextureObject_t atlas_tex = 0;           // optional fetch path of kernel3 (RR_TEX_OBJECTS=1): point-sampled pitch-2D view of d_atlas
    bool use_tex_objects = true;                 // measured equal or slightly faster than ld.global.nc (profiles/r2_texobj_ab.txt); RR_TEX_OBJECTS=0 selects the plain loads
    size_t atlas_texels = 0;
    uint32_t *d_nums = nullptr, *d_sizes = nullptr;
    uint32_t n_nums = 0, n_sizes = 0, mipmap_start = 0;
    uchar4* d_upload = nullptr;
    size_t upload_cap = 0;
    // lights
    std::vector<rr_light> lights;
    rr_light* d_lights = nullptr;
    LightLite* d_lightlite = nullptr;
    uint32_t n_shadow = 0, n_static = 0;
    uint32_t *d_shadow_dyn = nullptr, *d_shadow_static = nullptr;
    // second dynamic cubemap buffer (context-owned buffers only): while frame n shades from one, the other — last read by frame
    // n-1 — is cleared behind frame n's shadow pass on the shadow stream, so the clear (engine.cpp:1615) is off the critical path
    uint32_t* d_shadow_alt = nullptr;
    bool alt_clean = false;
    bool ext_shadow_dyn = false, ext_shadow_static = false;
    size_t shadow_dyn_words = 0, shadow_static_words = 0;
    // frame targets
    uint32_t* d_depth[2] = {nullptr, nullptr};
    uint32_t* d_ids[2] = {nullptr, nullptr};
    int cur = 0;
    uchar4* d_rgba8 = nullptr;          // colour target of the frame being / last drawn
    // ring of colour targets for pipelined read-back (rr_frame_e2e), like async_read.hpp's ring of host buffers:
    // d_ring[0] is the context's own target; the others are allocated when first used
    uchar4* d_ring[RR_RING_MAX] = {};
    int ring_depth = 2, ring_pos = 0;
    bool ext_rgba8 = false;
    cudaStream_t stream3 = nullptr;     // copy stream
    cudaEvent_t ev_draw_done = nullptr, ev_copy_done[RR_RING_MAX] = {};
    bool copy_pending[RR_RING_MAX] = {};
    ushort2* d_normals = nullptr;
    uchar4* d_post = nullptr;                    // second colour target of the post passes (rr_post_*)
    // what the post passes are handed besides the G-buffer: the camera of the last rr_frame_draw, the one before it
    // (object_context_data::c_pos_old, set at swap_buffers, object_context.cpp:23-24) and object_context_data::frame_id (engine.cpp:2024)
    CamParams cam_last = {}, cam_old = {};
    uint32_t frame_id = 0;
    uint8_t* d_seen = nullptr;                   // per object: visible in the frame being blurred (do_motion_blur's history update)
    uint32_t seen_cap = 0;
    uint32_t* d_shade_list = nullptr;            // compacted covered pixels
    uint2* d_samples = nullptr;                  // covered samples of inline-rasterised triangles (pixel, depth)
    uint32_t* d_sample_frag = nullptr;           // ... and the fragment index of each sample
    uint32_t cap_samples = 0;
    // raster storage
    uint32_t* d_frags = nullptr;
    uint32_t cap_frags = 0;
    float4* d_cutdown = nullptr;
    uint32_t cap_cut = 0;
    uint32_t* d_counters = nullptr;
    unsigned long long* d_lookback = nullptr;
    uint32_t lookback_blocks = 0;
    uint32_t* d_fragcnt = nullptr;               // per-fragment pixel-slot counts
    uint32_t *d_worklist = nullptr, *d_extra = nullptr;    // fragments kernel1 / kernel2 still have to walk (k_setup_main -> k_raster_warp_depth / k_ids_list)
    uint32_t* h_counters = nullptr;      // pinned (main counters, then shadow counters)
    // second raster workspace + stream: the shadow passes of a frame run concurrently with the main view's setup / depth /
    // id kernels (they only meet at shading). The reference shares g_tid_buf / g_cut_tri_mem between them and serialises.
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_shadow_done = nullptr;
    // side stream of kernel3's streaming stores (k_clear_next): forked at the top of rr_frame_draw, joined in front of the shading list
    cudaStream_t stream5 = nullptr;
    cudaEvent_t ev_fork_clear = nullptr, ev_clear_done = nullptr;
    // dirty-tile read-back of rr_frame_e2e (rr_set_readback_tiles): per ring slot, the tiles shaded in the frame being copied (now) and
    // in the frame last written into the slot's host buffer (prev), that buffer and the clear colour it was filled with
    bool tiles_on = false;
    uint8_t* d_tile_now[RR_RING_MAX] = {};
    uint8_t* d_tile_prev[RR_RING_MAX] = {};
    const void* tile_host[RR_RING_MAX] = {};
    uint64_t tile_seq[RR_RING_MAX] = {}, tile_calls = 0;   // when each slot last wrote its host buffer (a buffer that another slot has written since is stale for this one)
    float tile_clear[RR_RING_MAX][4] = {};
    bool tile_valid[RR_RING_MAX] = {};
    uint32_t* d_tile_sent = nullptr;             // [0]: tiles stored by k_tile_copy since rr_readback_tile_bytes; [1 + k]: tiles of slot k's last copy
    uint32_t* h_tile_last = nullptr;             // pinned: [k] = tiles that slot k's last copy needed (what sizes the next grids)
    uint32_t tile_estimate = 0;
    uint64_t tile_dma_bytes = 0;                 // whole frames copied by the copy engine while the tile mode is on (statistics)
    uint8_t* tile_mark_target = nullptr;         // set by rr_frame_e2e around rr_frame_draw: where the id resolve marks the frame's tiles
    bool tiles_marked = false;                   // ... and whether it did (otherwise k_tile_mark runs over the covered-pixel list)
    bool swapped = true;                         // rr_swap_buffers since the last rr_frame_draw: this frame's id image starts all zero
    // A/B knobs, read at rr_create
    bool split_clear = true;                     // RR_SPLIT_CLEAR=0: the stores stay in k_shade_pre4 on the main stream
    bool list_from_ids = true;                   // RR_LIST_FROM_IDS=0: the covered-pixel list comes from a pass over the screen (k_shade_list)
    int clear_at = 0;                            // RR_CLEAR_AT=1: the side stream forks behind k_setup_main instead of at the start of the frame
    int clear_grid = 2;                          // RR_CLEAR_GRID: CTAs per SM of k_clear_next (a small footprint: it runs beside other kernels)
    int tile_grid = 4;                           // RR_TILE_GRID: least number of CTAs of k_tile_copy (see the kernel: few on purpose)
    int raster_grid = 6;                         // RR_RASTER_GRID: CTAs per SM of k_raster_warp_depth (24 KB of sample stash per CTA of eight warps)
    bool shadow_pending = false;
    uint32_t *d_sfrags = nullptr, *d_sfragcnt = nullptr, *d_scounters = nullptr;
    float4* d_scutdown = nullptr;
    // host mirror of the descriptor array (what rr_frame_e2e uploads every frame, object_context::flush_locations)
    rr_obj_desc* h_objs_pinned = nullptr;
    // Page-locked staging for every asynchronous host-to-device upload on the main stream (patches, descriptor writes, the
    // per-frame descriptor upload of rr_frame_e2e): two halves used alternately; a half is reused only after the event recorded
    // behind its last copy has completed, so the host never overwrites bytes a pending DMA still has to read.
    struct Arena { char* base = nullptr; size_t half = 0, used = 0; int cur = 0; cudaEvent_t ev[2] = {nullptr, nullptr}; bool pending[2] = {false, false}; } arena;
    uint32_t last_overflow = 0;                  // what the last rr_sync saw (rr_get_timings reports it)
    // sort-first rows (rr_config.band_*): bounding ranges, and per-row tables in interleaved mode
    int own_lo = 0, own_hi = 0, need_lo = 0, need_hi = 0;
    bool banded = false;
    uint8_t* d_rowmask = nullptr;                // ROW_NEEDED | ROW_OWNED per row (interleaved mode only)
    int* d_rowpfx = nullptr;                     // [H + 1] prefix count of ROW_NEEDED rows
    // clusters of CLUSTER_TRIS consecutive triangles (k_cluster_bounds): per-frame culling units of the setup kernels
    uint32_t n_clusters = 0;
    ClusterBox* d_clusters = nullptr;
    uint8_t* d_cluster_vis = nullptr;            // main view: 0 = culled this frame (k_cluster_vis)
    uint32_t *d_active = nullptr, *d_skipped = nullptr;   // k_frame_prologue: surviving setup blocks, slots skipped in front of each
    bool stage_events = false;                   // rr_set_profiling: per-stage CUDA timing events for rr_get_timings (the reference's -DPROFILING)
    int shadow_pretest = 1;                      // k_shadow_setup's early back-face cull (RR_SHADOW_PRETEST=0 turns it off for A/B runs)
    int cluster_cull = 0;                        // rr_config.cluster_cull: 0 = when the frame is split (sort-first), 1 = always, -1 = never
    uint4* d_cluster_faces = nullptr;            // face sharding: per-cluster cube-face reach of the lights of the pass
    // multi-GPU exchange over peer memory (rr_mgpu_*)
    struct Mg {
        bool exported = false, connected = false, ipc = false;
        bool local_readback = false;                         // rr_mgpu_set_readback: rows stay local and go to the host over this GPU's own PCIe link
        int rank = 0, world = 1;
        uint32_t* shadow[2] = {nullptr, nullptr};            // local double-buffered dynamic cubemaps (one allocation)
        uchar4* fb[RR_RING_MAX] = {};                        // local colour-target ring (rank 0's are the composite targets)
        MgCtrl* ctrl = nullptr;
        size_t shadow_words = 0;                             // per buffer
        uint32_t* peer_shadow[MG_MAX_WORLD][2] = {};
        uchar4* fb0[RR_RING_MAX] = {};                       // rank 0's colour targets as seen from here
        MgCtrl* peer_ctrl[MG_MAX_WORLD] = {};
        void* opened[MG_MAX_WORLD][3] = {};                  // cudaIpcOpenMemHandle results to close
        uint8_t* prev_dirty[2] = {nullptr, nullptr};
        uint32_t shadow_epoch = 0, draw_epoch = 0;
        uint32_t* saved_shadow_dyn = nullptr; bool saved_ext_shadow = false; size_t saved_shadow_words = 0;
        uchar4* saved_rgba8 = nullptr; bool saved_ext_rgba8 = false;
    } mg;
    int mg_target = 0;                           // which of the colour-target pair this draw goes to (rr_frame_e2e alternates)
    // asynchronous scene rebuild (rr_scene_build_*)
    SceneBuf back;
    cudaStream_t stream4 = nullptr;              // upload stream of the rebuild (the reference's cqueue2)
    cudaEvent_t ev_built = nullptr, ev_retire = nullptr;
    bool building = false, retire_pending = false;
    // stats
    uint32_t launches = 0;
    FaceTable faces;
};

static void scene_free_fwd(rr_ctx* c);           // frees the back scene of an asynchronous rebuild (defined with rr_scene_build_*)

namespace {

template <class T>
int dev_alloc(T*& p, size_t count) {
    if (p) { cudaFree(p); p = nullptr; }
    if (count == 0) count = 1;
    CU(cudaMalloc((void**)&p, count * sizeof(T)));
    return RR_OK;
}

inline int grid_for(const rr_ctx* c, int per_sm) { return c->sm_count * per_sm; }

// n bytes of page-locked staging that stay untouched until the copies enqueued on `st` so far plus the one the caller is about
// to enqueue have executed. May block the host when both halves are still in flight (two arena halves of uploads ahead of the GPU).
int arena_take(rr_ctx* c, size_t n, cudaStream_t st, void** out) {
    rr_ctx::Arena& a = c->arena;
    n = (n + 15) & ~(size_t)15;
    if (!a.base || n > a.half) {                                  // first use, or an upload larger than a half: (re)allocate
        if (a.base) { CU(cudaStreamSynchronize(st)); cudaFreeHost(a.base); a.base = nullptr; }
        a.half = std::max<size_t>(n * 2, (size_t)1 << 20);
        CU(cudaMallocHost((void**)&a.base, a.half * 2));
        for (int i = 0; i < 2; i++) { if (!a.ev[i]) CU(cudaEventCreateWithFlags(&a.ev[i], cudaEventDisableTiming)); a.pending[i] = false; }
        a.used = 0; a.cur = 0;
    }
    if (a.used + n > a.half) {                                    // this half is full: everything staged in it is in the stream already
        CU(cudaEventRecord(a.ev[a.cur], st));
        a.pending[a.cur] = true;
        a.cur ^= 1;
        if (a.pending[a.cur]) { CU(cudaEventSynchronize(a.ev[a.cur])); a.pending[a.cur] = false; }
        a.used = 0;
    }
    *out = a.base + (size_t)a.cur * a.half + a.used;
    a.used += n;
    return RR_OK;
}

// host bytes -> device, asynchronously on the main stream, through the staging arena (the caller may reuse `src` on return)
int upload_staged(rr_ctx* c, void* dst_dev, const void* src, size_t n) {
    if (!n) return RR_OK;
    void* h = nullptr;
    int r = arena_take(c, n, c->stream, &h);
    if (r) return r;
    memcpy(h, src, n);
    CU(cudaMemcpyAsync(dst_dev, h, n, cudaMemcpyHostToDevice, c->stream));
    return RR_OK;
}

int fill_u32(rr_ctx* c, cudaStream_t st, uint32_t* p, size_t n, uint32_t v) {
    if (n == 0) return RR_OK;
    k_fill_u32<<<grid_for(c, 8), 256, 0, st>>>(p, n, v);
    c->launches++;
    CU(cudaGetLastError());
    return RR_OK;
}

int ensure_objlite(rr_ctx* c) {
    if (!c->objlite_dirty || c->n_objs == 0) return RR_OK;
    k_objlite<<<(c->n_objs + 127) / 128, 128, 0, c->stream>>>(c->d_objs, c->n_objs, c->d_objlite);
    c->launches++;
    CU(cudaGetLastError());
    c->objlite_dirty = false;
    return RR_OK;
}

// kernel1 for the work list (k_raster_warp_depth): depth + the covered samples kernel2 streams afterwards
int raster_depth(rr_ctx* c, cudaStream_t st, const RasterParams& rp, const SampleList& sl) {
    k_raster_warp_depth<<<grid_for(c, c->raster_grid), RW_WARPS * 32, 0, st>>>(rp, sl);
    c->launches++;
    CU(cudaGetLastError());
    return RR_OK;
}

void band_rows(const rr_ctx* c, int& band0, int& band1, int& row0, int& row1) {
    band0 = c->own_lo; band1 = c->own_hi; row0 = c->need_lo; row1 = c->need_hi;
}

// Rows this context shades (owned) and rasterises (needed = owned dilated by band_halo rows, for SSAO). Contiguous bands
// are fully described by the two ranges; the interleaved split (band_tile) also gets per-row tables on the device.
int setup_rows(rr_ctx* c) {
    const int H = c->H;
    const rr_config& g = c->cfg;
    std::vector<uint8_t> mask((size_t)H, 0);
    const bool tiled = g.band_tile > 0 && g.band_world > 1;
    if (tiled && (g.band_rank < 0 || g.band_rank >= g.band_world)) return fail(RR_ERR_INVALID, "rr_create: band_rank %d outside band_world %d", g.band_rank, g.band_world);
    const bool ranged = !tiled && g.band_y1 > g.band_y0;
    for (int y = 0; y < H; y++) {
        bool own = true;
        if (tiled) own = (y / g.band_tile) % g.band_world == g.band_rank;
        else if (ranged) own = y >= g.band_y0 && y < g.band_y1;
        if (own) mask[y] |= ROW_OWNED;
    }
    const bool whole = !tiled && !ranged;
    if (whole || g.band_halo < 0) { for (int y = 0; y < H; y++) mask[y] |= ROW_NEEDED; }
    else {
        int last_owned = -0x3FFFFFFF;
        for (int y = 0; y < H; y++) { if (mask[y] & ROW_OWNED) last_owned = y; if (y - last_owned <= g.band_halo) mask[y] |= ROW_NEEDED; }
        int next_owned = 0x3FFFFFFF;
        for (int y = H - 1; y >= 0; y--) { if (mask[y] & ROW_OWNED) next_owned = y; if (next_owned - y <= g.band_halo) mask[y] |= ROW_NEEDED; }
    }
    c->own_lo = c->need_lo = H; c->own_hi = c->need_hi = 0;
    for (int y = 0; y < H; y++) {
        if (mask[y] & ROW_OWNED) { c->own_lo = std::min(c->own_lo, y); c->own_hi = y + 1; }
        if (mask[y] & ROW_NEEDED) { c->need_lo = std::min(c->need_lo, y); c->need_hi = y + 1; }
    }
    if (c->own_hi == 0) { c->own_lo = 0; }                        // owns nothing (more contexts than tiles): empty ranges
    if (c->need_hi == 0) { c->need_lo = 0; }
    c->banded = !whole;
    if (tiled) {
        std::vector<int> pfx((size_t)H + 1, 0);
        for (int y = 0; y < H; y++) pfx[y + 1] = pfx[y] + ((mask[y] & ROW_NEEDED) ? 1 : 0);
        CU(cudaMalloc((void**)&c->d_rowmask, (size_t)H));
        CU(cudaMalloc((void**)&c->d_rowpfx, ((size_t)H + 1) * sizeof(int)));
        CU(cudaMemcpy(c->d_rowmask, mask.data(), (size_t)H, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_rowpfx, pfx.data(), ((size_t)H + 1) * sizeof(int), cudaMemcpyHostToDevice));
    }
    return RR_OK;
}

// (light, face) pair p of `total` is rendered by this context?
bool owns_pair(const rr_ctx* c, uint32_t p, uint32_t total) {
    if (c->cfg.face_world <= 1) return true;
    if (c->cfg.face_interleave) return (int)(p % (uint32_t)c->cfg.face_world) == c->cfg.face_rank;
    const uint32_t chunk = (total + c->cfg.face_world - 1) / c->cfg.face_world;
    return (p / chunk) == (uint32_t)c->cfg.face_rank;
}

// ---- multi-GPU helpers (rr_mgpu_*) ------------------------------------------------------------------------------------
unsigned long long mg_timeout_ns() {
    const char* e = getenv("RR_MGPU_TIMEOUT_MS");
    const double ms = e ? atof(e) : 20000.0;
    return (unsigned long long)(std::max(ms, 1.0) * 1e6);
}

// wait on `st` until every other context's flag (shadow or draw) in the LOCAL control block has reached `value`
int mg_wait(rr_ctx* c, cudaStream_t st, bool shadow, uint32_t value) {
    MgWait w;
    w.n = 0; w.value = value; w.error = &c->mg.ctrl->error; w.timeout_ns = mg_timeout_ns();
    for (int q = 0; q < c->mg.world; q++) {
        if (q == c->mg.rank) continue;
        w.flag[w.n++] = shadow ? &c->mg.ctrl->shadow_flag[q] : &c->mg.ctrl->draw_flag[q];
    }
    if (w.n == 0) return RR_OK;
    k_wait_flags<<<1, 32, 0, st>>>(w);
    c->launches++;
    CU(cudaGetLastError());
    return RR_OK;
}

int mg_owned_pairs(const rr_ctx* c, uint32_t* out) {
    const uint32_t total = 6u * c->n_shadow;
    int n = 0;
    for (uint32_t p = 0; p < total; p++) if (owns_pair(c, p, total)) out[n++] = p;
    return n;
}

// clear the owned faces of the epoch's buffer (the others are delivered by their owners and never cleared here)
int mg_fill_owned(rr_ctx* c, cudaStream_t st, uint32_t* buffer) {
    MgFillParams f;
    f.buffer = buffer; f.face_words = (uint32_t)c->L * (uint32_t)c->L;
    f.n_pairs = mg_owned_pairs(c, f.pair);
    if (f.n_pairs == 0) return RR_OK;
    k_fill_faces<<<grid_for(c, 8), 256, 0, st>>>(f);
    c->launches++;
    CU(cudaGetLastError());
    return RR_OK;
}

// copy the owned faces into every peer's buffer of the same parity and raise this context's flag there
int mg_push(rr_ctx* c, cudaStream_t st, int b, uint32_t epoch) {
    MgPushParams p;
    p.local = c->mg.shadow[b]; p.face_words = (uint32_t)c->L * (uint32_t)c->L; p.prev_dirty = c->mg.prev_dirty[b];
    p.n_pairs = mg_owned_pairs(c, p.pair);
    p.n_peers = 0; p.sig.n = 0; p.sig.counter = &c->mg.ctrl->push_done; p.sig.value = epoch; p.pushed = &c->mg.ctrl->pushed_chunks;
    for (int q = 0; q < c->mg.world; q++) {
        if (q == c->mg.rank) continue;
        p.peer[p.n_peers++] = c->mg.peer_shadow[q][b];
        p.sig.flag[p.sig.n++] = &c->mg.peer_ctrl[q]->shadow_flag[c->mg.rank];
    }
    if (p.n_peers == 0) return RR_OK;
    k_push_faces<<<grid_for(c, 4), 256, 0, st>>>(p);
    c->launches++;
    CU(cudaGetLastError());
    return RR_OK;
}

}  // namespace

// Every kernel of the library is loaded when the first context is created. With CUDA's default lazy module loading the first
// launch of a kernel loads it, and that load can wait for running kernels to finish: a k_wait_flags spinning on a flag whose
// writer (k_signal_flags, k_push_faces ...) has never been launched before would then never be released. Loading up front
// makes the spin-wait protocols independent of launch history.
static int preload_kernels() {
    static bool done = false;
    if (done) return RR_OK;
    const void* fns[] = {
        (const void*)k_repack, (const void*)k_objlite, (const void*)k_lightlite, (const void*)k_obj_rows, (const void*)k_cluster_bounds,
        (const void*)k_frame_prologue, (const void*)k_setup_main<false>, (const void*)k_setup_main<true>,
        (const void*)k_raster_warp_depth,
        (const void*)k_ids_list<true>, (const void*)k_ids_list<false>, (const void*)k_shadow_setup, (const void*)k_cluster_faces,
        (const void*)k_signal_flag, (const void*)k_signal_flags, (const void*)k_wait_flags, (const void*)k_push_faces, (const void*)k_fill_faces,
        (const void*)k_raster_shadow_warp, (const void*)k_fill_u32, (const void*)k_atlas_upload, (const void*)k_atlas_mip, (const void*)k_atlas_upload_batch, (const void*)k_atlas_mip_batch,
        (const void*)k_atlas_fill_colour, (const void*)k_atlas_from_raw,
        (const void*)k_shade_pre, (const void*)k_shade_pre4<true, true>, (const void*)k_shade_pre4<false, true>, (const void*)k_shade_list, (const void*)k_clear_next, (const void*)k_shade, (const void*)k_pseudo_aa, (const void*)k_motion_blur, (const void*)k_motion_history, (const void*)k_godrays, (const void*)k_copy_u32, (const void*)k_tile_mark, (const void*)k_tile_copy,
    };
    for (const void* f : fns) {
        cudaFuncAttributes a;
        CU(cudaFuncGetAttributes(&a, f));
    }
    done = true;
    return RR_OK;
}

extern "C" {

const char* rr_last_error(void) { return g_err; }
const char* rr_version(void) { return "openclrenderer_b200 0.1 (sm_100a)"; }

void rr_default_config(rr_config* cfg) {
    memset(cfg, 0, sizeof *cfg);
    cfg->width = 800; cfg->height = 600; cfg->light_dim = 1024;
    cfg->fov_const = 0.f; cfg->hfov_deg = 120.f;
    cfg->depth_icutoff = 20;
    cfg->ambient = 0.2f; cfg->ssao_rad = 5.f; cfg->ssao_div = 2.5f; cfg->mip_bias = 1.1f;
    cfg->shadow_bias = 50.f; cfg->shadow_exp = 1.f;
    cfg->test_linear = 0; cfg->use_linear_rendering = 1; cfg->no_ssao = 0;
    cfg->device = 0;
    cfg->band_y0 = cfg->band_y1 = 0; cfg->band_halo = -1;
    cfg->face_rank = 0; cfg->face_world = 0;
    cfg->max_fragments = 0; cfg->max_cutdown = 0;
}

float rr_fov_const_from_hfov(float hfov_deg, float screenwidth) {
    // engine.cpp:119-133: float fov_radians = (hfov/360.f)*2*M_PI; fov = (w/2)/tan(fov_radians/2); then the kernel sees
    // std::to_string(fov) + "f" (engine.cpp:474-477), i.e. the value rounded to 6 decimals and re-parsed as float.
    double fr = ((double)(hfov_deg / 360.f) * 2) * M_PI;
    float fov_radians = (float)fr;
    float triangle_angle = fov_radians / 2;
    float fov_constant = (float)((double)(screenwidth / 2) / tan((double)triangle_angle));
    char buf[64];
    snprintf(buf, sizeof buf, "%f", fov_constant);
    return strtof(buf, nullptr);
}

rr_ctx* rr_create(const rr_config* cfg) {
    if (!cfg || cfg->width <= 0 || cfg->height <= 0 || cfg->light_dim <= 0) { fail(RR_ERR_INVALID, "rr_create: bad config"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { fail(RR_ERR_CUDA, "rr_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e)); return nullptr; }
    if (cfg->device < 0 || cfg->device >= ndev) { fail(RR_ERR_INVALID, "rr_create: device %d out of range (%d devices)", cfg->device, ndev); return nullptr; }
    if (cudaSetDevice(cfg->device) != cudaSuccess) { fail(RR_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { fail(RR_ERR_CUDA, "cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major < 10) { fail(RR_ERR_CUDA, "rr_create: device is sm_%d%d; this build carries only sm_100a code", prop.major, prop.minor); return nullptr; }
    if (preload_kernels() != RR_OK) return nullptr;
    rr_ctx* c = new (std::nothrow) rr_ctx();
    if (!c) { fail(RR_ERR_OOM, "host alloc"); return nullptr; }
    c->cfg = *cfg;
    c->W = cfg->width; c->H = cfg->height; c->L = cfg->light_dim;
    c->fov = cfg->fov_const > 0 ? cfg->fov_const : rr_fov_const_from_hfov(cfg->hfov_deg, (float)cfg->width);
    c->sm_count = prop.multiProcessorCount;
    c->faces = make_face_table();
    c->cam_last.pos = c->cam_old.pos = make_float3(0.f, 0.f, 0.f);       // c_pos_old / c_rot_old start at zero (object_context.hpp:89-90): angle 0, not a zero matrix
    c->cam_last.rot = c->cam_old.rot = make_rotsc(0.f, 0.f, 0.f);
    c->cluster_cull = cfg->cluster_cull;
    if (const char* e = getenv("RR_SHADOW_PRETEST")) c->shadow_pretest = atoi(e) != 0;
    if (const char* e = getenv("RR_TEX_OBJECTS")) c->use_tex_objects = atoi(e) != 0;
    auto bail = [&](const char* what) { fail(RR_ERR_CUDA, "rr_create: %s: %s", what, cudaGetErrorString(cudaGetLastError())); rr_destroy(c); return (rr_ctx*)nullptr; };
    // RR_STREAM_PRIO (A/B knob): 1 = the main view's stream outranks the shadow stream, 2 = the other way round, 0 = equal
    int prio_lo = 0, prio_hi = 0, prio_mode = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);            // lo = least urgent (numerically largest)
    if (const char* e = getenv("RR_STREAM_PRIO")) prio_mode = atoi(e);
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_mode == 1 ? prio_hi : prio_lo) != cudaSuccess) return bail("stream");
    for (int i = 0; i < EV_COUNT; i++) if (cudaEventCreate(&c->ev[i]) != cudaSuccess) return bail("event");
    const size_t P = (size_t)c->W * c->H;
    for (int i = 0; i < 2; i++) {
        if (cudaMalloc((void**)&c->d_depth[i], P * 4) != cudaSuccess) return bail("depth");
        if (cudaMalloc((void**)&c->d_ids[i], P * 4) != cudaSuccess) return bail("ids");
    }
    if (cudaMalloc((void**)&c->d_rgba8, P * 4) != cudaSuccess) return bail("rgba8");
    if (cudaMalloc((void**)&c->d_normals, P * 4) != cudaSuccess) return bail("normals");
    if (cudaMalloc((void**)&c->d_shade_list, P * 4) != cudaSuccess) return bail("shade list");
    c->cap_samples = (uint32_t)std::min<size_t>(4 * P + (1u << 20), 0x7FFFFFFFu);
    if (const char* e = getenv("RR_SAMPLE_CAP")) c->cap_samples = std::max<uint32_t>(1u, std::min<uint32_t>(c->cap_samples, (uint32_t)strtoul(e, nullptr, 10)));   // tests: force the `extra` path of kernel2
    if (cudaMalloc((void**)&c->d_samples, (size_t)c->cap_samples * 8) != cudaSuccess) return bail("sample list");
    c->cap_frags = cfg->max_fragments ? cfg->max_fragments : (16u << 20);
    if (cudaMalloc((void**)&c->d_frags, (size_t)c->cap_frags * RR_FRAG_WORDS * 4) != cudaSuccess) return bail("fragment buffer");
    if (cudaMalloc((void**)&c->d_sample_frag, (size_t)c->cap_samples * 4) != cudaSuccess) return bail("sample list");
    if (cudaMalloc((void**)&c->d_fragcnt, (size_t)c->cap_frags * RR_FRAG_WORDS * 4 / RR_SFRAG_WORDS + 16) != cudaSuccess) return bail("slot counts");
    if (cudaMalloc((void**)&c->d_worklist, (size_t)c->cap_frags * 4 + 16) != cudaSuccess) return bail("work list");
    if (cudaMalloc((void**)&c->d_extra, (size_t)c->cap_frags * 4 + 16) != cudaSuccess) return bail("extra list");
    if (cudaMalloc((void**)&c->d_counters, CTR_COUNT * 4) != cudaSuccess) return bail("counters");
    if (cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_mode == 2 ? prio_hi : prio_lo) != cudaSuccess) return bail("stream2");
    if (cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking) != cudaSuccess) return bail("stream3");
    if (cudaEventCreateWithFlags(&c->ev_draw_done, cudaEventDisableTiming) != cudaSuccess) return bail("event");
    for (int i = 0; i < RR_RING_MAX; i++) if (cudaEventCreateWithFlags(&c->ev_copy_done[i], cudaEventDisableTiming) != cudaSuccess) return bail("event");
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) return bail("event");
    if (cudaEventCreateWithFlags(&c->ev_shadow_done, cudaEventDisableTiming) != cudaSuccess) return bail("event");
    if (cudaStreamCreateWithFlags(&c->stream5, cudaStreamNonBlocking) != cudaSuccess) return bail("stream5");
    if (cudaEventCreateWithFlags(&c->ev_fork_clear, cudaEventDisableTiming) != cudaSuccess) return bail("event");
    if (cudaEventCreateWithFlags(&c->ev_clear_done, cudaEventDisableTiming) != cudaSuccess) return bail("event");
    if (const char* e = getenv("RR_SPLIT_CLEAR")) c->split_clear = atoi(e) != 0;
    if (const char* e = getenv("RR_LIST_FROM_IDS")) c->list_from_ids = atoi(e) != 0;
    if (const char* e = getenv("RR_CLEAR_AT")) c->clear_at = atoi(e);
    if (const char* e = getenv("RR_CLEAR_GRID")) c->clear_grid = std::max(1, atoi(e));
    if (const char* e = getenv("RR_RASTER_GRID")) c->raster_grid = std::max(1, atoi(e));
    if (const char* e = getenv("RR_TILE_GRID")) c->tile_grid = std::max(1, atoi(e));
    {
        const size_t srec = (size_t)c->cap_frags * RR_FRAG_WORDS / RR_SFRAG_WORDS;     // shadow records are 4 words
        if (cudaMalloc((void**)&c->d_sfrags, (size_t)c->cap_frags * RR_FRAG_WORDS * 4) != cudaSuccess) return bail("shadow fragment buffer");
        if (cudaMalloc((void**)&c->d_sfragcnt, srec * 4 + 16) != cudaSuccess) return bail("shadow slot counts");
        if (cudaMalloc((void**)&c->d_scounters, CTR_COUNT * 4) != cudaSuccess) return bail("shadow counters");
        cudaMemsetAsync(c->d_scounters, 0, CTR_COUNT * 4, c->stream);
    }
    if (cudaMallocHost((void**)&c->h_counters, 2 * CTR_COUNT * 4) != cudaSuccess) return bail("pinned counters");
    cudaMemsetAsync(c->d_counters, 0, CTR_COUNT * 4, c->stream);
    // depth_buffer[0..1] start at UINT_MAX (object_context.cpp:43-52); the id image starts at 0
    for (int i = 0; i < 2; i++) {
        cudaMemsetAsync(c->d_depth[i], 0xFF, P * 4, c->stream);
        cudaMemsetAsync(c->d_ids[i], 0, P * 4, c->stream);
    }
    cudaMemsetAsync(c->d_rgba8, 0, P * 4, c->stream);
    cudaMemsetAsync(c->d_normals, 0, P * 4, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return bail("init sync");
    if (setup_rows(c) != RR_OK) { rr_destroy(c); return nullptr; }
    return c;
}

void rr_destroy(rr_ctx* c) {
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    if (c->stream3) cudaStreamSynchronize(c->stream3);
    rr_mgpu_disconnect(c);
    if (c->stream4) { cudaStreamSynchronize(c->stream4); cudaStreamDestroy(c->stream4); }
    if (c->ev_built) cudaEventDestroy(c->ev_built);
    if (c->ev_retire) cudaEventDestroy(c->ev_retire);
    scene_free_fwd(c);
    for (int i = 1; i < RR_RING_MAX; i++) if (c->d_ring[i]) cudaFree(c->d_ring[i]);
    if (!c->ext_rgba8 && c->d_ring[0]) c->d_rgba8 = c->d_ring[0];      // the context's own target (freed below)
    if (c->ev_draw_done) cudaEventDestroy(c->ev_draw_done);
    for (int i = 0; i < RR_RING_MAX; i++) if (c->ev_copy_done[i]) cudaEventDestroy(c->ev_copy_done[i]);
    if (c->stream3) cudaStreamDestroy(c->stream3);
    cudaFree(c->d_sfrags); cudaFree(c->d_sfragcnt); cudaFree(c->d_scounters);
    cudaFree(c->d_scutdown);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_shadow_done) cudaEventDestroy(c->ev_shadow_done);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream5) { cudaStreamSynchronize(c->stream5); cudaStreamDestroy(c->stream5); }
    for (int i = 0; i < RR_RING_MAX; i++) { cudaFree(c->d_tile_now[i]); cudaFree(c->d_tile_prev[i]); }
    cudaFree(c->d_tile_sent);
    if (c->h_tile_last) cudaFreeHost(c->h_tile_last);
    if (c->ev_fork_clear) cudaEventDestroy(c->ev_fork_clear);
    if (c->ev_clear_done) cudaEventDestroy(c->ev_clear_done);
    cudaFree(c->d_tris); cudaFree(c->d_pa); cudaFree(c->d_pb); cudaFree(c->d_pc); cudaFree(c->d_objs); cudaFree(c->d_objlite); cudaFree(c->d_obj_r2); cudaFree(c->d_obj_rows);
    cudaFree(c->d_clusters); cudaFree(c->d_cluster_vis); cudaFree(c->d_cluster_faces); cudaFree(c->d_active); cudaFree(c->d_skipped);
    cudaFree(c->d_rowmask); cudaFree(c->d_rowpfx);
    if (c->atlas_tex) cudaDestroyTextureObject(c->atlas_tex);
    cudaFree(c->d_atlas); cudaFree(c->d_nums); cudaFree(c->d_sizes); cudaFree(c->d_upload);
    cudaFree(c->d_lights); cudaFree(c->d_lightlite);
    if (!c->ext_shadow_dyn) cudaFree(c->d_shadow_dyn);
    cudaFree(c->d_shadow_alt);
    if (!c->ext_shadow_static) cudaFree(c->d_shadow_static);
    for (int i = 0; i < 2; i++) { cudaFree(c->d_depth[i]); cudaFree(c->d_ids[i]); }
    if (!c->ext_rgba8) cudaFree(c->d_rgba8);
    cudaFree(c->d_fragcnt); cudaFree(c->d_worklist); cudaFree(c->d_extra);
    cudaFree(c->d_post); cudaFree(c->d_seen);
    cudaFree(c->d_normals); cudaFree(c->d_shade_list); cudaFree(c->d_samples); cudaFree(c->d_sample_frag); cudaFree(c->d_frags); cudaFree(c->d_cutdown); cudaFree(c->d_counters); cudaFree(c->d_lookback);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_objs_pinned) cudaFreeHost(c->h_objs_pinned);
    if (c->arena.base) cudaFreeHost(c->arena.base);
    for (int i = 0; i < 2; i++) if (c->arena.ev[i]) cudaEventDestroy(c->arena.ev[i]);
    for (int i = 0; i < EV_COUNT; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ---- scene ---------------------------------------------------------------------------------------------------------
int rr_scene_alloc(rr_ctx* c, uint32_t n_tris, uint32_t n_objs) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    if (n_tris >= (1u << 26)) return fail(RR_ERR_INVALID, "rr_scene_alloc: %u triangles exceeds the 2^26 limit of the scan descriptor", n_tris);
    CU(cudaStreamSynchronize(c->stream));
    c->n_tris = n_tris; c->n_objs = n_objs; c->tri_cap = n_tris; c->obj_cap = n_objs;
    int r;
    if ((r = dev_alloc(c->d_tris, n_tris))) return r;
    if ((r = dev_alloc(c->d_pa, n_tris))) return r;
    if ((r = dev_alloc(c->d_pb, n_tris))) return r;
    if ((r = dev_alloc(c->d_pc, n_tris))) return r;
    if ((r = dev_alloc(c->d_objs, n_objs))) return r;
    if ((r = dev_alloc(c->d_objlite, n_objs))) return r;
    if ((r = dev_alloc(c->d_obj_r2, n_objs))) return r;
    if ((r = dev_alloc(c->d_obj_rows, n_objs))) return r;
    c->n_clusters = (n_tris + CLUSTER_TRIS - 1) / CLUSTER_TRIS;
    if ((r = dev_alloc(c->d_clusters, (size_t)c->n_clusters))) return r;
    if ((r = dev_alloc(c->d_cluster_vis, (size_t)c->n_clusters + 2))) return r;
    if ((r = dev_alloc(c->d_cluster_faces, (size_t)c->n_clusters))) return r;
    CU(cudaMemsetAsync(c->d_clusters, 0, std::max<size_t>(1, c->n_clusters) * sizeof(ClusterBox), c->stream));   // hi.w = 0: never culled until built
    CU(cudaMemsetAsync(c->d_obj_r2, 0, std::max<size_t>(1, n_objs) * 4, c->stream));
    // projected triangles: main pass needs <= 2T; a shadow pass (all lights at once) up to 12T per light in the worst
    // case (the reference allocates 12T, object_context.cpp:354). Default 6T + slack ; overflow is detected and reported.
    c->cap_cut = c->cfg.max_cutdown ? c->cfg.max_cutdown : (uint32_t)std::min<uint64_t>((uint64_t)n_tris * 6 + 1024, 0x7FFFFFFFu);
    if ((r = dev_alloc(c->d_cutdown, (size_t)c->cap_cut * 3))) return r;
    CU(cudaStreamSynchronize(c->stream2));
    if ((r = dev_alloc(c->d_scutdown, (size_t)c->cap_cut * 3))) return r;
    c->lookback_blocks = (n_tris + SETUP_THREADS - 1) / SETUP_THREADS;
    if ((r = dev_alloc(c->d_lookback, (size_t)c->lookback_blocks))) return r;
    if ((r = dev_alloc(c->d_active, (size_t)c->lookback_blocks))) return r;
    if ((r = dev_alloc(c->d_skipped, (size_t)c->lookback_blocks))) return r;
    if (c->h_objs_pinned) { cudaFreeHost(c->h_objs_pinned); c->h_objs_pinned = nullptr; }
    CU(cudaMallocHost((void**)&c->h_objs_pinned, std::max<size_t>(1, n_objs) * sizeof(rr_obj_desc)));
    c->objlite_dirty = true;
    return RR_OK;
}

int rr_scene_write_tris(rr_ctx* c, uint32_t first, uint32_t count, const rr_triangle* tris) {
    if (!c || !tris) return fail(RR_ERR_INVALID, "null argument");
    if ((uint64_t)first + count > c->n_tris) return fail(RR_ERR_INVALID, "rr_scene_write_tris: range [%u,+%u) outside %u", first, count, c->n_tris);
    if (count == 0) return RR_OK;
    CU(cudaMemcpyAsync(c->d_tris + first, tris, (size_t)count * sizeof(rr_triangle), cudaMemcpyHostToDevice, c->stream));
    k_repack<<<(count + 255) / 256, 256, 0, c->stream>>>(c->d_tris, first, count, c->d_pa, c->d_pb, c->d_pc, c->d_obj_r2, c->n_objs);
    c->launches++;
    {   // bounding boxes of the clusters this write touches
        const uint32_t cl0 = first / CLUSTER_TRIS, cl1 = (first + count - 1) / CLUSTER_TRIS;
        k_cluster_bounds<<<cl1 - cl0 + 1, CLUSTER_TRIS, 0, c->stream>>>(c->d_pa, c->d_pb, c->d_pc, c->n_tris, cl0, c->d_clusters);
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));      // caller may free `tris` on return (the reference keeps staging alive instead, object.cpp:729)
    return RR_OK;
}

int rr_scene_write_objs(rr_ctx* c, uint32_t first, uint32_t count, const rr_obj_desc* objs) {
    if (!c || !objs) return fail(RR_ERR_INVALID, "null argument");
    if ((uint64_t)first + count > c->n_objs) return fail(RR_ERR_INVALID, "rr_scene_write_objs: range outside %u", c->n_objs);
    if (count == 0) return RR_OK;
    memcpy(c->h_objs_pinned + first, objs, (size_t)count * sizeof(rr_obj_desc));          // host mirror (rr_frame_e2e uploads it every frame)
    int r = upload_staged(c, c->d_objs + first, objs, (size_t)count * sizeof(rr_obj_desc));
    if (r) return r;
    c->objlite_dirty = true;
    return RR_OK;
}

int rr_scene_patch_obj(rr_ctx* c, uint32_t obj_id, uint32_t byte_off, uint32_t nbytes, const void* src) {
    if (!c || !src) return fail(RR_ERR_INVALID, "null argument");
    if (obj_id >= c->n_objs || (uint64_t)byte_off + nbytes > sizeof(rr_obj_desc)) return fail(RR_ERR_INVALID, "rr_scene_patch_obj: out of range");
    memcpy((char*)(c->h_objs_pinned + obj_id) + byte_off, src, nbytes);
    int r = upload_staged(c, (char*)(c->d_objs + obj_id) + byte_off, src, nbytes);
    if (r) return r;
    c->objlite_dirty = true;
    return RR_OK;
}

// ---- asynchronous rebuild: object_context::build(async) + flip_buffers (object_context.cpp:520-797) ------------------
int rr_scene_read_objs(rr_ctx* c, uint32_t first, uint32_t count, rr_obj_desc* dst) {
    if (!c || (!dst && count)) return fail(RR_ERR_INVALID, "null argument");
    if ((uint64_t)first + count > c->n_objs) return fail(RR_ERR_INVALID, "rr_scene_read_objs: range outside the %u objects", c->n_objs);
    if (!count) return RR_OK;
    CU(cudaMemcpyAsync(dst, c->d_objs + first, (size_t)count * sizeof(rr_obj_desc), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return RR_OK;
}

static int join_shadows_fwd(rr_ctx* c);
namespace {
int back_alloc_bytes(void** p, size_t bytes) {   // plain cudaMalloc: only called when the back scene has to grow
    if (*p) { cudaFree(*p); *p = nullptr; }
    CU(cudaMalloc(p, std::max<size_t>(bytes, 1)));
    return RR_OK;
}
#define back_alloc(ptr, count) back_alloc_bytes((void**)&(ptr), (size_t)(count) * sizeof(*(ptr)))

void scene_swap(rr_ctx* c, SceneBuf& b) {
    std::swap(c->n_tris, b.n_tris); std::swap(c->n_objs, b.n_objs); std::swap(c->tri_cap, b.tri_cap); std::swap(c->obj_cap, b.obj_cap);
    std::swap(c->d_tris, b.d_tris); std::swap(c->d_pa, b.d_pa); std::swap(c->d_pb, b.d_pb); std::swap(c->d_pc, b.d_pc);
    std::swap(c->d_objs, b.d_objs); std::swap(c->d_objlite, b.d_objlite); std::swap(c->d_obj_r2, b.d_obj_r2); std::swap(c->d_obj_rows, b.d_obj_rows);
    std::swap(c->n_clusters, b.n_clusters); std::swap(c->d_clusters, b.d_clusters); std::swap(c->d_cluster_vis, b.d_cluster_vis);
    std::swap(c->d_cluster_faces, b.d_cluster_faces); std::swap(c->d_active, b.d_active); std::swap(c->d_skipped, b.d_skipped);
    std::swap(c->d_cutdown, b.d_cutdown); std::swap(c->d_scutdown, b.d_scutdown); std::swap(c->cap_cut, b.cap_cut);
    std::swap(c->d_lookback, b.d_lookback); std::swap(c->lookback_blocks, b.lookback_blocks); std::swap(c->h_objs_pinned, b.h_objs_pinned);
}

void scene_free(SceneBuf& b) {
    cudaFree(b.d_tris); cudaFree(b.d_pa); cudaFree(b.d_pb); cudaFree(b.d_pc); cudaFree(b.d_objs); cudaFree(b.d_objlite); cudaFree(b.d_obj_r2);
    cudaFree(b.d_obj_rows); cudaFree(b.d_clusters); cudaFree(b.d_cluster_vis); cudaFree(b.d_cluster_faces); cudaFree(b.d_active); cudaFree(b.d_skipped);
    cudaFree(b.d_cutdown); cudaFree(b.d_scutdown); cudaFree(b.d_lookback);
    if (b.h_objs_pinned) cudaFreeHost(b.h_objs_pinned);
    b = SceneBuf();
}
}  // namespace

}  // extern "C"
static void scene_free_fwd(rr_ctx* c) { scene_free(c->back); }
extern "C" {

int rr_scene_build_begin(rr_ctx* c, uint32_t n_tris, uint32_t n_objs) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    if (c->building) return fail(RR_ERR_INVALID, "rr_scene_build_begin: a rebuild is already in progress (commit it first)");
    if (n_tris >= (1u << 26)) return fail(RR_ERR_INVALID, "rr_scene_build_begin: %u triangles exceeds the 2^26 limit of the scan descriptor", n_tris);
    if (!c->stream4) {
        CU(cudaStreamCreateWithFlags(&c->stream4, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_built, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_retire, cudaEventDisableTiming));
    }
    // the back buffers were the front scene until the last commit: frames enqueued before it may still read them
    if (c->retire_pending) { CU(cudaStreamWaitEvent(c->stream4, c->ev_retire, 0)); }
    SceneBuf& b = c->back;
    int r;
    const uint32_t n_clusters = (n_tris + CLUSTER_TRIS - 1) / CLUSTER_TRIS, blocks = (n_tris + SETUP_THREADS - 1) / SETUP_THREADS;
    const uint32_t cap_cut = c->cfg.max_cutdown ? c->cfg.max_cutdown : (uint32_t)std::min<uint64_t>((uint64_t)n_tris * 6 + 1024, 0x7FFFFFFFu);
    if (n_tris > b.tri_cap || cap_cut > b.cap_cut || !b.d_tris) {    // grow (or first use: an empty scene still gets one element each) (cudaFree of the old buffers synchronises the device once)
        if (c->retire_pending) { CU(cudaEventSynchronize(c->ev_retire)); }
        if ((r = back_alloc(b.d_tris, n_tris))) return r;
        if ((r = back_alloc(b.d_pa, n_tris))) return r;
        if ((r = back_alloc(b.d_pb, n_tris))) return r;
        if ((r = back_alloc(b.d_pc, n_tris))) return r;
        if ((r = back_alloc(b.d_clusters, (size_t)n_clusters))) return r;
        if ((r = back_alloc(b.d_cluster_vis, (size_t)n_clusters + 2))) return r;
        if ((r = back_alloc(b.d_cluster_faces, (size_t)n_clusters))) return r;
        if ((r = back_alloc(b.d_cutdown, (size_t)cap_cut * 3))) return r;
        if ((r = back_alloc(b.d_scutdown, (size_t)cap_cut * 3))) return r;
        if ((r = back_alloc(b.d_lookback, (size_t)blocks))) return r;
        if ((r = back_alloc(b.d_active, (size_t)blocks))) return r;
        if ((r = back_alloc(b.d_skipped, (size_t)blocks))) return r;
        b.tri_cap = n_tris;
    }
    if (n_objs > b.obj_cap || !b.d_objs) {
        if (c->retire_pending) { CU(cudaEventSynchronize(c->ev_retire)); }
        if ((r = back_alloc(b.d_objs, n_objs))) return r;
        if ((r = back_alloc(b.d_objlite, n_objs))) return r;
        if ((r = back_alloc(b.d_obj_r2, n_objs))) return r;
        if ((r = back_alloc(b.d_obj_rows, n_objs))) return r;
        if (b.h_objs_pinned) cudaFreeHost(b.h_objs_pinned);
        b.h_objs_pinned = nullptr;
        CU(cudaMallocHost((void**)&b.h_objs_pinned, std::max<size_t>(1, n_objs) * sizeof(rr_obj_desc)));
        b.obj_cap = n_objs;
    }
    b.n_tris = n_tris; b.n_objs = n_objs; b.n_clusters = n_clusters; b.lookback_blocks = blocks; b.cap_cut = cap_cut;
    CU(cudaMemsetAsync(b.d_clusters, 0, std::max<size_t>(1, n_clusters) * sizeof(ClusterBox), c->stream4));
    CU(cudaMemsetAsync(b.d_obj_r2, 0, std::max<size_t>(1, n_objs) * 4, c->stream4));
    c->building = true;
    return RR_OK;
}

int rr_scene_build_write_objs(rr_ctx* c, uint32_t first, uint32_t count, const rr_obj_desc* objs) {
    if (!c || !objs) return fail(RR_ERR_INVALID, "null argument");
    if (!c->building) return fail(RR_ERR_INVALID, "rr_scene_build_write_objs outside rr_scene_build_begin / commit");
    SceneBuf& b = c->back;
    if ((uint64_t)first + count > b.n_objs) return fail(RR_ERR_INVALID, "rr_scene_build_write_objs: range outside %u", b.n_objs);
    if (count == 0) return RR_OK;
    memcpy(b.h_objs_pinned + first, objs, (size_t)count * sizeof(rr_obj_desc));
    CU(cudaMemcpyAsync(b.d_objs + first, b.h_objs_pinned + first, (size_t)count * sizeof(rr_obj_desc), cudaMemcpyHostToDevice, c->stream4));
    return RR_OK;
}

int rr_scene_build_write_tris(rr_ctx* c, uint32_t first, uint32_t count, const rr_triangle* tris) {
    if (!c || !tris) return fail(RR_ERR_INVALID, "null argument");
    if (!c->building) return fail(RR_ERR_INVALID, "rr_scene_build_write_tris outside rr_scene_build_begin / commit");
    SceneBuf& b = c->back;
    if ((uint64_t)first + count > b.n_tris) return fail(RR_ERR_INVALID, "rr_scene_build_write_tris: range [%u,+%u) outside %u", first, count, b.n_tris);
    if (count == 0) return RR_OK;
    // pageable `tris`: the runtime stages the copy before returning, the caller may free it; page-locked `tris`: a true
    // asynchronous DMA, the memory must stay valid until rr_scene_build_ready() (object.cpp:729-731 has the same rule)
    CU(cudaMemcpyAsync(b.d_tris + first, tris, (size_t)count * sizeof(rr_triangle), cudaMemcpyHostToDevice, c->stream4));
    k_repack<<<(count + 255) / 256, 256, 0, c->stream4>>>(b.d_tris, first, count, b.d_pa, b.d_pb, b.d_pc, b.d_obj_r2, b.n_objs);
    const uint32_t cl0 = first / CLUSTER_TRIS, cl1 = (first + count - 1) / CLUSTER_TRIS;
    k_cluster_bounds<<<cl1 - cl0 + 1, CLUSTER_TRIS, 0, c->stream4>>>(b.d_pa, b.d_pb, b.d_pc, b.n_tris, cl0, b.d_clusters);
    c->launches += 2;
    CU(cudaGetLastError());
    return RR_OK;
}

int rr_scene_build_ready(rr_ctx* c) {
    if (!c || !c->building) return 0;
    cudaError_t e = cudaStreamQuery(c->stream4);
    if (e == cudaSuccess) return 1;
    if (e != cudaErrorNotReady) fail(RR_ERR_CUDA, "rr_scene_build_ready: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 0;
}

int rr_scene_build_commit(rr_ctx* c) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    if (!c->building) return fail(RR_ERR_INVALID, "rr_scene_build_commit without rr_scene_build_begin");
    // frames enqueued so far read the current front scene: it becomes reusable once they are done
    if (c->shadow_pending) { int r = join_shadows_fwd(c); if (r) return r; }
    CU(cudaEventRecord(c->ev_retire, c->stream));
    c->retire_pending = true;
    // frames enqueued from now on read the new scene, and only after its uploads have landed (no host synchronisation)
    CU(cudaEventRecord(c->ev_built, c->stream4));
    CU(cudaStreamWaitEvent(c->stream, c->ev_built, 0));
    scene_swap(c, c->back);
    c->objlite_dirty = true;
    c->building = false;
    return RR_OK;
}

// ---- atlas ---------------------------------------------------------------------------------------------------------
int rr_atlas_alloc(rr_ctx* c, uint32_t n_slices, const uint32_t* nums, uint32_t n_nums, const uint32_t* sizes, uint32_t n_sizes, uint32_t mipmap_start) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    CU(cudaStreamSynchronize(c->stream));
    uint32_t slices = std::max(n_slices, 2u);                                      // clamped_array_len, texture_context.cpp:441
    c->atlas_texels = (size_t)slices * RR_ATLAS_DIM * RR_ATLAS_DIM;
    int r;
    if ((r = dev_alloc(c->d_atlas, c->atlas_texels))) return r;
    if ((r = dev_alloc(c->d_nums, std::max(n_nums, 2u)))) return r;
    if ((r = dev_alloc(c->d_sizes, std::max(n_sizes, 2u)))) return r;
    CU(cudaMemsetAsync(c->d_atlas, 0, c->atlas_texels * 4, c->stream));
    if (n_nums) CU(cudaMemcpyAsync(c->d_nums, nums, (size_t)n_nums * 4, cudaMemcpyHostToDevice, c->stream));
    if (n_sizes) CU(cudaMemcpyAsync(c->d_sizes, sizes, (size_t)n_sizes * 4, cudaMemcpyHostToDevice, c->stream));
    c->n_nums = n_nums; c->n_sizes = n_sizes; c->mipmap_start = mipmap_start;
    CU(cudaStreamSynchronize(c->stream));
    if (c->atlas_tex) { cudaDestroyTextureObject(c->atlas_tex); c->atlas_tex = 0; }
    if (c->use_tex_objects && (size_t)slices * RR_ATLAS_DIM <= 65000) {              // pitch-2D textures are limited to 65000 rows
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = c->d_atlas;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
        rd.res.pitch2D.width = RR_ATLAS_DIM; rd.res.pitch2D.height = (size_t)slices * RR_ATLAS_DIM; rd.res.pitch2D.pitchInBytes = (size_t)RR_ATLAS_DIM * 4;
        cudaTextureDesc td;
        memset(&td, 0, sizeof td);
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&c->atlas_tex, &rd, &td, nullptr));
    }
    return RR_OK;
}

int rr_atlas_upload(rr_ctx* c, uint32_t gpu_id, const uint8_t* rgba, uint32_t w, uint32_t h, int flip) {
    if (!c || !rgba) return fail(RR_ERR_INVALID, "null argument");
    if (!c->d_atlas) return fail(RR_ERR_INVALID, "rr_atlas_upload before rr_atlas_alloc");
    if (gpu_id * RR_MIP_LEVELS + c->mipmap_start + RR_MIP_LEVELS > c->n_nums) return fail(RR_ERR_INVALID, "rr_atlas_upload: texture %u has no mip descriptors", gpu_id);
    size_t n = (size_t)w * h;
    if (n > c->upload_cap) { int r = dev_alloc(c->d_upload, n); if (r) return r; c->upload_cap = n; }
    CU(cudaMemcpyAsync(c->d_upload, rgba, n * 4, cudaMemcpyHostToDevice, c->stream));
    dim3 grid((w + 15) / 16, (h + 15) / 16);
    k_atlas_upload<<<grid, 256, 0, c->stream>>>(c->d_upload, (int)w, (int)h, gpu_id, flip, c->d_atlas, c->d_nums, c->d_sizes);
    c->launches++;
    // texture::update_gpu_mipmaps, texture.cpp:465-493: generate_mips then generate_mip_mips for levels 0..2, all with the base image's global size
    uint32_t m0 = gpu_id * RR_MIP_LEVELS + c->mipmap_start;
    k_atlas_mip<<<grid, 256, 0, c->stream>>>(gpu_id, m0, (int)w, (int)h, c->d_atlas, c->d_nums, c->d_sizes);
    c->launches++;
    for (uint32_t i = 0; i < RR_MIP_LEVELS - 1; i++) {
        k_atlas_mip<<<grid, 256, 0, c->stream>>>(m0 + i, m0 + i + 1, (int)w, (int)h, c->d_atlas, c->d_nums, c->d_sizes);
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return RR_OK;
}

// texture_context::alloc_gpu's upload loop (texture_context.cpp:478-517) for a whole set of textures: one staged copy, five launches
int rr_atlas_upload_batch(rr_ctx* c, uint32_t n, const uint32_t* gpu_ids, const uint8_t* const* rgba, const uint32_t* w, const uint32_t* h, int flip) {
    if (!c || (n && (!gpu_ids || !rgba || !w || !h))) return fail(RR_ERR_INVALID, "null argument");
    if (!c->d_atlas) return fail(RR_ERR_INVALID, "atlas not allocated");
    if (n == 0) return RR_OK;
    std::vector<AtlasJob> jobs(n);
    size_t texels = 0;
    uint32_t tiles = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (gpu_ids[i] >= c->n_nums || !rgba[i]) return fail(RR_ERR_INVALID, "rr_atlas_upload_batch: bad texture %u", i);
        jobs[i] = AtlasJob{(unsigned long long)texels, w[i], h[i], gpu_ids[i], tiles};
        texels += (size_t)w[i] * h[i];
        const uint64_t t = (uint64_t)((w[i] + 15) / 16) * ((h[i] + 15) / 16);
        if (tiles + t > 0x7FFFFFFFull) return fail(RR_ERR_INVALID, "rr_atlas_upload_batch: batch too large");
        tiles += (uint32_t)t;
    }
    if (texels == 0 || tiles == 0) return RR_OK;
    int r;
    if (texels > c->upload_cap) { if ((r = dev_alloc(c->d_upload, texels))) return r; c->upload_cap = texels; }
    AtlasJob* d_jobs = nullptr;
    CU(cudaMalloc((void**)&d_jobs, (size_t)n * sizeof(AtlasJob)));
    uchar4* h_stage = nullptr;                                        // one pinned staging block: a single DMA instead of one write per texture
    if (cudaMallocHost((void**)&h_stage, texels * 4) != cudaSuccess) { cudaFree(d_jobs); return fail(RR_ERR_OOM, "rr_atlas_upload_batch: staging"); }
    for (uint32_t i = 0; i < n; i++) memcpy(h_stage + jobs[i].src_off, rgba[i], (size_t)w[i] * h[i] * 4);
    cudaError_t e = cudaMemcpyAsync(c->d_upload, h_stage, texels * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_jobs, jobs.data(), (size_t)n * sizeof(AtlasJob), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        k_atlas_upload_batch<<<tiles, 256, 0, c->stream>>>(d_jobs, n, c->d_upload, flip, c->d_atlas, c->d_nums, c->d_sizes);
        for (int level = 0; level < RR_MIP_LEVELS; level++)
            k_atlas_mip_batch<<<tiles, 256, 0, c->stream>>>(d_jobs, n, level, c->mipmap_start, c->d_atlas, c->d_nums, c->d_sizes);
        c->launches += 1 + RR_MIP_LEVELS;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFreeHost(h_stage);
    cudaFree(d_jobs);
    if (e != cudaSuccess) return fail(RR_ERR_CUDA, "rr_atlas_upload_batch: %s", cudaGetErrorString(e));
    return RR_OK;
}

// texture::update_gpu_texture_col, texture.cpp:445-463 (col in 0..255 units, launch size = the texture's image size)
int rr_atlas_fill_colour(rr_ctx* c, uint32_t gpu_id, const float col[4], uint32_t w, uint32_t h) {
    if (!c || !col) return fail(RR_ERR_INVALID, "null argument");
    if (!c->d_atlas || gpu_id >= c->n_nums) return fail(RR_ERR_INVALID, "atlas not allocated or bad texture id");
    if (w == 0 || h == 0) return RR_OK;
    dim3 grid((w + 15) / 16, (h + 15) / 16);
    k_atlas_fill_colour<<<grid, 256, 0, c->stream>>>(make_float4(col[0], col[1], col[2], col[3]), gpu_id, c->mipmap_start, (int)w, (int)h, c->d_atlas, c->d_nums, c->d_sizes);
    c->launches++;
    CU(cudaGetLastError());
    return RR_OK;
}

// texture::update_gpu_texture_mono, texture.cpp:554-584: len bytes, stride = len / height; asynchronous like there (the bytes are
// staged before the call returns)
int rr_atlas_upload_mono(rr_ctx* c, uint32_t gpu_id, const uint8_t* raw, uint32_t len, uint32_t w, uint32_t h, int flip) {
    (void)flip;                                                       // generate_from_raw ignores it (cl2.cl:1020-1026)
    if (!c || !raw) return fail(RR_ERR_INVALID, "null argument");
    if (!c->d_atlas || gpu_id >= c->n_nums) return fail(RR_ERR_INVALID, "atlas not allocated or bad texture id");
    if (w == 0 || h == 0 || len == 0) return RR_OK;
    const uint32_t stride = len / h;
    if ((uint64_t)(h - 1) * stride + w > len) return fail(RR_ERR_INVALID, "rr_atlas_upload_mono: %u bytes do not hold a %ux%u image", len, w, h);
    int r;
    const size_t need = ((size_t)len + 3) / 4;
    if (need > c->upload_cap) { if ((r = dev_alloc(c->d_upload, need))) return r; c->upload_cap = need; }
    if ((r = upload_staged(c, c->d_upload, raw, len))) return r;
    dim3 grid((w + 15) / 16, (h + 15) / 16);
    k_atlas_from_raw<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const unsigned char*>(c->d_upload), (int)stride, (int)w, (int)h, gpu_id, c->d_atlas, c->d_nums, c->d_sizes);
    c->launches++;
    CU(cudaGetLastError());
    return RR_OK;
}

int rr_atlas_write_raw(rr_ctx* c, const uint8_t* atlas, size_t nbytes) {
    if (!c || !atlas || nbytes > c->atlas_texels * 4) return fail(RR_ERR_INVALID, "rr_atlas_write_raw: bad size");
    CU(cudaMemcpyAsync(c->d_atlas, atlas, nbytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return RR_OK;
}
int rr_atlas_read_raw(rr_ctx* c, uint8_t* dst, size_t nbytes) {
    if (!c || !dst || nbytes > c->atlas_texels * 4) return fail(RR_ERR_INVALID, "rr_atlas_read_raw: bad size");
    CU(cudaMemcpyAsync(dst, c->d_atlas, nbytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return RR_OK;
}

// ---- lights --------------------------------------------------------------------------------------------------------
int rr_lights_write(rr_ctx* c, const rr_light* lights, uint32_t n_active) {
    if (!c || (!lights && n_active)) return fail(RR_ERR_INVALID, "null argument");
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->stream2));
    c->lights.assign(lights, lights + n_active);
    uint32_t ns = 0, nst = 0;
    for (uint32_t i = 0; i < n_active; i++) {
        if (lights[i].shadow == 1) ns++;                              // light.cpp:203-208
        if (lights[i].shadow && lights[i].is_static) nst++;
    }
    int r;
    if ((r = dev_alloc(c->d_lights, std::max(n_active, 1u)))) return r;   // clamped_num, light.cpp:172
    if ((r = dev_alloc(c->d_lightlite, std::max(n_active, 1u)))) return r;
    if (n_active) {
        CU(cudaMemcpyAsync(c->d_lights, lights, (size_t)n_active * sizeof(rr_light), cudaMemcpyHostToDevice, c->stream));
        k_lightlite<<<(n_active + 63) / 64, 64, 0, c->stream>>>(c->d_lights, n_active, c->d_lightlite);
        c->launches++;
    }
    const size_t slab = (size_t)6 * c->L * c->L;
    if (!c->ext_shadow_dyn && (ns != c->n_shadow || !c->d_shadow_dyn)) {
        c->shadow_dyn_words = std::max<size_t>(slab * ns, 4);
        if ((r = dev_alloc(c->d_shadow_dyn, c->shadow_dyn_words))) return r;
        if ((r = fill_u32(c, c->stream, c->d_shadow_dyn, c->shadow_dyn_words, 0xFFFFFFFFu))) return r;
        if ((r = dev_alloc(c->d_shadow_alt, c->shadow_dyn_words))) return r;
        if ((r = fill_u32(c, c->stream, c->d_shadow_alt, c->shadow_dyn_words, 0xFFFFFFFFu))) return r;
        c->alt_clean = true;
    }
    if (!c->ext_shadow_static && (nst != c->n_static || !c->d_shadow_static)) {
        c->shadow_static_words = std::max<size_t>(slab * nst, 4);
        if ((r = dev_alloc(c->d_shadow_static, c->shadow_static_words))) return r;
        if ((r = fill_u32(c, c->stream, c->d_shadow_static, c->shadow_static_words, 0xFFFFFFFFu))) return r;
    }
    if (c->ext_shadow_dyn && slab * ns > c->shadow_dyn_words) return fail(RR_ERR_INVALID, "bound dynamic shadow buffer too small");
    if (c->ext_shadow_static && slab * nst > c->shadow_static_words) return fail(RR_ERR_INVALID, "bound static shadow buffer too small");
    c->n_shadow = ns; c->n_static = nst;
    CU(cudaStreamSynchronize(c->stream));
    return RR_OK;
}

// ---- per frame -----------------------------------------------------------------------------------------------------
// one pass = all shadow lights of one kind (dynamic: only_static 0 into g_shadow_light_buffer; static: only_static 1
// into g_static_shadow_light_buffer), engine.cpp:1629-1784, in one setup launch + one scan + one raster pair
static int shadow_pass(rr_ctx* c, int only_static) {
    uint32_t* buffer = only_static ? c->d_shadow_static : c->d_shadow_dyn;
    const uint32_t total_pairs = 6u * (only_static ? c->n_static : c->n_shadow);
    const bool sharded = c->cfg.face_world > 1 && !only_static;      // static cubemaps are cached: every context renders all of theirs
    std::vector<ShadowLight> sel;
    uint32_t slab = 0;
    for (size_t i = 0; i < c->lights.size(); i++) {
        const rr_light& l = c->lights[i];
        const bool in_pass = only_static ? (l.shadow && l.is_static) : (l.shadow == 1);
        if (!in_pass) continue;
        // (light, face) pairs are owned either in contiguous chunks of ceil(total / face_world), so that the owned part of
        // the cubemap buffer is one contiguous range (in-place all-gather across contexts), or round-robin (peer-memory push)
        uint32_t mask = 0;
        for (uint32_t kk = 0; kk < 6; kk++)
            if (!sharded || owns_pair(c, slab * 6 + kk, total_pairs)) mask |= 1u << kk;
        if (mask) sel.push_back(ShadowLight{l.pos[0], l.pos[1], l.pos[2], slab, mask});
        slab++;
    }
    if (sel.empty() || c->n_tris == 0) return RR_OK;
    for (size_t first = 0; first < sel.size(); first += SHADOW_MAX_LIGHTS) {
        const int nl = (int)std::min<size_t>(SHADOW_MAX_LIGHTS, sel.size() - first);
        cudaStream_t st = c->stream2;
        ShadowSetupParams sp;
        sp.pa = c->d_pa; sp.pb = c->d_pb; sp.pc = c->d_pc; sp.objs = c->d_objlite; sp.n_tris = c->n_tris;
        sp.n_lights = nl;
        RasterParams dp;
        for (int k = 0; k < nl; k++) { sp.lights[k] = sel[first + k]; dp.slab_of_light[k] = sel[first + k].slab; }
        sp.faces = c->faces;
        sp.L = (float)c->L; sp.icut = (float)c->cfg.depth_icutoff;
        sp.only_static = only_static;
        sp.pretest = c->shadow_pretest;
        sp.far2_max = far2_threshold();
        sp.frags = c->d_sfrags; sp.cap_frags = (uint32_t)(((uint64_t)c->cap_frags * RR_FRAG_WORDS) / RR_SFRAG_WORDS);
        sp.fragcnt = c->d_sfragcnt;
        sp.cutdown = c->d_scutdown; sp.cap_cut = c->cap_cut; sp.counters = c->d_scounters;
        sp.buffer = buffer;
        {   // per-cluster cube-face reach of this pass's lights (a block of k_shadow_setup == one cluster): clusters that cannot reach a
            // face rendered here leave at once, clusters that reach a single face skip ret_cubeface; also zeroes the pass's counters
            ObjFacesParams fp;
            fp.n_lights = nl;
            for (int k = 0; k < nl; k++) fp.lights[k] = sel[first + k];
            k_cluster_faces<<<(c->n_clusters + 127) / 128, 128, 0, st>>>(c->d_clusters, c->n_clusters, c->d_objlite, c->n_objs, fp, c->d_cluster_faces, c->d_scounters);
            c->launches++;
            sp.cluster_faces = c->d_cluster_faces;
        }
        k_shadow_setup<<<(c->n_tris + 127) / 128, 128, 0, st>>>(sp);
        c->launches++;
        dp.frags = c->d_sfrags; dp.cutdown = c->d_scutdown; dp.fragcnt = c->d_sfragcnt; dp.counters = c->d_scounters; dp.cap_frags = sp.cap_frags;
        dp.worklist = nullptr; dp.extra = nullptr; dp.shade_list = nullptr; dp.shade_count = nullptr; dp.tile_now = nullptr; dp.tiles_x = 0;
        dp.n_index = CTR_S_NFRAG;
        dp.depth = buffer; dp.ids = nullptr; dp.width = (float)c->L; dp.height = (float)c->L; dp.W = c->L;
        dp.row_lo = 0; dp.row_hi = c->L; dp.rowmask = nullptr; dp.rowbit = 0;
        k_raster_shadow_warp<<<grid_for(c, 4), 256, 0, st>>>(dp);                         // the stored fragments (triangles larger than one small chunk)
        c->launches++;
    }
    CU(cudaGetLastError());
    return RR_OK;
}

int rr_frame_shadows(rr_ctx* c, int static_lights_dirty) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    int r;
    if ((r = ensure_objlite(c))) return r;
    // fork: the shadow work is ordered after everything already enqueued on the main stream (previous frame's shading
    // reads the cubemaps, uploads, k_objlite) and then runs on its own stream, concurrently with rr_frame_draw's
    // setup / depth / id kernels; rr_frame_draw joins right before shading
    CU(cudaEventRecord(c->ev_fork, c->stream));
    CU(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    if (c->stage_events) CU(cudaEventRecord(c->ev[EV_SH0], c->stream2));
    const size_t slab = (size_t)6 * c->L * c->L;
    int mgb = 0;
    uint32_t mg_epoch = 0;
    if (c->mg.connected && c->n_shadow) {    // (no shadow-casting light: nothing to exchange, no epoch)
        // peer-memory exchange: this epoch's cubemaps live in buffer (epoch & 1). Peers may only be written once they are
        // done with the frame that last read that buffer, which their flag for the previous epoch implies.
        mg_epoch = ++c->mg.shadow_epoch;
        mgb = (int)(mg_epoch & 1u);
        c->d_shadow_dyn = c->mg.shadow[mgb];
        if (mg_epoch > 1 && (r = mg_wait(c, c->stream2, true, mg_epoch - 1))) return r;
    }
    if (!c->lights.empty()) {                                                              // engine.cpp:1611-1626
        // a full clear (not only the owned faces) keeps the buffer defined for the all-gather that follows
        if (c->mg.connected) { if (c->n_shadow && mg_epoch && (r = mg_fill_owned(c, c->stream2, c->d_shadow_dyn))) return r; }
        else if (c->n_shadow) {
            if (!c->ext_shadow_dyn && c->d_shadow_alt && c->alt_clean) { std::swap(c->d_shadow_dyn, c->d_shadow_alt); c->alt_clean = false; }   // cleared behind the previous pass
            else if ((r = fill_u32(c, c->stream2, c->d_shadow_dyn, slab * c->n_shadow, 0xFFFFFFFFu))) return r;
        }
        if (static_lights_dirty && c->n_static && (r = fill_u32(c, c->stream2, c->d_shadow_static, slab * c->n_static, 0xFFFFFFFFu))) return r;
    }
    if (c->n_shadow && (r = shadow_pass(c, 0))) return r;                                  // engine.cpp:1629-1697
    if (static_lights_dirty && c->n_static && (r = shadow_pass(c, 1))) return r;           // engine.cpp:1699-1784
    if (c->mg.connected && c->n_shadow && mg_epoch && (r = mg_push(c, c->stream2, mgb, mg_epoch))) return r;
    if (c->stage_events) CU(cudaEventRecord(c->ev[EV_SH1], c->stream2));
    CU(cudaEventRecord(c->ev_shadow_done, c->stream2));
    if (!c->mg.connected && !c->ext_shadow_dyn && c->d_shadow_alt && c->n_shadow && !c->lights.empty()) {
        // the other buffer was last read by the previous frame's shading, which the fork above is ordered after: clear it now, behind
        // this frame's pass, for the next frame (it runs beside the main view's kernels and the shading, which are not memory-bound)
        if ((r = fill_u32(c, c->stream2, c->d_shadow_alt, slab * c->n_shadow, 0xFFFFFFFFu))) return r;
        c->alt_clean = true;
    }
    c->have_shadow_ev = c->stage_events;
    c->shadow_pending = true;
    return RR_OK;
}

static int join_shadows(rr_ctx* c) {
    if (c->shadow_pending) {
        CU(cudaStreamWaitEvent(c->stream, c->ev_shadow_done, 0));
        c->shadow_pending = false;
    }
    return RR_OK;
}
static int join_shadows_fwd(rr_ctx* c) { return join_shadows(c); }

int rr_frame_draw(rr_ctx* c, const float c_pos[4], const float c_rot[4], const float clear_rgba[4]) {
    if (!c || !c_pos || !c_rot) return fail(RR_ERR_INVALID, "null argument");
    if (c->n_tris == 0) return RR_OK;                                                      // engine.cpp:1806
    int r;
    if ((r = ensure_objlite(c))) return r;
    CamParams cam;
    cam.pos = make_float3(c_pos[0], c_pos[1], c_pos[2]);
    cam.rot = make_rotsc(c_rot[0], c_rot[1], c_rot[2]);
    int band0, band1, row0, row1;
    band_rows(c, band0, band1, row0, row1);

    const bool mg_composite = c->mg.connected && !c->mg.local_readback;      // rows of all contexts meet in rank 0's colour target
    if (c->mg.connected) ++c->mg.draw_epoch;
    if (mg_composite && c->mg.rank == 0 && c->mg.world > 1) {
        // Rank 0 owns the composite target. Everything that still reads the target this frame goes to (a synchronous read-back
        // of the previous frame, or — rr_frame_e2e — the copy of the frame that used this ring slot D frames ago, which the main
        // stream has just waited for) is ordered before this point of rank 0's stream: tell the peers they may store into it.
        MgFlagList fl;
        fl.n = 0; fl.value = c->mg.draw_epoch;
        for (int q = 1; q < c->mg.world; q++) fl.flag[fl.n++] = &c->mg.peer_ctrl[q]->fb_free;
        k_signal_flags<<<1, 32, 0, c->stream>>>(fl);
        c->launches++;
    }
    if (c->stage_events) CU(cudaEventRecord(c->ev[EV_F0], c->stream));
    // kernel3's parameters (needed first: its streaming stores are launched now, on the side stream)
    ShadeParams hp;
    hp.tris = c->d_tris; hp.objs = c->d_objs; hp.objlite = c->d_objlite; hp.lightlite = c->d_lightlite; hp.frags = c->d_frags; hp.cutdown = c->d_cutdown; hp.n_frags = c->d_counters + CTR_NFRAG;
    hp.depth = c->d_depth[c->cur]; hp.ids = c->d_ids[c->cur];
    hp.depth_next = c->d_depth[c->cur ^ 1]; hp.ids_next = c->d_ids[c->cur ^ 1];
    if (c->mg.connected) c->d_rgba8 = (c->mg.rank == 0 || c->mg.local_readback) ? c->mg.fb[c->mg_target] : c->mg.fb0[c->mg_target];
    hp.rgba8 = c->d_rgba8; hp.normals = c->d_normals;
    hp.atlas.texels = c->d_atlas; hp.atlas.nums = c->d_nums; hp.atlas.sizes = c->d_sizes; hp.atlas.mip_start = c->mipmap_start; hp.atlas.tex = c->atlas_tex;
    hp.lights = c->d_lights; hp.n_lights = (int)c->lights.size();
    hp.shadow_dyn = c->d_shadow_dyn; hp.shadow_static = c->d_shadow_static;
    hp.faces = c->faces; hp.cam = cam;
    hp.clear = clear_rgba ? make_float4(clear_rgba[0], clear_rgba[1], clear_rgba[2], clear_rgba[3]) : make_float4(0, 0, 0, 0);
    hp.W = c->W; hp.H = c->H; hp.L = c->L; hp.fov = c->fov;
    hp.ambient = c->cfg.ambient; hp.ssao_rad = c->cfg.ssao_rad; hp.ssao_div = c->cfg.ssao_div;
    hp.inv_mip_bias = 1.f / c->cfg.mip_bias;
    hp.shadow_bias = c->cfg.shadow_bias; hp.shadow_bias_max = powf(c->cfg.shadow_bias, c->cfg.shadow_exp);
    hp.linear = (c->cfg.test_linear && c->cfg.use_linear_rendering) ? 1 : 0;
    hp.no_ssao = c->cfg.no_ssao;
    hp.row0 = row0; hp.row1 = row1; hp.band_y0 = band0; hp.band_y1 = band1; hp.rowmask = c->d_rowmask;
    if (hp.n_lights > 0 && (!c->d_lights)) return fail(RR_ERR_INVALID, "lights not written");
    hp.shade_list = c->d_shade_list; hp.shade_count = c->d_counters + CTR_NSHADE;
    // The streaming stores of kernel3 — next frame's depth / id clear and the clear colour — depend on nothing this frame computes:
    // they run on a side stream (k_clear_next) and are joined in front of the shading list. Everything that last read those buffers
    // (the previous frame's shading and post passes, the copy of the ring slot) is ordered before this call on the main stream. The
    // fork sits at the start of the frame (RR_CLEAR_AT=1: behind k_setup_main, beside the latency-bound kernels that follow it —
    // measured equal or slightly slower). A peer of the composite target stores its clear colour later, behind the fb_free wait.
    const bool split_clear = c->split_clear && c->W % 4 == 0;
    const int clear_at = c->clear_at;
    auto fork_clear = [&]() -> int {
        CU(cudaEventRecord(c->ev_fork_clear, c->stream));
        CU(cudaStreamWaitEvent(c->stream5, c->ev_fork_clear, 0));
        k_clear_next<<<grid_for(c, c->clear_grid), 256, 0, c->stream5>>>(hp, (mg_composite && c->mg.rank != 0) ? 0 : 1);
        c->launches++;
        CU(cudaEventRecord(c->ev_clear_done, c->stream5));
        return RR_OK;
    };
    if (split_clear && clear_at == 0 && (r = fork_clear())) return r;
    // prearrange
    SetupMainParams sp;
    sp.pa = c->d_pa; sp.pb = c->d_pb; sp.pc = c->d_pc; sp.objs = c->d_objlite; sp.n_tris = c->n_tris;
    sp.cam = cam; sp.width = (float)c->W; sp.height = (float)c->H; sp.fov = c->fov; sp.icut = (float)c->cfg.depth_icutoff;
    sp.frags = c->d_frags; sp.cap_frags = c->cap_frags; sp.fragcnt = c->d_fragcnt; sp.cutdown = c->d_cutdown; sp.cap_cut = c->cap_cut;
    sp.counters = c->d_counters; sp.lookback = c->d_lookback;
    sp.depth = c->d_depth[c->cur]; sp.row_lo = row0; sp.row_hi = row1;
    sp.obj_rows = nullptr;
    sp.rowmask = c->d_rowmask;
    if ((row0 > 0 || row1 < c->H || c->d_rowpfx) && c->n_objs) {      // sort-first band: objects that cannot touch the rasterised rows are not set up here
        k_obj_rows<<<(c->n_objs + 127) / 128, 128, 0, c->stream>>>(c->d_objlite, c->d_obj_r2, c->n_objs, cam, (float)c->H, c->fov,
                                                                    (float)c->cfg.depth_icutoff, c->d_obj_rows, c->d_rowpfx, c->H);
        c->launches++;
        sp.obj_rows = c->d_obj_rows;
    }
    sp.rowpfx = c->d_rowpfx; sp.cull_rows = c->banded ? 1 : 0;
    sp.cluster_vis = nullptr; sp.active = nullptr; sp.skipped_before = nullptr;
    const bool cull = c->n_objs > 0 && (c->cluster_cull > 0 || (c->cluster_cull == 0 && c->banded));
    {   // one launch: zero the scan state and, when culling, classify the clusters (off-screen geometry; rows rasterised elsewhere)
        // and compact the setup blocks. No cudaMemsetAsync anywhere in the frame: memsets may be placed on the copy engine,
        // where they wait behind the previous frame's read-back DMA.
        PrologueParams pp;
        pp.cv.boxes = c->d_clusters; pp.cv.n_clusters = c->n_clusters; pp.cv.objs = c->d_objlite; pp.cv.n_objs = c->n_objs;
        pp.cv.cam = cam; pp.cv.width = (float)c->W; pp.cv.height = (float)c->H; pp.cv.fov = c->fov; pp.cv.icut = (float)c->cfg.depth_icutoff;
        pp.cv.rowpfx = c->d_rowpfx; pp.cv.row_lo = row0; pp.cv.row_hi = row1;
        pp.vis = c->d_cluster_vis; pp.n_tris = c->n_tris; pp.n_blocks = c->lookback_blocks;
        pp.active = c->d_active; pp.skipped_before = c->d_skipped; pp.counters = c->d_counters; pp.lookback = c->d_lookback;
        pp.cull = cull ? 1 : 0;
        k_frame_prologue<<<std::max(1u, (c->n_clusters + PROLOGUE_THREADS - 1) / PROLOGUE_THREADS), PROLOGUE_THREADS, 0, c->stream>>>(pp);
        c->launches++;
        if (cull) { sp.cluster_vis = c->d_cluster_vis; sp.active = c->d_active; sp.skipped_before = c->d_skipped; }
    }
    sp.sl.samples = c->d_samples; sp.sl.frag = c->d_sample_frag; sp.sl.cap = c->cap_samples; sp.sl.count = c->d_counters + CTR_NSAMPLES;
    sp.sl.extra = c->d_extra; sp.sl.extra_count = c->d_counters + CTR_NEXTRA; sp.sl.extra_cap = c->cap_frags; sp.sl.overflow = c->d_counters + CTR_OVERFLOW;
    sp.worklist = c->d_worklist;
    if (c->banded) k_setup_main<true><<<c->lookback_blocks, SETUP_THREADS, 0, c->stream>>>(sp);
    else k_setup_main<false><<<c->lookback_blocks, SETUP_THREADS, 0, c->stream>>>(sp);
    c->launches++;
    if (split_clear && clear_at != 0 && (r = fork_clear())) return r;
    if (c->stage_events) CU(cudaEventRecord(c->ev[EV_SETUP], c->stream));
    // kernel1 / kernel2
    RasterParams rp;
    rp.frags = c->d_frags; rp.cutdown = c->d_cutdown; rp.fragcnt = c->d_fragcnt; rp.counters = c->d_counters; rp.cap_frags = c->cap_frags;
    rp.worklist = c->d_worklist; rp.extra = c->d_extra;
    rp.n_index = CTR_NFRAG;
    rp.depth = c->d_depth[c->cur]; rp.ids = c->d_ids[c->cur];
    rp.width = (float)c->W; rp.height = (float)c->H; rp.W = c->W;
    rp.row_lo = row0; rp.row_hi = row1; rp.rowmask = c->d_rowmask; rp.rowbit = ROW_NEEDED;
    if ((r = raster_depth(c, c->stream, rp, sp.sl))) return r;
    if (c->stage_events) CU(cudaEventRecord(c->ev[EV_DEPTH], c->stream));
    rp.row_lo = band0; rp.row_hi = band1; rp.rowbit = ROW_OWNED;
    // kernel2: a stream over the samples both depth kernels recorded; with the streaming stores on the side stream it also builds
    // kernel3's covered-pixel list (the first sample to resolve a pixel appends it), so no pass over the screen is left in the frame
    // (a second draw without rr_swap_buffers in between finds the ids of the first one: "first to resolve" is then no criterion and
    // the list comes from the pass over the screen, as it did before)
    const bool list_from_ids = c->list_from_ids && split_clear && !(mg_composite && c->mg.rank != 0) && c->swapped;
    c->swapped = false;
    rp.shade_list = c->d_shade_list; rp.shade_count = c->d_counters + CTR_NSHADE;
    rp.tile_now = c->tile_mark_target; rp.tiles_x = (c->W + TILE_W - 1) / TILE_W;      // rr_frame_e2e with the dirty-tile read-back: the id resolve marks the tiles
    c->tiles_marked = list_from_ids && c->tile_mark_target != nullptr;
    if (list_from_ids) k_ids_list<true><<<grid_for(c, 6), 256, 0, c->stream>>>(sp.sl, rp);
    else k_ids_list<false><<<grid_for(c, 8), 256, 0, c->stream>>>(sp.sl, rp);
    c->launches++;
    if (c->stage_events) CU(cudaEventRecord(c->ev[EV_IDS], c->stream));
    // kernel3
    if (mg_composite && c->mg.rank != 0) {
        // first store of this frame into rank 0's colour target (the clear colour of k_shade_pre): not before rank 0 says the
        // target is free — the shadow-epoch wait further down comes too late for it and does not exist without shadow lights
        MgWait w;
        w.n = 1; w.flag[0] = &c->mg.ctrl->fb_free; w.value = c->mg.draw_epoch; w.error = &c->mg.ctrl->error; w.timeout_ns = mg_timeout_ns();
        k_wait_flags<<<1, 32, 0, c->stream>>>(w);
        c->launches++;
    }
    if (c->W % 4 == 0) {
        dim3 grid4((c->W + 127) / 128, (row1 - row0 + 7) / 8);
        if (!split_clear) { k_shade_pre4<true, true><<<grid4, 256, 0, c->stream>>>(hp); c->launches++; }
        else {
            CU(cudaStreamWaitEvent(c->stream, c->ev_clear_done, 0));
            if (mg_composite && c->mg.rank != 0) { k_shade_pre4<false, true><<<grid4, 256, 0, c->stream>>>(hp); c->launches++; }
            else if (!list_from_ids) { k_shade_list<<<dim3((c->W + 127) / 128, (row1 - row0 + 8 * SL_ROWS - 1) / (8 * SL_ROWS)), 256, 0, c->stream>>>(hp); c->launches++; }
        }
    } else {
        dim3 grid((c->W + 31) / 32, (row1 - row0 + 7) / 8);
        k_shade_pre<<<grid, 256, 0, c->stream>>>(hp);
        c->launches++;
    }
    if ((r = join_shadows(c))) return r;                                                   // the cubemaps must be complete before shading
    if (c->mg.connected && c->mg.shadow_epoch && (r = mg_wait(c, c->stream, true, c->mg.shadow_epoch))) return r;   // ... the peers' faces too
    k_shade<<<grid_for(c, 48), 128, 0, c->stream>>>(hp);      // 4x more CTAs than fit: the tail of the stride loop balances better (0.232 -> 0.219 ms on c3)
    c->launches++;
    if (mg_composite && c->mg.rank != 0) {                     // this context's rows are in rank 0's colour target
        k_signal_flag<<<1, 1, 0, c->stream>>>(&c->mg.peer_ctrl[0]->draw_flag[c->mg.rank], c->mg.draw_epoch);
        c->launches++;
    }
    if (mg_composite && c->mg.rank == 0 && (r = mg_wait(c, c->stream, false, c->mg.draw_epoch))) return r;             // composite complete
    if (c->stage_events) CU(cudaEventRecord(c->ev[EV_SHADE], c->stream));
    c->have_frame_ev = c->stage_events;
    c->cam_last = cam;
    c->frame_id++;                                                                         // engine.cpp:2024
    CU(cudaGetLastError());
    return RR_OK;
}

// ---- post passes on the G-buffer (after rr_frame_draw, before rr_swap_buffers) ------------------------------------------------
// Every pass reads the colour target and writes the second one (d_post), which then takes its place.
static int post_begin(rr_ctx* c, const char* who) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    if (c->mg.connected || c->banded) return fail(RR_ERR_INVALID, "%s: needs the whole frame's G-buffer on one context", who);
    if (!c->d_post) CU(cudaMalloc((void**)&c->d_post, (size_t)c->W * c->H * 4));
    return RR_OK;
}
static int post_publish(rr_ctx* c) {
    const size_t P = (size_t)c->W * c->H;
    for (int i = 0; i < RR_RING_MAX; i++) c->tile_valid[i] = false;      // a post pass may touch any pixel: the next tile read-back is a full one
    CU(cudaGetLastError());
    if (c->ext_rgba8) {                                               // caller-owned target: put the result back where the caller expects it
        k_copy_u32<<<grid_for(c, 8), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(c->d_post), reinterpret_cast<uint4*>(c->d_rgba8), P / 4,
                                                          reinterpret_cast<const uint32_t*>(c->d_post), reinterpret_cast<uint32_t*>(c->d_rgba8), P);
        c->launches++;
        CU(cudaGetLastError());
    } else {                                                          // own target: the post target becomes the colour target
        for (int i = 0; i < RR_RING_MAX; i++) if (c->d_ring[i] == c->d_rgba8) c->d_ring[i] = c->d_post;
        std::swap(c->d_rgba8, c->d_post);
    }
    return RR_OK;
}

// engine::do_pseudo_aa, engine.cpp:1513-1516
int rr_post_pseudo_aa(rr_ctx* c) {
    int r;
    if ((r = post_begin(c, "rr_post_pseudo_aa"))) return r;
    if (c->n_tris == 0) return RR_OK;
    const float aa_arg = 20.f * 2 * RR_PI_F / 360.f;
    const float cosrad = (float)cos((double)aa_arg);                  // cos() pinned: double on the host, rounded to float
    dim3 grid((c->W + 31) / 32, (c->H + 7) / 8);
    k_pseudo_aa<<<grid, 256, 0, c->stream>>>(c->d_rgba8, c->d_post, c->d_depth[c->cur], c->d_normals, c->W, c->H, cosrad);
    c->launches++;
    return post_publish(c);
}

// engine::do_motion_blur, engine.cpp:1518-1538
int rr_post_motion_blur(rr_ctx* c, float strength, float camera_contribution) {
    int r;
    if ((r = post_begin(c, "rr_post_motion_blur"))) return r;
    if (c->n_tris == 0 || c->n_objs == 0) return RR_OK;
    if (c->seen_cap < c->n_objs) {
        if (c->d_seen) cudaFree(c->d_seen);
        CU(cudaMalloc((void**)&c->d_seen, c->n_objs));
        CU(cudaMemsetAsync(c->d_seen, 0, c->n_objs, c->stream));
        c->seen_cap = c->n_objs;
    }
    MotionBlurParams mp;
    mp.in = c->d_rgba8; mp.out = c->d_post; mp.depth = c->d_depth[c->cur]; mp.ids = c->d_ids[c->cur]; mp.frags = c->d_frags; mp.n_frags = c->d_counters + CTR_NFRAG;
    mp.objs = c->d_objs; mp.n_objs = c->n_objs; mp.seen = c->d_seen;
    mp.cam = c->cam_last; mp.cam_old = c->cam_old;
    mp.W = c->W; mp.H = c->H; mp.fov = c->fov; mp.icut = (float)c->cfg.depth_icutoff; mp.strength = strength; mp.camera_contribution = camera_contribution;
    mp.frame_id = c->frame_id;
    dim3 grid((c->W + 31) / 32, (c->H + 7) / 8);
    k_motion_blur<<<grid, 256, 0, c->stream>>>(mp);
    k_motion_history<<<(c->n_objs + 127) / 128, 128, 0, c->stream>>>(c->d_objs, c->n_objs, c->d_seen, c->frame_id);
    c->launches += 2;
    return post_publish(c);
}

// engine::draw_godrays, engine.cpp:1463-1482 (nothing to do unless a light has godray_intensity > 0: light_data->any_godray)
int rr_post_godrays(rr_ctx* c) {
    int r;
    if ((r = post_begin(c, "rr_post_godrays"))) return r;
    bool any = false;
    for (const rr_light& l : c->lights) any = any || l.godray_intensity > 0;
    if (!any) return RR_OK;
    GodrayParams gp;
    gp.in = c->d_rgba8; gp.out = c->d_post; gp.depth = c->d_depth[c->cur]; gp.lights = c->d_lights; gp.n_lights = (int)c->lights.size();
    gp.cam = c->cam_last; gp.W = c->W; gp.H = c->H; gp.fov = c->fov;
    dim3 grid((c->W + 31) / 32, (c->H + 7) / 8);
    k_godrays<<<grid, 256, 0, c->stream>>>(gp);
    c->launches++;
    return post_publish(c);
}

int rr_swap_buffers(rr_ctx* c) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    c->cur ^= 1;                                           // depth_buffer.flip(), object_context.cpp:21
    c->swapped = true;
    c->cam_old = c->cam_last;                              // c_pos_old / c_rot_old, object_context.cpp:23-24
    return RR_OK;
}

int rr_sync(rr_ctx* c) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    int r;
    if ((r = join_shadows(c))) return r;
    CU(cudaMemcpyAsync(c->h_counters, c->d_counters, CTR_COUNT * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_counters + CTR_COUNT, c->d_scounters, CTR_COUNT * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->stream2));
    CU(cudaStreamSynchronize(c->stream3));
    for (int i = 0; i < RR_RING_MAX; i++) c->copy_pending[i] = false;
    if (c->mg.connected) {
        uint32_t perr = 0;
        CU(cudaMemcpy(&perr, &c->mg.ctrl->error, 4, cudaMemcpyDeviceToHost));
        if (perr) return fail(RR_ERR_PEER, "multi-GPU exchange: a peer context did not deliver its faces / rows within the time limit (rank %d of %d)", c->mg.rank, c->mg.world);
    }
    const uint32_t ovf = c->h_counters[CTR_OVERFLOW] | c->h_counters[CTR_STICKY] | c->h_counters[CTR_COUNT + CTR_OVERFLOW] | c->h_counters[CTR_COUNT + CTR_STICKY];
    c->last_overflow = ovf;
    if (ovf) {                                             // reported once: clear the current and the sticky words of both workspaces
        CU(cudaMemsetAsync(c->d_counters + CTR_OVERFLOW, 0, 4, c->stream));
        CU(cudaMemsetAsync(c->d_counters + CTR_STICKY, 0, 4, c->stream));
        CU(cudaMemsetAsync(c->d_scounters + CTR_OVERFLOW, 0, 4, c->stream));
        CU(cudaMemsetAsync(c->d_scounters + CTR_STICKY, 0, 4, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    if (ovf)
        return fail(RR_ERR_OVERFLOW, "raster storage exhausted (flags %u): fragments cap %u, projected-triangle cap %u", ovf, c->cap_frags, c->cap_cut);
    return RR_OK;
}

// ---- read-back -----------------------------------------------------------------------------------------------------
static int read_back(rr_ctx* c, void* dst, const void* src, size_t bytes) {
    if (!c || !dst) return fail(RR_ERR_INVALID, "null argument");
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return RR_OK;
}
int rr_read_depth(rr_ctx* c, uint32_t* dst) { return read_back(c, dst, c->d_depth[c->cur], (size_t)c->W * c->H * 4); }
int rr_read_ids(rr_ctx* c, uint32_t* dst) {
    int r = read_back(c, dst, c->d_ids[c->cur], (size_t)c->W * c->H * 4);
    if (r) return r;
    const size_t P = (size_t)c->W * c->H;                 // the device image holds fragment index + 1 (0 = unresolved)
    for (size_t i = 0; i < P; i++) dst[i] = dst[i] ? dst[i] - 1u : 0u;
    return RR_OK;
}
int rr_read_rgba8(rr_ctx* c, uint8_t* dst) { return read_back(c, dst, c->d_rgba8, (size_t)c->W * c->H * 4); }
int rr_read_normals(rr_ctx* c, uint16_t* dst) { return read_back(c, dst, c->d_normals, (size_t)c->W * c->H * 4); }
int rr_read_shadow(rr_ctx* c, int is_static, uint32_t slab_idx, uint32_t* dst) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    const size_t slab = (size_t)6 * c->L * c->L;
    if (slab_idx >= (is_static ? c->n_static : c->n_shadow)) return fail(RR_ERR_INVALID, "rr_read_shadow: slab %u out of range", slab_idx);
    { int jr = join_shadows(c); if (jr) return jr; }
    return read_back(c, dst, (is_static ? c->d_shadow_static : c->d_shadow_dyn) + slab * slab_idx, slab * 4);
}
int rr_read_fragments(rr_ctx* c, uint32_t* dst, uint32_t max_records, uint32_t* n_records) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    int r = rr_sync(c);
    if (r && r != RR_ERR_OVERFLOW) return r;
    uint32_t n = c->h_counters[CTR_NFRAG];
    if (n_records) *n_records = n;
    uint32_t k = std::min(std::min(n, max_records), c->cap_frags);
    if (dst && k) return read_back(c, dst, c->d_frags, (size_t)k * RR_FRAG_WORDS * 4);
    return RR_OK;
}
int rr_read_cutdown(rr_ctx* c, float* dst, uint32_t max_tris, uint32_t* n_tris) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    int r = rr_sync(c);
    if (r && r != RR_ERR_OVERFLOW) return r;
    uint32_t n = c->h_counters[CTR_NCUT];
    if (n_tris) *n_tris = n;
    uint32_t k = std::min(std::min(n, max_tris), c->cap_cut);
    if (dst && k) return read_back(c, dst, c->d_cutdown, (size_t)k * 48);
    return RR_OK;
}

int rr_set_profiling(rr_ctx* c, int on) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    c->stage_events = on != 0;
    if (!on) c->have_shadow_ev = c->have_frame_ev = false;
    return RR_OK;
}

int rr_get_timings(rr_ctx* c, rr_timings* t) {
    if (!c || !t) return fail(RR_ERR_INVALID, "null argument");
    memset(t, 0, sizeof *t);
    int r = rr_sync(c);
    if (r && r != RR_ERR_OVERFLOW) return r;
    if (c->have_shadow_ev) cudaEventElapsedTime(&t->shadow_depth_ms, c->ev[EV_SH0], c->ev[EV_SH1]);
    if (c->have_frame_ev) {
        cudaEventElapsedTime(&t->setup_ms, c->ev[EV_F0], c->ev[EV_SETUP]);
        cudaEventElapsedTime(&t->depth_ms, c->ev[EV_SETUP], c->ev[EV_DEPTH]);
        cudaEventElapsedTime(&t->id_ms, c->ev[EV_DEPTH], c->ev[EV_IDS]);
        cudaEventElapsedTime(&t->shade_ms, c->ev[EV_IDS], c->ev[EV_SHADE]);
        cudaEventElapsedTime(&t->frame_ms, c->ev[EV_F0], c->ev[EV_SHADE]);
    }
    t->n_cutdown = c->h_counters[CTR_NCUT];
    t->n_fragments = c->h_counters[CTR_NFRAG];
    t->n_shadow_fragments = c->h_counters[CTR_COUNT + CTR_S_NFRAG];    // stored (non-inlined) fragments of the last shadow pass
    t->overflow = c->last_overflow;
    t->launches = c->launches;
    return RR_OK;
}

// ---- multi-GPU hooks -----------------------------------------------------------------------------------------------
int rr_bind_external(rr_ctx* c, int which, void* p, size_t nbytes) {
    if (!c || !p) return fail(RR_ERR_INVALID, "null argument");
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->stream2));
    const size_t P = (size_t)c->W * c->H;
    switch (which) {
        case RR_BUF_RGBA8:
            if (nbytes < P * 4) return fail(RR_ERR_INVALID, "external RGBA8 buffer too small");
            CU(cudaStreamSynchronize(c->stream3));
            for (int i = 0; i < RR_RING_MAX; i++) c->copy_pending[i] = false;
            if (!c->ext_rgba8) {                           // own target(s): the ring of rr_frame_e2e may hold the current one
                bool in_ring = false;
                for (int i = 0; i < RR_RING_MAX; i++) {
                    if (!c->d_ring[i]) continue;
                    if (c->d_ring[i] == c->d_rgba8) in_ring = true;
                    cudaFree(c->d_ring[i]);
                    c->d_ring[i] = nullptr;
                }
                if (!in_ring) cudaFree(c->d_rgba8);
            }
            c->ring_pos = 0;
            c->d_rgba8 = (uchar4*)p; c->ext_rgba8 = true; return RR_OK;
        case RR_BUF_SHADOW_DYNAMIC:
            if (!c->ext_shadow_dyn) cudaFree(c->d_shadow_dyn);
            cudaFree(c->d_shadow_alt); c->d_shadow_alt = nullptr; c->alt_clean = false;
            c->d_shadow_dyn = (uint32_t*)p; c->ext_shadow_dyn = true; c->shadow_dyn_words = nbytes / 4; return RR_OK;
        case RR_BUF_SHADOW_STATIC:
            if (!c->ext_shadow_static) cudaFree(c->d_shadow_static);
            c->d_shadow_static = (uint32_t*)p; c->ext_shadow_static = true; c->shadow_static_words = nbytes / 4; return RR_OK;
        default: return fail(RR_ERR_INVALID, "rr_bind_external: buffer %d cannot be bound", which);
    }
}
void* rr_device_ptr(rr_ctx* c, int which) {
    if (!c) return nullptr;
    switch (which) {
        case RR_BUF_RGBA8: return c->d_rgba8;
        case RR_BUF_SHADOW_DYNAMIC: return c->d_shadow_dyn;
        case RR_BUF_SHADOW_STATIC: return c->d_shadow_static;
        case RR_BUF_DEPTH: return c->d_depth[c->cur];
        case RR_BUF_IDS: return c->d_ids[c->cur];
        default: return nullptr;
    }
}
void* rr_stream(rr_ctx* c) { return c ? (void*)c->stream : nullptr; }
void* rr_shadow_stream(rr_ctx* c) { return c ? (void*)c->stream2 : nullptr; }
int rr_shadows_done(rr_ctx* c) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    CU(cudaEventRecord(c->ev_shadow_done, c->stream2));      // everything enqueued on the shadow stream so far (e.g. the face all-gather)
    c->shadow_pending = true;
    return RR_OK;
}

// ---- multi-GPU exchange over peer memory ----------------------------------------------------------------------------
static int mg_alloc(rr_ctx* c) {
    if (c->mg.exported) return RR_OK;
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->stream2));
    const size_t P = (size_t)c->W * c->H;
    c->mg.shadow_words = std::max<size_t>((size_t)6 * c->L * c->L * c->n_shadow, 4);
    uint32_t* sh = nullptr;
    CU(cudaMalloc((void**)&sh, c->mg.shadow_words * 2 * 4));
    c->mg.shadow[0] = sh; c->mg.shadow[1] = sh + c->mg.shadow_words;
    CU(cudaMemset(sh, 0xFF, c->mg.shadow_words * 2 * 4));
    uchar4* fb = nullptr;
    CU(cudaMalloc((void**)&fb, P * RR_RING_MAX * 4));
    for (int i = 0; i < RR_RING_MAX; i++) c->mg.fb[i] = fb + (size_t)i * P;
    CU(cudaMemset(fb, 0, P * RR_RING_MAX * 4));
    CU(cudaMalloc((void**)&c->mg.ctrl, sizeof(MgCtrl)));
    CU(cudaMemset(c->mg.ctrl, 0, sizeof(MgCtrl)));
    c->mg.exported = true;
    return RR_OK;
}

int rr_mgpu_export(rr_ctx* c, rr_mgpu_handle* out) {
    if (!c || !out) return fail(RR_ERR_INVALID, "null argument");
    if (c->mg.connected) return fail(RR_ERR_INVALID, "rr_mgpu_export: already connected");
    int r;
    if ((r = mg_alloc(c))) return r;
    memset(out, 0, sizeof *out);
    static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle does not fit rr_mgpu_handle");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->mg.shadow[0])); memcpy(out->shadow, &h, sizeof h);
    CU(cudaIpcGetMemHandle(&h, c->mg.fb[0])); memcpy(out->fb, &h, sizeof h);
    CU(cudaIpcGetMemHandle(&h, c->mg.ctrl)); memcpy(out->ctrl, &h, sizeof h);
    out->shadow_bytes = c->mg.shadow_words * 2 * 4; out->fb_bytes = (uint64_t)c->W * c->H * RR_RING_MAX * 4;
    out->device = c->cfg.device; out->n_shadow = (int32_t)c->n_shadow; out->width = c->W; out->height = c->H; out->light_dim = c->L;
    return RR_OK;
}

// common tail of rr_mgpu_connect / rr_mgpu_connect_local once the peer pointers are in place
static int mg_finish_connect(rr_ctx* c, int rank, int world) {
    if (c->cfg.face_world != world || c->cfg.face_rank != rank)
        return fail(RR_ERR_INVALID, "rr_mgpu_connect: rr_config.face_rank/face_world (%d/%d) must equal rank/world (%d/%d)", c->cfg.face_rank, c->cfg.face_world, rank, world);
    c->mg.rank = rank; c->mg.world = world;
    uint32_t pairs[6 * SHADOW_MAX_LIGHTS];
    if (c->n_shadow > SHADOW_MAX_LIGHTS) return fail(RR_ERR_INVALID, "rr_mgpu_connect: more than %d shadow-casting lights", SHADOW_MAX_LIGHTS);
    const int np = mg_owned_pairs(c, pairs);
    const size_t chunks = std::max<size_t>(1, (size_t)np * ((size_t)c->L * c->L / MG_PUSH_CHUNK_WORDS));
    if (((size_t)c->L * c->L) % MG_PUSH_CHUNK_WORDS) return fail(RR_ERR_INVALID, "rr_mgpu_connect: light_dim^2 must be a multiple of %d", MG_PUSH_CHUNK_WORDS);
    for (int b = 0; b < 2; b++) {
        CU(cudaMalloc((void**)&c->mg.prev_dirty[b], chunks));
        CU(cudaMemset(c->mg.prev_dirty[b], 0, chunks));
    }
    // the context now renders into the exchange buffers
    c->mg.saved_shadow_dyn = c->d_shadow_dyn; c->mg.saved_ext_shadow = c->ext_shadow_dyn; c->mg.saved_shadow_words = c->shadow_dyn_words;
    c->mg.saved_rgba8 = c->d_rgba8; c->mg.saved_ext_rgba8 = c->ext_rgba8;
    c->d_shadow_dyn = c->mg.shadow[0]; c->ext_shadow_dyn = true; c->shadow_dyn_words = c->mg.shadow_words;
    c->d_rgba8 = rank == 0 ? c->mg.fb[0] : c->mg.fb0[0]; c->ext_rgba8 = true;
    c->mg.shadow_epoch = c->mg.draw_epoch = 0;
    c->mg.connected = true;
    return RR_OK;
}

int rr_mgpu_connect(rr_ctx* c, int rank, int world, const rr_mgpu_handle* handles) {
    if (!c || !handles) return fail(RR_ERR_INVALID, "null argument");
    if (world < 1 || world > MG_MAX_WORLD || rank < 0 || rank >= world) return fail(RR_ERR_INVALID, "rr_mgpu_connect: bad rank/world %d/%d", rank, world);
    if (!c->mg.exported) return fail(RR_ERR_INVALID, "rr_mgpu_connect before rr_mgpu_export");
    if (c->mg.connected) return fail(RR_ERR_INVALID, "rr_mgpu_connect: already connected");
    CU(cudaSetDevice(c->cfg.device));
    const size_t P = (size_t)c->W * c->H;
    for (int q = 0; q < world; q++) {
        const rr_mgpu_handle& h = handles[q];
        if (h.width != c->W || h.height != c->H || h.light_dim != c->L || h.n_shadow != (int32_t)c->n_shadow)
            return fail(RR_ERR_INVALID, "rr_mgpu_connect: rank %d was created with a different frame / cubemap / light configuration", q);
        if (q == rank) continue;
        cudaIpcMemHandle_t ih;
        void* p = nullptr;
        memcpy(&ih, h.shadow, sizeof ih);
        CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        c->mg.opened[q][0] = p;
        c->mg.peer_shadow[q][0] = (uint32_t*)p; c->mg.peer_shadow[q][1] = (uint32_t*)p + c->mg.shadow_words;
        memcpy(&ih, h.ctrl, sizeof ih);
        CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        c->mg.opened[q][1] = p;
        c->mg.peer_ctrl[q] = (MgCtrl*)p;
        if (q == 0) {
            memcpy(&ih, h.fb, sizeof ih);
            CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
            c->mg.opened[q][2] = p;
            for (int i = 0; i < RR_RING_MAX; i++) c->mg.fb0[i] = (uchar4*)p + (size_t)i * P;
        }
    }
    c->mg.ipc = true;
    return mg_finish_connect(c, rank, world);
}

int rr_mgpu_connect_local(rr_ctx* const* ctxs, int world) {
    if (!ctxs || world < 1 || world > MG_MAX_WORLD) return fail(RR_ERR_INVALID, "rr_mgpu_connect_local: bad arguments");
    int r;
    for (int k = 0; k < world; k++) {
        if (!ctxs[k]) return fail(RR_ERR_INVALID, "null ctx");
        if (ctxs[k]->mg.connected) return fail(RR_ERR_INVALID, "rr_mgpu_connect_local: context %d already connected", k);
        if (ctxs[k]->W != ctxs[0]->W || ctxs[k]->H != ctxs[0]->H || ctxs[k]->L != ctxs[0]->L || ctxs[k]->n_shadow != ctxs[0]->n_shadow)
            return fail(RR_ERR_INVALID, "rr_mgpu_connect_local: context %d has a different frame / cubemap / light configuration", k);
        if ((r = mg_alloc(ctxs[k]))) return r;
    }
    for (int k = 0; k < world; k++)
        for (int q = 0; q < world; q++) {
            if (q == k || ctxs[q]->cfg.device == ctxs[k]->cfg.device) continue;
            CU(cudaSetDevice(ctxs[k]->cfg.device));
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->cfg.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(RR_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctxs[k]->cfg.device, ctxs[q]->cfg.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
    for (int k = 0; k < world; k++) {
        rr_ctx* c = ctxs[k];
        for (int q = 0; q < world; q++) {
            if (q == k) continue;
            c->mg.peer_shadow[q][0] = ctxs[q]->mg.shadow[0]; c->mg.peer_shadow[q][1] = ctxs[q]->mg.shadow[1];
            c->mg.peer_ctrl[q] = ctxs[q]->mg.ctrl;
        }
        for (int i = 0; i < RR_RING_MAX; i++) c->mg.fb0[i] = ctxs[0]->mg.fb[i];
        c->mg.ipc = false;
        CU(cudaSetDevice(c->cfg.device));
        if ((r = mg_finish_connect(c, k, world))) return r;
    }
    return RR_OK;
}

int rr_mgpu_disconnect(rr_ctx* c) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    if (c->stream3) cudaStreamSynchronize(c->stream3);
    if (c->mg.connected) {
        c->d_shadow_dyn = c->mg.saved_shadow_dyn; c->ext_shadow_dyn = c->mg.saved_ext_shadow; c->shadow_dyn_words = c->mg.saved_shadow_words;
        c->d_rgba8 = c->mg.saved_rgba8; c->ext_rgba8 = c->mg.saved_ext_rgba8;
        for (int b = 0; b < 2; b++) { cudaFree(c->mg.prev_dirty[b]); c->mg.prev_dirty[b] = nullptr; }
        c->mg.connected = false;
    }
    for (int q = 0; q < MG_MAX_WORLD; q++)               // also after a connect that failed half-way
        for (int i = 0; i < 3; i++) if (c->mg.opened[q][i]) { cudaIpcCloseMemHandle(c->mg.opened[q][i]); c->mg.opened[q][i] = nullptr; }
    if (c->mg.exported) {
        cudaFree(c->mg.shadow[0]); cudaFree(c->mg.fb[0]); cudaFree(c->mg.ctrl);
        c->mg.shadow[0] = c->mg.shadow[1] = nullptr; for (int i = 0; i < RR_RING_MAX; i++) c->mg.fb[i] = nullptr;
        c->mg.ctrl = nullptr;
        c->mg.exported = false;
    }
    cudaGetLastError();
    return RR_OK;
}

// bytes this context's k_push_faces has stored into peer memory (over NVLink) since the last call: 512-byte chunks x peers
int rr_mgpu_pushed_bytes(rr_ctx* c, uint64_t* bytes) {
    if (!c || !bytes) return fail(RR_ERR_INVALID, "null argument");
    *bytes = 0;
    if (!c->mg.connected) return RR_OK;
    CU(cudaStreamSynchronize(c->stream2));
    uint32_t chunks = 0;
    CU(cudaMemcpy(&chunks, &c->mg.ctrl->pushed_chunks, 4, cudaMemcpyDeviceToHost));
    CU(cudaMemset(&c->mg.ctrl->pushed_chunks, 0, 4));
    *bytes = (uint64_t)chunks * MG_PUSH_CHUNK_WORDS * 4 * (uint64_t)std::max(0, c->mg.world - 1);
    return RR_OK;
}

int rr_mgpu_set_readback(rr_ctx* c, int distributed) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->stream3));
    c->mg.local_readback = distributed != 0;
    return RR_OK;
}

int rr_host_register(void* p, size_t nbytes) {
    if (!p || !nbytes) return fail(RR_ERR_INVALID, "rr_host_register: null argument");
    CU(cudaHostRegister(p, nbytes, cudaHostRegisterPortable));
    return RR_OK;
}
int rr_host_unregister(void* p) {
    if (!p) return fail(RR_ERR_INVALID, "rr_host_unregister: null argument");
    CU(cudaHostUnregister(p));
    return RR_OK;
}

void* rr_host_alloc(size_t nbytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, nbytes ? nbytes : 1) != cudaSuccess) { fail(RR_ERR_OOM, "rr_host_alloc(%zu) failed", nbytes); return nullptr; }
    return p;
}
void rr_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- host-to-host frame --------------------------------------------------------------------------------------------
int rr_set_pipeline_depth(rr_ctx* c, int depth) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    if (depth < 2 || depth > RR_RING_MAX) return fail(RR_ERR_INVALID, "rr_set_pipeline_depth: depth must be 2..%d", RR_RING_MAX);
    int r = rr_sync(c);
    if (r && r != RR_ERR_OVERFLOW) return r;
    c->ring_depth = depth; c->ring_pos = 0;
    for (int i = 0; i < RR_RING_MAX; i++) c->tile_valid[i] = false;      // (dirty-tile read-back: the slots' host buffers change)
    return RR_OK;
}

int rr_set_readback_tiles(rr_ctx* c, int on) {
    if (!c) return fail(RR_ERR_INVALID, "null ctx");
    int r = rr_sync(c);
    if (r && r != RR_ERR_OVERFLOW) return r;
    c->tiles_on = on != 0;
    c->tile_estimate = 0;
    for (int i = 0; i < RR_RING_MAX; i++) c->tile_valid[i] = false;
    return RR_OK;
}

int rr_readback_tile_bytes(rr_ctx* c, uint64_t* bytes) {
    if (!c || !bytes) return fail(RR_ERR_INVALID, "null argument");
    *bytes = 0;
    if (!c->d_tile_sent) return RR_OK;
    CU(cudaStreamSynchronize(c->stream3));
    uint32_t n = 0;
    CU(cudaMemcpy(&n, c->d_tile_sent, 4, cudaMemcpyDeviceToHost));
    CU(cudaMemset(c->d_tile_sent, 0, 4));
    *bytes = (uint64_t)n * TILE_W * TILE_H * 4 + c->tile_dma_bytes;
    c->tile_dma_bytes = 0;
    return RR_OK;
}

int rr_frame_e2e(rr_ctx* c, const float c_pos[4], const float c_rot[4], const float clear_rgba[4], int with_shadows, uint8_t* host_rgba8) {
    if (!c || !host_rgba8) return fail(RR_ERR_INVALID, "null argument");
    int r;
    // per-frame input: the object descriptors (object_context::flush_locations, object_context.cpp:819) from pinned memory
    // (snapshotted into the staging arena: the caller may patch the mirror for the next frame as soon as this call returns,
    // while this frame's copy has not executed yet — the ring keeps several frames in flight)
    if (c->n_objs) {
        if ((r = upload_staged(c, c->d_objs, c->h_objs_pinned, (size_t)c->n_objs * sizeof(rr_obj_desc)))) return r;
        c->objlite_dirty = true;
    }
    // Pipelined read-back over a ring of D colour targets (the reference keeps a ring of host buffers for the same reason,
    // async_read.hpp:30-144): frame n is drawn into target n % D while the copy stream is still moving earlier frames out of
    // the others. On return the host buffer passed D-1 calls ago is complete (D = 2: the previous call's); rr_sync()
    // completes all. The caller cycles through D host buffers.
    const int D = c->ring_depth, k = c->ring_pos;
    const bool mg = c->mg.connected;
    const bool pipelined = mg || !c->ext_rgba8;
    const size_t P = (size_t)c->W * c->H, rowb = (size_t)c->W * 4;
    if (mg) c->mg_target = k;                       // rr_frame_draw picks the local or rank-0 target of the ring
    else if (pipelined) {
        if (!c->d_ring[0]) c->d_ring[0] = c->d_rgba8;
        if (!c->d_ring[k]) CU(cudaMalloc((void**)&c->d_ring[k], P * 4));
        c->d_rgba8 = c->d_ring[k];
    }
    // target k is free again once the copy of frame n - D is done. Multi-GPU composite: rank 0's main stream waits for that and
    // then publishes the draw epoch in every peer's control block (fb_free, rr_frame_draw); a peer's first store into rank 0's
    // target waits for it.
    if (pipelined && c->copy_pending[k]) CU(cudaStreamWaitEvent(c->stream, c->ev_copy_done[k], 0));
    // dirty-tile read-back (rr_set_readback_tiles): the frame's tiles are marked by the id resolve (or, when that does not build the
    // pixel list, by k_tile_mark behind k_shade), the copy stream stores the tiles that need it into the mapped host buffer
    bool tiles_enqueued = false;
    uchar4* host_dev = nullptr;
    // whole frames of one context, or — distributed read-back of a connected context — its interleaved row tiles (whole multiples of
    // the 4-row read-back tiles, so that no tile belongs to two contexts)
    const bool tiles_whole = !mg && !c->banded && c->own_lo == 0 && c->own_hi == c->H;
    const bool tiles_rows = mg && c->mg.local_readback && c->cfg.band_tile > 0 && c->cfg.band_world > 1 && c->cfg.band_tile % TILE_H == 0;
    const bool tiles = c->tiles_on && (tiles_whole || tiles_rows) && pipelined && c->W % 4 == 0 && c->n_tris > 0 &&
                       cudaHostGetDevicePointer((void**)&host_dev, host_rgba8, 0) == cudaSuccess && host_dev;
    if (c->tiles_on && !tiles) (void)cudaGetLastError();        // (a host buffer that is not page-locked: the plain copy below)
    const int tiles_x = (c->W + TILE_W - 1) / TILE_W, tiles_y = (c->H + TILE_H - 1) / TILE_H;
    const uint32_t n_tiles = (uint32_t)tiles_x * (uint32_t)tiles_y;
    if (tiles) {
        if (!c->d_tile_sent) {
            CU(cudaMalloc((void**)&c->d_tile_sent, 4 * (1 + RR_RING_MAX))); CU(cudaMemsetAsync(c->d_tile_sent, 0, 4 * (1 + RR_RING_MAX), c->stream));
            CU(cudaMallocHost((void**)&c->h_tile_last, 4 * RR_RING_MAX)); memset(c->h_tile_last, 0, 4 * RR_RING_MAX);
        }
        if (!c->d_tile_now[k]) {
            CU(cudaMalloc((void**)&c->d_tile_now[k], n_tiles)); CU(cudaMalloc((void**)&c->d_tile_prev[k], n_tiles));
            CU(cudaMemsetAsync(c->d_tile_now[k], 0, n_tiles, c->stream)); CU(cudaMemsetAsync(c->d_tile_prev[k], 0, n_tiles, c->stream));
            c->tile_valid[k] = false;
        }
        c->tile_mark_target = c->d_tile_now[k];
    }
    c->tiles_marked = false;
    if (with_shadows && (r = rr_frame_shadows(c, 0))) { c->tile_mark_target = nullptr; return r; }
    r = rr_frame_draw(c, c_pos, c_rot, clear_rgba);
    c->tile_mark_target = nullptr;
    if (r) return r;
    if (tiles && !c->tiles_marked) {
        k_tile_mark<<<grid_for(c, 4), 256, 0, c->stream>>>(c->d_shade_list, c->d_counters + CTR_NSHADE, c->W, tiles_x, c->d_tile_now[k]);
        c->launches++;
    }
    CU(cudaEventRecord(c->ev_draw_done, c->stream));
    CU(cudaStreamWaitEvent(c->stream3, c->ev_draw_done, 0));
    const uint8_t* src = (const uint8_t*)c->d_rgba8;
    if (tiles) {
        const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
        const float* cl = clear_rgba ? clear_rgba : zero4;
        // the whole buffer (this context's rows of it) must be rewritten — first use of it in this slot, another clear colour — or the
        // last completed copy needed more than half of the tiles: the copy engine moves the frame as without this mode (below), the
        // kernel only keeps the books
        const uint32_t own_tiles = tiles_rows ? n_tiles / (uint32_t)c->cfg.band_world : n_tiles;
        bool all = !c->tile_valid[k] || c->tile_host[k] != (const void*)host_rgba8 || memcmp(c->tile_clear[k], cl, 16) != 0 ||
                   c->tile_estimate > own_tiles / 2;
        for (int j = 0; j < RR_RING_MAX; j++)                   // the same buffer written through another slot since: what this slot remembers of it is stale
            if (j != k && c->tile_host[j] == (const void*)host_rgba8 && c->tile_seq[j] > c->tile_seq[k]) all = true;
        c->tile_seq[k] = ++c->tile_calls;
        CU(cudaMemsetAsync(c->d_tile_sent + 1 + k, 0, 4, c->stream3));
        // grid: about one CTA per 2 000 tiles (1 MB) the last completed copy sent, at least c->tile_grid (RR_TILE_GRID, 4)
        const int grid = all ? 8 : std::min(64, std::max(c->tile_grid, (int)(c->tile_estimate / 2000u)));
        k_tile_copy<<<grid, 256, 0, c->stream3>>>(c->d_rgba8, all ? nullptr : host_dev, c->W, c->H, tiles_x, n_tiles, c->d_tile_now[k], c->d_tile_prev[k], 0,
                                                 c->d_tile_sent, c->d_tile_sent + 1 + k);
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(c->h_tile_last + k, c->d_tile_sent + 1 + k, 4, cudaMemcpyDeviceToHost, c->stream3));
        c->tile_host[k] = host_rgba8; memcpy(c->tile_clear[k], cl, 16); c->tile_valid[k] = true;
        tiles_enqueued = !all;                                  // the tiles are on their way; otherwise the copies below move the frame
        if (all) c->tile_dma_bytes += tiles_rows ? (P * 4) / (size_t)c->cfg.band_world : P * 4;
    }
    if (tiles_enqueued) {
        // (stored by k_tile_copy)
    } else if (mg && c->mg.local_readback) {
        // distributed read-back: this context's rows only, from its own target, over its own PCIe link
        if (c->cfg.band_tile > 0 && c->cfg.band_world > 1) {
            const int tile = c->cfg.band_tile, world = c->cfg.band_world, rank = c->cfg.band_rank;
            const int n_tiles = (c->H + tile - 1) / tile;
            int full = 0;                                       // owned tiles that are complete: one strided copy
            for (int t = rank; t < n_tiles; t += world) if ((t + 1) * tile <= c->H) full++;
            const size_t off = (size_t)rank * tile * rowb, pitch = (size_t)tile * world * rowb;
            if (full) CU(cudaMemcpy2DAsync(host_rgba8 + off, pitch, src + off, pitch, (size_t)tile * rowb, (size_t)full, cudaMemcpyDeviceToHost, c->stream3));
            const int t_last = n_tiles - 1;                     // a partial last tile, if it is ours
            if (t_last % world == rank && (t_last + 1) * tile > c->H) {
                const size_t o2 = (size_t)t_last * tile * rowb;
                CU(cudaMemcpyAsync(host_rgba8 + o2, src + o2, (size_t)(c->H - t_last * tile) * rowb, cudaMemcpyDeviceToHost, c->stream3));
            }
        } else if (c->own_hi > c->own_lo) {
            const size_t off = (size_t)c->own_lo * rowb;
            CU(cudaMemcpyAsync(host_rgba8 + off, src + off, (size_t)(c->own_hi - c->own_lo) * rowb, cudaMemcpyDeviceToHost, c->stream3));
        }
    } else if (mg) {
        // composite on rank 0 (every context stored its rows there over NVLink): rank 0 alone reads the frame back
        if (c->mg.rank == 0) CU(cudaMemcpyAsync(host_rgba8, src, P * 4, cudaMemcpyDeviceToHost, c->stream3));
    } else {
        const size_t off = (size_t)c->own_lo * rowb, len = (size_t)(c->own_hi - c->own_lo) * rowb;
        if (len) CU(cudaMemcpyAsync(host_rgba8 + off, src + off, len, cudaMemcpyDeviceToHost, c->stream3));   // direct DMA when host_rgba8 is page-locked
    }
    if (!tiles && pipelined) c->tile_valid[k] = false;           // (the tile path, when it comes back, starts from a full copy)
    CU(cudaEventRecord(c->ev_copy_done[k], c->stream3));
    c->copy_pending[k] = true;
    if (!pipelined) {                                     // one caller-owned target: nothing to overlap with
        CU(cudaEventSynchronize(c->ev_copy_done[k]));
        c->copy_pending[k] = false;
    } else {
        const int j = (k + 1) % D;                        // the frame issued D-1 calls ago
        if (c->copy_pending[j]) { CU(cudaEventSynchronize(c->ev_copy_done[j])); c->copy_pending[j] = false; }
        if (c->tiles_on && c->h_tile_last) c->tile_estimate = c->h_tile_last[j];      // tiles that copy needed: sizes the next grids
    }
    c->ring_pos = (k + 1) % D;
    return rr_swap_buffers(c);
}

// ---- micro-benchmarks ----------------------------------------------------------------------------------------------
int rr_microbench_atomic_min(rr_ctx* c, size_t footprint_bytes, uint64_t n_ops, float* ms_out) {
    if (!c || !ms_out) return fail(RR_ERR_INVALID, "null argument");
    size_t words = 1;
    while (words * 2 * 4 <= footprint_bytes) words *= 2;
    uint32_t* buf = nullptr;
    CU(cudaMalloc((void**)&buf, words * 4));
    CU(cudaMemsetAsync(buf, 0xFF, words * 4, c->stream));
    const int blocks = grid_for(c, 8), threads = 256;
    uint32_t iters = (uint32_t)std::max<uint64_t>(1, n_ops / ((uint64_t)blocks * threads));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_bench_atomic_min<<<blocks, threads, 0, c->stream>>>(buf, (uint32_t)(words - 1), 16);      // warm-up
    cudaEventRecord(a, c->stream);
    k_bench_atomic_min<<<blocks, threads, 0, c->stream>>>(buf, (uint32_t)(words - 1), iters);
    cudaEventRecord(b, c->stream);
    c->launches += 2;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(RR_ERR_CUDA, "atomic microbench: %s", cudaGetErrorString(e));
    // report the time normalised to exactly n_ops
    *ms_out = ms * (float)((double)n_ops / ((double)iters * blocks * threads));
    return RR_OK;
}

int rr_microbench_copy(rr_ctx* c, size_t nbytes, float* ms_out) {
    if (!c || !ms_out) return fail(RR_ERR_INVALID, "null argument");
    uint4 *src = nullptr, *dst = nullptr;
    size_t n16 = nbytes / 16;
    CU(cudaMalloc((void**)&src, n16 * 16));
    CU(cudaMalloc((void**)&dst, n16 * 16));
    CU(cudaMemsetAsync(src, 1, n16 * 16, c->stream));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_bench_copy<<<grid_for(c, 16), 256, 0, c->stream>>>(src, dst, n16);
    cudaEventRecord(a, c->stream);
    k_bench_copy<<<grid_for(c, 16), 256, 0, c->stream>>>(src, dst, n16);
    cudaEventRecord(b, c->stream);
    c->launches += 2;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaEventElapsedTime(ms_out, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(src); cudaFree(dst);
    if (e != cudaSuccess) return fail(RR_ERR_CUDA, "copy microbench: %s", cudaGetErrorString(e));
    return RR_OK;
}

}  // extern "C"

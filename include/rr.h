/*
 * rr.h — C ABI of the B200-native rasteriser that replaces OpenCLRenderer's OpenCL layer for ONE path:
 * the per-frame draw path  engine::generate_realtime_shadowing -> engine::draw_bulk_objs_n
 * (prearrange -> kernel1 -> kernel2 -> kernel3, plus the shadow-cubemap pair).
 *
 * Every entry point names the reference interface it replaces (file:line under the reference tree).
 * Conventions (same as the reference's single render thread):
 *   - plain C, opaque handle, `int` return: 0 = RR_OK, negative = error; rr_last_error() gives the text;
 *   - caller owns all host pointers; the context owns all device memory unless bound with rr_bind_external();
 *   - every call is asynchronous on the context's CUDA stream unless it is named rr_read_* / rr_sync;
 *   - a context is NOT thread-safe (neither is the reference: one cl::cqueue, main thread only);
 *   - there is no CPU fallback: rr_create() fails with RR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RR_H_INCLUDED
#define RR_H_INCLUDED

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_OK            0
#define RR_ERR_INVALID  (-1)   /* bad argument / call order */
#define RR_ERR_CUDA     (-2)   /* CUDA runtime failure (text in rr_last_error) */
#define RR_ERR_OOM      (-3)
#define RR_ERR_OVERFLOW (-4)   /* fragment / projected-triangle storage exhausted (reference: silent corruption, cl2.cl:4392) */
#define RR_ERR_PEER     (-5)   /* multi-GPU: a peer context did not deliver its part of the frame in time */

/* ---- device-visible PODs: byte-identical to the reference's host/device structs ------------------------------- */

/* struct vertex, cl2.cl:131-138 / vertex.hpp:33-37 — 48 bytes */
typedef struct rr_vertex {
    float    pos[4];
    float    normal[4];
    float    vt[2];
    uint32_t object_id;   /* host name `pad`; only vertex 0's is read (cl2.cl:4296) */
    uint32_t vertex_col;  /* RGBA8 packed r<<24|g<<16|b<<8|a, 0 = use texture (cl2.cl:5676-5686, 5933) */
} rr_vertex;

/* struct triangle, cl2.cl:148-151 / triangle.hpp:8-15 — 144 bytes */
typedef struct rr_triangle { rr_vertex vertices[3]; } rr_triangle;

/* struct obj_g_descriptor, cl2.cl:99-124 / obj_g_descriptor.hpp:9-36 — 136 bytes padded to 144 */
typedef struct rr_obj_desc {
    float    world_pos[4];
    float    world_rot_quat[4];
    float    old_world_pos_1[4];
    float    old_world_pos_2[4];
    float    old_world_rot_quat_1[4];
    float    old_world_rot_quat_2[4];
    float    scale;
    uint32_t tid, rid, ssid;
    uint32_t has_bump;
    float    specular, spec_mult, diffuse;
    int32_t  buffer_offset;
    int32_t  feature_flag;
    uint32_t _pad[2];
} rr_obj_desc;

/* enum object_feature_flag, cl2.cl:90-97 */
#define RR_FEATURE_SS_REFLECTIVE               1
#define RR_FEATURE_TWO_SIDED                   2
#define RR_FEATURE_OUTLINE                     4
#define RR_FEATURE_IS_STATIC                   8
#define RR_FEATURE_NO_DYNAMIC_SHADOWS         16

/* struct light, cl2.cl:77-87 / light.hpp:25-34 — 56 bytes padded to 64 */
typedef struct rr_light {
    float    pos[4];
    float    col[4];
    uint32_t shadow;
    float    brightness, radius, diffuse, godray_intensity;
    int32_t  is_static;
    uint32_t _pad[2];
} rr_light;

/* What the reference passes as OpenCL -D macros (ocl.h:227-236, main.cpp:80-85, cl2.cl:7-23) plus placement. */
typedef struct rr_config {
    int32_t width, height;         /* SCREENWIDTH / SCREENHEIGHT */
    int32_t light_dim;             /* LIGHTBUFFERDIM (engine::l_size, engine.cpp:471 = 1024) */
    float   fov_const;             /* FOV_CONST literal; <= 0 -> rr_fov_const_from_hfov(hfov_deg, width) */
    float   hfov_deg;              /* engine.hpp:115 default 120 */
    int32_t depth_icutoff;         /* 20 */
    float   ambient;               /* AMBIENT 0.2 */
    float   ssao_rad;              /* SSAO_RAD 5 (main.cpp uses 2) */
    float   ssao_div;              /* SSAO_DIV 2.5 */
    float   mip_bias;              /* MIP_BIAS 1.1 */
    float   shadow_bias;           /* SHADOWBIAS 50 */
    float   shadow_exp;            /* SHADOWEXP 1 */
    int32_t test_linear;           /* -D TEST_LINEAR */
    int32_t use_linear_rendering;  /* object_context_data::use_linear_rendering (object_context.hpp:87) */
    int32_t no_ssao;               /* -D NO_SSAO */
    int32_t device;                /* CUDA device ordinal */
    /* sort-first split (no reference counterpart). 0,0 = whole screen. */
    int32_t band_y0, band_y1;      /* this context resolves ids and shades rows [band_y0, band_y1) */
    int32_t band_halo;             /* depth rows rasterised outside the band for SSAO; < 0 = every row */
    int32_t face_rank, face_world; /* shadow (light,face) pair p = 6*slab+face is rendered here iff p / ceil(6S/face_world) == face_rank; 0,0 = all */
    uint32_t max_fragments;        /* fragment-record capacity; 0 = 16 Mi (reference: 2 Mi, engine.cpp:601) */
    uint32_t max_cutdown;          /* projected-triangle capacity; 0 = derived from the triangle count */
    /* interleaved sort-first split (load balance when the geometry sits in a few screen rows): */
    int32_t band_tile;             /* > 0: row y belongs to this context iff (y / band_tile) % band_world == band_rank; band_y0/y1 ignored */
    int32_t band_rank, band_world;
    int32_t face_interleave;       /* 1: pair p is rendered here iff p % face_world == face_rank (peer-memory exchange); 0: contiguous chunks */
    int32_t cluster_cull;          /* per-frame culling of 128-triangle clusters (off-screen; rows rasterised elsewhere) before triangle setup:
                                      0 = only when the frame is split across contexts, 1 = always (worth it when much of the scene is off screen), -1 = never */
} rr_config;

typedef struct rr_timings {      /* CUDA-event milliseconds of the last rr_frame_* calls (reference: -DPROFILING, engine.hpp:625-640) */
    float shadow_clear_ms, shadow_setup_ms, shadow_depth_ms;
    float setup_ms, depth_ms, id_ms, shade_ms, frame_ms;
    uint32_t n_cutdown, n_fragments;        /* totals of the main pass (id_cutdown_tris / id_buffer_atomc) */
    uint32_t n_shadow_fragments;            /* sum over shadow passes */
    uint32_t overflow;                      /* non-zero if any capacity was exceeded */
    uint32_t launches;                      /* kernels launched by this library since rr_create */
} rr_timings;

typedef struct rr_ctx rr_ctx;

enum rr_buffer {                 /* names for rr_bind_external / rr_device_ptr */
    RR_BUF_RGBA8 = 0,            /* uchar4[W*H] headless colour target (replaces cl_gl_interop_texture.hpp:273-331) */
    RR_BUF_SHADOW_DYNAMIC = 1,   /* uint32[6*L*L*n_shadow]  engine::g_shadow_light_buffer (light.cpp:236) */
    RR_BUF_SHADOW_STATIC = 2,    /* uint32[6*L*L*n_static]  engine::g_static_shadow_light_buffer (light.cpp:253) */
    RR_BUF_DEPTH = 3,            /* uint32[W*H] current depth buffer (object_context.cpp:47) */
    RR_BUF_IDS = 4               /* uint32[W*H] fragment-id image (object_context.cpp:36-38); on the device: fragment index + 1, 0 = unresolved */
};

/* ---- context -------------------------------------------------------------------------------------------------- */
void     rr_default_config(rr_config* cfg);                         /* reference defaults, SURVEY.md §5 */
float    rr_fov_const_from_hfov(float hfov_deg, float screenwidth); /* engine.cpp:119-133 + std::to_string literal, engine.cpp:474-477 */
rr_ctx*  rr_create(const rr_config* cfg);                           /* oclstuff()+build() ocl.h:185-537, engine::load buffers engine.cpp:600-666, ensure_screen_buffers object_context.cpp:27-65 */
void     rr_destroy(rr_ctx* ctx);
const char* rr_last_error(void);
const char* rr_version(void);

/* ---- scene (object_context::build -> alloc_gpu, object_context.cpp:346-484) ------------------------------------ */
int rr_scene_alloc(rr_ctx*, uint32_t n_tris, uint32_t n_objs);                                   /* g_tri_mem / g_obj_desc / g_cut_tri_mem allocation */
int rr_scene_write_tris(rr_ctx*, uint32_t first, uint32_t count, const rr_triangle* tris);       /* enqueue_write_buffer_async(g_tri_mem) object_context.cpp:409-441 (+ fill_ids cl2.cl:4231 is the caller's job: object_id must be set) */
int rr_scene_write_objs(rr_ctx*, uint32_t first, uint32_t count, const rr_obj_desc* objs);       /* alloc_object_descriptors object_context.cpp:460-484 */
int rr_scene_patch_obj(rr_ctx*, uint32_t obj_id, uint32_t byte_off, uint32_t nbytes, const void* src); /* object::g_flush partial writes object.cpp:652-857 */
int rr_scene_read_objs(rr_ctx*, uint32_t first, uint32_t count, rr_obj_desc* dst);                 /* synchronous read-back of the device descriptors (clEnqueueReadBuffer on g_obj_desc): do_motion_blur updates their motion history on the device */

/* ---- asynchronous rebuild (object_context::build(async) + flip_buffers, object_context.cpp:520-797) ------------------
 * The reference rebuilds a changed scene into `new_gpu_dat` on a second queue while the old one keeps rendering, and flips
 * when the uploads are done. Same here: begin sizes a BACK scene (buffers are reused when they are large enough; growing
 * them costs one device synchronisation), the writes go to it on an upload stream, commit makes every frame enqueued
 * afterwards use it — the render stream waits for the uploads on the device, the host never blocks. Frames already
 * enqueued finish on the old scene, which becomes the next back scene. */
int rr_scene_build_begin(rr_ctx*, uint32_t n_tris, uint32_t n_objs);
int rr_scene_build_write_objs(rr_ctx*, uint32_t first, uint32_t count, const rr_obj_desc* objs);
int rr_scene_build_write_tris(rr_ctx*, uint32_t first, uint32_t count, const rr_triangle* tris);   /* page-locked `tris` must stay valid until rr_scene_build_ready() */
int rr_scene_build_ready(rr_ctx*);                /* 1 once the uploads have landed (object_context::ready_to_flip); commit does not need it */
int rr_scene_build_commit(rr_ctx*);               /* flip_buffers */

/* ---- texture atlas (texture_context::alloc_gpu texture_context.cpp:350-517) ------------------------------------ */
int rr_atlas_alloc(rr_ctx*, uint32_t n_slices, const uint32_t* nums, uint32_t n_nums,
                   const uint32_t* sizes, uint32_t n_sizes, uint32_t mipmap_start);              /* g_texture_array / g_texture_nums / g_texture_sizes */
int rr_atlas_upload(rr_ctx*, uint32_t gpu_id, const uint8_t* rgba, uint32_t w, uint32_t h, int flip); /* texture::update_me_to_gpu texture.cpp:323-358 = update_gpu_tex + generate_mips + 3x generate_mip_mips */
int rr_atlas_upload_batch(rr_ctx*, uint32_t n, const uint32_t* gpu_ids, const uint8_t* const* rgba, const uint32_t* w, const uint32_t* h, int flip);
                                                                                                 /* the upload loop of texture_context::alloc_gpu texture_context.cpp:478-517 for n textures at once: one staged copy, one launch per phase (base + 4 mip levels); same atlas bytes as n rr_atlas_upload calls */
int rr_atlas_fill_colour(rr_ctx*, uint32_t gpu_id, const float col_0_255[4], uint32_t w, uint32_t h); /* texture::update_gpu_texture_col texture.cpp:445-463 -> update_gpu_tex_colour cl2.cl:955-984 (texture and its 4 mips) */
int rr_atlas_upload_mono(rr_ctx*, uint32_t gpu_id, const uint8_t* raw, uint32_t len, uint32_t w, uint32_t h, int flip); /* texture::update_gpu_texture_mono texture.cpp:554-584 -> generate_from_raw cl2.cl:1006-1031 (stride = len / h; mips untouched) */
int rr_atlas_write_raw(rr_ctx*, const uint8_t* atlas, size_t nbytes);                            /* test hook: inject a prebuilt atlas (decouples shading parity from atlas parity) */
int rr_atlas_read_raw(rr_ctx*, uint8_t* dst, size_t nbytes);

/* ---- lights (light::build light.cpp:145-276; engine::set_light_data engine.cpp:741) ---------------------------- */
int rr_lights_write(rr_ctx*, const rr_light* lights, uint32_t n_active);                         /* active lights in lightlist order; allocates the cubemap slabs */

/* ---- per frame ------------------------------------------------------------------------------------------------- */
int rr_frame_shadows(rr_ctx*, int static_lights_dirty);                                          /* engine::generate_realtime_shadowing engine.cpp:1601-1790 */
int rr_frame_draw(rr_ctx*, const float c_pos[4], const float c_rot[4], const float clear_rgba[4]); /* engine::draw_bulk_objs_n engine.cpp:2356 -> render_tris 1794-2025 */
int rr_post_motion_blur(rr_ctx*, float strength, float camera_contribution);                     /* engine::do_motion_blur engine.cpp:1518-1538 -> do_motion_blur cl2.cl:6714-6860 (object history in the descriptors advances as there) */
int rr_post_godrays(rr_ctx*);                                                                    /* engine::draw_godrays engine.cpp:1463-1482 -> screenspace_godrays cl2.cl:1792-1917; no-op unless a light has godray_intensity > 0 */
int rr_post_pseudo_aa(rr_ctx*);                                                                  /* engine::do_pseudo_aa engine.cpp:1513 -> do_pseudo_aa cl2.cl:6437-6657; after rr_frame_draw, whole-frame contexts only */
int rr_swap_buffers(rr_ctx*);                                                                    /* object_context_data::swap_buffers object_context.cpp:17-25 */
int rr_sync(rr_ctx*);                                                                            /* cl::cqueue.finish(); reports RR_ERR_OVERFLOW */

/* ---- read-back of the frame just drawn (replaces clEnqueueReadImage cl_gl_interop_texture.hpp:235, async_read.hpp:51) */
int rr_read_depth(rr_ctx*, uint32_t* dst);            /* W*H uint32, the depth buffer kernel1 filled */
int rr_read_ids(rr_ctx*, uint32_t* dst);              /* W*H uint32 fragment indices (0 where nothing was drawn or no fragment passed kernel2's depth window) */
int rr_read_rgba8(rr_ctx*, uint8_t* dst);             /* W*H*4, q = (uint8)(clamp(c,0,1)*255+0.5) */
int rr_read_normals(rr_ctx*, uint16_t* dst);          /* W*H*2 ushort2, screen_normals_optional cl2.cl:6390 */
int rr_read_shadow(rr_ctx*, int is_static, uint32_t slab, uint32_t* dst);  /* 6*L*L uint32 of one light's cubemap */
int rr_read_fragments(rr_ctx*, uint32_t* dst, uint32_t max_records, uint32_t* n_records);   /* 5 words each, g_tid_buf engine.cpp:601 */
int rr_read_cutdown(rr_ctx*, float* dst, uint32_t max_tris, uint32_t* n_tris);              /* 12 floats each, g_cut_tri_mem */
int rr_set_profiling(rr_ctx*, int on);      /* per-stage CUDA timing events (the reference's -DPROFILING, engine.hpp:625-640). Off by default: every
                                               timing event is a timestamp write to host memory that the stream waits for, which costs ~2 % of a frame
                                               and ~15 % while a read-back DMA occupies the PCIe link. Counts and `launches` are always available. */
int rr_get_timings(rr_ctx*, rr_timings* out);

/* ---- multi-GPU plumbing hooks ---------------------------------------------------------------------------------- */
int   rr_bind_external(rr_ctx*, int which /*enum rr_buffer*/, void* device_ptr, size_t nbytes); /* render straight into caller-owned (e.g. torch / peer-mapped) memory */
void* rr_device_ptr(rr_ctx*, int which /*enum rr_buffer*/);
void* rr_stream(rr_ctx*);                                                                       /* cudaStream_t the context launches on */
void* rr_shadow_stream(rr_ctx*);   /* cudaStream_t of the shadow passes: rr_frame_shadows runs there, concurrently with rr_frame_draw's
                                      setup/depth/id kernels; rr_frame_draw waits for it right before shading */
int   rr_shadows_done(rr_ctx*);    /* call after enqueueing extra work on the shadow stream (the cubemap-face all-gather): shading waits for it too */

/* ---- multi-GPU exchange over peer memory (NVLink P2P), one context per GPU --------------------------------------------
 * Replaces nothing in the reference (single device); it is the north star's "bands composited to GPU 0 over NVLink, cubemap
 * faces sharded" step done inside the producing kernels instead of with collectives afterwards:
 *   - every context renders the cubemap faces it owns and pushes them into all peers' cubemap buffers (k_push_faces);
 *   - every context's shading kernel stores its rows straight into rank 0's colour target;
 *   - frame-counter flags in a small control block order producers and consumers (no host synchronisation per frame).
 * Call order: rr_lights_write -> rr_mgpu_export on every context -> exchange the handles (any transport; the Python harness
 * uses torch.distributed) -> rr_mgpu_connect with all handles in rank order -> frames, called in lock-step on all contexts.
 * rr_mgpu_connect_local wires contexts living in ONE process (same device or peer-accessible devices) without IPC. */
typedef struct rr_mgpu_handle {
    uint8_t  shadow[64], fb[64], ctrl[64];   /* cudaIpcMemHandle_t of the cubemap pair, the colour-target pair, the control block */
    uint64_t shadow_bytes, fb_bytes;
    int32_t  device, n_shadow, width, height, light_dim, _pad;
} rr_mgpu_handle;
int rr_mgpu_export(rr_ctx*, rr_mgpu_handle* out);
int rr_mgpu_connect(rr_ctx*, int rank, int world, const rr_mgpu_handle* handles /* [world], rank order */);
int rr_mgpu_connect_local(rr_ctx* const* ctxs, int world);   /* ctxs[k] becomes rank k */
int rr_mgpu_disconnect(rr_ctx*);
/* Where the colour rows go. 0 (default): every context stores its rows into rank 0's colour target over NVLink (the composite
 * lives on GPU 0; rr_frame_e2e on rank 0 reads it back). 1: rows stay on the GPU that shaded them and rr_frame_e2e copies
 * exactly those rows into `host_rgba8` — pass every context the SAME host frame (memory shared between the processes and
 * page-locked in each with rr_host_register): each GPU then moves 1/world of the frame over its own PCIe link. */
int rr_mgpu_set_readback(rr_ctx*, int distributed);
int rr_mgpu_pushed_bytes(rr_ctx*, uint64_t* bytes);          /* statistics: bytes of cubemap faces this context stored into its peers (NVLink) since the last call */
int rr_host_register(void* p, size_t nbytes);     /* cudaHostRegister(portable): make caller-owned (e.g. shared) memory a DMA target */
int rr_host_unregister(void* p);

void* rr_host_alloc(size_t nbytes);            /* page-locked host memory for read-backs (CL_MEM_ALLOC_HOST_PTR role; async_read.hpp:30-60 host buffers) */
void  rr_host_free(void* p);

/* ---- host-to-host frame: upload the per-frame inputs (object descriptors), draw, read RGBA8 back ------------------
 * Pipelined like the reference's read-back ring (async_read.hpp:30-144) over D colour targets (D = 2 unless changed with
 * rr_set_pipeline_depth): the call returns once the host buffer passed D-1 calls ago is complete (D = 2: the PREVIOUS
 * call's); rr_sync completes all. Cycle through D host buffers (page-locked ones from rr_host_alloc make the copy a
 * direct DMA). Swaps buffers itself. D = 3 keeps two frames in flight, which hides the host's launch time behind the copy. */
#define RR_RING_MAX 4
int rr_frame_e2e(rr_ctx*, const float c_pos[4], const float c_rot[4], const float clear_rgba[4],
                 int with_shadows, uint8_t* host_rgba8);
int rr_set_pipeline_depth(rr_ctx*, int depth);     /* 2..RR_RING_MAX */
/* Dirty-tile read-back for rr_frame_e2e (off by default). on = 1: instead of copying 4*W*H bytes every frame, the library stores
 * into the host buffer only the 32x4-pixel tiles that can differ from what that buffer already holds: tiles with a shaded pixel
 * in this frame or in the frame it last wrote into the SAME host buffer with the same clear colour (everything else there is
 * already the clear colour). The host buffer is bit-identical to the device frame after every call, as with the full copy.
 * Contract: the host buffers are page-locked (rr_host_alloc / rr_host_register), the caller cycles through them in ring order and
 * does not write into them; a buffer the library has not seen in that ring slot, a changed clear colour, a post pass since the
 * last frame, rr_set_readback_tiles itself, or a previous copy that needed more than half of the tiles make the next copy a full
 * one (through the copy engine, as without this mode). Applies to whole frames of one context and to the distributed read-back of
 * connected contexts (rr_mgpu_set_readback(1), interleaved row tiles that are multiples of 4 rows: every context stores the tiles
 * of its own rows into the shared host frame); any other split keeps copying its rows. The reference has no counterpart: its frame stays in a GL texture
 * (cl_gl_interop_texture.hpp); this is the headless replacement's way around the PCIe link.
 * rr_readback_tile_bytes: bytes rr_frame_e2e has moved into host buffers in this mode since the last call — tiles stored plus the
 * whole frames that went through the copy engine (statistics). */
int rr_set_readback_tiles(rr_ctx*, int on);
int rr_readback_tile_bytes(rr_ctx*, uint64_t* bytes);

/* ---- roofline micro-benchmarks (R_atomic is not in MEASURED_PEAKS.json) ------------------------- */
int rr_microbench_atomic_min(rr_ctx*, size_t footprint_bytes, uint64_t n_ops, float* ms_out);
int rr_microbench_copy(rr_ctx*, size_t nbytes, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* RR_H_INCLUDED */

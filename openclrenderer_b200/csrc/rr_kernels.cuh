// rr_kernels.cuh — sm_100a kernels of the raster path (setup/bin, depth, id resolve, shade, shadow passes, atlas).
// Design notes live in DESIGN.md; every kernel names the cl2.cl kernel whose results it must reproduce.
#pragma once
#include "rr_math.cuh"
#include "../../include/rr.h"

namespace rr {

// ---- counters (one uint32 array per context) ----------------------------------------------------------------------
// The counters that many warps update with atomics while a kernel runs each sit on a 128-byte line of their own: atomics on one
// line retire one after the other in its L2 slice (about one per nanosecond), so hot counters that share a line share that rate.
enum {
    CTR_NCUT = 0,      // id_cutdown_tris of the main pass
    CTR_NFRAG = 1,     // id_buffer_atomc of the main pass
    CTR_OVERFLOW = 2,  // bit0 fragments, bit1 cutdown, bit2 look-back watchdog
    CTR_S_TOTAL = 6,   // running total of shadow fragments in this frame (statistics)
    CTR_NEXTRA = 7,    // fragments whose samples did not fit the sample list: kernel2 walks them (tail of k_ids_list)
    CTR_NDESC = 12,    // descriptors in the sample list
    CTR_NACTIVE = 13,  // k_frame_prologue: setup blocks with at least one cluster that is not culled
    CTR_CUT_SKIPPED = 14,  // ... and the projected-triangle slots of the blocks that were skipped
    CTR_PROLOGUE_TICKET = 15,  // last-CTA detection of k_frame_prologue (self-resetting)
    CTR_STICKY = 16,   // overflow flags of earlier frames / passes: the per-frame zeroing ORs CTR_OVERFLOW in here before clearing it,
                       // so a pipelined caller still sees an overflow at its next rr_sync (which reads and clears both words)
    // ---- one line each
    CTR_TICKET = 32,   // block ticket of the single-pass scan
    CTR_NWORK = 64,    // fragments k_setup_main did not rasterise inline: the work list of k_raster_warp_depth
    CTR_NSAMPLES = 96, // entries reserved in the sample list
    CTR_NSHADE = 128,  // covered pixels in the shading list
    CTR_S_NFRAG = 160, // shadow pass allocation counter, one 64-bit word at [160..161]: (projected triangles << 32) | fragments,
    CTR_S_NCUT = 161,  //   i.e. word 160 = fragment count, word 161 = projected-triangle count (little endian)
    CTR_COUNT = 192
};

// per-row ownership bits of the sort-first split (rr_config.band_tile): built on the host at rr_create
enum { ROW_NEEDED = 1, ROW_OWNED = 2 };

struct CamParams {
    float3 pos;
    RotSC rot;
};

struct FaceTable { RotSC r[6]; };

// =====================================================================================================================
// scene repack: AoS triangle (144 B) -> position SoA (40 B): pa = (v0.xyz, v1.x)  pb = (v1.yz, v2.xy)  pc = (v2.z, object id)
// Runs once per rr_scene_write_tris (the reference's fill_ids kernel, cl2.cl:4231, ran at the same point).
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_repack(const rr_triangle* __restrict__ tris, uint32_t first, uint32_t count,
                                                float4* __restrict__ pa, float4* __restrict__ pb, float2* __restrict__ pc,
                                                uint32_t* __restrict__ obj_r2_bits, uint32_t n_objs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const rr_triangle* t = tris + first + i;
    const float4 v0 = *reinterpret_cast<const float4*>(t->vertices[0].pos);
    const float4 v1 = *reinterpret_cast<const float4*>(t->vertices[1].pos);
    const float4 v2 = *reinterpret_cast<const float4*>(t->vertices[2].pos);
    uint32_t oid = t->vertices[0].object_id;
    pa[first + i] = make_float4(v0.x, v0.y, v0.z, v1.x);
    pb[first + i] = make_float4(v1.y, v1.z, v2.x, v2.y);
    pc[first + i] = make_float2(v2.z, __uint_as_float(oid));
    // bounding-sphere radius^2 of the object in object space (non-negative floats order like their bit patterns)
    const float r2 = fmaxf(fmaxf(v0.x * v0.x + v0.y * v0.y + v0.z * v0.z, v1.x * v1.x + v1.y * v1.y + v1.z * v1.z), v2.x * v2.x + v2.y * v2.y + v2.z * v2.z);
    if (oid < n_objs) atomicMax(obj_r2_bits + oid, __float_as_uint(r2));
}

// Per-object data the setup kernels need, 48 B instead of the 144 B descriptor: rebuilt when descriptors change.
struct ObjLite {
    float4 pos_scale;     // world_pos.xyz, scale
    float4 nquat;         // fast_normalize(world_rot_quat)  (rot_quat() normalises on every call, cl2.cl:352)
    float4 bquat;         // the normalised conjugate back_rot_quat() ends up rotating with (cl2.cl:359-370)
    int32_t feature_flag;
    int32_t _pad[3];
};

// Per-light constants of kernel3's light loop: the light colour after gamma_transform_approx (cl2.cl:6150-6153), which the
// reference recomputes per pixel per light.
struct LightLite { float4 col_linear; };

__global__ void k_lightlite(const rr_light* __restrict__ lights, uint32_t n, LightLite* __restrict__ out);

__global__ void __launch_bounds__(128) k_objlite(const rr_obj_desc* __restrict__ objs, uint32_t n, ObjLite* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rr_obj_desc& G = objs[i];
    ObjLite o;
    o.pos_scale = make_float4(G.world_pos[0], G.world_pos[1], G.world_pos[2], G.scale);
    const float4 q = make_float4(G.world_rot_quat[0], G.world_rot_quat[1], G.world_rot_quat[2], G.world_rot_quat[3]);
    o.nquat = normalize4(q);
    o.bquat = back_quat(q);
    o.feature_flag = G.feature_flag;
    o._pad[0] = o._pad[1] = o._pad[2] = 0;
    out[i] = o;
}

// Sort-first object culling (multi-GPU band mode only): conservative screen-row range of every object's bounding sphere
// for this frame's camera. A triangle of an object whose rows cannot touch [row_lo, row_hi) is not set up at all on this
// context — it cannot produce a fragment there. rows = (first, last) inclusive; (INT_MIN, INT_MAX) when in doubt.
__global__ void __launch_bounds__(128) k_obj_rows(const ObjLite* __restrict__ objs, const uint32_t* __restrict__ obj_r2_bits, uint32_t n,
                                                  CamParams cam, float height, float fov, float icut, int2* __restrict__ rows,
                                                  const int* __restrict__ rowpfx, int H) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ObjLite G = objs[i];
    const float R = sqrtf(__uint_as_float(obj_r2_bits[i])) * fabsf(G.pos_scale.w) * 1.001f + 1e-3f;
    const float3 c = rot(make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z), cam.pos, cam.rot);
    int2 r = make_int2(INT_MIN, INT_MAX);
    const float zn = c.z - R, zf = c.z + R;
    if (zn > icut + 1.f && isfinite(R) && isfinite(c.y) && isfinite(c.z)) {      // entirely in front of the near plane: no clipping involved
        const float a = (c.y - R) * fov, b = (c.y + R) * fov;
        const float lo = fminf(a / zn, a / zf) + height * 0.5f, hi = fmaxf(b / zn, b / zf) + height * 0.5f;
        r = make_int2((int)fmaxf(floorf(lo) - 4.f, -1e9f), (int)fminf(ceilf(hi) + 4.f, 1e9f));
        // interleaved bands: rowpfx[y] = number of rows < y this context rasterises; an object with none in its range is
        // reported as the empty range
        if (rowpfx) {
            const int a = max(r.x, 0), b = min(r.y, H - 1);
            if (a > b || rowpfx[b + 1] - rowpfx[a] == 0) r = make_int2(INT_MAX, INT_MIN);
        }
    }
    rows[i] = r;
}

// ---- clusters: 128 consecutive triangles with an object-space bounding box ---------------------------------------------
// Built once per rr_scene_write_tris. Every frame one thread per cluster transforms the 8 box corners exactly the way the
// setup kernels transform vertices and decides, conservatively (pixel / unit margins far above the rounding differences
// between corner and vertex arithmetic), that NONE of the cluster's triangles can produce a fragment this context needs:
//   main view : the box lies entirely in front of the near plane and inside depth_far (so every triangle is unclipped:
//               exactly one projected-triangle slot each, cl2.cl:4342) and its projection is off screen (the reference
//               rejects each of its triangles, cl2.cl:4356-4359) or touches no row this context rasterises;
//   shadows   : no point of the box can be assigned by ret_cubeface to a cube face rendered here.
// Culled clusters cost a block one load and one flag: this is what makes the replicated setup stage scale in the sort-
// first split, and it removes off-screen geometry on a single GPU. Clusters that span two objects are never culled.
#define CLUSTER_TRIS 128
struct ClusterBox { float4 lo; float4 hi; };   // lo.xyz / hi.xyz object space; lo.w = object id bits; hi.w = 1.f when usable

__global__ void __launch_bounds__(CLUSTER_TRIS) k_cluster_bounds(const float4* __restrict__ pa, const float4* __restrict__ pb, const float2* __restrict__ pc,
                                                                 uint32_t n_tris, uint32_t first_cluster, ClusterBox* __restrict__ out) {
    __shared__ float s_red[6][CLUSTER_TRIS / 32];
    __shared__ int s_ok[CLUSTER_TRIS / 32];
    const uint32_t cl = first_cluster + blockIdx.x;
    const uint32_t tri = cl * CLUSTER_TRIS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float INF = __int_as_float(0x7f800000);
    float mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
    const uint32_t oid0 = __float_as_uint(__ldg(pc + (size_t)cl * CLUSTER_TRIS).y);
    bool ok = true;
    if (tri < n_tris) {
        const float4 a = __ldg(pa + tri), b = __ldg(pb + tri);
        const float2 c = __ldg(pc + tri);
        const float vx[3] = {a.x, a.w, b.z}, vy[3] = {a.y, b.x, b.w}, vz[3] = {a.z, b.y, c.x};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            mn[0] = fminf(mn[0], vx[k]); mx[0] = fmaxf(mx[0], vx[k]);
            mn[1] = fminf(mn[1], vy[k]); mx[1] = fmaxf(mx[1], vy[k]);
            mn[2] = fminf(mn[2], vz[k]); mx[2] = fmaxf(mx[2], vz[k]);
            ok = ok && isfinite(vx[k]) && isfinite(vy[k]) && isfinite(vz[k]);
        }
        ok = ok && __float_as_uint(c.y) == oid0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
        }
    const bool wok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        for (int k = 0; k < 3; k++) { s_red[k][warp] = mn[k]; s_red[3 + k][warp] = mx[k]; }
        s_ok[warp] = wok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool all = true;
        for (int w = 0; w < CLUSTER_TRIS / 32; w++) {
            for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], s_red[k][w]); mx[k] = fmaxf(mx[k], s_red[3 + k][w]); }
            all = all && s_ok[w];
        }
        ClusterBox bx;
        bx.lo = make_float4(mn[0], mn[1], mn[2], __uint_as_float(oid0));
        bx.hi = make_float4(mx[0], mx[1], mx[2], all ? 1.f : 0.f);
        out[cl] = bx;
    }
}

// corner k (bit 0: x, bit 1: y, bit 2: z) of a cluster box in world space, computed like a vertex (cl2.cl:505-507)
__device__ __forceinline__ float3 cluster_corner_world(const ClusterBox& b, int k, const ObjLite& G) {
    const float3 v = make_float3((k & 1) ? b.hi.x : b.lo.x, (k & 2) ? b.hi.y : b.lo.y, (k & 4) ? b.hi.z : b.lo.z);
    return rot_quat_n(v * G.pos_scale.w, G.nquat) + make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z);
}

// main view: 0 when the cluster can be skipped by k_setup_main (its triangles still take their slots)
struct ClusterVisParams {
    const ClusterBox* boxes; uint32_t n_clusters; const ObjLite* objs; uint32_t n_objs;
    CamParams cam; float width, height, fov, icut;
    const int* rowpfx; int row_lo, row_hi;
};
__device__ __forceinline__ uint8_t cluster_visible(const ClusterVisParams& P, uint32_t i) {
    const ClusterBox* boxes = P.boxes; const ObjLite* objs = P.objs; const uint32_t n_objs = P.n_objs;
    const CamParams& cam = P.cam; const float width = P.width, height = P.height, fov = P.fov, icut = P.icut;
    const int* rowpfx = P.rowpfx; const int row_lo = P.row_lo, row_hi = P.row_hi;
    const ClusterBox b = boxes[i];
    const uint32_t oid = __float_as_uint(b.lo.w);
    uint8_t v = 1;
    if (b.hi.w == 1.f && oid < n_objs) {
        const ObjLite G = objs[oid];
        const float3 gpos = make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z);
        const float INF = __int_as_float(0x7f800000);
        float zmin = INF, zmax = -INF, xmin = INF, xmax = -INF, ymin = INF, ymax = -INF, amax = 0.f;
        bool fin = isfinite(length3(gpos - cam.pos)) && !(length3(gpos - cam.pos) > RR_DEPTH_FAR * 0.999f);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float3 q = rot(cluster_corner_world(b, k, G), cam.pos, cam.rot);
            zmin = fminf(zmin, q.z); zmax = fmaxf(zmax, q.z);
            amax = fmaxf(amax, fmaxf(fabsf(q.x), fmaxf(fabsf(q.y), fabsf(q.z))));
            const float kx = fov / fmaxf(q.z, 1e-3f);
            const float px = fmaf(q.x, kx, width * 0.5f), py = fmaf(q.y, kx, height * 0.5f);
            xmin = fminf(xmin, px); xmax = fmaxf(xmax, px); ymin = fminf(ymin, py); ymax = fmaxf(ymax, py);
            fin = fin && isfinite(q.x) && isfinite(q.y) && isfinite(q.z) && isfinite(px) && isfinite(py);
        }
        // camera-space slack for the arithmetic differences between corners and vertices, then its worst-case effect in pixels
        // (both go through the same rot_quat_n / rot code: the difference is a few ulps of the largest coordinate; 32 ulps kept)
        const float eps = 4e-6f * amax + 1e-4f;
        if (fin && zmin - eps > icut + 1.f && zmax + eps < RR_DEPTH_FAR * 0.999f) {
            const float slack = 4.f + 4.f * eps * fov / (zmin - eps) * (1.f + amax / (zmin - eps));
            if (isfinite(slack)) {
                bool culled = xmax + slack < 0.f || xmin - slack >= width || ymax + slack < 0.f || ymin - slack >= height;
                if (!culled) {
                    // rows the triangles' boxes [round(min) - 1, round(max)] can touch, clamped like calc_min_max
                    const int a = max((int)fmaxf(floorf(ymin - slack) - 2.f, -1e9f), 0), e = min((int)fminf(ceilf(ymax + slack) + 2.f, 1e9f), (int)height - 1);
                    if (a > e || e < row_lo || a >= row_hi) culled = true;
                    else if (rowpfx && rowpfx[e + 1] - rowpfx[a] == 0) culled = true;
                }
                v = culled ? 0 : 1;
            }
        }
    }
    return v;
}

// One clipped + projected triangle after culling. keep == false -> no storage written, no fragments.
struct SubTri { float3 p0, p1, p2; float rconst; int n_frag; int box; bool keep; };

// cull + bbox + fragment count, cl2.cl:4352-4377 (main) / 4571-4597 (shadow)
// Sort-first split (rows != nullptr): a triangle whose box touches no row this context rasterises is dropped like a culled one.
struct RowCull { const int* pfx; int lo, hi; bool on; };
__device__ __forceinline__ void classify(SubTri& s, bool two_sided, float ewidth, float eheight, float op_size, const RowCull rc = RowCull{nullptr, 0, 0, false}) {
    bool valid = two_sided || front_facing(s.p0, s.p1, s.p2);
    bool cond = (s.p0.x < 0 && s.p1.x < 0 && s.p2.x < 0) || (s.p0.x >= ewidth && s.p1.x >= ewidth && s.p2.x >= ewidth) ||
                (s.p0.y < 0 && s.p1.y < 0 && s.p2.y < 0) || (s.p0.y >= eheight && s.p1.y >= eheight && s.p2.y >= eheight);
    s.keep = valid && !cond;
    s.n_frag = 0;
    s.box = 0;
    s.rconst = 0.f;
    if (!s.keep) return;
    float3 xr = make_float3(roundf(s.p0.x), roundf(s.p1.x), roundf(s.p2.x));
    float3 yr = make_float3(roundf(s.p0.y), roundf(s.p1.y), roundf(s.p2.y));
    s.rconst = calc_rconstant_v(xr, yr);
    float4 mm = calc_min_max(xr, yr, ewidth, eheight);
    if (rc.on) {
        const int a = (int)mm.z, e = (int)mm.w - 1;               // the walk visits rows [mm.z, mm.w)
        if (e < a || e < rc.lo || a >= rc.hi || (rc.pfx && rc.pfx[e + 1] - rc.pfx[a] == 0)) { s.keep = false; return; }
    }
    float area = (mm.y - mm.x) * (mm.w - mm.z);
    s.n_frag = (int)ceilf(area / op_size);
    s.box = (int)area;                        // width * rows, exact (both are small integers)
}

// where the walk of a stored triangle ends: first k whose row counter reaches max_y (only needed for its slot counts)
__device__ __forceinline__ int subtri_walk_end(const SubTri& s, float ewidth, float eheight) {
    float3 xr = make_float3(roundf(s.p0.x), roundf(s.p1.x), roundf(s.p2.x));
    float3 yr = make_float3(roundf(s.p0.y), roundf(s.p1.y), roundf(s.p2.y));
    float4 mm = calc_min_max(xr, yr, ewidth, eheight);
    const int width = (int)(mm.y - mm.x), rows = (int)(mm.w - mm.z);
    return walk_end(width, rows, 1.f / (float)width, mm.z, mm.w);
}

// pixel slots chunk `a` of a triangle visits: k = a*op .. a*op+op, cut where the walk ends
__device__ __forceinline__ uint32_t chunk_slots(int kend, int a, int op) { return (uint32_t)min(max(kend - a * op, 0), op + 1); }

// camera-space triangle -> near-plane clip -> projection (cl2.cl:700-729 after the rotations). Returns num (0/1/2).
__device__ __forceinline__ int clip_project(float3 q0, float3 q1, float3 q2, float icut, float half_w, float half_h, float fovc, SubTri& s0, SubTri& s1) {
    float3 a0, a1, a2, b0, b1, b2;
    const int num = clip_near(q0, q1, q2, icut, a0, a1, a2, b0, b1, b2);
    if (num > 0) { s0.p0 = project(a0, half_w, half_h, fovc); s0.p1 = project(a1, half_w, half_h, fovc); s0.p2 = project(a2, half_w, half_h, fovc); }
    if (num > 1) { s1.p0 = project(b0, half_w, half_h, fovc); s1.p1 = project(b1, half_w, half_h, fovc); s1.p2 = project(b2, half_w, half_h, fovc); }
    return num;
}

__device__ __forceinline__ void subtri_clear(SubTri& s) { s.keep = false; s.n_frag = 0; s.box = 0; s.rconst = 0.f; }

// ---- inline rasterisation of a small single-chunk triangle from the setup kernels --------------------------------------
// A triangle whose whole walk is one chunk of at most RASTER_SMALL_MAX slots is rasterised by the warp that set it up
// (kernel1's work for it): no trip through the record and projected-triangle buffers and no second kernel.
#define RASTER_SMALL_MAX 48
#define FRAGCNT_DEPTH_DONE 0x80000000u      // flag in the per-fragment slot count: depth already written inline
#define FRAGCNT_MASK 0x3FFFFFFFu

__device__ __forceinline__ bool inline_candidate(const SubTri& s) { return s.keep && s.n_frag == 1 && s.box <= RASTER_SMALL_MAX; }

// Exactness test shared by the inline rasterisers: rounded vertex coordinates (integers) for which every product and partial
// sum of point_in_tri (cl2.cl:4798-4807) and of calc_rconstant_v (cl2.cl:408-411) is an integer below 2^24, hence exact in
// fp32 whatever the evaluation order: max|x| * max|y| + 64 (max|x| + max|y|) < 2^24 with extents of at most 64.
// (Always true at 4K and in a 2048^2 cube face; at 8K only towards the top-left of the screen.)
__device__ __forceinline__ bool tri_exact_arith(float3 xr, float3 yr) {
    const float mx = fmaxf(fmaxf(fabsf(xr.x), fabsf(xr.y)), fabsf(xr.z)), my = fmaxf(fmaxf(fabsf(yr.x), fabsf(yr.y)), fabsf(yr.z));
    const float ext = fmaxf(fmaxf(fmaxf(xr.x, xr.y), xr.z) - fminf(fminf(xr.x, xr.y), xr.z), fmaxf(fmaxf(yr.x, yr.y), yr.z) - fminf(fminf(yr.x, yr.y), yr.z));
    const float chk = ((xr.x + xr.y) + (xr.z + yr.x)) + (yr.y + yr.z);           // NaN / inf - inf guard (fmaxf skips NaNs)
    return chk == chk && ext <= 64.f && mx * my + 64.f * (mx + my) < 16000000.f;
}

// Per-warp queue in shared memory that compacts the small triangles of a warp (warp ballot + prefix) so that they are
// rasterised by 32 busy lanes: roughly half of a warp's triangles are culled and the survivors need different numbers of
// steps, so rasterising them in place would leave most lanes idle.
// The covered samples (pixel, depth, fragment index) of every triangle are appended to a sample list so that kernel2's work
// for these triangles is a stream over that list (k_ids_list) instead of a second walk. Coverage is determined first (a bit
// mask per triangle), so exactly popc(mask) entries are reserved, with one warp-aggregated atomicAdd per drain; a triangle
// whose samples do not fit goes to the `extra` list and is walked again by the tail of k_ids_list.
#define IQ_SLOTS 64
#define IQ_FIELDS 12            // rounded x of the 3 vertices, rounded y, camera z, rconst, fragment index, flags
#define IQ_EXACT 1u
struct InlineQueue { float f[IQ_FIELDS][IQ_SLOTS]; };

struct SampleList { uint2* samples; uint32_t* frag; uint32_t cap; uint32_t* count;      // (pixel, depth), fragment index
                    uint32_t* extra; uint32_t* extra_count; uint32_t extra_cap; uint32_t* overflow; };   // fragments whose samples did not fit

// a fragment kernel2 has to walk again; the list itself is bounded (a fragment can be listed once per failed reservation)
__device__ __forceinline__ void extra_push(const SampleList& sl, uint32_t f) {
    const uint32_t at = atomicAdd(sl.extra_count, 1u);
    if (at < sl.extra_cap) sl.extra[at] = f; else atomicOr(sl.overflow, 1u);
}

struct RowFilter { int lo, hi; const uint8_t* mask; };     // rows rasterised here: [lo, hi), and bit 0 of mask[y] when mask != nullptr
__device__ __forceinline__ bool row_wanted(const RowFilter& rf, int y) { return y >= rf.lo && y < rf.hi && (!rf.mask || (rf.mask[y] & 1)); }

// The literal path of the inline rasteriser (triangles that fail tri_exact_arith, have a lagging row counter or a tight box of
// more than 32 pixels): the reference's walk (scan_chunk) with point_in_tri, in two passes over the same state machine —
// first the coverage as a bit per walk step (a chunk of at most 48 slots takes at most 49 steps), then the emission.
__device__ __noinline__ unsigned long long inline_literal_mask(float3 xr, float3 yr, float width, float height, int op, RowFilter rf) {
    const float4 mm = calc_min_max(xr, yr, width, height);
    unsigned long long mask = 0ull, bit = 1ull;
    scan_chunk(mm, op, 0u, [&](float x, float y) {
        if (row_wanted(rf, (int)y) && point_in_tri(x, y, xr.x, yr.x, xr.y, yr.y, xr.z, yr.z)) mask |= bit;
        bit <<= 1;
    });
    return mask;
}
__device__ __noinline__ void inline_literal_emit(float3 xr, float3 yr, float A, float B, float C, float width, float height, int op, unsigned long long mask,
                                                 uint32_t* __restrict__ target, uint2* __restrict__ samples, uint32_t* __restrict__ sfrag, uint32_t fidx) {
    const float4 mm = calc_min_max(xr, yr, width, height);
    uint32_t n = 0;
    scan_chunk(mm, op, 0u, [&](float x, float y) {
        if (mask & 1ull) {
            const float fd = fmaf(A, x, fmaf(B, y, C));
            const uint32_t d = sat_u32(RR_U32MAXF / fd);
            const uint32_t px = (uint32_t)((int)(y * width) + (int)x);
            red_min_u32(target + px, d);
            if (samples) { samples[n] = make_uint2(px, d); sfrag[n] = fidx; n++; }
        }
        mask >>= 1;
    });
}

template <bool MASKED>
struct InlineRaster {
    InlineQueue* q; int count;  // count is warp-uniform
    int op; float width, height; uint32_t* target; RowFilter rf;
    SampleList sl;

    // executed by all 32 lanes; lanes with !valid only take part in the warp collectives.
    // Fast path: see shadow_raster_small (k_shadow_setup's stage C) — exact integer edge functions stepped over the vertices' own
    // bounding box; here the covered samples are also recorded.
    __device__ __forceinline__ void run(int slot, bool valid) const {
        const int lane = threadIdx.x & 31;
        const float* f = &q->f[0][valid ? slot : 0];
        const float3 xr = make_float3(f[0], f[IQ_SLOTS], f[2 * IQ_SLOTS]), yr = make_float3(f[3 * IQ_SLOTS], f[4 * IQ_SLOTS], f[5 * IQ_SLOTS]);
        const uint32_t fidx = __float_as_uint(f[10 * IQ_SLOTS]), flags = __float_as_uint(f[11 * IQ_SLOTS]);
        const float4 mm = calc_min_max(xr, yr, width, height);
        const float x0 = fmaxf(mm.x, fminf(fminf(xr.x, xr.y), xr.z)), y0 = fmaxf(mm.z, fminf(fminf(yr.x, yr.y), yr.z));
        const int wt = (int)(mm.y - x0), ht = (int)(mm.w - y0), rows = (int)(mm.w - mm.z);
        bool fast = valid && (flags & IQ_EXACT) && wt * ht <= 32;
        if (fast && (float)(rows - 1) > mm.z) {             // a box at the very top of the screen: look for a lagging row (rr_math.cuh walk_row)
            const int bw = (int)(mm.y - mm.x);
            const float iw = 1.f / (float)bw;
            for (int r = 1; r < rows; r++) fast = fast && walk_row(r * bw, iw, mm.z) == mm.z + (float)r;
        }
        uint32_t mask = 0u;
        unsigned long long lmask = 0ull;
        if (fast) {
            if (wt > 0 && ht > 0) {
                const float Ah = 0.5f * (-yr.y * xr.z + yr.x * (-xr.y + xr.z) + xr.x * (yr.y - yr.z) + xr.y * yr.z);
                const float sign = Ah < 0 ? -1.f : 1.f;
                const float lim = 2.0001f * Ah * sign;
                const float as = (yr.z - yr.x) * sign, bs = (xr.x - xr.z) * sign, at = (yr.x - yr.y) * sign, bt = (xr.y - xr.x) * sign;
                float s_row = (yr.x * xr.z - xr.x * yr.z) * sign + as * x0 + bs * y0;
                float t_row = (xr.x * yr.y - yr.x * xr.y) * sign + at * x0 + bt * y0;
                uint32_t bit = 1u;
                for (int r = 0; r < ht; r++) {
                    float s = s_row, t = t_row;
                    const bool row_ok = !MASKED || row_wanted(rf, (int)y0 + r);
                    for (int c = 0; c < wt; c++) {
                        if (row_ok && s > -0.0001f && t > -0.0001f && (s + t) < lim) mask |= bit;
                        bit <<= 1;
                        s += as; t += at;
                    }
                    s_row += bs; t_row += bt;
                }
            }
        } else if (valid) {
            lmask = inline_literal_mask(xr, yr, width, height, op, MASKED ? rf : RowFilter{0, 0x7FFFFFFF, nullptr});
        }
        const int n = fast ? __popc(mask) : __popcll(lmask);
        // reserve n sample entries per lane with one atomic per warp
        int inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        if (total == 0) return;
        uint32_t b0 = 0;
        if (lane == 31) b0 = atomicAdd(sl.count, (uint32_t)total);
        b0 = __shfl_sync(0xffffffffu, b0, 31);
        const bool rec = (unsigned long long)b0 + (unsigned long long)total <= (unsigned long long)sl.cap;
        if (n == 0) return;
        const uint32_t first = b0 + (uint32_t)(inc - n);
        if (!rec) {                                                               // sample list full: kernel2 walks this fragment again
            extra_push(sl, fidx);
            // the part of the failed reservation that lies inside the list is streamed by k_ids_list: mark it as holding no sample
            for (uint32_t k = 0; k < (uint32_t)n && first + k < sl.cap; k++) sl.samples[first + k] = make_uint2(0xFFFFFFFFu, 0u);
        }
        // plane of 1 / (z / far), cl2.cl:5029-5040
        float3 d = make_float3(f[6 * IQ_SLOTS] / RR_DEPTH_FAR, f[7 * IQ_SLOTS] / RR_DEPTH_FAR, f[8 * IQ_SLOTS] / RR_DEPTH_FAR);     // dcalc
        d = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);                                                                        // native_recip
        float A, B, C;
        interpolate_get_const(d, xr, yr, f[9 * IQ_SLOTS], A, B, C);
        if (fast) {
            const float iwt = __fdividef(1.f, (float)wt);
            uint32_t at = first;
            while (mask) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1u;
                const int r = (int)(((float)i + 0.5f) * iwt);   // i / wt: i < 32, wt <= 32, so (i + 0.5) / wt is at least 1/64 away from an integer
                const float x = x0 + (float)(i - r * wt), y = y0 + (float)r;
                const float fd = fmaf(A, x, fmaf(B, y, C));
                const uint32_t dd = sat_u32(RR_U32MAXF / fd);
                const uint32_t px = (uint32_t)((int)(y * width) + (int)x);
                red_min_u32(target + px, dd);
                if (rec) { sl.samples[at] = make_uint2(px, dd); sl.frag[at] = fidx; at++; }
            }
        } else {
            inline_literal_emit(xr, yr, A, B, C, width, height, op, lmask, target, rec ? sl.samples + first : nullptr, sl.frag + first, fidx);
        }
    }
    // called by all 32 lanes: queue the triangle (rounded here), unless it provably covers nothing
    __device__ __forceinline__ void push(bool has, const SubTri& s, uint32_t fidx) {
        const int lane = threadIdx.x & 31;
        float3 xr, yr;
        uint32_t flags = 0;
        if (has) {
            xr = make_float3(roundf(s.p0.x), roundf(s.p1.x), roundf(s.p2.x));
            yr = make_float3(roundf(s.p0.y), roundf(s.p1.y), roundf(s.p2.y));
            if (tri_exact_arith(xr, yr)) {
                flags = IQ_EXACT;
                if (isinf(s.rconst)) has = false;       // collinear after rounding (determinant exactly 0): point_in_tri's s + t < 0 can never hold
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (!m) return;
        if (has) {
            float* f = &q->f[0][count + __popc(m & ((1u << lane) - 1u))];
            f[0] = xr.x; f[IQ_SLOTS] = xr.y; f[2 * IQ_SLOTS] = xr.z;
            f[3 * IQ_SLOTS] = yr.x; f[4 * IQ_SLOTS] = yr.y; f[5 * IQ_SLOTS] = yr.z;
            f[6 * IQ_SLOTS] = s.p0.z; f[7 * IQ_SLOTS] = s.p1.z; f[8 * IQ_SLOTS] = s.p2.z;
            f[9 * IQ_SLOTS] = s.rconst; f[10 * IQ_SLOTS] = __uint_as_float(fidx); f[11 * IQ_SLOTS] = __uint_as_float(flags);
        }
        count += __popc(m);
        __syncwarp();
    }
    // called by all 32 lanes: rasterise full warps' worth; with `all`, whatever is left as well
    __device__ __forceinline__ void drain(bool all) {
        const int lane = threadIdx.x & 31;
        while (count >= 32 || (all && count > 0)) {
            const int n = min(count, 32);
            count -= n;
            run(count + lane, lane < n);
            __syncwarp();
        }
    }
};

// =====================================================================================================================
// k_setup_main == prearrange (cl2.cl:4272-4409), single pass.
//
// The reference allocates projected-triangle slots and fragment slots with two global atomic_add per triangle; which
// triangle gets which slot depends on scheduling, and the id buffer stores those slot numbers. Here both allocations
// are exclusive prefix sums in triangle order (= the reference run one work-item at a time), computed in the same pass
// with a decoupled look-back scan: block scan in shared memory, one 64-bit descriptor per block
// (flag:2 | cut:27 | frag:35), tickets so a block only ever waits on blocks that are already resident.
// Records are then written cooperatively by the whole block, one 32-bit word per thread per step, so the stores of a
// warp are consecutive addresses regardless of how many fragments each triangle produced.
// =====================================================================================================================
#define SETUP_THREADS 256

__device__ __forceinline__ unsigned long long lb_pack(uint32_t flag, uint32_t c, uint32_t f) {
    return ((unsigned long long)flag << 62) | ((unsigned long long)c << 35) | (unsigned long long)f;
}

// Decoupled look-back of k_setup_main's single-pass scan, executed by one full warp of block `bid`: publishes the block's
// aggregate, sums its predecessors' and publishes the inclusive prefix. The last block also writes the totals.
__device__ __forceinline__ void setup_lookback(unsigned long long* lookback, uint32_t* counters, uint32_t bid, bool is_last, uint32_t cut_extra,
                                               uint32_t tot_c, uint32_t tot_f, uint32_t& base_c, uint32_t& base_f) {
    const int lane = threadIdx.x & 31;
    volatile unsigned long long* desc = lookback;
    base_c = 0; base_f = 0;
    if (bid == 0) {
        if (lane == 0) { desc[0] = lb_pack(2, tot_c, tot_f); }
    } else {
        if (lane == 0) { desc[bid] = lb_pack(1, tot_c, tot_f); }
        __threadfence();
        int look = (int)bid - 1;
        uint32_t watchdog = 0;
        while (true) {
            int idx = look - lane;
            unsigned long long v = lb_pack(2, 0, 0);
            if (idx >= 0) {
                do {
                    v = desc[idx];
                    if (++watchdog > (1u << 26)) { atomicOr(&counters[CTR_OVERFLOW], 4u); v = lb_pack(2, 0, 0); break; }
                } while ((v >> 62) == 0);
            }
            const uint32_t flag = (uint32_t)(v >> 62);
            const unsigned incl_mask = __ballot_sync(0xffffffffu, flag == 2);
            const int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
            uint32_t vc = (lane <= first_incl) ? (uint32_t)((v >> 35) & 0x7FFFFFFull) : 0u;
            uint32_t vf = (lane <= first_incl) ? (uint32_t)(v & 0x7FFFFFFFFull) : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { vc += __shfl_xor_sync(0xffffffffu, vc, d); vf += __shfl_xor_sync(0xffffffffu, vf, d); }
            base_c += vc; base_f += vf;
            if (incl_mask) break;
            look -= 32;
        }
        if (lane == 0) { desc[bid] = lb_pack(2, base_c + tot_c, base_f + tot_f); }
    }
    if (lane == 0 && is_last) { counters[CTR_NCUT] = base_c + tot_c + cut_extra; counters[CTR_NFRAG] = base_f + tot_f; }
}

// k_frame_prologue: everything k_setup_main needs before it starts, in ONE launch (it replaces three memsets and two
// kernels): zero the scan descriptors and the frame counters, classify the clusters (cluster_visible), and — in the CTA
// that finishes last — compact the setup blocks: a block of SETUP_THREADS triangles whose clusters are all culled
// contributes exactly its triangle count to the projected-triangle numbering and nothing else, so k_setup_main never
// runs it. active[i] = i-th surviving block, skipped_before[i] = slots of the culled blocks in front of it.
#define PROLOGUE_THREADS 256
struct PrologueParams {
    ClusterVisParams cv;
    uint8_t* vis;
    uint32_t n_tris, n_blocks;
    uint32_t* active; uint32_t* skipped_before;
    uint32_t* counters; unsigned long long* lookback;
    int cull;                                // 0: zeroing only
};

__global__ void __launch_bounds__(PROLOGUE_THREADS) k_frame_prologue(const PrologueParams P) {
    __shared__ uint32_t s_a[PROLOGUE_THREADS / 32], s_s[PROLOGUE_THREADS / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t i = blockIdx.x * blockDim.x + tid;
    for (uint32_t j = i; j < P.n_blocks; j += gridDim.x * blockDim.x) P.lookback[j] = 0ull;
    if (!P.cull) {                           // whole frame, no culling: only the zeroing (cudaMemsetAsync would queue behind a
        if (i == 0) {                        // read-back DMA on the copy engine and stall the frame)
            P.counters[CTR_STICKY] |= P.counters[CTR_OVERFLOW];
            P.counters[CTR_NCUT] = 0u; P.counters[CTR_NFRAG] = 0u; P.counters[CTR_OVERFLOW] = 0u; P.counters[CTR_TICKET] = 0u;
            P.counters[CTR_NSAMPLES] = 0u; P.counters[CTR_NDESC] = 0u; P.counters[CTR_NSHADE] = 0u;
            P.counters[CTR_NEXTRA] = 0u; P.counters[CTR_NWORK] = 0u;
        }
        return;
    }
    if (i < P.cv.n_clusters) P.vis[i] = cluster_visible(P.cv, i);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&P.counters[CTR_PROLOGUE_TICKET], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA: ordered compaction over all setup blocks, 8 * PROLOGUE_THREADS blocks per round: a thread takes eight
    // consecutive blocks (their sixteen cluster flags are one 16-byte load: a single round trip per round), the CTA scans the
    // per-thread (surviving blocks, skipped slots) pairs, every thread places its survivors.
    const uint32_t n_clusters = P.cv.n_clusters, n_tris = P.n_tris, n_blocks = P.n_blocks;
    const unsigned short* vis16 = reinterpret_cast<const unsigned short*>(P.vis);      // written by other CTAs of this launch: ld.cg
    constexpr uint32_t NW = PROLOGUE_THREADS / 32, PER_T = 8;
    const uint32_t last_count = n_tris - (n_blocks - 1) * 2u * CLUSTER_TRIS;           // only the last block can be short
    uint32_t ta = 0, ts = 0;                                                           // running totals (CTA-uniform)
    for (uint32_t chunk = 0; chunk < n_blocks; chunk += PROLOGUE_THREADS * PER_T) {
        const uint32_t b0 = chunk + (uint32_t)tid * PER_T;
        uint32_t v[PER_T];
        if (b0 + PER_T <= n_blocks) {
            const uint4 q = __ldcg(reinterpret_cast<const uint4*>(vis16 + b0));       // b0 is a multiple of 8: 16-byte aligned
            v[0] = q.x & 0xFFFFu; v[1] = q.x >> 16; v[2] = q.y & 0xFFFFu; v[3] = q.y >> 16;
            v[4] = q.z & 0xFFFFu; v[5] = q.z >> 16; v[6] = q.w & 0xFFFFu; v[7] = q.w >> 16;
        } else {
#pragma unroll
            for (uint32_t k = 0; k < PER_T; k++) v[k] = b0 + k < n_blocks ? (uint32_t)__ldcg(vis16 + b0 + k) : 0xFFFFu;
        }
        uint32_t cmask = 0, na = 0, ns = 0;                                            // bit k: block b0 + k is culled
#pragma unroll
        for (uint32_t k = 0; k < PER_T; k++) {
            const uint32_t b = b0 + k;
            if (b >= n_blocks) break;
            const bool cu = (v[k] & 0xFFu) == 0u && (2 * b + 1 >= n_clusters || (v[k] >> 8) == 0u);
            if (cu) { cmask |= 1u << k; ns += (b == n_blocks - 1) ? last_count : 2u * CLUSTER_TRIS; } else na++;
        }
        uint32_t ia = na, is = ns;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, ia, d), y = __shfl_up_sync(0xffffffffu, is, d); if (lane >= d) { ia += x; is += y; } }
        __syncthreads();                                                               // (s_a / s_s of the previous round are consumed)
        if (lane == 31) { s_a[warp] = ia; s_s[warp] = is; }
        __syncthreads();
        uint32_t at = ta + ia - na, sk = ts + is - ns, ca = 0, cs = 0;
        for (int w = 0; w < (int)NW; w++) { if (w < warp) { at += s_a[w]; sk += s_s[w]; } ca += s_a[w]; cs += s_s[w]; }
#pragma unroll
        for (uint32_t k = 0; k < PER_T; k++) {
            const uint32_t b = b0 + k;
            if (b >= n_blocks) break;
            if (cmask & (1u << k)) sk += (b == n_blocks - 1) ? last_count : 2u * CLUSTER_TRIS;
            else { P.active[at] = b; P.skipped_before[at] = sk; at++; }
        }
        ta += ca; ts += cs;
    }
    if (tid == 0) {
        P.counters[CTR_PROLOGUE_TICKET] = 0u;
        P.counters[CTR_STICKY] |= P.counters[CTR_OVERFLOW];
        P.counters[CTR_NCUT] = ta == 0 ? ts : 0u;                 // nothing survives: k_setup_main's blocks all leave at once
        P.counters[CTR_NFRAG] = 0u; P.counters[CTR_OVERFLOW] = 0u; P.counters[CTR_TICKET] = 0u;
        P.counters[CTR_NSAMPLES] = 0u; P.counters[CTR_NDESC] = 0u; P.counters[CTR_NSHADE] = 0u;
        P.counters[CTR_NEXTRA] = 0u; P.counters[CTR_NWORK] = 0u;
        P.counters[CTR_NACTIVE] = ta;
        P.counters[CTR_CUT_SKIPPED] = ts;
    }
}

struct SetupMainParams {
    const float4* pa; const float4* pb; const float2* pc;
    const ObjLite* objs;
    uint32_t n_tris;
    CamParams cam;
    float width, height, fov, icut;
    uint32_t* frags; uint32_t cap_frags;
    uint32_t* fragcnt;                       // pixel slots per fragment (+ FRAGCNT_DEPTH_DONE)
    uint32_t* worklist;                      // fragments that are not rasterised inline (count: CTR_NWORK)
    float4* cutdown; uint32_t cap_cut;
    uint32_t* counters;
    unsigned long long* lookback;
    uint32_t* depth; int row_lo, row_hi;     // inline depth of small triangles
    SampleList sl;                           // their covered samples, for k_ids_list
    const int2* obj_rows;                    // band mode: per-object row range (k_obj_rows); nullptr = no object culling
    const uint8_t* rowmask;                  // interleaved bands: per-row ownership bits (ROW_NEEDED | ROW_OWNED); nullptr = contiguous
    const uint8_t* cluster_vis;              // k_cluster_vis: 0 = the cluster's triangles take their slots but produce nothing here
    const uint32_t* active; const uint32_t* skipped_before;   // k_frame_prologue (valid when cluster_vis != nullptr)
    const int* rowpfx; int cull_rows;        // sort-first split: per-triangle row culling (prefix count of rasterised rows; nullptr = [row_lo, row_hi))
};

// BANDED: sort-first split (row masks, per-triangle row culling); false compiles those paths out for the whole-frame case
#ifndef RR_LB_SETUP_MAIN
#define RR_LB_SETUP_MAIN 4               // 64 registers (see examples/lb_sweep.sh for the A/B harness of these knobs)
#endif
template <bool BANDED>
__global__ void __launch_bounds__(SETUP_THREADS, RR_LB_SETUP_MAIN) k_setup_main(const SetupMainParams P) {
    __shared__ uint32_t s_bid;
    __shared__ uint32_t s_warp_c[SETUP_THREADS / 32], s_warp_f[SETUP_THREADS / 32];
    __shared__ uint32_t s_base_c, s_base_f, s_tot_f;
    __shared__ InlineQueue s_iq[SETUP_THREADS / 32];

    static_assert(SETUP_THREADS == 2 * CLUSTER_TRIS, "one k_setup_main block == two clusters");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t n_active = gridDim.x, cut_skipped_total = 0;
    if (P.cluster_vis) { n_active = P.counters[CTR_NACTIVE]; cut_skipped_total = P.counters[CTR_CUT_SKIPPED]; }   // independent of the ticket
    if (tid == 0) s_bid = atomicAdd(&P.counters[CTR_TICKET], 1u);
    __syncthreads();
    const uint32_t bid = s_bid;                                   // position in the scan
    uint32_t tblock = bid, cut_skipped = 0;
    if (bid >= n_active) return;                                  // only the blocks k_frame_prologue kept run; the others' slots are added in
    if (n_active != gridDim.x) { tblock = P.active[bid]; cut_skipped = P.skipped_before[bid]; }   // nothing culled: identity, no indirection
    const bool is_last = bid == n_active - 1;
    const uint32_t tri = tblock * SETUP_THREADS + tid;
    const bool cl_culled = P.cluster_vis && tri < P.n_tris && !P.cluster_vis[tri / CLUSTER_TRIS];

    SubTri st0, st1;
    int num = 0;
    uint32_t oid = 0;
    subtri_clear(st0); subtri_clear(st1);
    if (tri < P.n_tris) {
        const float4 a = __ldg(P.pa + tri), b = __ldg(P.pb + tri);
        const float2 c = __ldg(P.pc + tri);
        oid = __float_as_uint(c.y);
        const ObjLite G = P.objs[oid];
        const float3 gpos = make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z);
        bool in_band = true;
        if (BANDED && P.obj_rows) { const int2 rw = __ldg(P.obj_rows + oid); in_band = !(rw.y < P.row_lo || rw.x >= P.row_hi); }
        if (cl_culled) num = 1;                                                  // slot taken (cl2.cl:4342), nothing kept
        else if (in_band && !(length3(gpos - P.cam.pos) > RR_DEPTH_FAR)) {       // cl2.cl:4321
            const float sc = G.pos_scale.w;
            const float3 q0 = rot(rot_quat_n(make_float3(a.x, a.y, a.z) * sc, G.nquat) + gpos, P.cam.pos, P.cam.rot);
            const float3 q1 = rot(rot_quat_n(make_float3(a.w, b.x, b.y) * sc, G.nquat) + gpos, P.cam.pos, P.cam.rot);
            const float3 q2 = rot(rot_quat_n(make_float3(b.z, b.w, c.x) * sc, G.nquat) + gpos, P.cam.pos, P.cam.rot);
            num = clip_project(q0, q1, q2, P.icut, P.width / 2.f, P.height / 2.f, P.fov, st0, st1);
            const bool two_sided = (G.feature_flag & RR_FEATURE_TWO_SIDED) > 0;
            const RowCull rc{P.rowpfx, P.row_lo, P.row_hi, BANDED && P.cull_rows != 0};
            if (num > 0) classify(st0, two_sided, P.width, P.height, (float)RR_OP_SIZE, rc);
            if (num > 1) classify(st1, two_sided, P.width, P.height, (float)RR_OP_SIZE, rc);
        }
    }
    const bool inl0 = num > 0 && inline_candidate(st0), inl1 = num > 1 && inline_candidate(st1);
    // slot counts are only needed for what the raster kernels will walk: find the end of those walks now
    const int kend0 = (st0.n_frag > 0) ? (inl0 ? st0.box : subtri_walk_end(st0, P.width, P.height)) : 0;
    const int kend1 = (st1.n_frag > 0) ? (inl1 ? st1.box : subtri_walk_end(st1, P.width, P.height)) : 0;
    const uint32_t my_c = (uint32_t)num;                                        // slots are taken before culling (cl2.cl:4342)
    const uint32_t my_f0 = (num > 0) ? (uint32_t)st0.n_frag : 0u;
    const uint32_t my_f1 = (num > 1) ? (uint32_t)st1.n_frag : 0u;
    const uint32_t my_f = my_f0 + my_f1;

    // block exclusive scan of (my_c, my_f)
    uint32_t inc_c = my_c, inc_f = my_f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t tc = __shfl_up_sync(0xffffffffu, inc_c, d), tf = __shfl_up_sync(0xffffffffu, inc_f, d);
        if (lane >= d) { inc_c += tc; inc_f += tf; }
    }
    if (lane == 31) { s_warp_c[warp] = inc_c; s_warp_f[warp] = inc_f; }
    __syncthreads();
    uint32_t woff_c = 0, woff_f = 0, tot_c = 0, tot_f = 0;
#pragma unroll
    for (int w = 0; w < SETUP_THREADS / 32; w++) {
        uint32_t wc = s_warp_c[w], wf = s_warp_f[w];
        if (w < warp) { woff_c += wc; woff_f += wf; }
        tot_c += wc; tot_f += wf;
    }
    const uint32_t ex_c = woff_c + inc_c - my_c, ex_f = woff_f + inc_f - my_f;

    // decoupled look-back (warp 0)
    if (warp == 0) {
        uint32_t base_c, base_f;
        setup_lookback(P.lookback, P.counters, bid, is_last, cut_skipped_total, tot_c, tot_f, base_c, base_f);
        if (lane == 0) { s_base_c = base_c + cut_skipped; s_base_f = base_f; s_tot_f = tot_f; }
    }
    __syncthreads();
    const uint32_t base_c = s_base_c, base_f = s_base_f;
    const uint32_t cid0 = base_c + ex_c;

    // projected triangles: (x_px, y_px, z_cam, 0) unrounded, cl2.cl:4384-4386
    bool cut_ok = (base_c + tot_c) <= P.cap_cut;
    if (!cut_ok && tid == 0) atomicOr(&P.counters[CTR_OVERFLOW], 2u);
    if (cut_ok) {
        if (num > 0 && st0.keep) {
            float4* dst = P.cutdown + (size_t)cid0 * 3;
            dst[0] = make_float4(st0.p0.x, st0.p0.y, st0.p0.z, 0.f);
            dst[1] = make_float4(st0.p1.x, st0.p1.y, st0.p1.z, 0.f);
            dst[2] = make_float4(st0.p2.x, st0.p2.y, st0.p2.z, 0.f);
        }
        if (num > 1 && st1.keep) {
            float4* dst = P.cutdown + (size_t)(cid0 + 1) * 3;
            dst[0] = make_float4(st1.p0.x, st1.p0.y, st1.p0.z, 0.f);
            dst[1] = make_float4(st1.p1.x, st1.p1.y, st1.p1.z, 0.f);
            dst[2] = make_float4(st1.p2.x, st1.p2.y, st1.p2.z, 0.f);
        }
    }

    // fragment records {tri id, chunk, c_id, bits(rconst), o_id}, cl2.cl:4394-4406, the per-fragment slot counts, and the work
    // list of the fragments kernel1 / kernel2 still have to walk (everything that is not rasterised inline below).
    const uint32_t totf = s_tot_f;
    const bool frag_ok = (unsigned long long)base_f + totf <= (unsigned long long)P.cap_frags;
    if (!frag_ok && tid == 0) atomicOr(&P.counters[CTR_OVERFLOW], 1u);
#pragma unroll 1
    for (int i = 0; i < 2; i++) {
        const uint32_t n = frag_ok ? (i ? my_f1 : my_f0) : 0u;
        const uint32_t fi = base_f + ex_f + (i ? my_f0 : 0u), cid = cid0 + (uint32_t)i;
        const float rc = i ? st1.rconst : st0.rconst;
        const int kend = i ? kend1 : kend0;
        const bool inl = i ? inl1 : inl0;
        // one fragment (nearly every triangle): the owning lane writes its record; offsets grow with the lane, so a warp's stores
        // fall into one short address range
        const bool walk1 = n == 1 && !inl && chunk_slots(kend, 0, RR_OP_SIZE) > 0;
        const unsigned mw = __ballot_sync(0xffffffffu, walk1);
        uint32_t wbase = 0;
        if (mw) {
            if (lane == 0) wbase = atomicAdd(&P.counters[CTR_NWORK], (uint32_t)__popc(mw));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
        }
        if (n == 1) {
            uint32_t* o = P.frags + (size_t)fi * RR_FRAG_WORDS;
            o[0] = tri; o[1] = 0u; o[2] = cid; o[3] = __float_as_uint(rc); o[4] = oid;
            P.fragcnt[fi] = chunk_slots(kend, 0, RR_OP_SIZE) | (inl ? FRAGCNT_DEPTH_DONE : 0u);
            if (walk1) P.worklist[wbase + __popc(mw & ((1u << lane) - 1u))] = fi;
        }
        // several chunks: the warp writes them together, one 32-bit word per lane per step
        for (unsigned m = __ballot_sync(0xffffffffu, n >= 2); m; m &= m - 1u) {
            const int src = __ffs(m) - 1;
            const uint32_t bn = __shfl_sync(0xffffffffu, n, src), bfi = __shfl_sync(0xffffffffu, fi, src), btri = __shfl_sync(0xffffffffu, tri, src);
            const uint32_t bcid = __shfl_sync(0xffffffffu, cid, src), brc = __shfl_sync(0xffffffffu, __float_as_uint(rc), src), boid = __shfl_sync(0xffffffffu, oid, src);
            const int bkend = __shfl_sync(0xffffffffu, kend, src);
            const uint32_t n_walk = min(bn, (uint32_t)((bkend + RR_OP_SIZE - 1) / RR_OP_SIZE));     // chunks that visit at least one slot
            uint32_t lbase = 0;
            if (lane == 0 && n_walk) lbase = atomicAdd(&P.counters[CTR_NWORK], n_walk);
            lbase = __shfl_sync(0xffffffffu, lbase, 0);
            uint32_t* o = P.frags + (size_t)bfi * RR_FRAG_WORDS;
            for (uint32_t w = lane; w < bn * RR_FRAG_WORDS; w += 32u) {
                const uint32_t r = w / RR_FRAG_WORDS, field = w - r * RR_FRAG_WORDS;
                o[w] = field == 0 ? btri : (field == 1 ? r : (field == 2 ? bcid : (field == 3 ? brc : boid)));
            }
            for (uint32_t r = lane; r < bn; r += 32u) {
                P.fragcnt[bfi + r] = chunk_slots(bkend, (int)r, RR_OP_SIZE);
                if (r < n_walk) P.worklist[lbase + r] = bfi + r;
            }
        }
    }
    // kernel1's work for the small single-chunk triangles, after everything other blocks wait for has been published
    InlineRaster<BANDED> ir;
    ir.q = &s_iq[warp]; ir.count = 0; ir.op = RR_OP_SIZE; ir.width = P.width; ir.height = P.height; ir.target = P.depth;
    ir.rf = RowFilter{P.row_lo, P.row_hi, P.rowmask}; ir.sl = P.sl;
#pragma unroll 1
    for (int i = 0; i < 2; i++) {                                      // (one copy of the rasteriser for both clip halves)
        const bool has = (i ? inl1 : inl0) && frag_ok;
        ir.push(has, i ? st1 : st0, base_f + ex_f + (i ? my_f0 : 0u));  // single-chunk triangles: their one fragment's index
        ir.drain(i == 1);
    }
}

// =====================================================================================================================
// Rasterisation: kernel1 (cl2.cl:4986-5127), kernel2 (5391-5546), kernel1_realtime_shadowing (5130-5246).
//
// The reference gives one work-item one <=501-pixel chunk and walks it sequentially, so a pass lasts as long as its
// longest walk (measured: 82 us for a pass whose total work is ~10 us). Here:
//   * a triangle whose whole walk is one chunk of at most RASTER_SMALL_MAX slots is rasterised by the setup kernel itself,
//     through a per-warp queue (InlineRaster / k_shadow_setup's stage C): no record round trip, no second kernel;
//   * every other fragment goes to a work list and gets a whole warp (k_raster_warp_depth / k_raster_shadow_warp), which
//     resolves its pixel slots 32 at a time with the closed form of the walk (rr_math.cuh walk_pixel).
// Depth goes out as red.global.min.u32 on the L2-resident buffer; ids as red.global.max.u32 (canonical last writer).
// =====================================================================================================================
enum { RM_DEPTH = 0, RM_IDS = 1, RM_SHADOW = 2 };
#define TILE_W 32               // tiles of the dirty-tile read-back (k_tile_mark / k_tile_copy below)
#define TILE_H 4

struct RasterParams {
    const uint32_t* frags; const float4* cutdown; const uint32_t* fragcnt; const uint32_t* counters; uint32_t cap_frags;
    const uint32_t* worklist; const uint32_t* extra;      // main view: fragments to walk (CTR_NWORK), plus — kernel2 only — CTR_NEXTRA unlisted inline ones
    int n_index;                    // CTR_NFRAG or CTR_S_NFRAG
    uint32_t* depth; uint32_t* ids; // RM_SHADOW: depth = base of the cubemap buffer of this pass
    uint32_t slab_of_light[16];     // RM_SHADOW: record word 0 = light << 8 | face; slab index of each light of the pass
    float width, height; int W;
    int row_lo, row_hi;             // rows this context needs (band +- halo)
    const uint8_t* rowmask; uint32_t rowbit;   // interleaved bands: rows whose mask has `rowbit` set (nullptr: all of [row_lo, row_hi))
    uint32_t* shade_list; uint32_t* shade_count;   // k_ids_list<true>: the covered-pixel list of kernel3, appended to by whoever resolves a pixel first
    uint8_t* tile_now; int tiles_x;                // ... which also marks the pixel's tile for the dirty-tile read-back (nullptr: off)
};

template <int MODE>
__device__ __forceinline__ void emit_sample(const RasterParams& P, float x, float y, float A, float B, float C, uint32_t face, uint32_t f) {
    const float fd = fmaf(A, x, fmaf(B, y, C));
    const uint32_t d = sat_u32(RR_U32MAXF / fd);
    if (MODE == RM_DEPTH) {
        red_min_u32(P.depth + ((int)(y * P.width) + (int)x), d);
    } else if (MODE == RM_SHADOW) {
        red_min_u32(P.depth + ((size_t)P.slab_of_light[face >> 8] * 6 + (face & 0xFF)) * P.W * P.W + ((int)(y * P.width) + (int)x), d);
    } else {
        const int px = (int)y * P.W + (int)x;
        const uint32_t val = P.depth[px];
        // racing plain stores in the reference; canonical winner = highest fragment index (last writer in id order).
        // The id image holds fragment index + 1: 0 = no fragment passed the test (e.g. depth < 20, where the reference's
        // unsigned window wraps, cl2.cl:5534) — such pixels are shaded like uncovered ones instead of with a stale id (q7)
        if (d > val - RR_BUF_ERROR && d < val + RR_BUF_ERROR) red_max_u32(P.ids + px, f + 1u);
    }
}

template <int MODE>
__device__ __forceinline__ void load_fragment(const RasterParams& P, uint32_t f, uint32_t& face, uint32_t& distance, FragGeom& g) {
    uint32_t ctri;
    float rconst;
    face = 0;
    if (MODE == RM_SHADOW) {
        const uint4 rec = __ldg(reinterpret_cast<const uint4*>(P.frags) + f);
        face = rec.x; distance = rec.y; ctri = rec.z; rconst = __uint_as_float(rec.w);
    } else {
        const uint32_t* rec = P.frags + (size_t)f * RR_FRAG_WORDS;
        distance = __ldg(rec + 1); ctri = __ldg(rec + 2); rconst = __uint_as_float(__ldg(rec + 3));
    }
    const float4 c0 = __ldg(P.cutdown + (size_t)ctri * 3), c1 = __ldg(P.cutdown + (size_t)ctri * 3 + 1), c2 = __ldg(P.cutdown + (size_t)ctri * 3 + 2);
    g = frag_geom(xyz(c0), xyz(c1), xyz(c2), rconst, P.width, P.height);
}

// rows a chunk can touch, conservatively (for the sort-first band cull)
__device__ __forceinline__ bool chunk_rows_outside(const float4 mm, int op_size, uint32_t distance, int row_lo, int row_hi) {
    const int width = (int)(mm.y - mm.x);
    if (width <= 0) return true;
    const int k0 = op_size * (int)distance;
    const int y_lo = (int)mm.z + k0 / width - 2, y_hi = (int)mm.z + (k0 + op_size) / width + 2;
    return y_hi < row_lo || y_lo >= row_hi;
}

// kernel1 (cl2.cl:4986-5127) for the fragments k_setup_main did not rasterise inline: one warp per fragment of the work list, the
// lanes take the chunk's pixel slots 32 at a time through the closed form of the walk (rr_math.cuh walk_pixel). A chunk has at
// most 501 slots, so a warp's work is bounded and consecutive chunks of a large triangle land on different warps. Depth goes out
// as red.global.min.u32 on the L2-resident buffer. The covered samples (pixel, depth, fragment) are collected in a per-warp stash
// in shared memory and appended to the sample list, so that kernel2 never walks the fragment again: its work is a stream over the
// list (k_ids_list). Fragments whose samples do not fit the list go to the `extra` list instead.
//  * No vote inside the walk: a lane that found a sample takes its place in the stash with a shared-memory atomic and goes on, so
//    the divergent paths of a step interleave like independent warps.
//  * A warp keeps the samples of consecutive fragments in its stash and reserves list space for all of them at once, when the
//    next 128 slots might not fit or at the end of its work. With one reservation per fragment the 17 k atomicAdds of config 3 on
//    the one list counter took as long as the whole walk: same-address atomics retire at about one per nanosecond and every warp
//    waited for its own to come back (34 us instead of 18; profiles/r2b_raster_depth_atomics.txt).
#define RW_WARPS 8
#define RW_STASH 256            // samples a warp holds before it flushes
#define RW_BLOCK 128            // slots walked between two capacity checks
struct RasterStash { uint2 sample[RW_STASH]; uint32_t frag[RW_STASH]; };

__device__ __forceinline__ void raster_stash_flush(const SampleList& sl, RasterStash& st, uint32_t* s_run, int lane) {
    __syncwarp();
    const uint32_t run = *s_run;
    __syncwarp();
    if (run == 0) return;
    uint32_t b0 = 0;
    if (lane == 0) { b0 = atomicAdd(sl.count, run); *s_run = 0u; }
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if ((unsigned long long)b0 + run <= (unsigned long long)sl.cap) {
        for (uint32_t q = lane; q < run; q += 32u) { sl.samples[b0 + q] = st.sample[q]; sl.frag[b0 + q] = st.frag[q]; }
    } else {                                                    // sample list full: kernel2 walks these fragments again
        for (uint32_t q = lane; q < run; q += 32u) {
            if (q == 0 || st.frag[q] != st.frag[q - 1]) extra_push(sl, st.frag[q]);       // (a fragment listed twice is walked twice: harmless)
            if (b0 + q < sl.cap) sl.samples[b0 + q] = make_uint2(0xFFFFFFFFu, 0u);                                  // (no sample here)
        }
    }
    __syncwarp();                                               // the stash is reused
}

__global__ void __launch_bounds__(RW_WARPS * 32) k_raster_warp_depth(const RasterParams P, const SampleList sl) {
    __shared__ RasterStash s_stash[RW_WARPS];
    __shared__ uint32_t s_run[RW_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    RasterStash& st = s_stash[wid];
    if (lane == 0) s_run[wid] = 0u;
    __syncwarp();
    const uint32_t n_work = min(P.counters[CTR_NWORK], P.cap_frags);
    const uint32_t warps = gridDim.x * RW_WARPS;
    uint32_t held = 0;                                          // upper bound of the samples in the stash (warp-uniform)
    for (uint32_t i = blockIdx.x * RW_WARPS + wid; i < n_work; i += warps) {
        const uint32_t f = __ldg(P.worklist + i);
        const uint32_t cnt = __ldg(P.fragcnt + f) & FRAGCNT_MASK;
        if (cnt == 0) continue;
        uint32_t face, distance;
        FragGeom g;
        load_fragment<RM_DEPTH>(P, f, face, distance, g);      // the same record for every lane: broadcast loads
        if (chunk_rows_outside(g.mm, RR_OP_SIZE, distance, P.row_lo, P.row_hi)) continue;
        const int width = (int)(g.mm.y - g.mm.x), k0 = RR_OP_SIZE * (int)distance;
        const float iw = 1.f / (float)width;
        for (uint32_t s0 = 0; s0 < cnt; s0 += RW_BLOCK) {
            const uint32_t s1 = min(cnt, s0 + (uint32_t)RW_BLOCK);
            if (held + (s1 - s0) > (uint32_t)RW_STASH) { raster_stash_flush(sl, st, &s_run[wid], lane); held = 0; }
            held += s1 - s0;
            for (uint32_t s = s0 + lane; s < s1; s += 32u) {
                float x, y;
                if (!walk_pixel(k0 + (int)s, k0, width, iw, g.mm, x, y)) continue;
                const int iy = (int)y;
                if (iy < P.row_lo || iy >= P.row_hi) continue;
                if (P.rowmask && !(P.rowmask[iy] & P.rowbit)) continue;
                if (!point_in_tri(x, y, g.xr.x, g.yr.x, g.xr.y, g.yr.y, g.xr.z, g.yr.z)) continue;
                const float fd = fmaf(g.A, x, fmaf(g.B, y, g.C));
                const uint32_t d = sat_u32(RR_U32MAXF / fd);
                const uint32_t px = (uint32_t)((int)(y * P.width) + (int)x);
                red_min_u32(P.depth + px, d);
                const uint32_t at = atomicAdd(&s_run[wid], 1u);     // at most `held` <= RW_STASH samples
                st.sample[at] = make_uint2(px, d);
                st.frag[at] = f;
            }
        }
    }
    raster_stash_flush(sl, st, &s_run[wid], lane);
}

// kernel2 (cl2.cl:5391-5546). Every covered sample of the frame was recorded by the kernels that wrote its depth (the inline
// rasteriser of k_setup_main, k_raster_warp_depth), so the id resolve is a fully parallel stream over the sample list: a sample
// within +-20 of the final depth proposes its fragment, red.global.max.u32 keeps the canonical winner (highest fragment index).
// The fragments whose samples did not fit the list (`extra`, normally none) are walked again, one warp each.
// LIST: the kernel also builds kernel3's list of covered pixels. The id image is all zero when the frame's id resolve starts, so the
// sample that turns a pixel's id from 0 into something is the first to resolve it: it appends the pixel (a covered pixel is one
// with a resolved id, k_shade_pre4). The appends of a CTA's step are collected in shared memory and reserved with one atomicAdd.
// The list then follows the order of the samples (triangle order) instead of screen order, which k_shade does not care about.
#define IL_STEP 4
template <bool LIST>
__global__ void __launch_bounds__(256) k_ids_list(const SampleList sl, const RasterParams P) {
    __shared__ uint32_t s_px[LIST ? 256 * IL_STEP : 1];
    __shared__ uint32_t s_n, s_base;
    const uint32_t* __restrict__ depth = P.depth;
    uint32_t* __restrict__ ids = P.ids;
    const uint32_t n = min(*sl.count, sl.cap);
    // IL_STEP samples per thread and step: the loads of a step are independent of each other (sample -> depth is the only chain)
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t c0 = blockIdx.x * blockDim.x; c0 < n; c0 += IL_STEP * stride) {        // (CTA-uniform trip count: barriers inside)
        const uint32_t i0 = c0 + threadIdx.x;
        uint2 sm[IL_STEP];
        uint32_t fr[IL_STEP], val[IL_STEP];
        bool ok[IL_STEP];
        if (LIST) { if (threadIdx.x == 0) s_n = 0u; __syncthreads(); }
#pragma unroll
        for (int k = 0; k < IL_STEP; k++) {
            const uint32_t i = i0 + (uint32_t)k * stride;
            ok[k] = i < n;
            sm[k] = ok[k] ? sl.samples[i] : make_uint2(0xFFFFFFFFu, 0u);
            fr[k] = ok[k] ? sl.frag[i] : 0u;
        }
#pragma unroll
        for (int k = 0; k < IL_STEP; k++) {
            const int row = (int)(sm[k].x / (uint32_t)P.W);      // (0xFFFFFFFF marks the inside part of a reservation that did not fit: row >= H)
            ok[k] = ok[k] && row >= P.row_lo && row < P.row_hi && (!P.rowmask || (P.rowmask[row] & ROW_OWNED));
            val[k] = ok[k] ? depth[sm[k].x] : 0u;
        }
#pragma unroll
        for (int k = 0; k < IL_STEP; k++) {
            if (!(ok[k] && sm[k].y > val[k] - RR_BUF_ERROR && sm[k].y < val[k] + RR_BUF_ERROR)) continue;
            if (!LIST) red_max_u32(ids + sm[k].x, fr[k] + 1u);
            else if (atomicMax(ids + sm[k].x, fr[k] + 1u) == 0u && val[k] != 0xFFFFFFFFu) {
                s_px[atomicAdd(&s_n, 1u)] = sm[k].x;
                if (P.tile_now) { const uint32_t y = sm[k].x / (uint32_t)P.W; P.tile_now[(y / TILE_H) * (uint32_t)P.tiles_x + (sm[k].x - y * (uint32_t)P.W) / TILE_W] = 1; }
            }
        }
        if (LIST) {
            __syncthreads();
            const uint32_t cn = s_n;
            if (cn) {
                if (threadIdx.x == 0) s_base = atomicAdd(P.shade_count, cn);
                __syncthreads();
                for (uint32_t q = threadIdx.x; q < cn; q += blockDim.x) P.shade_list[s_base + q] = s_px[q];
            }
            __syncthreads();
        }
    }
    const uint32_t n_extra = min(*sl.extra_count, P.cap_frags);
    if (n_extra == 0) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_extra; i += warps) {
        const uint32_t f = __ldg(sl.extra + i);
        const uint32_t cnt = __ldg(P.fragcnt + f) & FRAGCNT_MASK;
        if (cnt == 0) continue;
        uint32_t face, distance;
        FragGeom g;
        load_fragment<RM_IDS>(P, f, face, distance, g);
        if (chunk_rows_outside(g.mm, RR_OP_SIZE, distance, P.row_lo, P.row_hi)) continue;
        const int width = (int)(g.mm.y - g.mm.x), k0 = RR_OP_SIZE * (int)distance;
        const float iw = 1.f / (float)width;
        for (uint32_t s = lane; s < cnt; s += 32u) {
            float x, y;
            if (!walk_pixel(k0 + (int)s, k0, width, iw, g.mm, x, y)) continue;
            const int iy = (int)y;
            if (iy < P.row_lo || iy >= P.row_hi) continue;
            if (P.rowmask && !(P.rowmask[iy] & P.rowbit)) continue;
            if (!point_in_tri(x, y, g.xr.x, g.yr.x, g.xr.y, g.yr.y, g.xr.z, g.yr.z)) continue;
            if (!LIST) emit_sample<RM_IDS>(P, x, y, g.A, g.B, g.C, face, f);
            else {                                              // emit_sample<RM_IDS> with the first-resolver append
                const float fd = fmaf(g.A, x, fmaf(g.B, y, g.C));
                const uint32_t d = sat_u32(RR_U32MAXF / fd);
                const int px = (int)y * P.W + (int)x;
                const uint32_t val = P.depth[px];
                if (d > val - RR_BUF_ERROR && d < val + RR_BUF_ERROR && atomicMax(P.ids + px, f + 1u) == 0u && val != 0xFFFFFFFFu) {
                    P.shade_list[atomicAdd(P.shade_count, 1u)] = (uint32_t)px;
                    if (P.tile_now) P.tile_now[((uint32_t)iy / TILE_H) * (uint32_t)P.tiles_x + (uint32_t)x / TILE_W] = 1;
                }
            }
        }
    }
}

// =====================================================================================================================
// shadow passes
// =====================================================================================================================
#define SHADOW_MAX_LIGHTS 16
struct ShadowLight { float x, y, z; uint32_t slab; uint32_t face_mask; };

struct ShadowSetupParams {
    const float4* pa; const float4* pb; const float2* pc;
    const ObjLite* objs;
    uint32_t n_tris;
    int n_lights;
    ShadowLight lights[SHADOW_MAX_LIGHTS];
    FaceTable faces;
    float L, icut;
    float far2_max;                          // largest float t with sqrtf(t) <= depth_far: length(v) > depth_far <=> dot(v, v) > far2_max (sqrt.rn is monotone)
    int only_static;
    int pretest;                             // 1: cull clearly back-facing triangles before they are transformed (see shadow_clearly_back)
    uint32_t* frags; uint32_t cap_frags;
    uint32_t* fragcnt;
    float4* cutdown; uint32_t cap_cut;
    uint32_t* counters;
    uint32_t* buffer;                        // cubemap buffer of this pass (inline raster of small triangles)
    const uint4* cluster_faces;              // byte li of cluster_faces[block] = cube faces of light li the cluster's bounding box can
                                             // reach (k_cluster_faces); nullptr = every face
};

// reserve `nc` projected-triangle slots and `nf` fragment records for this lane; one atomic per warp
__device__ __forceinline__ void warp_alloc2(uint32_t* counters, uint32_t nc, uint32_t nf, uint32_t& cbase, uint32_t& fbase) {
    const int lane = threadIdx.x & 31;
    const unsigned long long mine = ((unsigned long long)nc << 32) | nf;
    unsigned long long inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    const unsigned long long total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned long long base = 0;
    if (lane == 31 && total) base = atomicAdd(reinterpret_cast<unsigned long long*>(counters + CTR_S_NFRAG), total);
    base = __shfl_sync(0xffffffffu, base, 31) + inc - mine;
    cbase = (uint32_t)(base >> 32);
    fbase = (uint32_t)(base & 0xFFFFFFFFull);
}

// Minimum resident CTAs per SM the register allocation is asked to allow (build-time knobs for A/B runs:
// -DRR_LB_SHADOW_SETUP=n, -DRR_LB_SHADE=n, RR_LIB=<other build>; examples/lb_sweep.sh).
#ifndef RR_LB_SHADOW_SETUP
#define RR_LB_SHADOW_SETUP 7
#endif
#define RR_LB_SHADOW_SETUP_ATTR __launch_bounds__(128, RR_LB_SHADOW_SETUP)
#ifdef RR_LB_SHADE
#define RR_LB_SHADE_ATTR __launch_bounds__(128, RR_LB_SHADE)
#else
#define RR_LB_SHADE_ATTR __launch_bounds__(128)
#endif

// Per-warp work queues of k_shadow_setup (shared memory). The kernel is a three-stage pipeline inside each warp, with a
// compaction between the stages so that every stage runs with (nearly) all 32 lanes busy:
//   A  per (triangle, light): which cube faces does the triangle mark, is it clearly back-facing   -> items (lane, face, light)
//   B  per item: rotate into the face's camera, clip, project, cull, classify                      -> small triangles / records
//   C  per small triangle: rasterise its box straight into the cubemap (kernel1_realtime_shadowing's work for it)
#define SQ_GROUP 4              // lights per stage-A round (one byte of face bits per light in a 32-bit word)
#define SQ_ITEMS 448            // stage A -> B: up to 3 faces x 4 lights x 32 lanes arrive at once on top of < 32 waiting items (+ 32 requeued clip halves)
#define SR_SLOTS 64             // stage B -> C: 32 arrive on top of < 32 waiting
#define SR_FIELDS 11            // rounded x of the 3 vertices, rounded y, camera z, rconst, face slab index
struct ShadowWarpQueue {
    float f[SR_FIELDS][SR_SLOTS];
    unsigned short item[SQ_ITEMS];          // src lane (5 bits) | face << 5 (3) | light << 8 (4) | two_sided << 12 | second clip half << 13
};

// Stage A's early cull. A triangle that is not two-sided is dropped by prearrange_realtime_shadowing when its projection is
// not front-facing, i.e. when cross(p1 - p0, p2 - p0).z >= 0 (cl2.cl:491-494, 4571). For an unclipped triangle that sign is
// the sign of det(q0, q1, q2) (camera-space vertices, all z > 0), every face camera is a proper rotation about the light, so
// it is the sign of D = (w0 - light) . n with n = (w1 - w0) x (w2 - w0) — the same for all six faces. The cull fires only
// when the sign is far from being decided by rounding and the triangle cannot be clipped in any face it marks:
//   * D > 0 and D^2 > 0.0025 |w0 - light|^2 |n|^2      (the view ray is more than ~2.9 degrees off the triangle's plane)
//   * longest edge <= 0.1 |w0 - light|                  (then every vertex has z >= 0.32 |w0 - light| in a face that holds one of them)
//   * 10 (icut + 1)^2 <= |w0 - light|^2 <= 0.8 far^2    (so 0.32 |w0 - light| > icut and z < far: no near / far clipping)
//   * |n| fov >= 0.22 longest edge |w0 - light|         (projected height >= 5e-3 px at that angle: far above the ~5e-4 px noise
//                                                        of the projected coordinates the reference decides with)
// Everything else goes through the exact test of stage B. emax2 = squared longest edge, nn = |n|^2, fov2 = (L/2)^2.
__device__ __forceinline__ bool shadow_clearly_back(float3 rel0, float3 n, float nn, float emax2, float fov2, float r2lo, float r2hi) {
    const float D = dot3(rel0, n), r2 = dot3(rel0, rel0);
    return D > 0.f && D * D > 0.0025f * r2 * nn && emax2 <= 0.01f * r2 && r2 >= r2lo && r2 <= r2hi && nn * fov2 >= 0.0484f * emax2 * r2;
}

// The literal path of stage C: replays the reference's walk (scan_chunk) with point_in_tri. Only taken by triangles the exact
// fast path below cannot take (coordinates beyond 2047, a clamped box, a lagging row counter); kept out of line so that its
// registers do not count against the kernel.
__device__ __noinline__ void shadow_raster_small_literal(float3 xr, float3 yr, float3 zc, float rconst, float L, uint32_t* __restrict__ target) {
    const float4 mm = calc_min_max(xr, yr, L, L);
    float3 d = make_float3(zc.x / RR_DEPTH_FAR, zc.y / RR_DEPTH_FAR, zc.z / RR_DEPTH_FAR);     // dcalc
    d = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);                                     // native_recip
    float A, B, C;
    interpolate_get_const(d, xr, yr, rconst, A, B, C);
    scan_chunk(mm, RR_OP_SIZE_LIGHT, 0u, [&](float x, float y) {
        if (point_in_tri(x, y, xr.x, yr.x, xr.y, yr.y, xr.z, yr.z)) {
            const float fd = fmaf(A, x, fmaf(B, y, C));
            red_min_u32(target + ((int)(y * L) + (int)x), sat_u32(RR_U32MAXF / fd));
        }
    });
}

#define SR_EXACT 0x80000000u    // flag in a queued triangle's face word: stage B found its coordinates small enough for exact arithmetic

// Stage C for one queued small triangle (single chunk, box of at most RASTER_SMALL_MAX pixel slots).
// Fast path — every operation below is exact in fp32, so it yields bit for bit what the reference's walk + point_in_tri do:
//   * SR_EXACT: the rounded vertex coordinates are integers with |v| <= 2047 and extents <= 64, so every product and partial
//     sum of point_in_tri (cl2.cl:4798-4807) is an integer below 2^24: the edge functions can be stepped incrementally, and
//     "inside" means inside the CLOSED rounded triangle (s >= 0, t >= 0, s + t <= 2|A|; 2.0001f * |A| < 2|A| + 1 at this size).
//     Hence only pixels of the vertices' own bounding box can be covered: the walk's extra first row and column
//     (bbox = [round(min) - 1, round(max)), cl2.cl:420-441) are skipped, the last ones it never visits are left out as well.
//   * no row of the walk lags: its float row counter floor(fma(k, 1/width, min_y)) is exact at a row start r * width when
//     r <= min_y (the error r * 2^-24 stays below half an ulp of min_y + r); the few boxes within `rows` of the top edge are
//     checked row by row. Without a lagging row the walk visits exactly the box (rr_math.cuh scan_chunk).
// Covered pixels are collected in a bit mask first; the depth plane (six IEEE divisions) is only paid for triangles with a hit.
__device__ __forceinline__ void shadow_raster_small(const ShadowWarpQueue& Q, int slot, bool valid, float L, uint32_t* __restrict__ buffer) {
    if (!valid) return;
    const float* f = &Q.f[0][slot];
    const float3 xr = make_float3(f[0], f[SR_SLOTS], f[2 * SR_SLOTS]), yr = make_float3(f[3 * SR_SLOTS], f[4 * SR_SLOTS], f[5 * SR_SLOTS]);
    const uint32_t fw = __float_as_uint(f[10 * SR_SLOTS]);
    uint32_t* target = buffer + (size_t)(fw & ~SR_EXACT) * (size_t)(L * L);
    const float4 mm = calc_min_max(xr, yr, L, L);
    // tight box: the vertices' own bounding box inside the walk's box
    const float x0 = fmaxf(mm.x, fminf(fminf(xr.x, xr.y), xr.z)), y0 = fmaxf(mm.z, fminf(fminf(yr.x, yr.y), yr.z));
    const int wt = (int)(mm.y - x0), ht = (int)(mm.w - y0), rows = (int)(mm.w - mm.z);
    bool fast = (fw & SR_EXACT) && wt * ht <= 32;
    if (fast && (float)(rows - 1) > mm.z) {                 // a box at the very top of the face: look for a lagging row
        const int width = (int)(mm.y - mm.x);
        const float iw = 1.f / (float)width;
        for (int r = 1; r < rows; r++) fast = fast && walk_row(r * width, iw, mm.z) == mm.z + (float)r;
    }
    if (!fast) {
        shadow_raster_small_literal(xr, yr, make_float3(f[6 * SR_SLOTS], f[7 * SR_SLOTS], f[8 * SR_SLOTS]), f[9 * SR_SLOTS], L, target);
        return;
    }
    if (wt <= 0 || ht <= 0) return;
    const float Ah = 0.5f * (-yr.y * xr.z + yr.x * (-xr.y + xr.z) + xr.x * (yr.y - yr.z) + xr.y * yr.z);
    const float sign = Ah < 0 ? -1.f : 1.f;
    const float lim = 2.0001f * Ah * sign;
    const float as = (yr.z - yr.x) * sign, bs = (xr.x - xr.z) * sign, at = (yr.x - yr.y) * sign, bt = (xr.y - xr.x) * sign;
    float s_row = (yr.x * xr.z - xr.x * yr.z) * sign + as * x0 + bs * y0;
    float t_row = (xr.x * yr.y - yr.x * xr.y) * sign + at * x0 + bt * y0;
    uint32_t mask = 0u, bit = 1u;
    for (int r = 0; r < ht; r++) {
        float s = s_row, t = t_row;
        for (int c = 0; c < wt; c++) {
            if (s > -0.0001f && t > -0.0001f && (s + t) < lim) mask |= bit;
            bit <<= 1;
            s += as; t += at;
        }
        s_row += bs; t_row += bt;
    }
    if (!mask) return;
    float3 d = make_float3(f[6 * SR_SLOTS] / RR_DEPTH_FAR, f[7 * SR_SLOTS] / RR_DEPTH_FAR, f[8 * SR_SLOTS] / RR_DEPTH_FAR);     // dcalc, cl2.cl:5029-5040
    d = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);                                                                        // native_recip
    float A, B, C;
    interpolate_get_const(d, xr, yr, f[9 * SR_SLOTS], A, B, C);
    const float iwt = __fdividef(1.f, (float)wt);
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1u;
        const int r = (int)(((float)i + 0.5f) * iwt);       // i / wt: i < 32, wt <= 32, so (i + 0.5) / wt is at least 1/64 away from an integer
        const float x = x0 + (float)(i - r * wt), y = y0 + (float)r;
        const float fd = fmaf(A, x, fmaf(B, y, C));
        red_min_u32(target + ((int)(y * L) + (int)x), sat_u32(RR_U32MAXF / fd));
    }
}

// =====================================================================================================================
// shadow passes: k_shadow_setup == prearrange_realtime_shadowing (cl2.cl:4420-4636) for ALL lights of a pass in one launch,
// plus kernel1_realtime_shadowing (cl2.cl:5130-5246) for every triangle whose walk is one small chunk (97 % of config 3's).
// The reference launches prearrange once per light and re-reads / re-transforms every triangle each time. Here a thread
// loads its triangle once and brings it to world space once (bit-identical: that part of full_rotate_quat does not depend
// on the light); the per-light, per-face work then flows through the warp's queues (ShadowWarpQueue). Slot numbers of a
// shadow pass are never observable (only the atomic_min result is), so the few triangles that need records get them from
// ONE warp-aggregated 64-bit atomicAdd that reserves projected-triangle slots and fragment records together.
// Records: {light << 8 | face, chunk, c_id, bits(rconst)} (cl2.cl:4626-4631 with the light folded into word 0); they are
// rasterised by k_raster_shadow_warp.
// =====================================================================================================================
// The rare tail of stage B: a triangle that is larger than one small chunk gets a projected-triangle slot and one record per chunk
// (rasterised by k_raster_shadow_warp). Called by all 32 lanes (the allocation is one warp-aggregated atomic); out of line so that
// its registers and code stay out of the hot loop.
__device__ __noinline__ void shadow_store_big(uint32_t* __restrict__ counters, float4* __restrict__ cutdown, uint32_t cap_cut, uint32_t* __restrict__ frags, uint32_t cap_frags,
                                              uint32_t* __restrict__ fragcnt, bool big, float3 p0, float3 p1, float3 p2, float3 xr, float3 yr, float rconst, float area,
                                              float L, uint32_t word0) {
    const int n_frag = big ? (int)ceilf(area / (float)RR_OP_SIZE_LIGHT) : 0;
    uint32_t cid, fbase;
    warp_alloc2(counters, big ? 1u : 0u, (uint32_t)n_frag, cid, fbase);
    if (!big) return;
    if (cid + 1u > cap_cut) { atomicOr(&counters[CTR_OVERFLOW], 2u); return; }
    if ((unsigned long long)fbase + (uint32_t)n_frag > (unsigned long long)cap_frags) { atomicOr(&counters[CTR_OVERFLOW], 1u); return; }
    float4* dst = cutdown + (size_t)cid * 3;
    dst[0] = make_float4(p0.x, p0.y, p0.z, 0.f);
    dst[1] = make_float4(p1.x, p1.y, p1.z, 0.f);
    dst[2] = make_float4(p2.x, p2.y, p2.z, 0.f);
    const float4 mm = calc_min_max(xr, yr, L, L);
    const int width = (int)(mm.y - mm.x), nrows = (int)(mm.w - mm.z);
    const int kend = walk_end(width, nrows, 1.f / (float)width, mm.z, mm.w);
    uint4* rec = reinterpret_cast<uint4*>(frags) + fbase;
    for (int a = 0; a < n_frag; a++) {
        rec[a] = make_uint4(word0, (uint32_t)a, cid, __float_as_uint(rconst));
        fragcnt[fbase + a] = chunk_slots(kend, a, RR_OP_SIZE_LIGHT);
    }
}

__global__ void RR_LB_SHADOW_SETUP_ATTR k_shadow_setup(const ShadowSetupParams P) {
    __shared__ ShadowWarpQueue s_q[128 / 32];
    __shared__ RotSC s_face[6];
    __shared__ float4 s_light[SHADOW_MAX_LIGHTS];    // position, w = bits of (slab << 8 | face_mask)
    static_assert(CLUSTER_TRIS == 128, "one k_shadow_setup block == one cluster");
    static_assert(SHADOW_MAX_LIGHTS <= 16, "face bits of a pass: one byte per light in two 64-bit words");
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t tri = blockIdx.x * blockDim.x + tid;
    uint4 reach4 = make_uint4(0x3F3F3F3Fu, 0x3F3F3F3Fu, 0x3F3F3F3Fu, 0x3F3F3F3Fu);
    if (P.cluster_faces) reach4 = __ldg(P.cluster_faces + blockIdx.x);
    auto reach_of = [&](int li) -> uint32_t {
        const uint32_t word = li < 4 ? reach4.x : (li < 8 ? reach4.y : (li < 12 ? reach4.z : reach4.w));
        return (word >> ((li & 3) * 8)) & 0x3Fu;
    };
    uint32_t light_reach = 0u;               // bit li: the cluster can reach a face of light li that is rendered here (block-uniform)
    for (int li = 0; li < P.n_lights; li++) if (reach_of(li) & P.lights[li].face_mask) light_reach |= 1u << li;
    if (!light_reach) return;                // none of this context's faces can see the cluster
    if (tid < 6) s_face[tid] = P.faces.r[tid];
    if (tid < P.n_lights) s_light[tid] = make_float4(P.lights[tid].x, P.lights[tid].y, P.lights[tid].z, __uint_as_float((P.lights[tid].slab << 8) | P.lights[tid].face_mask));
    __syncthreads();

    ShadowWarpQueue& Q = s_q[tid >> 5];
    const float L = P.L, half = P.L / 2.f;
    // ---- the triangle in world space, and (stage A's arithmetic) the cube faces of every light it goes to: byte li of (flo, fhi)
    float3 w0 = make_float3(0, 0, 0), w1 = w0, w2 = w0;
    unsigned long long flo = 0ull, fhi = 0ull;
    bool two_sided = false;
    if (tri < P.n_tris) {
        const float4 a = __ldg(P.pa + tri), b = __ldg(P.pb + tri);
        const float2 c = __ldg(P.pc + tri);
        const ObjLite G = P.objs[__float_as_uint(c.y)];
        const bool is_static = (G.feature_flag & RR_FEATURE_IS_STATIC) > 0;
        if (!((!P.only_static && is_static) || (P.only_static && !is_static))) {               // cl2.cl:4460-4464
            const float3 gpos = make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z);
            const float sc = G.pos_scale.w;
            w0 = rot_quat_n(make_float3(a.x, a.y, a.z) * sc, G.nquat) + gpos;                  // cl2.cl:4524-4525 == 505-507
            w1 = rot_quat_n(make_float3(a.w, b.x, b.y) * sc, G.nquat) + gpos;
            w2 = rot_quat_n(make_float3(b.z, b.w, c.x) * sc, G.nquat) + gpos;
            two_sided = (G.feature_flag & RR_FEATURE_TWO_SIDED) > 0;
            const float3 e1 = w1 - w0, e2 = w2 - w0, e3 = w2 - w1;
            const float3 nrm = cross3(e1, e2);
            const float nn = dot3(nrm, nrm), emax2 = fmaxf(fmaxf(dot3(e1, e1), dot3(e2, e2)), dot3(e3, e3));
            const bool pretest = P.pretest && !two_sided && isfinite(nn) && isfinite(emax2);
            const float r2lo = 10.f * (P.icut + 1.f) * (P.icut + 1.f), r2hi = 0.8f * RR_DEPTH_FAR * RR_DEPTH_FAR, fov2 = half * half;
#pragma unroll 1
            for (int li = 0; li < P.n_lights; li++) {
                if (!((light_reach >> li) & 1u)) continue;
                const float4 l4 = s_light[li];
                const float3 lpos = make_float3(l4.x, l4.y, l4.z);
                const float3 gl = gpos - lpos;
                if (dot3(gl, gl) > P.far2_max) continue;                                        // length(...) > depth_far, cl2.cl:4472 (far2_max: see ShadowSetupParams)
                const uint32_t reach = reach_of(li);
                uint32_t faces;
                // a cluster whose box reaches a single face: every vertex is assigned that face (the reach is a superset)
                if ((reach & (reach - 1u)) == 0u) faces = reach;
                else faces = (1u << ret_cubeface(w0, lpos)) | (1u << ret_cubeface(w1, lpos)) | (1u << ret_cubeface(w2, lpos));   // cl2.cl:4520-4539
                faces &= __float_as_uint(l4.w) & 0x3Fu;
                if (faces && pretest && shadow_clearly_back(w0 - lpos, nrm, nn, emax2, fov2, r2lo, r2hi)) faces = 0;
                if (li < 8) flo |= (unsigned long long)faces << (li * 8); else fhi |= (unsigned long long)faces << ((li - 8) * 8);
            }
        }
    }
    const unsigned lt = (1u << lane) - 1u;
    const int n_groups = (P.n_lights + SQ_GROUP - 1) / SQ_GROUP;

    int qa = 0, qr = 0, g = 0;               // items waiting for stage B / small triangles waiting for stage C / next group of lights (warp-uniform)
    while (true) {
        // ---- stage A: the items of the next four lights, once fewer than 32 are waiting
        if (g < n_groups && qa < 32) {
            const uint32_t word = (uint32_t)((g < 2 ? flo : fhi) >> ((g & 1) * 32));
            const int cnt = __popc(word);
            int inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            int at = qa + inc - cnt;
            const uint32_t base_item = (uint32_t)lane | ((uint32_t)(g * SQ_GROUP) << 8) | (two_sided ? 1u << 12 : 0u);
            for (uint32_t m = word; m; m &= m - 1u) {
                const uint32_t bpos = (uint32_t)__ffs(m) - 1u;                                  // bit = 8 * light-in-group + face
                Q.item[at++] = (unsigned short)(base_item + ((bpos >> 3) << 8) + ((bpos & 7u) << 5));
            }
            qa += __shfl_sync(0xffffffffu, inc, 31);
            g++;
            __syncwarp();
        }
        const bool a_done = g >= n_groups;
        // ---- stage B: 32 items at a time (whatever is left after the last light)
        const bool do_b = qa >= 32 || (a_done && qa > 0);
        if (do_b) {
            const int take = min(qa, 32);
            qa -= take;
            const bool have = lane < take;
            const uint32_t item = have ? Q.item[qa + lane] : 0u;
            __syncwarp();
            const int src = item & 31, kk = (item >> 5) & 7, l = (item >> 8) & 15;
            const bool ts = (item >> 12) & 1u, second = (item >> 13) & 1u;
            float3 v0, v1, v2;               // the source lane's world-space triangle
            v0.x = __shfl_sync(0xffffffffu, w0.x, src); v0.y = __shfl_sync(0xffffffffu, w0.y, src); v0.z = __shfl_sync(0xffffffffu, w0.z, src);
            v1.x = __shfl_sync(0xffffffffu, w1.x, src); v1.y = __shfl_sync(0xffffffffu, w1.y, src); v1.z = __shfl_sync(0xffffffffu, w1.z, src);
            v2.x = __shfl_sync(0xffffffffu, w2.x, src); v2.y = __shfl_sync(0xffffffffu, w2.y, src); v2.z = __shfl_sync(0xffffffffu, w2.z, src);
            bool requeue = false, small = false, big = false;
            float3 p0, p1, p2, xr, yr;       // (only read by lanes that set them: no defaults, so no moves at the branch joins)
            float rconst, area;
            uint32_t faceword;
            if (have) {
                const float4 l4 = s_light[l];
                const float3 lpos = make_float3(l4.x, l4.y, l4.z);
                faceword = (__float_as_uint(l4.w) >> 8) * 6u + (uint32_t)kk;
                const RotSC fr = s_face[kk];
                float3 a0, a1, a2;
                const int num = clip_near_one(rot(v0, lpos, fr), rot(v1, lpos, fr), rot(v2, lpos, fr), P.icut, second, a0, a1, a2);
                requeue = !second && num > 1;                                                    // the other half of a clipped triangle: same item again
                if (num > (second ? 1 : 0)) {
                    p0 = project(a0, half, half, half); p1 = project(a1, half, half, half); p2 = project(a2, half, half, half);
                    // cull + bbox + fragment count, cl2.cl:4571-4597 (classify() with the bookkeeping stage C needs)
                    const bool valid = ts || front_facing(p0, p1, p2);
                    // all three on one outer side of the viewport (cl2.cl:4583-4586): min / max form unless a coordinate is NaN
                    // (fminf / fmaxf skip NaNs, the reference's comparisons do not); a NaN survives the sum, inf - inf makes one
                    const float chk = ((p0.x + p1.x) + (p2.x + p0.y)) + (p1.y + p2.y);
                    bool cond;
                    if (chk == chk) {
                        const float xmax = fmaxf(fmaxf(p0.x, p1.x), p2.x), xmin = fminf(fminf(p0.x, p1.x), p2.x);
                        const float ymax = fmaxf(fmaxf(p0.y, p1.y), p2.y), ymin = fminf(fminf(p0.y, p1.y), p2.y);
                        cond = xmax < 0 || xmin >= L || ymax < 0 || ymin >= L;
                    } else {
                        cond = (p0.x < 0 && p1.x < 0 && p2.x < 0) || (p0.x >= L && p1.x >= L && p2.x >= L) || (p0.y < 0 && p1.y < 0 && p2.y < 0) || (p0.y >= L && p1.y >= L && p2.y >= L);
                    }
                    if (valid && !cond) {
                        xr = make_float3(roundf(p0.x), roundf(p1.x), roundf(p2.x));
                        yr = make_float3(roundf(p0.y), roundf(p1.y), roundf(p2.y));
                        const float det = xr.y * yr.z + xr.x * (yr.y - yr.z) - xr.z * yr.y + (xr.z - xr.y) * yr.x;      // calc_rconstant_v
                        rconst = recip_or_inf(det);
                        const float4 mm = calc_min_max(xr, yr, L, L);
                        area = (mm.y - mm.x) * (mm.w - mm.z);
                        small = area >= 1.f && area <= (float)RASTER_SMALL_MAX;                  // one chunk (ceil(area / 300) == 1) of at most 48 slots
                        big = area > (float)RASTER_SMALL_MAX;                                    // (an empty or NaN box produces nothing)
                        if (small && tri_exact_arith(xr, yr)) {
                            faceword |= SR_EXACT;
                            if (det == 0.f) small = false;                                       // collinear after rounding: point_in_tri's s + t < 0 can never hold
                        }
                    }
                }
            }
            const unsigned mq = __ballot_sync(0xffffffffu, requeue);
            if (mq) {
                if (requeue) Q.item[qa + __popc(mq & lt)] = (unsigned short)(item | (1u << 13));
                qa += __popc(mq);
            }
            // small triangles: queued for stage C, nothing is stored for them
            const unsigned ms = __ballot_sync(0xffffffffu, small);
            if (small) {
                float* f = &Q.f[0][qr + __popc(ms & lt)];
                f[0] = xr.x; f[SR_SLOTS] = xr.y; f[2 * SR_SLOTS] = xr.z;
                f[3 * SR_SLOTS] = yr.x; f[4 * SR_SLOTS] = yr.y; f[5 * SR_SLOTS] = yr.z;
                f[6 * SR_SLOTS] = p0.z; f[7 * SR_SLOTS] = p1.z; f[8 * SR_SLOTS] = p2.z;
                f[9 * SR_SLOTS] = rconst; f[10 * SR_SLOTS] = __uint_as_float(faceword);
            }
            qr += __popc(ms);
            // the rest gets records
            if (__any_sync(0xffffffffu, big))
                shadow_store_big(P.counters, P.cutdown, P.cap_cut, P.frags, P.cap_frags, P.fragcnt, big, p0, p1, p2, xr, yr, rconst, area, L, ((uint32_t)l << 8) | (uint32_t)kk);
            __syncwarp();
        }
        // ---- stage C: 32 small triangles at a time (whatever is left once stages A and B are done)
        const bool do_c = qr >= 32 || (a_done && qa == 0 && qr > 0);
        if (do_c) {
            const int take = min(qr, 32);
            qr -= take;
            shadow_raster_small(Q, qr + lane, lane < take, L, P.buffer);
            __syncwarp();
        }
        if (a_done && !do_b && !do_c) break;
    }
}

// kernel1_realtime_shadowing (cl2.cl:5130-5246) for the fragments k_shadow_setup stored: one warp per fragment, the lanes
// take the chunk's pixel slots 32 at a time through the closed form of the walk (walk_pixel). A chunk has at most 301
// slots, so the work per warp is bounded and no prefix sum over slot counts is needed; config 3 stores about a thousand
// fragments per frame, shadow-heavy scenes (a ground plane seen from a light at L = 2048) tens of thousands.
__global__ void __launch_bounds__(256) k_raster_shadow_warp(const RasterParams P) {
    const int lane = threadIdx.x & 31;
    const uint32_t n = min(P.counters[CTR_S_NFRAG], P.cap_frags);
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < n; f += warps) {
        const uint32_t cnt = __ldg(P.fragcnt + f) & FRAGCNT_MASK;
        if (cnt == 0) continue;
        uint32_t face, distance;
        FragGeom g;
        load_fragment<RM_SHADOW>(P, f, face, distance, g);      // the same record for every lane: broadcast loads
        const int width = (int)(g.mm.y - g.mm.x), k0 = RR_OP_SIZE_LIGHT * (int)distance;
        if (width <= 0) continue;
        const float iw = 1.f / (float)width;
        for (uint32_t s = lane; s < cnt; s += 32u) {
            float x, y;
            if (!walk_pixel(k0 + (int)s, k0, width, iw, g.mm, x, y)) continue;
            if (!point_in_tri(x, y, g.xr.x, g.yr.x, g.xr.y, g.yr.y, g.xr.z, g.yr.z)) continue;
            emit_sample<RM_SHADOW>(P, x, y, g.A, g.B, g.C, face, 0u);
        }
    }
}

// Per-cluster cube-face reach for the face-sharded shadow pass: the faces of each light that ANY point of the cluster's
// world-space bounding box can be assigned to by ret_cubeface (cl2.cl:1745-1790), by interval arithmetic on light-relative
// coordinates. A conservative superset of what the per-vertex test of prearrange_realtime_shadowing (cl2.cl:4520-4539)
// marks, so skipping a cluster whose reach misses the faces rendered here cannot change a cubemap texel.
// Byte li of out[cluster] = 6-bit face mask for light li of the pass.
struct ObjFacesParams { ShadowLight lights[SHADOW_MAX_LIGHTS]; int n_lights; };

__device__ __forceinline__ float iv_min_abs(float lo, float hi) { return (lo <= 0.f && hi >= 0.f) ? 0.f : fminf(fabsf(lo), fabsf(hi)); }
__device__ __forceinline__ float iv_max_abs(float lo, float hi) { return fmaxf(fabsf(lo), fabsf(hi)); }

__global__ void __launch_bounds__(128) k_cluster_faces(const ClusterBox* __restrict__ boxes, uint32_t n_clusters, const ObjLite* __restrict__ objs, uint32_t n_objs,
                                                       const ObjFacesParams P, uint4* __restrict__ out, uint32_t* __restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {                            // first launch of a shadow pass: its allocation state (no cudaMemsetAsync inside a frame, see k_frame_prologue)
        counters[CTR_STICKY] |= counters[CTR_OVERFLOW];           // one pass's overflow stays visible until rr_sync has reported it
        counters[CTR_OVERFLOW] = 0u;
        counters[CTR_S_NFRAG] = 0u; counters[CTR_S_NCUT] = 0u;
    }
    if (i >= n_clusters) return;
    const ClusterBox b = boxes[i];
    const uint32_t oid = __float_as_uint(b.lo.w);
    uint32_t w[4] = {0x3F3F3F3Fu, 0x3F3F3F3Fu, 0x3F3F3F3Fu, 0x3F3F3F3Fu};
    if (b.hi.w == 1.f && oid < n_objs) {
        const ObjLite G = objs[oid];
        const float INF = __int_as_float(0x7f800000);
        float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
        bool fin = true;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float3 q = cluster_corner_world(b, k, G);
            lo[0] = fminf(lo[0], q.x); hi[0] = fmaxf(hi[0], q.x);
            lo[1] = fminf(lo[1], q.y); hi[1] = fmaxf(hi[1], q.y);
            lo[2] = fminf(lo[2], q.z); hi[2] = fmaxf(hi[2], q.z);
            fin = fin && isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
        }
        if (fin) {
            w[0] = w[1] = w[2] = w[3] = 0u;
            for (int li = 0; li < P.n_lights; li++) {
                const float l[3] = {P.lights[li].x, P.lights[li].y, P.lights[li].z};
                float rl[3], rh[3];
                float amax = 0.f;
#pragma unroll
                for (int k = 0; k < 3; k++) { rl[k] = lo[k] - l[k]; rh[k] = hi[k] - l[k]; amax = fmaxf(amax, iv_max_abs(rl[k], rh[k])); }
                const float e = 1e-5f * amax + 1e-3f;               // corner-vs-vertex arithmetic slack (~100 ulps)
#pragma unroll
                for (int k = 0; k < 3; k++) { rl[k] -= e; rh[k] += e; }
                const float nx = iv_min_abs(rl[0], rh[0]), ny = iv_min_abs(rl[1], rh[1]), nz = iv_min_abs(rl[2], rh[2]);
                const float xx = iv_max_abs(rl[0], rh[0]), xy = iv_max_abs(rl[1], rh[1]), xz = iv_max_abs(rl[2], rh[2]);
                // (ret_cubeface's fall-through to face 0 needs a NaN: for finite coordinates one of its three tests holds)
                uint32_t m = isfinite(amax) ? 0u : 0x3Fu;
                if (xx >= ny && xx >= nz) m |= (rl[0] < 0.f ? 1u << 4 : 0u) | (rh[0] >= 0.f ? 1u << 5 : 0u);
                if (xy >= nx && xy >= nz) m |= (rl[1] < 0.f ? 1u << 1 : 0u) | (rh[1] >= 0.f ? 1u << 3 : 0u);
                if (xz >= nx && xz >= ny) m |= (rl[2] < 0.f ? 1u << 2 : 0u) | (rh[2] >= 0.f ? 1u << 0 : 0u);
                w[li >> 2] |= m << ((li & 3) * 8);
            }
        }
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// =====================================================================================================================
// Multi-GPU exchange over peer memory (NVLink P2P). One process per GPU; every context maps its peers'
// cubemap buffers, rank 0's colour target and a small control block (cudaIpc*, rr_mgpu_connect).
//   shadow faces : each context rasterises the (light, face) pairs it owns into its own cubemap buffer, then k_push_faces
//                  copies them into every peer's buffer with 16-byte stores and raises a flag in every peer's control
//                  block. Only 512-byte chunks that hold a sample now, or held one the last time this buffer was used, are
//                  sent (cubemaps are mostly empty): receivers never clear the faces they do not own.
//   colour bands : k_shade / k_shade_pre* store straight into rank 0's frame buffer; the last CTA of k_shade raises the
//                  context's flag on rank 0.
// Flags are monotonically increasing frame counters, written after a system-scope fence by the last CTA of the producing
// kernel and polled by k_wait_flags (one CTA) in front of the consuming kernel. Waits are bounded by a wall-clock limit.
// =====================================================================================================================
#define MG_MAX_WORLD 16
#define MG_PUSH_CHUNK_WORDS 128             // one warp iteration: 32 lanes x uint4 = 512 bytes
struct MgCtrl {
    uint32_t shadow_flag[MG_MAX_WORLD];     // [q] = last shadow epoch whose faces rank q has delivered here
    uint32_t draw_flag[MG_MAX_WORLD];       // [q] = last draw epoch rank q has finished storing into this context's colour target
    uint32_t push_done, shade_done;         // last-CTA counters of the local producing kernels
    uint32_t error;                         // bit 0: a wait timed out
    uint32_t pushed_chunks;                 // statistics: 512-byte chunks k_push_faces has sent to each peer since the last rr_mgpu_pushed_bytes
    uint32_t fb_free;                       // written by rank 0: last draw epoch whose colour target on rank 0 may be stored into (the copy of
                                            // the frame that used the ring slot before has left it); peers wait for it before their first store
    uint32_t _pad[11];
};

struct MgSignal {                           // raised by the last CTA of a kernel; n == 0: nothing to do
    uint32_t* counter;                      // local last-CTA counter
    uint32_t* flag[MG_MAX_WORLD];           // remote (or local) flag words
    int n;
    uint32_t value;
};

__device__ __forceinline__ void mg_signal_tail(const MgSignal& S) {
    if (S.n == 0) return;
    __threadfence_system();                 // this thread's stores (peer memory included) before the counter
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(S.counter, 1u);
        if (t == gridDim.x * gridDim.y - 1) {
            *S.counter = 0u;
            __threadfence_system();
            for (int i = 0; i < S.n; i++) *reinterpret_cast<volatile uint32_t*>(S.flag[i]) = S.value;
        }
    }
}

// Raise one flag after everything enqueued before this launch on the same stream has completed (stream order makes those
// kernels' stores, peer memory included, visible before this kernel starts; the fence orders them before the flag).
// Used after k_shade: a tail inside k_shade would cost it 9 registers and an occupancy step.
__global__ void k_signal_flag(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(flag) = value;
}

// same for several flags at once (rank 0 -> every peer's fb_free), one lane per flag
struct MgFlagList { uint32_t* flag[MG_MAX_WORLD]; int n; uint32_t value; };
__global__ void __launch_bounds__(32) k_signal_flags(const MgFlagList S) {
    __threadfence_system();
    if ((int)threadIdx.x < S.n) *reinterpret_cast<volatile uint32_t*>(S.flag[threadIdx.x]) = S.value;
}

struct MgWait { const uint32_t* flag[MG_MAX_WORLD]; int n; uint32_t value; uint32_t* error; unsigned long long timeout_ns; };

__global__ void __launch_bounds__(32) k_wait_flags(const MgWait Wt) {
    const int lane = threadIdx.x;
    if (lane >= Wt.n) return;
    const volatile uint32_t* f = Wt.flag[lane];
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t spins = 0;
    while ((int32_t)(*f - Wt.value) < 0) {
        __nanosleep(200);
        if ((++spins & 1023u) == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > Wt.timeout_ns) { atomicOr(Wt.error, 1u); break; }
        }
    }
    __threadfence_system();                 // acquire side: the kernels that follow read what the flag's writer stored before it
}

struct MgPushParams {
    const uint32_t* local;                  // this context's cubemap buffer of the epoch
    uint32_t* peer[MG_MAX_WORLD];           // the same buffer on every other context
    int n_peers;
    uint32_t pair[6 * SHADOW_MAX_LIGHTS];   // owned (light, face) pairs: face slabs p * L * L
    int n_pairs;
    uint32_t face_words;                    // L * L
    uint8_t* prev_dirty;                    // [n_pairs * face_words / MG_PUSH_CHUNK_WORDS] of this buffer
    uint32_t* pushed;                       // statistics counter (MgCtrl::pushed_chunks of this context)
    MgSignal sig;
};

__global__ void __launch_bounds__(256) k_push_faces(const MgPushParams P) {
    const uint32_t chunks_per_face = P.face_words / MG_PUSH_CHUNK_WORDS;
    const uint32_t total = chunks_per_face * (uint32_t)P.n_pairs;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ch < total; ch += warps) {
        const uint32_t pi = ch / chunks_per_face, c = ch - pi * chunks_per_face;
        const size_t off = (size_t)P.pair[pi] * P.face_words + (size_t)c * MG_PUSH_CHUNK_WORDS + (size_t)lane * 4;
        const uint4 v = *reinterpret_cast<const uint4*>(P.local + off);
        const bool dirty = __any_sync(0xffffffffu, (v.x & v.y & v.z & v.w) != 0xFFFFFFFFu);
        const bool was = P.prev_dirty[ch] != 0;
        if (dirty || was) {
            for (int q = 0; q < P.n_peers; q++) *reinterpret_cast<uint4*>(P.peer[q] + off) = v;
            if (lane == 0) atomicAdd(P.pushed, 1u);
        }
        __syncwarp();
        if (lane == 0 && dirty != was) P.prev_dirty[ch] = dirty ? 1 : 0;
    }
    mg_signal_tail(P.sig);
}

struct MgFillParams { uint32_t* buffer; uint32_t pair[6 * SHADOW_MAX_LIGHTS]; int n_pairs; uint32_t face_words; };

// clear of the owned cubemap faces only (the others are delivered by their owners)
__global__ void __launch_bounds__(256) k_fill_faces(const MgFillParams P) {
    const size_t per_face4 = P.face_words / 4;
    const size_t total = per_face4 * (size_t)P.n_pairs;
    const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pi = i / per_face4, k = i - pi * per_face4;
        reinterpret_cast<uint4*>(P.buffer + (size_t)P.pair[pi] * P.face_words)[k] = ones;
    }
}

// =====================================================================================================================
// fills (clEnqueueFillBuffer, engine.cpp:1615-1624) — 128-bit stores, grid-stride
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* __restrict__ p, size_t n, uint32_t v) {
    size_t n4 = n / 4;
    uint4* p4 = reinterpret_cast<uint4*>(p);
    const uint4 vv = make_uint4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) p4[i] = vv;
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) p[n4 * 4 + threadIdx.x] = v;
}

// =====================================================================================================================
// texture atlas: update_gpu_tex (cl2.cl:923-953), generate_mips (1071-1129), generate_mip_mips (1132-1189)
// =====================================================================================================================
struct AtlasView { const uchar4* texels; const uint32_t* nums; const uint32_t* sizes; uint32_t mip_start;
                   cudaTextureObject_t tex; };      // tex != 0: fetch texels through a point-sampled texture object over the same memory (rr_config-free A/B: RR_TEX_OBJECTS)

// read_tex_array, cl2.cl:785-821
__device__ __forceinline__ float4 read_tex_array(float cx, float cy, uint32_t tid, const uchar4* __restrict__ atlas, const uint32_t* __restrict__ nums,
                                                 const uint32_t* __restrict__ sizes) {
    int nv = (int)nums[tid];
    int slice = nv >> 16, which = nv & 0xFFFF;
    float width = (float)sizes[slice];
    float hnum = floorf(2048.f / width);
    float tnumy = floorf((float)which / hnum);
    float tnumx = fmaf(-tnumy, hnum, (float)which);
    cx = clampf(cx, 0.001f, width - 0.001f);
    cy = clampf(cy, 0.001f, width - 0.001f);
    int ix = (int)fmaf(tnumx, width, cx), iy = (int)fmaf(tnumy, width, cy);
    uchar4 t = atlas[(size_t)slice * RR_ATLAS_DIM * RR_ATLAS_DIM + (size_t)iy * RR_ATLAS_DIM + ix];
    return make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
}

// write_tex_array, cl2.cl:856-889
__device__ __forceinline__ void write_tex_array(uchar4 v, float cx, float cy, uint32_t tid, uchar4* __restrict__ atlas, const uint32_t* __restrict__ nums,
                                                const uint32_t* __restrict__ sizes) {
    int nv = (int)nums[tid];
    int slice = nv >> 16, which = nv & 0xFFFF;
    float width = (float)sizes[slice];
    float hnum = floorf(2048.f / width);
    float tnumy = floorf((float)which / hnum);
    float tnumx = fmaf(-tnumy, hnum, (float)which);
    float tx = tnumx * width, ty = tnumy * width;
    cx = fmodf(cx, width); cy = fmodf(cy, width);
    cx = clampf(cx, 0.001f, width - 0.001f);
    cy = clampf(cy, 0.001f, width - 0.001f);
    int ix = (int)(tx + cx), iy = (int)(ty + cy);
    atlas[(size_t)slice * RR_ATLAS_DIM * RR_ATLAS_DIM + (size_t)iy * RR_ATLAS_DIM + ix] = v;
}

// update_gpu_tex, cl2.cl:923-953, for texel (x, y) of a w x h source image
__device__ __forceinline__ void atlas_upload_texel(const uchar4* __restrict__ src, int w, int h, int x, int y, uint32_t tex_id, int flip, uchar4* __restrict__ atlas,
                                                   const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    if (x >= w || y >= h) return;
    uchar4 s = src[(size_t)y * w + x];
    // read_imagef on CL_UNORM_INT8 then *255 then truncating convert (cl2.cl:940-944), pinned as ((float)c/255.f)*255.f
    uchar4 o;
    o.x = (unsigned char)(uint32_t)(((float)s.x / 255.f) * 255.f);
    o.y = (unsigned char)(uint32_t)(((float)s.y / 255.f) * 255.f);
    o.z = (unsigned char)(uint32_t)(((float)s.z / 255.f) * 255.f);
    o.w = (unsigned char)(uint32_t)(((float)s.w / 255.f) * 255.f);
    int slice = (int)(nums[tex_id] >> 16);
    float width = (float)sizes[slice];
    int yy = y;
    if (flip) yy = (int)(width - (float)y);                         // cl2.cl:949-950
    write_tex_array(o, (float)x, (float)yy, tex_id, atlas, nums, sizes);
}

// generate_mips (cl2.cl:1071-1129) / generate_mip_mips (1132-1189) for work-item (x, y) of a gw x gh launch
__device__ __forceinline__ void atlas_mip_texel(uint32_t src_id, uint32_t dst_id, int gw, int gh, int x, int y, uchar4* __restrict__ atlas,
                                                const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    if (x >= gw || y >= gh) return;
    int slice = (int)(nums[src_id] >> 16);
    float width = (float)sizes[slice];
    if ((float)x >= width || (float)y >= width) return;
    float4 accum = make_float4(0, 0, 0, 0);
    float div = 0.f;
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
        for (int i = -1; i <= 1; i++) {
            const float g = (float)((2 - (i < 0 ? -i : i)) * (2 - (j < 0 ? -j : j)));     // {1,2,1;2,4,2;1,2,1}
            float4 col = read_tex_array((float)(x * 2 + i), (float)(y * 2 + j), src_id, atlas, nums, sizes);
            col.w /= 255.f;
            col.x *= col.w; col.y *= col.w; col.z *= col.w;
            accum = accum + col * g;
            div += g;
        }
    accum = accum / div;
    if (accum.w > 0.00000001f) { accum.x /= accum.w; accum.y /= accum.w; accum.z /= accum.w; }
    accum.w *= 255.f;
    int w2 = (int)(nums[dst_id] >> 16);
    float nwidth = (float)sizes[w2];
    float yx_x = ((float)(x * 2) / width) * nwidth, yx_y = ((float)(y * 2) / width) * nwidth;
    if (yx_x >= nwidth || yx_y >= nwidth) return;
    uchar4 o = make_uchar4((unsigned char)sat_u32(accum.x), (unsigned char)sat_u32(accum.y), (unsigned char)sat_u32(accum.z), (unsigned char)sat_u32(accum.w));
    write_tex_array(o, yx_x, yx_y, dst_id, atlas, nums, sizes);
}

__global__ void __launch_bounds__(256) k_atlas_upload(const uchar4* __restrict__ src, int w, int h, uint32_t tex_id, int flip, uchar4* __restrict__ atlas,
                                                      const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    atlas_upload_texel(src, w, h, blockIdx.x * 16 + (threadIdx.x & 15), blockIdx.y * 16 + (threadIdx.x >> 4), tex_id, flip, atlas, nums, sizes);
}

__global__ void __launch_bounds__(256) k_atlas_mip(uint32_t src_id, uint32_t dst_id, int gw, int gh, uchar4* __restrict__ atlas,
                                                   const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    atlas_mip_texel(src_id, dst_id, gw, gh, blockIdx.x * 16 + (threadIdx.x & 15), blockIdx.y * 16 + (threadIdx.x >> 4), atlas, nums, sizes);
}

// ---- batched atlas build (texture_context::alloc_gpu, texture_context.cpp:478-517, uploads every texture with its own write,
// kernel and four mip kernels, serially on one queue). Here all textures of a batch go through ONE launch per phase: the base
// upload, then each of the four mip levels (a level reads the one before it, so the levels stay separate launches).
// A job = one texture; the grid is the concatenation of the jobs' 16x16 tiles, `tile0` its exclusive prefix.
struct AtlasJob { unsigned long long src_off; uint32_t w, h, tex_id, tile0; };      // src_off: texel offset of the image in the staged batch

__device__ __forceinline__ bool atlas_job_of_block(const AtlasJob* __restrict__ jobs, uint32_t n_jobs, uint32_t block, AtlasJob& job, int& x, int& y) {
    uint32_t lo = 0, hi = n_jobs;                    // last job with tile0 <= block
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(&jobs[mid].tile0) <= block) lo = mid; else hi = mid; }
    job = jobs[lo];
    const uint32_t t = block - job.tile0, tw = (job.w + 15) / 16;
    if (t >= tw * ((job.h + 15) / 16)) return false;
    x = (int)((t % tw) * 16 + (threadIdx.x & 15));
    y = (int)((t / tw) * 16 + (threadIdx.x >> 4));
    return true;
}

__global__ void __launch_bounds__(256) k_atlas_upload_batch(const AtlasJob* __restrict__ jobs, uint32_t n_jobs, const uchar4* __restrict__ staged, int flip,
                                                            uchar4* __restrict__ atlas, const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    AtlasJob job; int x, y;
    if (!atlas_job_of_block(jobs, n_jobs, blockIdx.x, job, x, y)) return;
    atlas_upload_texel(staged + job.src_off, (int)job.w, (int)job.h, x, y, job.tex_id, flip, atlas, nums, sizes);
}

// level 0: generate_mips (base -> first mip); level 1..3: generate_mip_mips (mip level-1 -> mip level), texture.cpp:465-493
__global__ void __launch_bounds__(256) k_atlas_mip_batch(const AtlasJob* __restrict__ jobs, uint32_t n_jobs, int level, uint32_t mipmap_start,
                                                         uchar4* __restrict__ atlas, const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    AtlasJob job; int x, y;
    if (!atlas_job_of_block(jobs, n_jobs, blockIdx.x, job, x, y)) return;
    const uint32_t m0 = job.tex_id * RR_MIP_LEVELS + mipmap_start;
    atlas_mip_texel(level == 0 ? job.tex_id : m0 + (uint32_t)level - 1u, m0 + (uint32_t)level, (int)job.w, (int)job.h, x, y, atlas, nums, sizes);
}

// update_gpu_tex_colour, cl2.cl:955-984 (texture::update_gpu_texture_col, texture.cpp:445-463): a flat colour into the texture and
// its four mips. col is in 0..255 units; convert_uint4 pinned as the saturating truncation used everywhere else.
__global__ void __launch_bounds__(256) k_atlas_fill_colour(float4 col, uint32_t tex_id, uint32_t mipmap_start, int gw, int gh, uchar4* __restrict__ atlas,
                                                           const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x >= gw || y >= gh) return;
    const int slice = (int)(nums[tex_id] >> 16);
    const float width = (float)sizes[slice];
    if ((float)x >= width || (float)y >= width) return;
    const uchar4 ucol = make_uchar4((unsigned char)sat_u32(col.x), (unsigned char)sat_u32(col.y), (unsigned char)sat_u32(col.z), (unsigned char)sat_u32(col.w));
    write_tex_array(ucol, (float)x, (float)y, tex_id, atlas, nums, sizes);
    for (int i = 0; i < RR_MIP_LEVELS; i++) {
        const uint32_t mtexid = tex_id * RR_MIP_LEVELS + mipmap_start + (uint32_t)i;
        const float nwidth = (float)sizes[nums[mtexid] >> 16];
        write_tex_array(ucol, ((float)x / width) * nwidth, ((float)y / width) * nwidth, mtexid, atlas, nums, sizes);
    }
}

// generate_from_raw, cl2.cl:1006-1031 (texture::update_gpu_texture_mono, texture.cpp:554-584): one byte per texel, replicated to the
// four channels; the mips are not touched (the call is commented out there) and `flip` is ignored by the kernel
__global__ void __launch_bounds__(256) k_atlas_from_raw(const unsigned char* __restrict__ raw, int stride, int dimx, int dimy, uint32_t tex_id,
                                                        uchar4* __restrict__ atlas, const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x >= dimx || y >= dimy) return;
    const int width = (int)sizes[nums[tex_id] >> 16];
    if (x >= width || y >= width) return;
    const unsigned char v = raw[(size_t)y * stride + x];
    write_tex_array(make_uchar4(v, v, v, v), (float)x, (float)y, tex_id, atlas, nums, sizes);
}

// =====================================================================================================================
// k_shade == kernel3 (cl2.cl:5795-6408). One thread per pixel, 32x8 tiles (a warp = 32 consecutive pixels of a row so
// depth / id loads and the RGBA8 / normal / clear stores are full 128-byte lines).
// =====================================================================================================================
struct ShadeParams {
    const rr_triangle* tris; const rr_obj_desc* objs; const ObjLite* objlite; const LightLite* lightlite;
    const uint32_t* frags; const float4* cutdown; const uint32_t* n_frags;
    uint32_t* shade_list; uint32_t* shade_count;     // covered pixels of this frame (k_shade_pre -> k_shade)
    const uint32_t* depth; const uint32_t* ids;
    uint32_t* depth_next; uint32_t* ids_next;        // cleared for the next frame (to_clear, cl2.cl:5820)
    uchar4* rgba8; ushort2* normals;
    AtlasView atlas;
    const rr_light* lights; int n_lights;
    const uint32_t* shadow_dyn; const uint32_t* shadow_static;
    FaceTable faces;
    CamParams cam;
    float4 clear;
    int W, H, L;
    float fov;
    float ambient, ssao_rad, ssao_div, inv_mip_bias, shadow_bias, shadow_bias_max;
    int linear, no_ssao;
    int row0, row1;          // rows covered by the grid (band +- halo): cleared for the next frame
    int band_y0, band_y1;    // rows actually shaded
    const uint8_t* rowmask;  // interleaved bands: ROW_NEEDED rows are cleared, ROW_OWNED rows are shaded (nullptr: the ranges above)
};

// read_tex_array_all_precalculated, cl2.cl:823-851
__device__ __forceinline__ float4 read_tex_pre(float cx, float cy, int which, int slice, float width, const uchar4* __restrict__ atlas, cudaTextureObject_t tex = 0) {
    const float ihnum = width * (1.f / 2048);
    float tnumy = floorf((float)which * ihnum);
    float tnumx = (float)which - div_pow2(tnumy, ihnum);                // (ihnum = tile size / 2048)
    cx = clampf(cx, 0.001f, width - 0.001f);
    cy = clampf(cy, 0.001f, width - 0.001f);
    int ix = (int)fmaf(tnumx, width, cx), iy = (int)fmaf(tnumy, width, cy);
    // integer texels, blended in the kernel exactly as the reference does (SURVEY.md §7 hard part 7). Two fetch paths over the same
    // pitch-linear atlas: the read-only data path (ld.global.nc) or a point-sampled texture object (element read mode, unnormalised
    // coordinates) — same bytes either way; measured side by side in profiles/r2_texobj_ab.txt
    uchar4 t;
    if (tex) t = tex2D<uchar4>(tex, (float)ix + 0.5f, (float)(iy + slice * RR_ATLAS_DIM) + 0.5f);
    else t = __ldg(atlas + (size_t)slice * RR_ATLAS_DIM * RR_ATLAS_DIM + (size_t)iy * RR_ATLAS_DIM + ix);
    return make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
}

// return_bilinear_col_all_precalculated, cl2.cl:1426-1455
__device__ __forceinline__ float4 bilinear_pre(float mx, float my, int which, int slice, float width, const uchar4* __restrict__ atlas, cudaTextureObject_t tex) {
    float px = floorf(mx), py = floorf(my);
    float4 c0 = read_tex_pre(px, py, which, slice, width, atlas, tex);
    float4 c1 = read_tex_pre(px + 1, py, which, slice, width, atlas, tex);
    float4 c2 = read_tex_pre(px, py + 1, which, slice, width, atlas, tex);
    float4 c3 = read_tex_pre(px + 1, py + 1, which, slice, width, atlas, tex);
    float ux = mx - px, uy = my - py;
    float bx = 1.f - ux, by = 1.f - uy;
    return mad4(c0, bx, c1 * ux) * by + mad4(c2, bx, c3 * ux) * uy;
}

// texture_filter_diff, cl2.cl:1511-1573
__device__ __forceinline__ float4 texture_filter_diff(float2 vt, float2 vtdiff, int tid2, const AtlasView& av) {
    int nv = (int)__ldg(av.nums + tid2);
    int slice = nv >> 16;
    int tsize = (int)__ldg(av.sizes + slice);
    float vx = texture_mod1(vt.x), vy = texture_mod1(vt.y);
    float sx = vtdiff.x * (float)tsize, sy = vtdiff.y * (float)tsize;
    float worst = sqrtf(sx * sx + sy * sy);
    float worst_id_frac = fmaxf(log2_approx(worst), 0.f);
    float mip_lower = clampf(floorf(worst_id_frac), 0.f, (float)RR_MIP_LEVELS);
    float fmd = worst_id_frac - mip_lower;
    int tid_lower = mip_lower == 0 ? tid2 : (int)(mip_lower - 1 + (float)av.mip_start + (float)(tid2 * RR_MIP_LEVELS));
    int tid_higher = (int)(clampf(mip_lower, 0.f, RR_MIP_LEVELS - 1.f) + (float)av.mip_start + (float)(tid2 * RR_MIP_LEVELS));
    int lower_nv = (int)__ldg(av.nums + tid_lower), higher_nv = (int)__ldg(av.nums + tid_higher);
    int slice_lower = lower_nv >> 16, slice_higher = higher_nv >> 16;
    int which_lower = lower_nv & 0xFFFF, which_higher = higher_nv & 0xFFFF;
    float size_lower = (float)__ldg(av.sizes + slice_lower), size_higher = (float)__ldg(av.sizes + slice_higher);
    float4 col1 = bilinear_pre(vx * size_lower, vy * size_lower, which_lower, slice_lower, size_lower, av.texels, av.tex);
    float4 col2 = bilinear_pre(vx * size_higher, vy * size_higher, which_higher, slice_higher, size_higher, av.texels, av.tex);
    float4 fc = col1 + (col2 - col1) * fmd;
    return fc * (1.f / 255.f);
}

// generate_ssao, cl2.cl:2194-2260
__device__ __forceinline__ float generate_ssao(int sx, int sy, const uint32_t* __restrict__ depth_buffer, int W, int H, float fov, float ssao_rad, float ssao_div) {
    uint32_t seed1 = wang_hash((uint32_t)sx + (uint32_t)W * (uint32_t)H * (uint32_t)sy);
    uint32_t seed2 = rand_xorshift(seed1);
    float foffset = (float)seed2 * RR_INV_U32MAXF;
    float depth = ((float)__ldg(depth_buffer + sy * W + sx) * RR_INV_U32MAXF) * RR_DEPTH_FAR;
    float rad = ssao_rad + foffset / 2.f;
    float world_rad = rad * fov / depth;
    // the five thresholds depth + z are per-pixel constants; acc only ever adds 1.f, so an integer count is the same value
    const float t0 = depth + -2.f, t1 = depth + -1.f, t2 = depth + 0.f, t3 = depth + 1.f, t4 = depth + 2.f;
    int cnt = 0;
    for (int y = -2; y <= 2; y++) {
        const float oy = roundf((float)y * world_rad);
        const float wy = clampf((float)sy + oy, 1.f, (float)H - 2.f);
        const uint32_t* row = depth_buffer + ((int)wy) * W;
#pragma unroll
        for (int x = -2; x <= 2; x++) {
            const float ox = roundf((float)x * world_rad);
            const float wx = clampf((float)sx + ox, 1.f, (float)W - 2.f);
            const float d2 = ((float)__ldg(row + (int)wx) * RR_INV_U32MAXF) * RR_DEPTH_FAR;
            cnt += (int)(d2 > t0) + (int)(d2 > t1) + (int)(d2 > t2) + (int)(d2 > t3) + (int)(d2 > t4);
        }
    }
    float acc = (float)cnt;
#ifdef RR_SHADE_EXACT
    acc = div_pos(acc, 125.f);          // pow(samples*2+1, 3)
    return 1.f - (1.f - acc) / ssao_div;
#else
    acc = acc * (1.f / 125.f);          // the occlusion factor only scales the colour sums (colour-only arithmetic, rr_math.cuh)
    return 1.f - div_c(1.f - acc, ssao_div);
#endif
}

// generate_hard_occlusion, cl2.cl:2536-2701 (SMOOTH_SHADOWS)
__device__ __forceinline__ float hard_occlusion(float3 lpos, float3 normal, float3 position_to_light, const uint32_t* __restrict__ light_depth_buffer,
                                                int which_cubeface, float3 global_position, int shnum, const ShadeParams& P) {
    const int L = P.L;
    const float Lf = (float)L;
    position_to_light = normalize3(position_to_light);
    float3 local_pos = rot(global_position, lpos, P.faces.r[which_cubeface]);
    float3 pp = project(local_pos, Lf / 2.f, Lf / 2.f, Lf / 2.0f);
    float dpth = pp.z;
    const uint32_t* ldepth_map = light_depth_buffer + (size_t)(which_cubeface + shnum * 6) * L * L;
    pp.x = clampf(pp.x, 3.f, Lf - 4.f);
    pp.y = clampf(pp.y, 3.f, Lf - 4.f);
    int ipx = (int)pp.x, ipy = (int)pp.y;
    float acos_res = rational_acos(clampf(dot3(normal, position_to_light), 0.05f, 0.95f));
    float bias = P.shadow_bias * tanf(acos_res);
    bias = clampf(bias, 0.1f * P.shadow_bias, P.shadow_bias_max);
    float cnd[16];
#pragma unroll
    for (int y = -1; y <= 2; y++)
#pragma unroll
        for (int x = -1; x <= 2; x++) {
            float ldp1 = ((float)__ldg(ldepth_map + (ipy + y) * L + ipx + x) * RR_INV_U32MAXF) * RR_DEPTH_FAR;
            cnd[(y + 1) * 4 + x + 1] = dpth > ldp1 + bias ? 1.f : 0.f;
        }
#ifdef RR_SHADE_EXACT
    float shadow = 0.f;
#pragma unroll
    for (int y = -1; y <= 1; y++)
#pragma unroll
        for (int x = -1; x <= 1; x++)
            shadow += bilinear_interpolate(pp.x + 0.5f + (float)x, pp.y + 0.5f + (float)y, cnd[(y + 1) * 4 + x + 1], cnd[(y + 1) * 4 + x + 2],
                                           cnd[(y + 2) * 4 + x + 1], cnd[(y + 2) * 4 + x + 2]);
    return div_pos(shadow, 9.f);
#else
    // The nine 2x2 blends share their weights up to rounding (the fraction of pp + k is the fraction of pp), so their sum is a
    // separable 4x4 kernel with weights (1 - u, 1, 1, u) per axis. The 16 compare results above are exact; only the blend of
    // those 0 / 1 values is evaluated differently (colour-only arithmetic, rr_math.cuh).
    const float ux = pp.x - floorf(pp.x), uy = pp.y - floorf(pp.y), bx = 1.f - ux, by = 1.f - uy;
    float rowsum[4];
#pragma unroll
    for (int y = 0; y < 4; y++) rowsum[y] = fmaf(bx, cnd[y * 4], fmaf(ux, cnd[y * 4 + 3], cnd[y * 4 + 1] + cnd[y * 4 + 2]));
    return fmaf(by, rowsum[0], fmaf(uy, rowsum[3], rowsum[1] + rowsum[2])) * (1.f / 9.f);
#endif
}

__device__ __forceinline__ unsigned short to_ushort_sat(float v) {
    if (!(v > 0.f)) return 0;
    if (v >= 65535.f) return 65535;
    return (unsigned short)v;
}

__device__ __forceinline__ float4 vertex_col_f(uint32_t c) {        // cl2.cl:5676-5686
    return make_float4((float)(c >> 24), (float)((c >> 16) & 0xFF), (float)((c >> 8) & 0xFF), (float)(c & 0xFF)) / 255.f;
}

__device__ __forceinline__ unsigned char quant8(float c) { return (unsigned char)(clampf(c, 0.f, 1.f) * 255.f + 0.5f); }

// Shading of one covered pixel (everything of kernel3 after the depth == UINT_MAX early-out, cl2.cl:5864-6390).
__device__ __forceinline__ void shade_pixel(const ShadeParams& P, const int x, const int y) {
    const int W = P.W, H = P.H;
    const size_t px = (size_t)y * W + x;
    const uint32_t d = P.depth[px];
    const uint32_t idv = P.ids[px] - 1u;                             // the id image holds fragment index + 1 (k_shade_pre lists resolved pixels only)
    const uint32_t* rec = P.frags + (size_t)idv * RR_FRAG_WORDS;
    const uint32_t tri_global = __ldg(rec + 0), ctri = __ldg(rec + 2);
    const float rconst = __uint_as_float(__ldg(rec + 3));
    const int o_id = (int)__ldg(rec + 4);
    const rr_triangle* T = P.tris + tri_global;
    const rr_obj_desc* G = P.objs + o_id;
    const float4 pv0 = __ldg(reinterpret_cast<const float4*>(T->vertices[0].pos));
    const float4 pv1 = __ldg(reinterpret_cast<const float4*>(T->vertices[1].pos));
    const float4 pv2 = __ldg(reinterpret_cast<const float4*>(T->vertices[2].pos));
    const float4 nv0 = __ldg(reinterpret_cast<const float4*>(T->vertices[0].normal));
    const float4 nv1 = __ldg(reinterpret_cast<const float4*>(T->vertices[1].normal));
    const float4 nv2 = __ldg(reinterpret_cast<const float4*>(T->vertices[2].normal));
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(T->vertices[0].vt));   // vt.xy, object_id, vertex_col
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(T->vertices[1].vt));
    const float4 t2 = __ldg(reinterpret_cast<const float4*>(T->vertices[2].vt));
    const float2 vt1 = make_float2(t0.x, t0.y), vt2 = make_float2(t1.x, t1.y), vt3 = make_float2(t2.x, t2.y);
    const uint32_t vc0 = __float_as_uint(t0.w), vc1 = __float_as_uint(t1.w), vc2 = __float_as_uint(t2.w);
    const float4 Gpos4 = __ldg(reinterpret_cast<const float4*>(G->world_pos));
    const float Gscale = __ldg(&G->scale);
    const float3 Gpos = xyz(Gpos4);
    const float4 Gqn = __ldg(&P.objlite[o_id].nquat), Gqb = __ldg(&P.objlite[o_id].bquat);     // per-object constants, hoisted (k_objlite)
    const float3 p1 = xyz(pv0) * Gscale, p2 = xyz(pv1) * Gscale, p3 = xyz(pv2) * Gscale;
    const float fov = P.fov;
    const float3 zero3 = make_float3(0, 0, 0);

    const float ldepth = ((float)d * RR_INV_U32MAXF) * RR_DEPTH_FAR;  // cl2.cl:5897 (x / 2^32 == x * 2^-32 exactly)
    float3 local_position = make_float3((((float)x - W / 2.0f) * ldepth / fov), (((float)y - H / 2.0f) * ldepth / fov), ldepth);
    float3 global_position = back_rot(local_position, zero3, P.cam.rot);
    global_position = global_position + P.cam.pos;
    float3 object_local = rot_quat_n(global_position - Gpos, Gqb);
    float l1, l2, l3;
    get_barycentric(object_local, p1, p2, p3, l1, l2, l3);
    const float2 vt = mad2(vt1, l1, mad2(vt2, l2, vt3 * l3));
    float3 normal = mad3(xyz(nv0), l1, mad3(xyz(nv1), l2, xyz(nv2) * l3));
    normal = rot_quat_n(normal, Gqn);

    const float4 ct0 = __ldg(P.cutdown + (size_t)ctri * 3), ct1 = __ldg(P.cutdown + (size_t)ctri * 3 + 1), ct2 = __ldg(P.cutdown + (size_t)ctri * 3 + 2);
    float4 col;
    if (vc0 != 0) {
        col = mad4(vertex_col_f(vc0), l1, mad4(vertex_col_f(vc1), l2, vertex_col_f(vc2) * l3));        // cl2.cl:5933-5938
    } else {
        // get_vtdiff, cl2.cl:5691-5760
        const float fx = (float)x, fy = (float)y;
        float3 xr = make_float3(roundf(ct0.x), roundf(ct1.x), roundf(ct2.x));
        float3 yr = make_float3(roundf(ct0.y), roundf(ct1.y), roundf(ct2.y));
        float3 depths = make_float3(1.0f / ct0.z, 1.0f / ct1.z, 1.0f / ct2.z);
        float DA, DB, DC;
        interpolate_get_const(depths, xr, yr, rconst, DA, DB, DC);
        float dmx = fmaf(DA, fx + 1, fmaf(DB, fy, DC));
        float dmy = fmaf(DA, fx, fmaf(DB, fy + 1, DC));
        float3 lmx = make_float3((fx + 1 - W / 2.f) / fov, (fy - H / 2.f) / fov, 1.f);
        float3 lmy = make_float3((fx - W / 2.f) / fov, (fy + 1 - H / 2.f) / fov, 1.f);
        lmx = lmx / dmx;
        lmy = lmy / dmy;
        float3 gmx = rot_quat_n(back_rot(lmx, zero3, P.cam.rot) + P.cam.pos - Gpos, Gqb);
        float3 gmy = rot_quat_n(back_rot(lmy, zero3, P.cam.rot) + P.cam.pos - Gpos, Gqb);
        float lx1, lx2, lx3, ly1, ly2, ly3;
        get_barycentric(gmx, p1, p2, p3, lx1, lx2, lx3);
        get_barycentric(gmy, p1, p2, p3, ly1, ly2, ly3);
        float2 vtx = mad2(vt1, lx1, mad2(vt2, lx2, vt3 * lx3));
        float2 vty = mad2(vt1, ly1, mad2(vt2, ly2, vt3 * ly3));
        float2 vdx = vtx - vt, vdy = vty - vt;
        float2 vtdiff = make_float2(fabsf(vdx.x) + fabsf(vdy.x), fabsf(vdx.y) + fabsf(vdy.y)) * P.inv_mip_bias;
        col = texture_filter_diff(vt, vtdiff, (int)__ldg(&G->tid), P.atlas);
    }
    if (P.linear) { col.x = gamma_fwd(col.x); col.y = gamma_fwd(col.y); col.z = gamma_fwd(col.z); }

    const uint32_t seed1 = wang_hash((uint32_t)x + (uint32_t)y * (uint32_t)W * (uint32_t)H);    // cl2.cl:5965 (wraps mod 2^32)
    const uint32_t seed2 = rand_xorshift(seed1), seed3 = rand_xorshift(seed2), seed4 = rand_xorshift(seed3);
    float3 rseed = make_float3((float)seed2 * RR_INV_U32MAXF, (float)seed3 * RR_INV_U32MAXF, (float)seed4 * RR_INV_U32MAXF);
    rseed = make_float3((rseed.x - 0.5f) * 2, (rseed.y - 0.5f) * 2, (rseed.z - 0.5f) * 2);

    float3 diffuse_sum = zero3, specular_sum = zero3;
    float3 l2p = normalize3_c(P.cam.pos - global_position);
    const int feature_flag = __ldg(&G->feature_flag);
    const bool is_two_sided = (feature_flag & RR_FEATURE_TWO_SIDED) > 0;
    const bool receives_dynamic_shadows = !((feature_flag & RR_FEATURE_NO_DYNAMIC_SHADOWS) > 0);
    const bool is_front = front_facing(xyz(ct0), xyz(ct1), xyz(ct2));
    if (!is_front && is_two_sided) normal = -normal;
    const float ssao = P.no_ssao ? 1.f : generate_ssao(x, y, P.depth, W, H, fov, P.ssao_rad, P.ssao_div);
    normal = normalize3(normal);
#ifdef RR_SHADE_EXACT
    const float3 lighting_normal = normalize3(normal + rseed / 100.f);
#else
    const float3 lighting_normal = normalize3_c(mad3(rseed, 0.01f, normal));
#endif
    const float ambient = P.linear ? gamma_fwd(P.ambient) : P.ambient;
    const float Gdiffuse = __ldg(&G->diffuse), Gspecular = __ldg(&G->specular), Gspec_mult = __ldg(&G->spec_mult);

    int shnum = 0, static_num = 0;
    for (int i = 0; i < P.n_lights; i++) {                                                       // cl2.cl:6115-6278
        const rr_light* l = P.lights + i;
        const float4 lp4 = __ldg(reinterpret_cast<const float4*>(l->pos));
        const float4 lc4 = __ldg(reinterpret_cast<const float4*>(l->col));
        const uint32_t lshadow = __ldg(&l->shadow);
        const int lstatic = __ldg(&l->is_static);
        const float3 lpos = xyz(lp4);
        float3 point_to_light = lpos - global_position;
        float occlusion = 1.f;
        if (lshadow && lstatic) {
            int face = ret_cubeface(global_position, lpos);
            occlusion = 1.f - hard_occlusion(lpos, normal, point_to_light, P.shadow_static, face, global_position, static_num, P);
            static_num++;
        }
        float distance = length3(point_to_light);
        const float dr1 = (distance / __ldg(&l->radius)) + 1.f;
        float illumination = __ldg(&l->brightness) / (dr1 * dr1);                             // pow(x, 2)
        const float cutoff = 0.1f;
        illumination -= cutoff;
        illumination *= 1.f / (1.f - cutoff);
        if (illumination <= 0) continue;
        const float3 light_col = P.linear ? xyz(__ldg(&P.lightlite[i].col_linear)) : xyz(lc4);   // gamma of the light colour hoisted (k_lightlite)
        if (lshadow && receives_dynamic_shadows) {
            int face = ret_cubeface(global_position, lpos);
            float dyn = 1.f - hard_occlusion(lpos, normal, point_to_light, P.shadow_dyn, face, global_position, shnum, P);
            occlusion = fminf(occlusion, dyn);
            shnum++;
        }
        point_to_light = normalize3_c(point_to_light);
        float light = dot3(point_to_light, lighting_normal);
        light *= occlusion;
        light = fmaxf(light, 0.f);
        float diffuse = (1.0f - ambient) * light;
        diffuse_sum = diffuse_sum + light_col * ((diffuse + ambient) * __ldg(&l->diffuse) * Gdiffuse * illumination);
        float3 Hh = normalize3_c(l2p + point_to_light);
        const float kS = 0.4f;
        float ndh = fmaxf(0.f, dot3(normal, Hh));
        float ndv = fmaxf(0.f, dot3(normal, l2p));
        float vdh = fmaxf(0.f, dot3(l2p, Hh));
        float ndl = fmaxf(0.f, dot3(normal, point_to_light));
        const float F0 = 0.4f;
        const float omv = 1.f - vdh, omv2 = omv * omv;
        float fresnel = F0 + (1 - F0) * (omv2 * omv2 * omv);                                  // native_powr(x, 5)
        float rough = clampf(1.f - Gspecular, 0.001f, 10.f);
        float alpha = rational_acos_c(ndh);
        float microfacet = 0.8346f * exp_c(div_c(-alpha * alpha, rough * rough));
        const float sv_num = 2 * ndh;
#ifdef RR_SHADE_EXACT
        float sv = (sv_num == 0.f && vdh > 0.f) ? sv_num : sv_num / vdh;                       // 0 / positive == 0 (0 / 0 stays NaN, q17)
#else
        float sv = div_c(sv_num, vdh);
#endif
        float c1 = sv * ndv, c2 = sv * ndl;
        float geometric = fminf(fminf(1.f, c1), c2);
        const float spec_num = fresnel * microfacet * geometric, spec_den = RR_PI_F * ndv;
#ifdef RR_SHADE_EXACT
        float spec = (spec_num == 0.f && spec_den > 0.f) ? spec_num : spec_num / spec_den;   // 0 / positive == 0 without the division's slow path
#else
        float spec = div_c(spec_num, spec_den);
#endif
        specular_sum = specular_sum + light_col * (spec * kS * illumination) * Gspec_mult;
        specular_sum = make_float3(fmaxf(specular_sum.x, 0.f), fmaxf(specular_sum.y, 0.f), fmaxf(specular_sum.z, 0.f));
        specular_sum = specular_sum * occlusion;
    }
    specular_sum = specular_sum * ssao;
    diffuse_sum = diffuse_sum * ssao;
    const float rsc = 0.7f;
    float3 colclamp = make_float3(col.x, col.y, col.z) + zero3 + specular_sum * rsc;
    float3 fc = make_float3(fmaf(colclamp.x, diffuse_sum.x, specular_sum.x * (1.f - rsc)), fmaf(colclamp.y, diffuse_sum.y, specular_sum.y * (1.f - rsc)),
                            fmaf(colclamp.z, diffuse_sum.z, specular_sum.z * (1.f - rsc)));
    if (P.linear) fc = make_float3(gamma_inv_c(fc.x), gamma_inv_c(fc.y), gamma_inv_c(fc.z));
    P.rgba8[px] = make_uchar4(quant8(clampf(fc.x, 0.f, 1.f)), quant8(clampf(fc.y, 0.f, 1.f)), quant8(clampf(fc.z, 0.f, 1.f)), quant8(col.w));

    // encode_normal + float_to_short, cl2.cl:5588-5628
    float3 nn = normal;
    if (nn.x * nn.x + nn.y * nn.y < 0.0001f) nn.x = 0.01f;
    float ln = sqrtf(nn.x * nn.x + nn.y * nn.y);
    float k = sqrtf(fmaxf(nn.z * 0.5f + 0.5f, 0.f));
    float rx = (nn.x / ln) * k, ry = (nn.y / ln) * k;
    P.normals[px] = make_ushort2(to_ushort_sat(((rx + 1) / 2) * 65536 - 1), to_ushort_sat(((ry + 1) / 2) * 65536 - 1));
}

// k_shade_pre: the streaming part of kernel3 for every pixel of the 32x8 tile (clear the next frame's depth and id,
// cl2.cl:5820; write the clear colour where nothing was drawn, 5835-5862) and a device-wide list of the covered pixels:
// warp ballot + block prefix in shared memory, ONE atomicAdd per tile to reserve its range, coalesced index stores.
__global__ void __launch_bounds__(256) k_shade_pre(const ShadeParams P) {
    __shared__ int s_warp[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int y = P.row0 + blockIdx.y * 8 + warp;
    bool covered = false;
    uint32_t px = 0;
    if (x < P.W && y < P.row1) {
        px = (uint32_t)y * (uint32_t)P.W + (uint32_t)x;
        const uint32_t d = P.depth[px];
        const uint32_t idv = (d != 0xFFFFFFFFu) ? P.ids[px] : 0u;
        const uint32_t rbits = P.rowmask ? P.rowmask[y] : (uint32_t)(ROW_NEEDED | ROW_OWNED);
        if (rbits & ROW_NEEDED) {
            P.depth_next[px] = 0xFFFFFFFFu;
            P.ids_next[px] = 0u;                                           // id image of the next frame (atomicMax needs a clean slate)
        }
        if (y >= P.band_y0 && y < P.band_y1 && (rbits & ROW_OWNED)) {
            // idv holds fragment index + 1; 0: unresolved; > n_frags: stale id (buffers not swapped since an earlier frame) -
            // never index past this frame's records (q7)
            covered = d != 0xFFFFFFFFu && idv != 0u && idv <= P.n_frags[0];
            if (!covered) P.rgba8[px] = make_uchar4(quant8(P.clear.x), quant8(P.clear.y), quant8(P.clear.z), quant8(P.clear.w));
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, covered);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const int c = s_warp[w]; if (w < warp) off += c; total += c; }
    if (total == 0) return;
    if (tid == 0) s_base = atomicAdd(P.shade_count, (uint32_t)total);
    __syncthreads();
    if (covered) P.shade_list[s_base + off + __popc(m & ((1u << lane) - 1u))] = px;
}

// k_shade_pre4: same, four consecutive pixels per thread with 128-bit loads and stores (W % 4 == 0). Tile = 128 x 8.
// CLEAR_NEXT / CLEAR_RGBA: whether this launch also does the streaming stores (next frame's depth / id clear, clear colour of
// uncovered pixels). <true, true> is all of kernel3's streaming part in one launch; the frame driver normally runs the stores
// in k_clear_next on a side stream, beside the issue-bound setup kernels, and launches <false, false> here: a read-only pass
// over the depth buffer (ids only where something was drawn) that builds the list.
template <bool CLEAR_NEXT, bool CLEAR_RGBA>
__global__ void __launch_bounds__(256) k_shade_pre4(const ShadeParams P) {
    __shared__ int s_warp[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x = blockIdx.x * 128 + lane * 4;
    const int y = P.row0 + blockIdx.y * 8 + warp;
    unsigned cov = 0;                   // bit i: pixel x+i is covered
    uint32_t px = 0;
    if (x < P.W && y < P.row1) {
        px = (uint32_t)y * (uint32_t)P.W + (uint32_t)x;
        const uint32_t rbits = P.rowmask ? P.rowmask[y] : (uint32_t)(ROW_NEEDED | ROW_OWNED);
        uint4 d = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        uint4 idv = make_uint4(0, 0, 0, 0);
        if (rbits & ROW_NEEDED) {
            d = *reinterpret_cast<const uint4*>(P.depth + px);
            const bool any = (d.x & d.y & d.z & d.w) != 0xFFFFFFFFu;
            if (any) idv = *reinterpret_cast<const uint4*>(P.ids + px);
            if (CLEAR_NEXT) {
                *reinterpret_cast<uint4*>(P.depth_next + px) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                *reinterpret_cast<uint4*>(P.ids_next + px) = make_uint4(0, 0, 0, 0);
            }
        }
        if (y >= P.band_y0 && y < P.band_y1 && (rbits & ROW_OWNED)) {
            const uint32_t nfr = P.n_frags[0];
            // ids hold fragment index + 1 (0 = unresolved): covered <=> 1 <= id <= nfr
            cov = (unsigned)(d.x != 0xFFFFFFFFu && idv.x - 1u < nfr) | ((unsigned)(d.y != 0xFFFFFFFFu && idv.y - 1u < nfr) << 1) |
                  ((unsigned)(d.z != 0xFFFFFFFFu && idv.z - 1u < nfr) << 2) | ((unsigned)(d.w != 0xFFFFFFFFu && idv.w - 1u < nfr) << 3);
            if (CLEAR_RGBA && cov != 0xFu) {          // covered pixels are overwritten by k_shade afterwards
                const uchar4 c = make_uchar4(quant8(P.clear.x), quant8(P.clear.y), quant8(P.clear.z), quant8(P.clear.w));
                const uint32_t cw = *reinterpret_cast<const uint32_t*>(&c);
                *reinterpret_cast<uint4*>(P.rgba8 + px) = make_uint4(cw, cw, cw, cw);
            }
        }
    }
    const int mine = __popc(cov);
    int inc = mine;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, dlt); if (lane >= dlt) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const int c = s_warp[w]; if (w < warp) off += c; total += c; }
    if (total == 0) return;
    if (tid == 0) s_base = atomicAdd(P.shade_count, (uint32_t)total);
    __syncthreads();
    uint32_t at = s_base + off + inc - mine;
#pragma unroll
    for (int i = 0; i < 4; i++) if (cov & (1u << i)) P.shade_list[at++] = px + i;
}

// k_shade_list: the covered-pixel list alone (what k_shade_pre4<false, false> computes), shaped for a read-only pass: a thread takes
// four pixels in each of four rows, all four depth loads in flight at once, ids only where something was drawn; a tile is
// 128 x 32 pixels with one atomicAdd. Nine tenths of config 3's screen is empty: such threads do four loads and leave.
#define SL_ROWS 4
__global__ void __launch_bounds__(256) k_shade_list(const ShadeParams P) {
    __shared__ int s_warp[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x = blockIdx.x * 128 + lane * 4;
    const int y0 = P.row0 + (blockIdx.y * 8 + warp) * SL_ROWS;
    uint4 d[SL_ROWS];
    bool use[SL_ROWS];
    const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
    for (int r = 0; r < SL_ROWS; r++) {
        const int y = y0 + r;
        use[r] = x < P.W && y < P.row1 && y >= P.band_y0 && y < P.band_y1 && (!P.rowmask || (P.rowmask[y] & ROW_OWNED));
        d[r] = use[r] ? *reinterpret_cast<const uint4*>(P.depth + (size_t)y * P.W + x) : ones;
    }
    const uint32_t nfr = P.n_frags[0];
    unsigned cov = 0;                   // bit 4 r + i: pixel (x + i, y0 + r) is covered
#pragma unroll
    for (int r = 0; r < SL_ROWS; r++) {
        if ((d[r].x & d[r].y & d[r].z & d[r].w) == 0xFFFFFFFFu) continue;
        const uint4 idv = *reinterpret_cast<const uint4*>(P.ids + (size_t)(y0 + r) * P.W + x);
        // ids hold fragment index + 1 (0 = unresolved): covered <=> 1 <= id <= nfr
        cov |= ((unsigned)(d[r].x != 0xFFFFFFFFu && idv.x - 1u < nfr) | ((unsigned)(d[r].y != 0xFFFFFFFFu && idv.y - 1u < nfr) << 1) |
                ((unsigned)(d[r].z != 0xFFFFFFFFu && idv.z - 1u < nfr) << 2) | ((unsigned)(d[r].w != 0xFFFFFFFFu && idv.w - 1u < nfr) << 3)) << (4 * r);
    }
    const int mine = __popc(cov);
    int inc = mine;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, dlt); if (lane >= dlt) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const int c = s_warp[w]; if (w < warp) off += c; total += c; }
    if (total == 0) return;
    if (tid == 0) s_base = atomicAdd(P.shade_count, (uint32_t)total);
    __syncthreads();
    uint32_t at = s_base + off + inc - mine;
    for (unsigned m = cov; m; m &= m - 1u) {
        const int b = __ffs(m) - 1;
        P.shade_list[at++] = (uint32_t)(y0 + (b >> 2)) * (uint32_t)P.W + (uint32_t)(x + (b & 3));
    }
}

// k_clear_next: the streaming stores of kernel3 — the next frame's depth buffer back to UINT_MAX and its id image to 0 (to_clear,
// cl2.cl:5820), and the clear colour (cl2.cl:5835-5862; k_shade overwrites the covered pixels afterwards). None of them depends on
// anything this frame computes, so the frame driver runs them on a side stream at the start of the frame: 100 MB of stores at 4K that
// overlap the issue-bound setup kernels instead of sitting on the critical path in front of the shading. W % 4 == 0.
__global__ void __launch_bounds__(256) k_clear_next(const ShadeParams P, int with_rgba) {
    const uint32_t w4 = (uint32_t)P.W / 4u, total = (uint32_t)(P.row1 - P.row0) * w4;
    const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), zero = make_uint4(0, 0, 0, 0);
    const uchar4 c = make_uchar4(quant8(P.clear.x), quant8(P.clear.y), quant8(P.clear.z), quant8(P.clear.w));
    const uint32_t cw = *reinterpret_cast<const uint32_t*>(&c);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t r = i / w4;
        const int y = P.row0 + (int)r;
        const size_t px = (size_t)y * P.W + (size_t)(i - r * w4) * 4;
        const uint32_t rbits = P.rowmask ? P.rowmask[y] : (uint32_t)(ROW_NEEDED | ROW_OWNED);
        if (rbits & ROW_NEEDED) {
            *reinterpret_cast<uint4*>(P.depth_next + px) = ones;
            *reinterpret_cast<uint4*>(P.ids_next + px) = zero;
        }
        if (with_rgba && y >= P.band_y0 && y < P.band_y1 && (rbits & ROW_OWNED)) *reinterpret_cast<uint4*>(P.rgba8 + px) = make_uint4(cw, cw, cw, cw);
    }
}

// k_shade: the expensive part of kernel3 (~6000 instructions per covered pixel) over the compacted list — every warp
// is dense whatever the screen coverage looks like. Persistent grid, stride over the list.
__global__ void RR_LB_SHADE_ATTR k_shade(const ShadeParams P) {
    const uint32_t n = *P.shade_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t px = __ldg(P.shade_list + i);
        const int y = (int)(px / (uint32_t)P.W);
        shade_pixel(P, (int)(px - (uint32_t)y * (uint32_t)P.W), y);
    }
}

// =====================================================================================================================
// Dirty-tile read-back (rr_set_readback_tiles): the finished frame goes to the host as the tiles that can differ from what the host
// buffer already holds, instead of as 4·W·H bytes over PCIe every frame (33 MB at 4K: 0.59 ms at the 56 GB/s of the link, against
// 0.39 ms of rendering). A tile (32 x 4 pixels, 512 B) is sent when it holds a shaded pixel in this frame, or held one in the frame
// that was last written into the same host buffer; every other tile of the host buffer already holds the clear colour from an
// earlier copy. The host buffer ends up bit-identical to the device frame either way (tests/test_gpu_parity.py).
//   k_tile_mark : main stream, behind k_shade — the tiles of the covered-pixel list
//   k_tile_copy : copy stream — a warp stores a tile into the page-locked, mapped host buffer with 16-byte stores (four 128-byte
//                 row segments per tile) and remembers the tile's state for the next use of this ring slot. FEW CTAs: a store
//                 into host memory waits for the PCIe link, and an SM whose store queue is full of them holds up every other warp
//                 it runs — with the kernel on all 148 SMs the frames beside it slowed down by the kernel's whole duration
//                 (0.486 ms per frame end to end; 4 CTAs: 0.417). The host sizes the grid from the number of tiles the last
//                 completed copy sent; a frame whose buffer must be rewritten entirely (first use, new clear colour, more than half
//                 of the tiles) goes through the copy engine as before and the kernel only does the bookkeeping (host == nullptr).
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_tile_mark(const uint32_t* __restrict__ shade_list, const uint32_t* __restrict__ shade_count, int W, int tiles_x,
                                                   uint8_t* __restrict__ now) {
    const uint32_t n = *shade_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t px = __ldg(shade_list + i);
        const uint32_t y = px / (uint32_t)W, x = px - y * (uint32_t)W;
        now[(y / TILE_H) * (uint32_t)tiles_x + x / TILE_W] = 1;
    }
}

__global__ void __launch_bounds__(256) k_tile_copy(const uchar4* __restrict__ frame, uchar4* __restrict__ host, int W, int H, int tiles_x, uint32_t n_tiles,
                                                   uint8_t* __restrict__ now, uint8_t* __restrict__ prev, int all, uint32_t* __restrict__ sent, uint32_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    uint32_t mine = 0;
    // a warp takes 32 consecutive tiles: lane l owns the flags of tile t0 + l (one coalesced load each), then the warp copies the
    // tiles that need it one after the other, lane = (row of the tile, 16-byte column)
    for (uint32_t t0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; t0 < n_tiles; t0 += warps * 32u) {
        const uint32_t t = t0 + (uint32_t)lane;
        bool need = false;
        if (t < n_tiles) {
            const uint8_t d = now[t], p = prev[t];
            prev[t] = d; now[t] = 0;
            need = all || d || p;
        }
        unsigned m = __ballot_sync(0xffffffffu, need);
        mine += (uint32_t)__popc(m);
        if (!host) continue;                                    // flags only: the pixels of this frame travel as one DMA copy
        for (; m; m &= m - 1u) {
            const uint32_t tile = t0 + (uint32_t)(__ffs(m) - 1);
            const int ty = (int)(tile / (uint32_t)tiles_x), tx = (int)(tile - (uint32_t)ty * (uint32_t)tiles_x);
            const int y = ty * TILE_H + (lane >> 3), x = tx * TILE_W + (lane & 7) * 4;
            if (y < H && x < W) {
                const size_t o = (size_t)y * W + x;
                *reinterpret_cast<uint4*>(host + o) = *reinterpret_cast<const uint4*>(frame + o);      // W % 4 == 0: never straddles the row end
            }
        }
    }
    if (lane == 0 && mine) { if (host) atomicAdd(sent, mine); atomicAdd(cnt, mine); }
}

// =====================================================================================================================
// k_pseudo_aa == do_pseudo_aa (cl2.cl:6437-6657): edge smoothing on the G-buffer after kernel3. A pixel whose 3x3
// neighbourhood has exactly one horizontal and one vertical neighbour (plus at least one diagonal) on the other side of a
// depth step (> 100 units) or of a normal crease (> 20 degrees) becomes 0.65 * mean(same side) + 0.35 * mean(other side).
// The reference runs it in place on one image (engine.cpp:1854-1856), so its result depends on scheduling; here every read
// sees kernel3's frame (`in`) and every pixel is written to `out`. Streaming: 12 B/pixel in, 4 B/pixel out.
// =====================================================================================================================
__device__ __forceinline__ float3 decode_normal(ushort2 s) {            // short_to_float + decode_normal, cl2.cl:5598-5647
    float vx = (float)s.x, vy = (float)s.y;
    vx = vx / 65535.f; vy = vy / 65535.f;
    vx = vx * 2.f; vy = vy * 2.f;
    vx = vx - 1.f; vy = vy - 1.f;
    const float d = vx * vx + vy * vy;
    const float z = d * 2.f - 1.f;
    const float l = sqrtf(d);
    const float k = sqrtf(fmaxf(1.f - z * z, 0.f));
    return make_float3((vx / l) * k, (vy / l) * k, z);
}

__global__ void __launch_bounds__(256) k_pseudo_aa(const uchar4* __restrict__ in, uchar4* __restrict__ out, const uint32_t* __restrict__ depth_buffer,
                                                   const ushort2* __restrict__ normals, int W, int H, float cosrad) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t px = (size_t)y * W + x;
    const uchar4 mine = in[px];
    uchar4 res = mine;
    const uint32_t my_depth_raw = depth_buffer[px];
    if (x >= 1 && y >= 1 && x < W - 1 && y < H - 1 && my_depth_raw != 0xFFFFFFFFu) {
        const float3 my_normal = normalize3(decode_normal(normals[px]));
        const float my_depth = ((float)my_depth_raw * RR_INV_U32MAXF) * RR_DEPTH_FAR;      // idcalc(x / mulint)
        int num_x[2] = {0, 0}, num_y[2] = {0, 0}, num_corner[2] = {0, 0};
        float3 my_accum[2] = {make_float3(0, 0, 0), make_float3(0, 0, 0)}, their_accum[2] = {make_float3(0, 0, 0), make_float3(0, 0, 0)};
        float my_samples[2] = {0.f, 0.f}, their_samples[2] = {0.f, 0.f};
#pragma unroll
        for (int j = -1; j < 2; j++) {
#pragma unroll
            for (int i = -1; i < 2; i++) {
                if (i == 0 && j == 0) continue;
                const size_t q = (size_t)(y + j) * W + (x + i);
                const float depth = ((float)depth_buffer[q] * RR_INV_U32MAXF) * RR_DEPTH_FAR;
                const float3 found_normal = decode_normal(normals[q]);
                const uchar4 cv = in[q];
                const float3 val = make_float3((float)cv.x / 255.f, (float)cv.y / 255.f, (float)cv.z / 255.f);
                const bool tests[2] = {fabsf(depth - my_depth) > 100.f, dot3(my_normal, found_normal) < cosrad};
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    if (tests[kk]) {
                        if (i == j || i == -j) num_corner[kk]++;
                        else if (i == 1 || i == -1) num_x[kk]++;
                        else num_y[kk]++;
                        their_accum[kk] = their_accum[kk] + val;
                        their_samples[kk] += 1.f;
                    } else {
                        my_accum[kk] = my_accum[kk] + val;
                        my_samples[kk] += 1.f;
                    }
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
            if (num_x[kk] == 1 && num_y[kk] == 1 && num_corner[kk] >= 1) {
                const float3 ma = my_accum[kk] / my_samples[kk], ta = their_accum[kk] / their_samples[kk];
                const float3 accum = ma * 0.65f + ta * 0.35f;
                res = make_uchar4(quant8(accum.x), quant8(accum.y), quant8(accum.z), 255);
                break;
            }
        }
    }
    out[px] = res;
}

// =====================================================================================================================
// do_motion_blur (cl2.cl:6714-6860) and screenspace_godrays (cl2.cl:1792-1917): the other post passes on the G-buffer.
// Both read the frame kernel3 produced (`in`) through the CLK_FILTER_LINEAR formula of the OpenCL specification (§8.2,
// unnormalised coordinates: i0 = floor(u - 0.5), a = frac(u - 0.5), four taps, clamped to the image) and write every pixel
// to a second target (`out`), which then becomes the colour target — the reference runs godrays in place on one image
// (engine.cpp:1471-1476), motion blur from gl_screen[1] into gl_screen[0] (engine.cpp:1520-1521). Streaming gather kernels,
// HBM / L2 bound: 4 B/pixel in and out plus the taps.
// =====================================================================================================================
__device__ __forceinline__ float4 sample_linear_rgba8(const uchar4* __restrict__ img, int W, int H, float u, float v) {
    const float fu = u - 0.5f, fv = v - 0.5f;
    const float i0f = floorf(fu), j0f = floorf(fv);
    const float a = fu - i0f, b = fv - j0f;
    const int i0 = (int)clampf(i0f, 0.f, (float)(W - 1)), i1 = (int)clampf(i0f + 1.f, 0.f, (float)(W - 1));
    const int j0 = (int)clampf(j0f, 0.f, (float)(H - 1)), j1 = (int)clampf(j0f + 1.f, 0.f, (float)(H - 1));
    auto T = [&](int i, int j) { const uchar4 t = __ldg(img + (size_t)j * W + i); return make_float4((float)t.x / 255.f, (float)t.y / 255.f, (float)t.z / 255.f, (float)t.w / 255.f); };
    const float4 t00 = T(i0, j0), t10 = T(i1, j0), t01 = T(i0, j1), t11 = T(i1, j1);
    return (t00 * (1.f - a) + t10 * a) * (1.f - b) + (t01 * (1.f - a) + t11 * a) * b;
}
__device__ __forceinline__ uchar4 quant_rgba8(float4 c) { return make_uchar4(quant8(c.x), quant8(c.y), quant8(c.z), quant8(c.w)); }

struct MotionBlurParams {
    const uchar4* in; uchar4* out;
    const uint32_t* depth; const uint32_t* ids; const uint32_t* frags; const uint32_t* n_frags;
    const rr_obj_desc* objs; uint32_t n_objs; uint8_t* seen;
    CamParams cam, cam_old;
    int W, H; float fov, icut, strength, camera_contribution;
    uint32_t frame_id;
};

__global__ void __launch_bounds__(256) k_motion_blur(const MotionBlurParams P) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.W || y >= P.H) return;
    const int W = P.W, H = P.H;
    const float Wf = (float)W, Hf = (float)H, fov = P.fov;
    const size_t px = (size_t)y * W + x;
    const uchar4 mine = P.in[px];
    const uint32_t dbuf_val = P.depth[px];
    const uint32_t idv = dbuf_val != 0xFFFFFFFFu ? P.ids[px] : 0u;
    uint32_t o_id = 0xFFFFFFFFu;
    if (idv != 0u && idv <= P.n_frags[0]) o_id = __ldg(P.frags + (size_t)(idv - 1u) * RR_FRAG_WORDS + 4);
    if (o_id >= P.n_objs) { P.out[px] = mine; return; }                  // nothing drawn here (or an unresolved id, q7): the frame's colour stays
    const rr_obj_desc* G = P.objs + o_id;
    const float actual_depth = ((float)dbuf_val * RR_INV_U32MAXF) * RR_DEPTH_FAR;
    const float3 local_position = make_float3((((float)x - Wf / 2.0f) * actual_depth / fov), (((float)y - Hf / 2.0f) * actual_depth / fov), actual_depth);
    float3 global_position = back_rot(local_position, make_float3(0, 0, 0), P.cam.rot);
    global_position = global_position + P.cam.pos;
    float3 object_local = global_position - make_float3(G->world_pos[0], G->world_pos[1], G->world_pos[2]);
    object_local = rot_quat_n(object_local, back_quat(make_float4(G->world_rot_quat[0], G->world_rot_quat[1], G->world_rot_quat[2], G->world_rot_quat[3])));
    const bool even = (P.frame_id & 1u) == 0u;
    const float* owp = even ? G->old_world_pos_1 : G->old_world_pos_2;
    const float* owq = even ? G->old_world_rot_quat_1 : G->old_world_rot_quat_2;
    P.seen[o_id] = 1;                                                    // its history advances after the pass (k_motion_history)
    float3 last_frame_pos = rot_quat(object_local, make_float4(owq[0], owq[1], owq[2], owq[3]));
    last_frame_pos = last_frame_pos + make_float3(owp[0], owp[1], owp[2]);
    float3 last_frame_no_camera = rot(last_frame_pos, P.cam.pos, P.cam.rot);
    last_frame_pos = rot(last_frame_pos, P.cam_old.pos, P.cam_old.rot);
    last_frame_no_camera = project(last_frame_no_camera, Wf / 2.f, Hf / 2.f, fov);
    last_frame_pos = project(last_frame_pos, Wf / 2.f, Hf / 2.f, fov);
    if (last_frame_pos.z < P.icut) { P.out[px] = quant_rgba8(sample_linear_rgba8(P.in, W, H, (float)x + 0.5f, (float)y + 0.5f)); return; }
    const float2 current_screen_pos = make_float2((float)x, (float)y);
    float2 to_me_vector = current_screen_pos - make_float2(last_frame_pos.x, last_frame_pos.y);
    const float2 to_me_nocamera = current_screen_pos - make_float2(last_frame_no_camera.x, last_frame_no_camera.y);
    to_me_vector = to_me_vector * P.camera_contribution + to_me_nocamera * (1.f - P.camera_contribution);
    to_me_vector = to_me_vector * P.strength;
    int n = (int)(fmaxf(fabsf(to_me_vector.x), fabsf(to_me_vector.y)) + 1);
    const int bound = 50;
    if (n > bound) {
        to_me_vector = make_float2(to_me_vector.x / (float)n, to_me_vector.y / (float)n);
        to_me_vector = to_me_vector * (float)bound;
        n = bound;
    }
    float2 diff = make_float2(0, 0);
    if (n != 0) diff = make_float2(to_me_vector.x / (float)n, to_me_vector.y / (float)n);
    float2 current = current_screen_pos - make_float2(to_me_vector.x / 2.f, to_me_vector.y / 2.f);
    float4 accum = make_float4(0, 0, 0, 0);
    float fcount = 0;
    for (int i = 0; i < n; i++, current = current + diff) {
        if (current.x < 0 || current.x >= Wf || current.y < 0 || current.y >= Hf) continue;
        accum = accum + sample_linear_rgba8(P.in, W, H, current.x + 0.5f, current.y + 0.5f) * 1.f;
        fcount += 1.f;
    }
    if (fcount != 0) accum = accum / fcount;
    P.out[px] = quant_rgba8(accum);
}

// the history half of do_motion_blur (cl2.cl:6768-6784): objects seen in at least one pixel store their current placement in
// the slot the NEXT frame reads
__global__ void __launch_bounds__(128) k_motion_history(rr_obj_desc* __restrict__ objs, uint32_t n_objs, uint8_t* __restrict__ seen, uint32_t frame_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_objs || !seen[i]) return;
    seen[i] = 0;
    rr_obj_desc& G = objs[i];
    float* dp = (frame_id & 1u) == 0u ? G.old_world_pos_2 : G.old_world_pos_1;
    float* dq = (frame_id & 1u) == 0u ? G.old_world_rot_quat_2 : G.old_world_rot_quat_1;
    dp[0] = G.world_pos[0]; dp[1] = G.world_pos[1]; dp[2] = G.world_pos[2];
    dq[0] = G.world_rot_quat[0]; dq[1] = G.world_rot_quat[1]; dq[2] = G.world_rot_quat[2]; dq[3] = G.world_rot_quat[3];
}

struct GodrayParams {
    const uchar4* in; uchar4* out; const uint32_t* depth;
    const rr_light* lights; int n_lights;
    CamParams cam; int W, H; float fov;
};

__global__ void __launch_bounds__(256) k_godrays(const GodrayParams P) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.W || y >= P.H) return;
    const int W = P.W, H = P.H;
    const float Wf = (float)W, Hf = (float)H;
    const size_t px = (size_t)y * W + x;
    const float samples = 80.f;
    const uint32_t my_depth = P.depth[px];
    float4 my_col = sample_linear_rgba8(P.in, W, H, (float)x - 0.25f, (float)y - 0.25f);
    const float decay_factor = 0.97f, weight = 0.01f, max_length = 400.f;
    for (int i = 0; i < P.n_lights; i++) {
        const rr_light* l = P.lights + i;
        const float ray_intensity = __ldg(&l->godray_intensity);
        if (ray_intensity <= 0) continue;
        float idecay = 1.f;
        float3 iter_col = make_float3(0, 0, 0);
        float3 slpos = rot(make_float3(__ldg(&l->pos[0]), __ldg(&l->pos[1]), __ldg(&l->pos[2])), P.cam.pos, P.cam.rot);
        slpos = project(slpos, Wf / 2.f, Hf / 2.f, P.fov);
        float3 current_pos = make_float3((float)x, (float)y, ((float)my_depth * RR_INV_U32MAXF) * RR_DEPTH_FAR);
        float3 destination_pos = slpos;
        const float3 original = current_pos;
        if (slpos.z < 0) destination_pos = (current_pos - destination_pos) + current_pos;
        const float vx = fabsf(current_pos.x - destination_pos.x), vy = fabsf(current_pos.y - destination_pos.y);
        const float mnum = vx > vy ? vx : vy;
        float3 dir = (destination_pos - current_pos) / mnum;
        dir = dir * (max_length / samples);
        const float3 col = make_float3(__ldg(&l->col[0]), __ldg(&l->col[1]), __ldg(&l->col[2]));
        for (int j = 0; (float)j < mnum && (float)j < samples; j++) {
            if (current_pos.x < 0 || current_pos.y < 0 || current_pos.x >= Wf - 1 || current_pos.y >= Hf - 1) continue;
            const uint32_t cdepth = __ldg(P.depth + (size_t)((int)current_pos.y) * W + (int)current_pos.x);
            const float fdepth = ((float)cdepth * RR_INV_U32MAXF) * RR_DEPTH_FAR;
            float3 val = make_float3(0, 0, 0);
            if (fdepth < original.z - 5 && cdepth != 0xFFFFFFFFu) {
                idecay *= 0.9f;
                val = col * ray_intensity;
            }
            val = val * idecay * weight;
            iter_col = iter_col + val;
            idecay *= decay_factor;
            current_pos = current_pos + dir;
        }
        my_col.x += iter_col.x; my_col.y += iter_col.y; my_col.z += iter_col.z;
    }
    my_col.w = 1;
    P.out[px] = quant_rgba8(my_col * 0.99f);
}

__global__ void k_lightlite(const rr_light* __restrict__ lights, uint32_t n, LightLite* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i].col_linear = make_float4(gamma_fwd(lights[i].col[0]), gamma_fwd(lights[i].col[1]), gamma_fwd(lights[i].col[2]), lights[i].col[3]);
}

// =====================================================================================================================
// roofline micro-benchmarks
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_bench_atomic_min(uint32_t* buf, uint32_t n_words_mask, uint32_t iters) {
    uint32_t s = wang_hash(blockIdx.x * blockDim.x + threadIdx.x + 1u);
    for (uint32_t i = 0; i < iters; i++) {
        s = rand_xorshift(s);
        red_min_u32(buf + (s & n_words_mask), s >> 3);
    }
}

__global__ void __launch_bounds__(256) k_copy_u32(const uint4* __restrict__ src4, uint4* __restrict__ dst4, size_t n4,
                                                  const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst4[i] = src4[i];
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = src[n4 * 4 + threadIdx.x];
}

__global__ void __launch_bounds__(256) k_bench_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace rr

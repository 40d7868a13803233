// rr_kernels.cuh — sm_100a kernels of the raster path (setup/bin, depth, id resolve, shade, shadow passes, atlas).
// Design notes live in DESIGN.md; every kernel names the cl2.cl kernel whose results it must reproduce.
#pragma once
#include "rr_math.cuh"
#include "../../include/rr.h"

namespace rr {

// ---- counters (one uint32 array per context) ----------------------------------------------------------------------
enum {
    CTR_NCUT = 0,      // id_cutdown_tris of the main pass
    CTR_NFRAG = 1,     // id_buffer_atomc of the main pass
    CTR_OVERFLOW = 2,  // bit0 fragments, bit1 cutdown, bit2 look-back watchdog
    CTR_TICKET = 3,    // block ticket of the single-pass scan
    CTR_S_NCUT = 4,    // shadow pass counters (reset per pass)
    CTR_S_NFRAG = 5,
    CTR_S_TOTAL = 6,   // running total of shadow fragments in this frame (statistics)
    CTR_COUNT = 8
};

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
                                                float4* __restrict__ pa, float4* __restrict__ pb, float2* __restrict__ pc) {
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
}

// Per-object data the setup kernels need, 48 B instead of the 144 B descriptor: rebuilt when descriptors change.
struct ObjLite {
    float4 pos_scale;     // world_pos.xyz, scale
    float4 nquat;         // fast_normalize(world_rot_quat)  (rot_quat() normalises on every call, cl2.cl:352)
    int32_t feature_flag;
    int32_t _pad[3];
};

__global__ void __launch_bounds__(128) k_objlite(const rr_obj_desc* __restrict__ objs, uint32_t n, ObjLite* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rr_obj_desc& G = objs[i];
    ObjLite o;
    o.pos_scale = make_float4(G.world_pos[0], G.world_pos[1], G.world_pos[2], G.scale);
    o.nquat = normalize4(make_float4(G.world_rot_quat[0], G.world_rot_quat[1], G.world_rot_quat[2], G.world_rot_quat[3]));
    o.feature_flag = G.feature_flag;
    o._pad[0] = o._pad[1] = o._pad[2] = 0;
    out[i] = o;
}

// One clipped + projected triangle after culling. keep == false -> no storage written, no fragments.
struct SubTri { float3 p0, p1, p2; float rconst; int n_frag; bool keep; };

// cull + bbox + fragment count, cl2.cl:4352-4377 (main) / 4571-4597 (shadow)
__device__ __forceinline__ void classify(SubTri& s, bool two_sided, float ewidth, float eheight, float op_size) {
    bool valid = two_sided || front_facing(s.p0, s.p1, s.p2);
    bool cond = (s.p0.x < 0 && s.p1.x < 0 && s.p2.x < 0) || (s.p0.x >= ewidth && s.p1.x >= ewidth && s.p2.x >= ewidth) ||
                (s.p0.y < 0 && s.p1.y < 0 && s.p2.y < 0) || (s.p0.y >= eheight && s.p1.y >= eheight && s.p2.y >= eheight);
    s.keep = valid && !cond;
    s.n_frag = 0;
    s.rconst = 0.f;
    if (!s.keep) return;
    float3 xr = make_float3(roundf(s.p0.x), roundf(s.p1.x), roundf(s.p2.x));
    float3 yr = make_float3(roundf(s.p0.y), roundf(s.p1.y), roundf(s.p2.y));
    s.rconst = calc_rconstant_v(xr, yr);
    float4 mm = calc_min_max(xr, yr, ewidth, eheight);
    float area = (mm.y - mm.x) * (mm.w - mm.z);
    s.n_frag = (int)ceilf(area / op_size);
}

// object -> world -> camera for the three vertices, then clip + project. cl2.cl:700-729. Returns num (0/1/2).
__device__ __forceinline__ int transform_clip_project(float3 v0, float3 v1, float3 v2, const ObjLite& G, float3 cam_pos, const RotSC& cam_rot,
                                                      float icut, float half_w, float half_h, float fovc, SubTri (&out)[2]) {
    const float scale = G.pos_scale.w;
    const float3 gpos = make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z);
    float3 pr[3];
    pr[0] = rot(rot_quat_n(v0 * scale, G.nquat) + gpos, cam_pos, cam_rot);
    pr[1] = rot(rot_quat_n(v1 * scale, G.nquat) + gpos, cam_pos, cam_rot);
    pr[2] = rot(rot_quat_n(v2 * scale, G.nquat) + gpos, cam_pos, cam_rot);
    float3 cl[2][3];
    int num = clip_near(pr, icut, cl);
    for (int i = 0; i < num; i++) {
        out[i].p0 = project(cl[i][0], half_w, half_h, fovc);
        out[i].p1 = project(cl[i][1], half_w, half_h, fovc);
        out[i].p2 = project(cl[i][2], half_w, half_h, fovc);
    }
    return num;
}

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

struct SetupMainParams {
    const float4* pa; const float4* pb; const float2* pc;
    const ObjLite* objs;
    uint32_t n_tris;
    CamParams cam;
    float width, height, fov, icut;
    uint32_t* frags; uint32_t cap_frags;
    float4* cutdown; uint32_t cap_cut;
    uint32_t* counters;
    unsigned long long* lookback;
};

__global__ void __launch_bounds__(SETUP_THREADS) k_setup_main(const SetupMainParams P) {
    __shared__ uint32_t s_bid;
    __shared__ uint32_t s_warp_c[SETUP_THREADS / 32], s_warp_f[SETUP_THREADS / 32];
    __shared__ uint32_t s_base_c, s_base_f, s_tot_f;
    __shared__ uint32_t s_fexcl[2 * SETUP_THREADS];      // exclusive fragment offset of slot (2*thread + i) inside the block
    __shared__ uint32_t s_cid[2 * SETUP_THREADS];
    __shared__ float s_rconst[2 * SETUP_THREADS];
    __shared__ uint32_t s_oid[SETUP_THREADS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(&P.counters[CTR_TICKET], 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t tri = bid * SETUP_THREADS + tid;

    SubTri st[2];
    int num = 0;
    uint32_t oid = 0;
    st[0].keep = st[1].keep = false; st[0].n_frag = st[1].n_frag = 0; st[0].rconst = st[1].rconst = 0.f;
    if (tri < P.n_tris) {
        const float4 a = __ldg(P.pa + tri), b = __ldg(P.pb + tri);
        const float2 c = __ldg(P.pc + tri);
        oid = __float_as_uint(c.y);
        const ObjLite G = P.objs[oid];
        const float3 gpos = make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z);
        if (!(length3(gpos - P.cam.pos) > RR_DEPTH_FAR)) {                      // cl2.cl:4321
            num = transform_clip_project(make_float3(a.x, a.y, a.z), make_float3(a.w, b.x, b.y), make_float3(b.z, b.w, c.x), G, P.cam.pos,
                                         P.cam.rot, P.icut, P.width / 2.f, P.height / 2.f, P.fov, st);
            const bool two_sided = (G.feature_flag & RR_FEATURE_TWO_SIDED) > 0;
            for (int i = 0; i < num; i++) classify(st[i], two_sided, P.width, P.height, (float)RR_OP_SIZE);
        }
    }
    const uint32_t my_c = (uint32_t)num;                                        // slots are taken before culling (cl2.cl:4342)
    const uint32_t my_f0 = (num > 0) ? (uint32_t)st[0].n_frag : 0u;
    const uint32_t my_f1 = (num > 1) ? (uint32_t)st[1].n_frag : 0u;
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
        volatile unsigned long long* desc = P.lookback;
        uint32_t base_c = 0, base_f = 0;
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
                        if (++watchdog > (1u << 26)) { atomicOr(&P.counters[CTR_OVERFLOW], 4u); v = lb_pack(2, 0, 0); break; }
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
        if (lane == 0) {
            s_base_c = base_c; s_base_f = base_f; s_tot_f = tot_f;
            if (bid == gridDim.x - 1) { P.counters[CTR_NCUT] = base_c + tot_c; P.counters[CTR_NFRAG] = base_f + tot_f; }
        }
    }
    // stage per-slot data for the cooperative record write
    const uint32_t cid0 = ex_c;    // block-relative; global added below
    s_fexcl[2 * tid] = ex_f;
    s_fexcl[2 * tid + 1] = ex_f + my_f0;
    s_rconst[2 * tid] = st[0].rconst;
    s_rconst[2 * tid + 1] = st[1].rconst;
    s_oid[tid] = oid;
    __syncthreads();
    const uint32_t base_c = s_base_c, base_f = s_base_f;
    s_cid[2 * tid] = base_c + cid0;
    s_cid[2 * tid + 1] = base_c + cid0 + 1;

    // projected triangles: (x_px, y_px, z_cam, 0) unrounded, cl2.cl:4384-4386
    bool cut_ok = (base_c + tot_c) <= P.cap_cut;
    if (!cut_ok && tid == 0) atomicOr(&P.counters[CTR_OVERFLOW], 2u);
    if (cut_ok) {
        for (int i = 0; i < num; i++) {
            if (!st[i].keep) continue;
            float4* dst = P.cutdown + (size_t)(base_c + cid0 + i) * 3;
            dst[0] = make_float4(st[i].p0.x, st[i].p0.y, st[i].p0.z, 0.f);
            dst[1] = make_float4(st[i].p1.x, st[i].p1.y, st[i].p1.z, 0.f);
            dst[2] = make_float4(st[i].p2.x, st[i].p2.y, st[i].p2.z, 0.f);
        }
    }
    __syncthreads();

    // fragment records {tri id, chunk, c_id, bits(rconst), o_id}, cl2.cl:4394-4406 — block-cooperative, word-coalesced
    const uint32_t totf = s_tot_f;
    if (totf == 0) return;
    if ((unsigned long long)base_f + totf > (unsigned long long)P.cap_frags) { if (tid == 0) atomicOr(&P.counters[CTR_OVERFLOW], 1u); return; }
    uint32_t* out = P.frags + (size_t)base_f * RR_FRAG_WORDS;
    const uint32_t nwords = totf * RR_FRAG_WORDS;
    for (uint32_t w = tid; w < nwords; w += SETUP_THREADS) {
        const uint32_t r = w / RR_FRAG_WORDS, field = w - r * RR_FRAG_WORDS;
        // last slot whose exclusive offset is <= r (slots with zero fragments share offsets; the last one owns r)
        int lo = 0, hi = 2 * SETUP_THREADS - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (s_fexcl[mid] <= r) lo = mid; else hi = mid - 1;
        }
        const int slot = lo;
        uint32_t val;
        switch (field) {
            case 0: val = bid * SETUP_THREADS + (uint32_t)(slot >> 1); break;
            case 1: val = r - s_fexcl[slot]; break;
            case 2: val = s_cid[slot]; break;
            case 3: val = __float_as_uint(s_rconst[slot]); break;
            default: val = s_oid[slot >> 1]; break;
        }
        out[w] = val;
    }
}

// =====================================================================================================================
// k_depth == kernel1 (cl2.cl:4986-5127) and k_ids == kernel2 (cl2.cl:5391-5546).
// Persistent grid (multiple of the SM count), grid-stride over the fragment records whose count lives on the device,
// so the host never reads the count back (the reference sizes the launch from a stale host copy, engine.cpp:1836,1899).
// One thread replays one chunk's pixel walk; depth goes out as red.global.min.u32 on the L2-resident depth buffer.
// =====================================================================================================================
struct RasterParams {
    const uint32_t* frags; const float4* cutdown; const uint32_t* counters; uint32_t cap_frags;
    uint32_t* depth; uint32_t* ids;
    float width, height; int W;
    int row_lo, row_hi;     // rows this context needs rasterised (band +- halo); chunks entirely outside are skipped
};

__device__ __forceinline__ bool chunk_rows_outside(const float4 mm, int op_size, uint32_t distance, int row_lo, int row_hi) {
    int width = (int)(mm.y - mm.x);
    if (width <= 0) return true;
    int k0 = op_size * (int)distance;
    int y_lo = (int)mm.z + k0 / width - 2;
    int y_hi = (int)mm.z + (k0 + op_size) / width + 2;
    return y_hi < row_lo || y_lo >= row_hi;
}

__global__ void __launch_bounds__(256) k_depth(const RasterParams P) {
    const uint32_t n = min(P.counters[CTR_NFRAG], P.cap_frags);
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        const uint32_t* rec = P.frags + (size_t)f * RR_FRAG_WORDS;
        const uint32_t distance = __ldg(rec + 1), ctri = __ldg(rec + 2);
        const float rconst = __uint_as_float(__ldg(rec + 3));
        const float4 c0 = __ldg(P.cutdown + (size_t)ctri * 3), c1 = __ldg(P.cutdown + (size_t)ctri * 3 + 1), c2 = __ldg(P.cutdown + (size_t)ctri * 3 + 2);
        const FragGeom g = frag_geom(xyz(c0), xyz(c1), xyz(c2), rconst, P.width, P.height);
        if (chunk_rows_outside(g.mm, RR_OP_SIZE, distance, P.row_lo, P.row_hi)) continue;
        uint32_t* depth = P.depth;
        const float ew = P.width;
        scan_chunk(g.mm, RR_OP_SIZE, distance, [&](float x, float y) {
            if (point_in_tri(x, y, g.xr.x, g.yr.x, g.xr.y, g.yr.y, g.xr.z, g.yr.z)) {
                float fd = fmaf(g.A, x, fmaf(g.B, y, g.C));
                uint32_t d = sat_u32(RR_U32MAXF / fd);
                atomicMin(depth + ((int)(y * ew) + (int)x), d);
            }
        });
    }
}

__global__ void __launch_bounds__(256) k_ids(const RasterParams P) {
    const uint32_t n = min(P.counters[CTR_NFRAG], P.cap_frags);
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        const uint32_t* rec = P.frags + (size_t)f * RR_FRAG_WORDS;
        const uint32_t distance = __ldg(rec + 1), ctri = __ldg(rec + 2);
        const float rconst = __uint_as_float(__ldg(rec + 3));
        const float4 c0 = __ldg(P.cutdown + (size_t)ctri * 3), c1 = __ldg(P.cutdown + (size_t)ctri * 3 + 1), c2 = __ldg(P.cutdown + (size_t)ctri * 3 + 2);
        const FragGeom g = frag_geom(xyz(c0), xyz(c1), xyz(c2), rconst, P.width, P.height);
        if (chunk_rows_outside(g.mm, RR_OP_SIZE, distance, P.row_lo, P.row_hi)) continue;
        const uint32_t* depth = P.depth;
        uint32_t* ids = P.ids;
        const int W = P.W;
        scan_chunk(g.mm, RR_OP_SIZE, distance, [&](float x, float y) {
            if (x < g.mm.x || y < g.mm.z) return;                                   // cl2.cl:5503
            if (point_in_tri(x, y, g.xr.x, g.yr.x, g.xr.y, g.yr.y, g.xr.z, g.yr.z)) {
                float fd = fmaf(g.A, x, fmaf(g.B, y, g.C));
                uint32_t d = sat_u32(RR_U32MAXF / fd);
                const int px = (int)y * W + (int)x;
                uint32_t val = depth[px];
                // racing plain stores in the reference; canonical winner = highest fragment index (last writer in id order)
                if (d > val - RR_BUF_ERROR && d < val + RR_BUF_ERROR) atomicMax(ids + px, f);
            }
        });
    }
}

// =====================================================================================================================
// shadow passes: k_shadow_setup == prearrange_realtime_shadowing (cl2.cl:4420-4636),
//                k_shadow_depth == kernel1_realtime_shadowing (cl2.cl:5130-5246).
// Slot numbers of a shadow pass are never observable (only the atomic_min result is), so allocation uses one
// warp-aggregated atomicAdd per counter per warp instead of the scan. Faces not owned by this context are skipped.
// =====================================================================================================================
struct ShadowSetupParams {
    const float4* pa; const float4* pb; const float2* pc;
    const ObjLite* objs;
    uint32_t n_tris;
    float3 lpos;
    FaceTable faces;
    float L, icut;
    int only_static;
    uint32_t face_mask;          // bit kk set -> this context renders face kk of this light
    uint32_t* frags; uint32_t cap_frags;
    float4* cutdown; uint32_t cap_cut;
    uint32_t* counters;
};

__device__ __forceinline__ uint32_t warp_alloc(uint32_t* counter, uint32_t mine) {
    const int lane = threadIdx.x & 31;
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + inc - mine;
}

__global__ void __launch_bounds__(256) k_shadow_setup(const ShadowSetupParams P) {
    const uint32_t tri = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = tri < P.n_tris;
    float3 v0, v1, v2;
    ObjLite G;
    uint32_t faces = 0;
    if (active) {
        const float4 a = __ldg(P.pa + tri), b = __ldg(P.pb + tri);
        const float2 c = __ldg(P.pc + tri);
        G = P.objs[__float_as_uint(c.y)];
        const bool is_static = (G.feature_flag & RR_FEATURE_IS_STATIC) > 0;
        const float3 gpos = make_float3(G.pos_scale.x, G.pos_scale.y, G.pos_scale.z);
        if ((!P.only_static && is_static) || (P.only_static && !is_static)) active = false;     // cl2.cl:4460-4464
        else if (length3(gpos - P.lpos) > RR_DEPTH_FAR) active = false;                          // cl2.cl:4472
        else {
            v0 = make_float3(a.x, a.y, a.z); v1 = make_float3(a.w, b.x, b.y); v2 = make_float3(b.z, b.w, c.x);
            const float s = G.pos_scale.w;
            faces |= 1u << ret_cubeface(rot_quat_n(v0 * s, G.nquat) + gpos, P.lpos);             // cl2.cl:4520-4539
            faces |= 1u << ret_cubeface(rot_quat_n(v1 * s, G.nquat) + gpos, P.lpos);
            faces |= 1u << ret_cubeface(rot_quat_n(v2 * s, G.nquat) + gpos, P.lpos);
            faces &= P.face_mask;
        }
    }
    if (!active) faces = 0;
    const bool two_sided = active && (G.feature_flag & RR_FEATURE_TWO_SIDED) > 0;
    // warp-uniform loop over the six faces so the warp-level allocation stays converged
    for (int kk = 0; kk < 6; kk++) {
        const bool mine = (faces >> kk) & 1u;
        if (!__any_sync(0xffffffffu, mine)) continue;
        SubTri st[2];
        int num = 0;
        st[0].keep = st[1].keep = false; st[0].n_frag = st[1].n_frag = 0; st[0].rconst = st[1].rconst = 0.f;
        if (mine) {
            num = transform_clip_project(v0, v1, v2, G, P.lpos, P.faces.r[kk], P.icut, P.L / 2.f, P.L / 2.f, P.L / 2.0f, st);
            for (int i = 0; i < num; i++) classify(st[i], two_sided, P.L, P.L, (float)RR_OP_SIZE_LIGHT);
        }
        const uint32_t nk = (st[0].keep ? 1u : 0u) + (st[1].keep ? 1u : 0u);
        const uint32_t nf0 = st[0].keep ? (uint32_t)st[0].n_frag : 0u, nf1 = st[1].keep ? (uint32_t)st[1].n_frag : 0u;
        uint32_t cbase = warp_alloc(&P.counters[CTR_S_NCUT], nk);
        uint32_t fbase = warp_alloc(&P.counters[CTR_S_NFRAG], nf0 + nf1);
        if (nk == 0) continue;
        if (cbase + nk > P.cap_cut) { atomicOr(&P.counters[CTR_OVERFLOW], 2u); continue; }
        if ((unsigned long long)fbase + nf0 + nf1 > (unsigned long long)P.cap_frags) { atomicOr(&P.counters[CTR_OVERFLOW], 1u); continue; }
        uint32_t cid = cbase;
        for (int i = 0; i < 2; i++) {
            if (!st[i].keep) continue;
            float4* dst = P.cutdown + (size_t)cid * 3;
            dst[0] = make_float4(st[i].p0.x, st[i].p0.y, st[i].p0.z, 0.f);
            dst[1] = make_float4(st[i].p1.x, st[i].p1.y, st[i].p1.z, 0.f);
            dst[2] = make_float4(st[i].p2.x, st[i].p2.y, st[i].p2.z, 0.f);
            uint4* rec = reinterpret_cast<uint4*>(P.frags) + fbase;                 // {face, chunk, c_id, bits(rconst)} cl2.cl:4626-4631
            for (int a = 0; a < st[i].n_frag; a++) rec[a] = make_uint4((uint32_t)kk, (uint32_t)a, cid, __float_as_uint(st[i].rconst));
            fbase += (uint32_t)st[i].n_frag;
            cid++;
        }
    }
}

struct ShadowDepthParams {
    const uint32_t* frags; const float4* cutdown; const uint32_t* counters; uint32_t cap_frags;
    uint32_t* slab;      // this light's 6*L*L cubemap
    float L; int Li;
};

__global__ void __launch_bounds__(256) k_shadow_depth(const ShadowDepthParams P) {
    const uint32_t n = min(P.counters[CTR_S_NFRAG], P.cap_frags);
    const uint4* recs = reinterpret_cast<const uint4*>(P.frags);
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        const uint4 rec = __ldg(recs + f);
        const uint32_t face = rec.x, distance = rec.y, ctri = rec.z;
        const float rconst = __uint_as_float(rec.w);
        const float4 c0 = __ldg(P.cutdown + (size_t)ctri * 3), c1 = __ldg(P.cutdown + (size_t)ctri * 3 + 1), c2 = __ldg(P.cutdown + (size_t)ctri * 3 + 2);
        const FragGeom g = frag_geom(xyz(c0), xyz(c1), xyz(c2), rconst, P.L, P.L);
        uint32_t* depth = P.slab + (size_t)face * P.Li * P.Li;
        const float ew = P.L;
        scan_chunk(g.mm, RR_OP_SIZE_LIGHT, distance, [&](float x, float y) {
            if (point_in_tri(x, y, g.xr.x, g.yr.x, g.xr.y, g.yr.y, g.xr.z, g.yr.z)) {
                float fd = fmaf(g.A, x, fmaf(g.B, y, g.C));
                uint32_t d = sat_u32(RR_U32MAXF / fd);
                atomicMin(depth + ((int)(y * ew) + (int)x), d);
            }
        });
    }
}

// accumulate statistics of a finished shadow pass and reset its counters for the next one
__global__ void k_shadow_pass_end(uint32_t* counters) {
    counters[CTR_S_TOTAL] += counters[CTR_S_NFRAG];
    counters[CTR_S_NCUT] = 0;
    counters[CTR_S_NFRAG] = 0;
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
struct AtlasView { const uchar4* texels; const uint32_t* nums; const uint32_t* sizes; uint32_t mip_start; };

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

__global__ void __launch_bounds__(256) k_atlas_upload(const uchar4* __restrict__ src, int w, int h, uint32_t tex_id, int flip, uchar4* __restrict__ atlas,
                                                      const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
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

__global__ void __launch_bounds__(256) k_atlas_mip(uint32_t src_id, uint32_t dst_id, int gw, int gh, uchar4* __restrict__ atlas,
                                                   const uint32_t* __restrict__ nums, const uint32_t* __restrict__ sizes) {
    int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
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

// =====================================================================================================================
// k_shade == kernel3 (cl2.cl:5795-6408). One thread per pixel, 32x8 tiles (a warp = 32 consecutive pixels of a row so
// depth / id loads and the RGBA8 / normal / clear stores are full 128-byte lines).
// =====================================================================================================================
struct ShadeParams {
    const rr_triangle* tris; const rr_obj_desc* objs;
    const uint32_t* frags; const float4* cutdown;
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
};

// read_tex_array_all_precalculated, cl2.cl:823-851
__device__ __forceinline__ float4 read_tex_pre(float cx, float cy, int which, int slice, float width, const uchar4* __restrict__ atlas) {
    const float ihnum = width * (1.f / 2048);
    float tnumy = floorf((float)which * ihnum);
    float tnumx = (float)which - tnumy / ihnum;
    cx = clampf(cx, 0.001f, width - 0.001f);
    cy = clampf(cy, 0.001f, width - 0.001f);
    int ix = (int)fmaf(tnumx, width, cx), iy = (int)fmaf(tnumy, width, cy);
    uchar4 t = __ldg(atlas + (size_t)slice * RR_ATLAS_DIM * RR_ATLAS_DIM + (size_t)iy * RR_ATLAS_DIM + ix);
    return make_float4((float)t.x, (float)t.y, (float)t.z, (float)t.w);
}

// return_bilinear_col_all_precalculated, cl2.cl:1426-1455
__device__ __forceinline__ float4 bilinear_pre(float mx, float my, int which, int slice, float width, const uchar4* __restrict__ atlas) {
    float px = floorf(mx), py = floorf(my);
    float4 c0 = read_tex_pre(px, py, which, slice, width, atlas);
    float4 c1 = read_tex_pre(px + 1, py, which, slice, width, atlas);
    float4 c2 = read_tex_pre(px, py + 1, which, slice, width, atlas);
    float4 c3 = read_tex_pre(px + 1, py + 1, which, slice, width, atlas);
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
    float4 col1 = bilinear_pre(vx * size_lower, vy * size_lower, which_lower, slice_lower, size_lower, av.texels);
    float4 col2 = bilinear_pre(vx * size_higher, vy * size_higher, which_higher, slice_higher, size_higher, av.texels);
    float4 fc = col1 + (col2 - col1) * fmd;
    return fc * (1.f / 255.f);
}

// generate_ssao, cl2.cl:2194-2260
__device__ __forceinline__ float generate_ssao(int sx, int sy, const uint32_t* __restrict__ depth_buffer, int W, int H, float fov, float ssao_rad, float ssao_div) {
    uint32_t seed1 = wang_hash((uint32_t)sx + (uint32_t)W * (uint32_t)H * (uint32_t)sy);
    uint32_t seed2 = rand_xorshift(seed1);
    float foffset = (float)seed2 / RR_U32MAXF;
    float depth = ((float)__ldg(depth_buffer + sy * W + sx) / RR_U32MAXF) * RR_DEPTH_FAR;
    float rad = ssao_rad + foffset / 2.f;
    float world_rad = rad * fov / depth;
    float acc = 0.f;
    for (int y = -2; y <= 2; y++)
        for (int x = -2; x <= 2; x++) {
            float ox = roundf((float)x * world_rad), oy = roundf((float)y * world_rad);
            float wx = clampf((float)sx + ox, 1.f, (float)W - 2.f), wy = clampf((float)sy + oy, 1.f, (float)H - 2.f);
            float d2 = ((float)__ldg(depth_buffer + ((int)wy) * W + (int)wx) / RR_U32MAXF) * RR_DEPTH_FAR;
#pragma unroll
            for (int z = -2; z <= 2; z++)
                if (d2 > depth + (float)z) acc += 1.f;
        }
    acc /= 125.f;                       // pow(samples*2+1, 3)
    return 1.f - (1.f - acc) / ssao_div;
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
            float ldp1 = ((float)__ldg(ldepth_map + (ipy + y) * L + ipx + x) / RR_U32MAXF) * RR_DEPTH_FAR;
            cnd[(y + 1) * 4 + x + 1] = dpth > ldp1 + bias ? 1.f : 0.f;
        }
    float shadow = 0.f;
#pragma unroll
    for (int y = -1; y <= 1; y++)
#pragma unroll
        for (int x = -1; x <= 1; x++)
            shadow += bilinear_interpolate(pp.x + 0.5f + (float)x, pp.y + 0.5f + (float)y, cnd[(y + 1) * 4 + x + 1], cnd[(y + 1) * 4 + x + 2],
                                           cnd[(y + 2) * 4 + x + 1], cnd[(y + 2) * 4 + x + 2]);
    return shadow / 9.f;
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

__global__ void __launch_bounds__(256) k_shade(const ShadeParams P) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = P.row0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.W || y >= P.row1) return;
    const int W = P.W, H = P.H;
    const size_t px = (size_t)y * W + x;
    const uint32_t d = P.depth[px];
    P.depth_next[px] = 0xFFFFFFFFu;                                    // to_clear, cl2.cl:5820
    P.ids_next[px] = 0u;                                               // id image of the next frame (atomicMax needs a clean slate)
    if (y < P.band_y0 || y >= P.band_y1) return;
    if (d == 0xFFFFFFFFu) {                                            // cl2.cl:5835-5862
        P.rgba8[px] = make_uchar4(quant8(P.clear.x), quant8(P.clear.y), quant8(P.clear.z), quant8(P.clear.w));
        return;
    }
    const uint32_t idv = P.ids[px];
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
    const float4 Gq = __ldg(reinterpret_cast<const float4*>(G->world_rot_quat));
    const float Gscale = __ldg(&G->scale);
    const float3 Gpos = xyz(Gpos4);
    const float4 Gqn = normalize4(Gq), Gqb = back_quat(Gq);
    const float3 p1 = xyz(pv0) * Gscale, p2 = xyz(pv1) * Gscale, p3 = xyz(pv2) * Gscale;
    const float fov = P.fov;
    const float3 zero3 = make_float3(0, 0, 0);

    const float ldepth = ((float)d / RR_U32MAXF) * RR_DEPTH_FAR;      // cl2.cl:5897
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
    float3 rseed = make_float3((float)seed2 / RR_U32MAXF, (float)seed3 / RR_U32MAXF, (float)seed4 / RR_U32MAXF);
    rseed = make_float3((rseed.x - 0.5f) * 2, (rseed.y - 0.5f) * 2, (rseed.z - 0.5f) * 2);

    float3 diffuse_sum = zero3, specular_sum = zero3;
    float3 l2p = normalize3(P.cam.pos - global_position);
    const int feature_flag = __ldg(&G->feature_flag);
    const bool is_two_sided = (feature_flag & RR_FEATURE_TWO_SIDED) > 0;
    const bool receives_dynamic_shadows = !((feature_flag & RR_FEATURE_NO_DYNAMIC_SHADOWS) > 0);
    const bool is_front = front_facing(xyz(ct0), xyz(ct1), xyz(ct2));
    if (!is_front && is_two_sided) normal = -normal;
    const float ssao = P.no_ssao ? 1.f : generate_ssao(x, y, P.depth, W, H, fov, P.ssao_rad, P.ssao_div);
    normal = normalize3(normal);
    const float3 lighting_normal = normalize3(normal + rseed / 100.f);
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
        float illumination = __ldg(&l->brightness) / powf((distance / __ldg(&l->radius)) + 1.f, 2.f);
        const float cutoff = 0.1f;
        illumination -= cutoff;
        illumination *= 1.f / (1.f - cutoff);
        if (illumination <= 0) continue;
        float3 light_col = xyz(lc4);
        if (P.linear) light_col = make_float3(gamma_fwd(light_col.x), gamma_fwd(light_col.y), gamma_fwd(light_col.z));
        if (lshadow && receives_dynamic_shadows) {
            int face = ret_cubeface(global_position, lpos);
            float dyn = 1.f - hard_occlusion(lpos, normal, point_to_light, P.shadow_dyn, face, global_position, shnum, P);
            occlusion = fminf(occlusion, dyn);
            shnum++;
        }
        point_to_light = normalize3(point_to_light);
        float light = dot3(point_to_light, lighting_normal);
        light *= occlusion;
        light = fmaxf(light, 0.f);
        float diffuse = (1.0f - ambient) * light;
        diffuse_sum = diffuse_sum + light_col * ((diffuse + ambient) * __ldg(&l->diffuse) * Gdiffuse * illumination);
        float3 Hh = normalize3(l2p + point_to_light);
        const float kS = 0.4f;
        float ndh = fmaxf(0.f, dot3(normal, Hh));
        float ndv = fmaxf(0.f, dot3(normal, l2p));
        float vdh = fmaxf(0.f, dot3(l2p, Hh));
        float ndl = fmaxf(0.f, dot3(normal, point_to_light));
        const float F0 = 0.4f;
        float fresnel = F0 + (1 - F0) * powf((1.f - vdh), 5.f);
        float rough = clampf(1.f - Gspecular, 0.001f, 10.f);
        float alpha = rational_acos(ndh);
        float microfacet = 0.8346f * expf(-alpha * alpha / (rough * rough));
        float sv = 2 * ndh / vdh;
        float c1 = sv * ndv, c2 = sv * ndl;
        float geometric = fminf(fminf(1.f, c1), c2);
        float spec = (fresnel * microfacet * geometric) / (RR_PI_F * ndv);
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
    if (P.linear) fc = make_float3(gamma_inv(fc.x), gamma_inv(fc.y), gamma_inv(fc.z));
    P.rgba8[px] = make_uchar4(quant8(clampf(fc.x, 0.f, 1.f)), quant8(clampf(fc.y, 0.f, 1.f)), quant8(clampf(fc.z, 0.f, 1.f)), quant8(col.w));

    // encode_normal + float_to_short, cl2.cl:5588-5628
    float3 nn = normal;
    if (nn.x * nn.x + nn.y * nn.y < 0.0001f) nn.x = 0.01f;
    float ln = sqrtf(nn.x * nn.x + nn.y * nn.y);
    float k = sqrtf(fmaxf(nn.z * 0.5f + 0.5f, 0.f));
    float rx = (nn.x / ln) * k, ry = (nn.y / ln) * k;
    P.normals[px] = make_ushort2(to_ushort_sat(((rx + 1) / 2) * 65536 - 1), to_ushort_sat(((ry + 1) / 2) * 65536 - 1));
}

// =====================================================================================================================
// roofline micro-benchmarks
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_bench_atomic_min(uint32_t* buf, uint32_t n_words_mask, uint32_t iters) {
    uint32_t s = wang_hash(blockIdx.x * blockDim.x + threadIdx.x + 1u);
    for (uint32_t i = 0; i < iters; i++) {
        s = rand_xorshift(s);
        atomicMin(buf + (s & n_words_mask), s >> 3);
    }
}

__global__ void __launch_bounds__(256) k_bench_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace rr

// oracle.cpp — TEST INFRASTRUCTURE ONLY. CPU restatement of OpenCLRenderer's per-frame raster path.
//
// This file is the parity oracle for the CUDA product in openclrenderer_b200/csrc. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load it. The product never does.
//
// It follows /root/reference/cl2.cl literally, one OpenCL work-item per loop iteration, under the arithmetic pinned in
// SURVEY.md §7/§8c (the reference is built with -cl-fast-relaxed-math and cannot be compiled here, so one arithmetic
// has to be chosen): IEEE-RN + - * / sqrt, mad() == single-rounding fmaf, no implicit contraction (-ffp-contract=off),
// native_divide == /, native_recip(x) == 1.0f/x, fast_normalize(v) == v / sqrt(dot) (dot summed x->w left to right),
// camera / cube-face sin,cos computed once on the host in double and rounded to float, float->uint conversions saturate,
// round() == half away from zero, min/max == fminf/fmaxf, allocation order == sequential global-id order (so cutdown
// ids and fragment ids are exclusive prefix sums in triangle order, and the id winner is the highest fragment index).
//
// Parity pin: the reference holds no golden vectors or tests for this path. The pin is the reference
// itself: oracle/ref_opencl.py runs the unmodified cl2.cl through the NVIDIA OpenCL ICD on the GPU box
// (tests/test_gpu_reference_cl.py) and tests/golden/ holds the outputs it produced (tests/golden/make_golden.py).
//
// Each function cites the cl2.cl lines it restates.

#include "../include/rr.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr float DEPTH_FAR = 350000.0f;       // cl2.cl:17
constexpr float U32MAXF   = 4294967296.0f;   // (float)UINT_MAX, cl2.cl:19
constexpr float CL_M_PI   = 3.1415927f;      // cl2.cl:14
constexpr int   OP_SIZE = 500;               // cl2.cl:4247
constexpr int   OP_SIZE_LIGHT = 300;         // cl2.cl:4249
constexpr int   FRAG_MUL = 5;                // cl2.cl:4252
constexpr int   FIDM1 = 4;                   // cl2.cl:4411
constexpr uint32_t BUF_ERROR = 20;           // cl2.cl:5387
constexpr int   MIP_LEVELS = 4;              // cl2.cl:5
constexpr int   ATLAS_DIM = 2048;            // texture_context.hpp:16

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f4 operator+(f4 a, f4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline f4 operator-(f4 a, f4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline f4 operator*(f4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline f4 operator/(f4 a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
inline f2 operator+(f2 a, f2 b) { return {a.x + b.x, a.y + b.y}; }
inline f2 operator-(f2 a, f2 b) { return {a.x - b.x, a.y - b.y}; }
inline f2 operator*(f2 a, float s) { return {a.x * s, a.y * s}; }

inline float cl_min(float a, float b) { return fminf(a, b); }
inline float cl_max(float a, float b) { return fmaxf(a, b); }
inline float cl_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot4(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline f3 cross3(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float fast_length3(f3 a) { return sqrtf(dot3(a, a)); }
inline f3 fast_normalize3(f3 a) { float l = sqrtf(dot3(a, a)); return a / l; }
inline f4 fast_normalize4(f4 a) { float l = sqrtf(dot4(a, a)); return a / l; }
inline f3 mad3(f3 a, float b, f3 c) { return {fmaf(a.x, b, c.x), fmaf(a.y, b, c.y), fmaf(a.z, b, c.z)}; }
inline f2 mad2(f2 a, float b, f2 c) { return {fmaf(a.x, b, c.x), fmaf(a.y, b, c.y)}; }
inline f4 mad4(f4 a, float b, f4 c) { return {fmaf(a.x, b, c.x), fmaf(a.y, b, c.y), fmaf(a.z, b, c.z), fmaf(a.w, b, c.w)}; }
inline f3 v3(const float* p) { return {p[0], p[1], p[2]}; }

inline uint32_t sat_u32(float f) {           // pinned float->uint (SURVEY.md §7 hard part 4)
    if (!(f > 0.0f)) return 0u;               // NaN, negative, zero
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}
inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct rotsc { f3 s, c; };                    // host-computed native_sin / native_cos of an euler triple
inline rotsc make_rotsc(float rx, float ry, float rz) {
    rotsc r;
    r.s = {(float)sin((double)rx), (float)sin((double)ry), (float)sin((double)rz)};
    r.c = {(float)cos((double)rx), (float)cos((double)ry), (float)cos((double)rz)};
    return r;
}

// cl2.cl:220-271 rot()
inline f3 rot(f3 point, f3 c_pos, const rotsc& r) {
    const f3 c = r.c, s = r.s;
    f3 rel = point - c_pos;
    float t = fmaf(s.z, rel.y, c.z * rel.x);
    float u = fmaf(c.y, rel.z, s.y * t);
    float v = fmaf(c.z, rel.y, -(s.z * rel.x));
    f3 ret;
    ret.x = fmaf(c.y, t, -(s.y * rel.z));
    ret.y = fmaf(s.x, u, c.x * v);
    ret.z = fmaf(c.x, u, -(s.x * v));
    return ret;
}

// cl2.cl:275-348 back_rot() — "transposed, with mads" factorisation at 325-345
inline f3 back_rot(f3 point, f3 c_pos, const rotsc& r) {
    const f3 c = r.c, s = r.s;
    f3 rel = point - c_pos;
    f3 ret;
    ret.x = c.z * (fmaf(c.y, rel.x, fmaf(s.x, s.y * rel.y, c.x * s.y * rel.z))) + s.z * (s.x * rel.z - c.x * rel.y);
    ret.y = fmaf(s.z, c.y * rel.x, fmaf(fmaf(c.x, c.z, s.x * s.y * s.z), rel.y, (fmaf(-s.x, c.z, c.x * s.y * s.z) * rel.z)));
    ret.z = fmaf(-s.y, rel.x, c.y * fmaf(s.x, rel.y, c.x * rel.z));
    return ret;
}

// cl2.cl:350-357
inline f3 rot_quat(f3 point, f4 quat) {
    quat = fast_normalize4(quat);
    f3 q = {quat.x, quat.y, quat.z};
    f3 t = 2.f * cross3(q, point);
    return point + quat.w * t + cross3(q, t);
}

// cl2.cl:359-370
inline f3 back_rot_quat(f3 point, f4 quat) {
    f4 conj = {-quat.x, -quat.y, -quat.z, quat.w};
    float len_sq = dot4(conj, conj);
    return rot_quat(point, conj / len_sq);
}

// cl2.cl:408-411
inline float calc_rconstant_v(f3 x, f3 y) {
    return 1.0f / (x.y * y.z + x.x * (y.y - y.z) - x.z * y.y + (x.z - x.y) * y.x);
}

// cl2.cl:413-418
inline void interpolate_get_const(f3 f, f3 x, f3 y, float rconstant, float* A, float* B, float* C) {
    *A = ((f.y * y.z + f.x * (y.y - y.z) - f.z * y.y + (f.z - f.y) * y.x) * rconstant);
    *B = (-(f.y * x.z + f.x * (x.y - x.z) - f.z * x.y + (f.z - f.y) * x.x) * rconstant);
    *C = f.x - (*A) * x.x - (*B) * y.x;
}

// cl2.cl:420-441 calc_min_max() (443-459 calc_min_max_p is the same arithmetic)
inline void calc_min_max(const f3 p[3], float width, float height, float ret[4]) {
    float x[3], y[3];
    for (int i = 0; i < 3; i++) { x[i] = roundf(p[i].x); y[i] = roundf(p[i].y); }
    ret[0] = cl_min(cl_min(x[0], x[1]), x[2]) - 1;
    ret[1] = cl_max(cl_max(x[0], x[1]), x[2]);
    ret[2] = cl_min(cl_min(y[0], y[1]), y[2]) - 1;
    ret[3] = cl_max(cl_max(y[0], y[1]), y[2]);
    ret[0] = cl_clamp(ret[0], 0.0f, width - 1);
    ret[1] = cl_clamp(ret[1], 0.0f, width - 1);
    ret[2] = cl_clamp(ret[2], 0.0f, height - 1);
    ret[3] = cl_clamp(ret[3], 0.0f, height - 1);
}

// cl2.cl:491-494
inline int backface_cull_expanded(f3 p0, f3 p1, f3 p2) { return cross3(p1 - p0, p2 - p0).z < 0.f; }

// cl2.cl:577-663 generate_new_triangles()
inline void generate_new_triangles(const f3 points[3], int icut, int* num, f3 ret[2][3]) {
    int id_valid = 0;
    int ids_behind[3];
    int n_behind = 0;
    for (int i = 0; i < 3; i++) {
        if (points[i].z <= (float)icut || points[i].z > DEPTH_FAR) { ids_behind[n_behind] = i; n_behind++; }
        else id_valid = i;
    }
    if (n_behind > 2) { *num = 0; return; }
    if (n_behind == 0) {
        ret[0][0] = points[0]; ret[0][1] = points[1]; ret[0][2] = points[2];
        *num = 1; return;
    }
    int g1 = 0, g2 = 0, g3 = 0;
    if (n_behind == 1) {
        int id = ids_behind[0];
        g1 = id;
        g2 = (id + 1) >= 3 ? id - 2 : id + 1;
        g3 = (id + 2) >= 3 ? id - 1 : id + 2;
    }
    if (n_behind == 2) { g2 = ids_behind[0]; g3 = ids_behind[1]; g1 = id_valid; }
    f3 p1 = points[g2] + (((float)icut - points[g2].z) * (points[g1] - points[g2])) / (points[g1].z - points[g2].z);
    f3 p2 = points[g3] + (((float)icut - points[g3].z) * (points[g1] - points[g3])) / (points[g1].z - points[g3].z);
    if (n_behind == 1) {
        f3 c1 = points[g2], c2 = points[g3];
        ret[0][0] = p1; ret[0][1] = c1; ret[0][2] = c2;
        ret[1][0] = p1; ret[1][1] = c2; ret[1][2] = p2;
        *num = 2;
    } else {
        f3 c1 = points[g1];
        ret[0][ids_behind[0]] = p1;
        ret[0][ids_behind[1]] = p2;
        ret[0][id_valid] = c1;
        *num = 1;
    }
}

// cl2.cl:535-544 depth_project()
inline void depth_project(const f3 rotated[3], float width, float height, float fovc, f3 ret[3]) {
    for (int i = 0; i < 3; i++) {
        float k = fovc / rotated[i].z;
        ret[i].x = fmaf(rotated[i].x, k, width / 2.f);
        ret[i].y = fmaf(rotated[i].y, k, height / 2.f);
        ret[i].z = rotated[i].z;
    }
}
// cl2.cl:546-570
inline f3 depth_project_singular(f3 rotated, float width, float height, float fovc) {
    float k = fovc / rotated.z;
    return {fmaf(rotated.x, k, width / 2.f), fmaf(rotated.y, k, height / 2.f), rotated.z};
}

// cl2.cl:700-729 full_rotate_quat() (503-508 rot_quat_with_offset)
inline void full_rotate_quat(f3 v1, f3 v2, f3 v3, f3 passback[2][3], int* num, f3 c_pos, const rotsc& c_rot, f3 offset,
                             f4 rotation_offset, float scale, float fovc, float width, float height, int icut) {
    f3 tris[2][3];
    f3 pr[3];
    pr[0] = rot(rot_quat(v1 * scale, rotation_offset) + offset, c_pos, c_rot);
    pr[1] = rot(rot_quat(v2 * scale, rotation_offset) + offset, c_pos, c_rot);
    pr[2] = rot(rot_quat(v3 * scale, rotation_offset) + offset, c_pos, c_rot);
    int n = 0;
    generate_new_triangles(pr, icut, &n, tris);
    *num = n;
    if (n == 0) return;
    depth_project(tris[0], width, height, fovc, passback[0]);
    if (n == 2) depth_project(tris[1], width, height, fovc, passback[1]);
}

// cl2.cl:4798-4807
inline bool point_in_tri(f2 p, f2 p0, f2 p1, f2 p2) {
    float A = 0.5f * (-p1.y * p2.x + p0.y * (-p1.x + p2.x) + p0.x * (p1.y - p2.y) + p1.x * p2.y);
    float sign = A < 0 ? -1.f : 1.f;
    float s = (p0.y * p2.x - p0.x * p2.y + (p2.y - p0.y) * p.x + (p0.x - p2.x) * p.y) * sign;
    float t = (p0.x * p1.y - p0.y * p1.x + (p0.y - p1.y) * p.x + (p1.x - p0.x) * p.y) * sign;
    return s > -0.0001f && t > -0.0001f && (s + t) < 2.0001f * A * sign;
}

// cl2.cl:1745-1790
inline int ret_cubeface(f3 point, f3 light) {
    f3 rel = point - light;
    f3 arel = {fabsf(rel.x), fabsf(rel.y), fabsf(rel.z)};
    if (arel.x >= arel.y && arel.x >= arel.z) return rel.x < 0 ? 4 : 5;
    if (arel.y > arel.x && arel.y >= arel.z) return rel.y < 0 ? 1 : 3;
    if (arel.z > arel.x && arel.z > arel.y) { if (rel.z < 0) return 2; }
    return 0;
}

// cl2.cl:1919-1936
inline uint32_t wang_hash(uint32_t seed) {
    seed = (seed ^ 61) ^ (seed >> 16);
    seed *= 9;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2d;
    seed = seed ^ (seed >> 15);
    return seed;
}
inline uint32_t rand_xorshift(uint32_t s) { s ^= (s << 13); s ^= (s >> 17); s ^= (s << 5); return s; }

// cl2.cl:2469-2477
inline float rational_acos(float x) {
    float a = -0.939115566365855f, b = 0.9217841528914573f, c = -1.2845906244690837f, d = 0.295624144969963174f;
    return CL_M_PI / 2.f + (a * x + b * x * x * x) / (1.f + c * x * x + d * powf(x, 4.f));
}

// The pixel walk of kernel1 / kernel2 / kernel1_realtime_shadowing, cl2.cl:5042-5095 (== 5184-5227 == 5447-5508).
// Calls f(x, y) for every pixel the reference's state machine tests.
template <class F>
inline void scan_fragment(const float mm[4], int op_size, uint32_t distance, F&& f) {
    int width = (int)(mm[1] - mm[0]);
    if (width <= 0) return;                     // cannot happen for an emitted fragment (area > 0)
    int pixel_along = op_size * (int)distance;
    int pcount = -1;
    float x = (float)((pixel_along + 0) % width) + mm[0] - 1;
    float y = floorf((float)(pixel_along + pcount) / (float)width) + mm[2];
    float iwidth = 1.f / (float)width;
    float running_width_mod = (float)((pixel_along + pcount) % width);
    while (pcount < op_size) {
        pcount++;
        x += 1;
        running_width_mod += 1;
        if (running_width_mod >= (float)width) running_width_mod = 0;
        float ty = y;
        y = floorf(fmaf((float)(pixel_along + pcount), iwidth, mm[2]));
        x = y != ty ? running_width_mod + mm[0] : x;
        if (y >= mm[3]) break;
        bool oob = x >= mm[1];
        if (oob) continue;
        f(x, y);
    }
}

inline void atomic_min_u32(uint32_t* p, uint32_t v) {
    uint32_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}
inline void atomic_max_u32(uint32_t* p, uint32_t v) {
    uint32_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

thread_local char g_err[256] = "";

}  // namespace

struct orc_ctx {
    rr_config cfg;
    float fov;
    int W, H, L;
    int threads;
    std::vector<rr_triangle> tris, back_tris;
    std::vector<rr_obj_desc> objs, back_objs;
    std::vector<rr_light> lights;
    std::vector<uint8_t> atlas;       // uchar4[2048*2048*slices]
    std::vector<uint32_t> nums, sizes;
    uint32_t mipmap_start = 0;
    std::vector<uint32_t> depth[2];
    int cur = 0;                      // depth_buffer[0] of the n_buffer (object_context.hpp); flip() advances it
    std::vector<uint32_t> ids;
    std::vector<uint8_t> rgba8;
    std::vector<float> colour;        // float4 per pixel, what write_imagef received
    std::vector<uint16_t> normals;
    std::vector<uint32_t> frags;      // main-pass fragment records, 5 words
    std::vector<f4> cutdown;          // main-pass projected triangles, 3 float4 each
    uint32_t n_frags = 0, n_cut = 0;
    std::vector<uint32_t> shadow_dyn, shadow_static;
    uint32_t n_shadow = 0, n_static = 0;
    rr_timings tm;
    uint64_t depth_samples = 0;       // covered depth samples of the last main pass (A_depth)
    uint64_t shadow_samples = 0;      // A_shadow
    uint64_t sat_events = 0;          // float->uint conversions that saturated (hard part 4)
    // camera of the last orc_frame_draw and of the frame before it (object_context_data::c_pos_old, object_context.cpp:23-24),
    // and object_context_data::frame_id (engine.cpp:2024) — what the post passes are handed
    float cam_pos[4] = {0, 0, 0, 0}, cam_rot[4] = {0, 0, 0, 0}, cam_pos_old[4] = {0, 0, 0, 0}, cam_rot_old[4] = {0, 0, 0, 0};
    uint32_t frame_id = 0;
};

namespace {

// One projected sub-triangle produced by triangle setup.
struct sub_tri { f3 p[3]; float rconst; float mm[4]; int n_frag; bool keep; };

// cl2.cl:4352-4377 (main) / 4571-4597 (shadow): cull + bbox + fragment count for one clipped triangle.
inline void classify(sub_tri& s, bool two_sided, float ewidth, float eheight, int op_size) {
    const f3* tp = s.p;
    int valid = two_sided || backface_cull_expanded(tp[0], tp[1], tp[2]);
    int cond = (tp[0].x < 0 && tp[1].x < 0 && tp[2].x < 0) ||
               (tp[0].x >= ewidth && tp[1].x >= ewidth && tp[2].x >= ewidth) ||
               (tp[0].y < 0 && tp[1].y < 0 && tp[2].y < 0) ||
               (tp[0].y >= eheight && tp[1].y >= eheight && tp[2].y >= eheight);
    s.keep = !(!valid || cond);
    s.n_frag = 0;
    if (!s.keep) return;
    f3 xpv = {roundf(tp[0].x), roundf(tp[1].x), roundf(tp[2].x)};
    f3 ypv = {roundf(tp[0].y), roundf(tp[1].y), roundf(tp[2].y)};
    s.rconst = calc_rconstant_v(xpv, ypv);
    calc_min_max(tp, ewidth, eheight, s.mm);
    float area = (s.mm[1] - s.mm[0]) * (s.mm[3] - s.mm[2]);
    s.n_frag = (int)ceilf(area / (float)op_size);
}

struct tri_setup { int num; sub_tri s[2]; bool skipped; };

// ---- main view: prearrange, cl2.cl:4272-4409 -------------------------------------------------------------------
void prearrange(orc_ctx* c, const float c_pos[4], const rotsc& crot) {
    const uint32_t T = (uint32_t)c->tris.size();
    const float ewidth = (float)c->W, eheight = (float)c->H, efov = c->fov;
    const f3 cpos = v3(c_pos);
    std::vector<tri_setup> st(T);
#pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t id = 0; id < (int64_t)T; id++) {
        const rr_triangle& Tt = c->tris[id];
        tri_setup& o = st[id];
        o.num = 0; o.skipped = false;
        int o_id = (int)Tt.vertices[0].object_id;
        const rr_obj_desc& G = c->objs[o_id];
        f3 g_world_pos = v3(G.world_pos);
        if (fast_length3(g_world_pos - cpos) > DEPTH_FAR) { o.skipped = true; continue; }   // cl2.cl:4321
        f3 proj[2][3];
        int num = 0;
        f4 q = {G.world_rot_quat[0], G.world_rot_quat[1], G.world_rot_quat[2], G.world_rot_quat[3]};
        full_rotate_quat(v3(Tt.vertices[0].pos), v3(Tt.vertices[1].pos), v3(Tt.vertices[2].pos), proj, &num, cpos, crot,
                         g_world_pos, q, G.scale, efov, ewidth, eheight, c->cfg.depth_icutoff);
        o.num = num;
        bool two_sided = (G.feature_flag & RR_FEATURE_TWO_SIDED) > 0;
        for (int i = 0; i < num; i++) {
            for (int k = 0; k < 3; k++) o.s[i].p[k] = proj[i][k];
            classify(o.s[i], two_sided, ewidth, eheight, OP_SIZE);
        }
    }
    // sequential semantics of atomic_add(id_cutdown_tris, num) (4342) and atomic_add(id_buffer_atomc, thread_num) (4388)
    std::vector<uint32_t> cbase(T + 1), fbase(T + 1);
    uint32_t cc = 0, fc = 0;
    for (uint32_t id = 0; id < T; id++) {
        cbase[id] = cc; fbase[id] = fc;
        cc += (uint32_t)st[id].num;
        for (int i = 0; i < st[id].num; i++) fc += (uint32_t)st[id].s[i].n_frag;
    }
    cbase[T] = cc; fbase[T] = fc;
    c->n_cut = cc; c->n_frags = fc;
    c->cutdown.assign((size_t)cc * 3, f4{0, 0, 0, 0});
    c->frags.assign((size_t)fc * FRAG_MUL, 0u);
#pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t id = 0; id < (int64_t)T; id++) {
        const tri_setup& o = st[id];
        uint32_t f = fbase[id] * FRAG_MUL;
        uint32_t o_id = c->tris[id].vertices[0].object_id;
        for (int i = 0; i < o.num; i++) {
            if (!o.s[i].keep) continue;
            uint32_t c_id = cbase[id] + (uint32_t)i;
            for (int k = 0; k < 3; k++) c->cutdown[(size_t)c_id * 3 + k] = {o.s[i].p[k].x, o.s[i].p[k].y, o.s[i].p[k].z, 0.f};
            for (int a = 0; a < o.s[i].n_frag; a++) {
                c->frags[f++] = (uint32_t)id;
                c->frags[f++] = (uint32_t)a;
                c->frags[f++] = c_id;
                c->frags[f++] = as_uint(o.s[i].rconst);
                c->frags[f++] = o_id;
            }
        }
    }
}

struct frag_geom { f3 xpv, ypv; float A, B, C; float mm[4]; };

// common prologue of kernel1 / kernel2, cl2.cl:5011-5057
inline frag_geom frag_prologue(const f4* cut, uint32_t ctri, float rconst, float ewidth, float eheight) {
    frag_geom g;
    f3 tp[3];
    for (int k = 0; k < 3; k++) tp[k] = {cut[(size_t)ctri * 3 + k].x, cut[(size_t)ctri * 3 + k].y, cut[(size_t)ctri * 3 + k].z};
    calc_min_max(tp, ewidth, eheight, g.mm);
    g.xpv = {roundf(tp[0].x), roundf(tp[1].x), roundf(tp[2].x)};
    g.ypv = {roundf(tp[0].y), roundf(tp[1].y), roundf(tp[2].y)};
    f3 depths = {tp[0].z / DEPTH_FAR, tp[1].z / DEPTH_FAR, tp[2].z / DEPTH_FAR};
    depths = {1.0f / depths.x, 1.0f / depths.y, 1.0f / depths.z};
    interpolate_get_const(depths, g.xpv, g.ypv, rconst, &g.A, &g.B, &g.C);
    return g;
}

// ---- kernel1, cl2.cl:4986-5127 ------------------------------------------------------------------------------------
void kernel1(orc_ctx* c) {
    uint32_t* depth = c->depth[c->cur].data();
    const float ewidth = (float)c->W, eheight = (float)c->H;
    uint64_t samples = 0, sats = 0;
#pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads) reduction(+ : samples, sats)
    for (int64_t id = 0; id < (int64_t)c->n_frags; id++) {
        uint32_t distance = c->frags[id * FRAG_MUL + 1];
        uint32_t ctri = c->frags[id * FRAG_MUL + 2];
        float rconst = as_float(c->frags[id * FRAG_MUL + 3]);
        frag_geom g = frag_prologue(c->cutdown.data(), ctri, rconst, ewidth, eheight);
        scan_fragment(g.mm, OP_SIZE, distance, [&](float x, float y) {
            bool cond = point_in_tri({x, y}, {g.xpv.x, g.ypv.x}, {g.xpv.y, g.ypv.y}, {g.xpv.z, g.ypv.z});
            if (cond) {
                float fmydepth = fmaf(g.A, x, fmaf(g.B, y, g.C));
                float q = U32MAXF / fmydepth;
                uint32_t mydepth = sat_u32(q);
                if (!(q > 0.0f) || q >= U32MAXF) sats++;
                atomic_min_u32(&depth[(int)(y * ewidth) + (int)x], mydepth);
                samples++;
            }
        });
    }
    c->depth_samples = samples;
    c->sat_events += sats;
}

// ---- kernel2, cl2.cl:5391-5546 (canonical winner = highest fragment index, SURVEY.md §7 hard part 3) ----------------
void kernel2(orc_ctx* c) {
    const uint32_t* depth = c->depth[c->cur].data();
    uint32_t* ids = c->ids.data();
    const float ewidth = (float)c->W, eheight = (float)c->H;
    const int W = c->W;
#pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads)
    for (int64_t id = 0; id < (int64_t)c->n_frags; id++) {
        uint32_t distance = c->frags[id * FRAG_MUL + 1];
        uint32_t ctri = c->frags[id * FRAG_MUL + 2];
        float rconst = as_float(c->frags[id * FRAG_MUL + 3]);
        frag_geom g = frag_prologue(c->cutdown.data(), ctri, rconst, ewidth, eheight);
        scan_fragment(g.mm, OP_SIZE, distance, [&](float x, float y) {
            if (x < g.mm[0] || y < g.mm[2]) return;   // extra oob guard, cl2.cl:5503
            bool cond = point_in_tri({x, y}, {g.xpv.x, g.ypv.x}, {g.xpv.y, g.ypv.y}, {g.xpv.z, g.ypv.z});
            if (cond) {
                float fmydepth = fmaf(g.A, x, fmaf(g.B, y, g.C));
                uint32_t mydepth = sat_u32(U32MAXF / fmydepth);
                uint32_t val = depth[(int)y * W + (int)x];
                int c2 = mydepth > val - BUF_ERROR && mydepth < val + BUF_ERROR;   // unsigned wrap kept
                // the id image holds fragment index + 1: 0 = no fragment passed the window (q7: such pixels are shaded like uncovered ones)
                if (c2) atomic_max_u32(&ids[(int)y * W + (int)x], (uint32_t)id + 1u);
            }
        });
    }
}

// ---- shadow passes: prearrange_realtime_shadowing cl2.cl:4420-4636 + kernel1_realtime_shadowing 5130-5246 ----------
struct face_tab { rotsc r[6]; };
face_tab make_face_tab() {       // r_struct, cl2.cl:4487-4511 / 2538-2558, float arithmetic on M_PI = 3.1415927f
    face_tab t;
    const float PI = CL_M_PI;
    float e[6][3] = {{0, 0, 0}, {PI / 2.0f, 0, 0}, {0, PI, 0}, {3.0f * PI / 2.0f, 0, 0}, {0, 3.0f * PI / 2.0f, 0}, {0, PI / 2.0f, 0}};
    for (int k = 0; k < 6; k++) t.r[k] = make_rotsc(e[k][0], e[k][1], e[k][2]);
    return t;
}

void shadow_pass(orc_ctx* c, const float lpos4[4], int only_static, uint32_t* slab, uint32_t pair_base) {
    // (light, face) pair ownership of the sort-first split (include/rr.h rr_config.face_rank/face_world)
    bool owned[6];
    {
        const uint32_t total = 6u * (only_static ? c->n_static : c->n_shadow);
        const uint32_t world = c->cfg.face_world > 1 ? (uint32_t)c->cfg.face_world : 1u;
        const uint32_t chunk = (total + world - 1) / world;
        for (int kk = 0; kk < 6; kk++) owned[kk] = world <= 1 || ((pair_base + kk) / chunk) == (uint32_t)c->cfg.face_rank;
    }
    const uint32_t T = (uint32_t)c->tris.size();
    const float L = (float)c->L;
    const float efov = L / 2.0f;
    const f3 lpos = v3(lpos4);
    const face_tab ft = make_face_tab();
    struct rec { uint32_t face, a; f3 p[3]; float rconst; };
    // Allocation order does not influence the cubemap (atomic_min), so records are kept per triangle.
    uint64_t samples = 0, nfr = 0, sats = 0;
    const int LL = c->L;
#pragma omp parallel for schedule(dynamic, 256) num_threads(c->threads) reduction(+ : samples, nfr, sats)
    for (int64_t id = 0; id < (int64_t)T; id++) {
        const rr_triangle& Tt = c->tris[id];
        int o_id = (int)Tt.vertices[0].object_id;
        const rr_obj_desc& G = c->objs[o_id];
        int feature_flag = G.feature_flag;
        bool is_static = (feature_flag & RR_FEATURE_IS_STATIC) > 0;
        if (!only_static && is_static) continue;    // 4460
        if (only_static && !is_static) continue;    // 4463
        f3 g_world_pos = v3(G.world_pos);
        if (fast_length3(g_world_pos - lpos) > DEPTH_FAR) continue;   // 4472
        f4 q = {G.world_rot_quat[0], G.world_rot_quat[1], G.world_rot_quat[2], G.world_rot_quat[3]};
        int skip_structure[6] = {0, 0, 0, 0, 0, 0};
        for (int kk = 0; kk < 3; kk++) {            // 4520-4539
            f3 rotated = rot_quat(v3(Tt.vertices[kk].pos) * G.scale, q);
            rotated = rotated + g_world_pos;
            skip_structure[ret_cubeface(rotated, lpos)] = 1;
        }
        bool two_sided = (feature_flag & RR_FEATURE_TWO_SIDED) > 0;
        for (int kk = 0; kk < 6; kk++) {
            if (!skip_structure[kk] || !owned[kk]) continue;
            f3 proj[2][3];
            int num = 0;
            full_rotate_quat(v3(Tt.vertices[0].pos), v3(Tt.vertices[1].pos), v3(Tt.vertices[2].pos), proj, &num, lpos, ft.r[kk],
                             g_world_pos, q, G.scale, efov, L, L, c->cfg.depth_icutoff);
            for (int i = 0; i < num; i++) {
                sub_tri s;
                for (int k = 0; k < 3; k++) s.p[k] = proj[i][k];
                classify(s, two_sided, L, L, OP_SIZE_LIGHT);
                if (!s.keep) continue;
                // kernel1_realtime_shadowing for each of this sub-triangle's fragments (5130-5246)
                f3 xpv = {roundf(s.p[0].x), roundf(s.p[1].x), roundf(s.p[2].x)};
                f3 ypv = {roundf(s.p[0].y), roundf(s.p[1].y), roundf(s.p[2].y)};
                f3 depths = {s.p[0].z / DEPTH_FAR, s.p[1].z / DEPTH_FAR, s.p[2].z / DEPTH_FAR};
                depths = {1.0f / depths.x, 1.0f / depths.y, 1.0f / depths.z};
                float A, B, C;
                interpolate_get_const(depths, xpv, ypv, s.rconst, &A, &B, &C);
                for (int a = 0; a < s.n_frag; a++) {
                    nfr++;
                    scan_fragment(s.mm, OP_SIZE_LIGHT, (uint32_t)a, [&](float x, float y) {
                        bool cond = point_in_tri({x, y}, {xpv.x, ypv.x}, {xpv.y, ypv.y}, {xpv.z, ypv.z});
                        if (cond) {
                            float fmydepth = fmaf(A, x, fmaf(B, y, C));
                            float qd = U32MAXF / fmydepth;
                            uint32_t mydepth = sat_u32(qd);
                            if (!(qd > 0.0f) || qd >= U32MAXF) sats++;
                            atomic_min_u32(&slab[(int)(y * L) + (int)x + kk * LL * LL], mydepth);
                            samples++;
                        }
                    });
                }
            }
        }
    }
    c->shadow_samples += samples;
    c->tm.n_shadow_fragments += (uint32_t)nfr;
    c->sat_events += sats;
}

// ---- texture atlas ------------------------------------------------------------------------------------------------
// read_tex_array, cl2.cl:785-821
inline f4 read_tex_array(f2 coords, uint32_t tid, const orc_ctx* c) {
    int nv = (int)c->nums[tid];
    int slice = nv >> 16;
    int which = nv & 0x0000FFFF;
    const float max_tex_size = 2048;
    float width = (float)c->sizes[slice];
    float hnum = floorf(max_tex_size / width);
    float tnumy = floorf((float)which / hnum);
    float tnumx = fmaf(-tnumy, hnum, (float)which);
    coords.x = cl_clamp(coords.x, 0.001f, width - 0.001f);
    coords.y = cl_clamp(coords.y, 0.001f, width - 0.001f);
    float rx = fmaf(tnumx, width, coords.x), ry = fmaf(tnumy, width, coords.y);
    int ix = (int)rx, iy = (int)ry;
    const uint8_t* p = &c->atlas[((size_t)slice * ATLAS_DIM * ATLAS_DIM + (size_t)iy * ATLAS_DIM + ix) * 4];
    return {(float)p[0], (float)p[1], (float)p[2], (float)p[3]};
}

// write_tex_array, cl2.cl:856-889
inline void write_tex_array(const uint32_t to_write[4], f2 coords, uint32_t tid, orc_ctx* c) {
    int nv = (int)c->nums[tid];
    int slice = nv >> 16;
    int which = nv & 0x0000FFFF;
    const float max_tex_size = 2048;
    float width = (float)c->sizes[slice];
    float hnum = floorf(max_tex_size / width);
    float tnumy = floorf((float)which / hnum);
    float tnumx = fmaf(-tnumy, hnum, (float)which);
    float tx = tnumx * width, ty = tnumy * width;
    coords.x = fmodf(coords.x, width);
    coords.y = fmodf(coords.y, width);
    coords.x = cl_clamp(coords.x, 0.001f, width - 0.001f);
    coords.y = cl_clamp(coords.y, 0.001f, width - 0.001f);
    int ix = (int)(tx + coords.x), iy = (int)(ty + coords.y);
    uint8_t* p = &c->atlas[((size_t)slice * ATLAS_DIM * ATLAS_DIM + (size_t)iy * ATLAS_DIM + ix) * 4];
    for (int k = 0; k < 4; k++) p[k] = (uint8_t)to_write[k];    // convert_uchar4, cl2.cl:756
}

// update_gpu_tex, cl2.cl:923-953
void update_gpu_tex(orc_ctx* c, uint32_t tex_id, const uint8_t* rgba, int w, int h, int flip) {
    int slice = (int)(c->nums[tex_id] >> 16);
    float width = (float)c->sizes[slice];
    for (int y0 = 0; y0 < h; y0++)
        for (int x = 0; x < w; x++) {
            int y = y0;
            uint32_t ucol[4];
            for (int k = 0; k < 4; k++) {
                float col = (float)rgba[((size_t)y0 * w + x) * 4 + k] / 255.f;    // read_imagef on CL_UNORM_INT8 (pinned)
                col *= 255.f;
                ucol[k] = (uint32_t)col;                                            // convert_uint4 truncates
            }
            if (flip) y = (int)(width - (float)y);
            write_tex_array(ucol, {(float)x, (float)y}, tex_id, c);
        }
}

// generate_mips cl2.cl:1071-1129 (src = tex_id, dst = tex_id*4 + mipmap_start) and
// generate_mip_mips cl2.cl:1132-1189 (src = proper_id, dst = proper_id + 1); gw/gh = global size = base image size
void mip_pass(orc_ctx* c, uint32_t src_id, uint32_t dst_id, int gw, int gh) {
    int slice = (int)(c->nums[src_id] >> 16);
    float width = (float)c->sizes[slice];
    const float gauss[3][3] = {{1, 2, 1}, {2, 4, 2}, {1, 2, 1}};
    int w2 = (int)(c->nums[dst_id] >> 16);
    float nwidth = (float)c->sizes[w2];
    for (int y = 0; y < gh; y++)
        for (int x = 0; x < gw; x++) {
            if ((float)x >= width || (float)y >= width) continue;
            f4 accum = {0, 0, 0, 0};
            float div = 0.f;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    f4 col = read_tex_array({(float)(x * 2 + i), (float)(y * 2 + j)}, src_id, c);
                    col.w /= 255.f;
                    col.x *= col.w; col.y *= col.w; col.z *= col.w;
                    accum = accum + col * gauss[j + 1][i + 1];
                    div += gauss[j + 1][i + 1];
                }
            accum = accum / div;
            if (accum.w > 0.00000001f) { accum.x /= accum.w; accum.y /= accum.w; accum.z /= accum.w; }
            accum.w *= 255.f;
            f2 yx = {((float)(x * 2) / width) * nwidth, ((float)(y * 2) / width) * nwidth};
            if (yx.x >= nwidth || yx.y >= nwidth) continue;
            uint32_t out[4] = {sat_u32(accum.x), sat_u32(accum.y), sat_u32(accum.z), sat_u32(accum.w)};
            write_tex_array(out, yx, dst_id, c);
        }
}

// read_tex_array_all_precalculated, cl2.cl:823-851
inline f4 read_tex_pre(f2 coords, int which, int slice, float width, const orc_ctx* c) {
    const float imax_tex_size = 1.f / 2048;
    float ihnum = width * imax_tex_size;
    float tnumy = floorf((float)which * ihnum);
    float tnumx = (float)which - tnumy / ihnum;
    coords.x = cl_clamp(coords.x, 0.001f, width - 0.001f);
    coords.y = cl_clamp(coords.y, 0.001f, width - 0.001f);
    float rx = fmaf(tnumx, width, coords.x), ry = fmaf(tnumy, width, coords.y);
    int ix = (int)rx, iy = (int)ry;
    const uint8_t* p = &c->atlas[((size_t)slice * ATLAS_DIM * ATLAS_DIM + (size_t)iy * ATLAS_DIM + ix) * 4];
    return {(float)p[0], (float)p[1], (float)p[2], (float)p[3]};
}

// return_bilinear_col_all_precalculated, cl2.cl:1426-1455
inline f4 bilinear_pre(f2 mcoord, int which, int slice, float width, const orc_ctx* c) {
    f2 pos = {floorf(mcoord.x), floorf(mcoord.y)};
    f2 co[4] = {{pos.x, pos.y}, {pos.x + 1, pos.y}, {pos.x, pos.y + 1}, {pos.x + 1, pos.y + 1}};
    f4 col[4];
    for (int i = 0; i < 4; i++) col[i] = read_tex_pre(co[i], which, slice, width, c);
    f2 uvratio = mcoord - pos;
    f2 buvr = {1.f - uvratio.x, 1.f - uvratio.y};
    return mad4(col[0], buvr.x, col[1] * uvratio.x) * buvr.y + mad4(col[2], buvr.x, col[3] * uvratio.x) * uvratio.y;
}

// texture_mod, cl2.cl:1457-1468
inline f2 texture_mod(f2 in) {
    f2 vtm = in;
    vtm.x = vtm.x >= 1 ? 1.0f - (vtm.x - floorf(vtm.x)) : vtm.x;
    vtm.y = vtm.y >= 1 ? 1.0f - (vtm.y - floorf(vtm.y)) : vtm.y;
    vtm.x = vtm.x < 0 ? 1.0f + fabsf(vtm.x) - fabsf(floorf(vtm.x)) : vtm.x;
    vtm.y = vtm.y < 0 ? 1.0f + fabsf(vtm.y) - fabsf(floorf(vtm.y)) : vtm.y;
    return vtm;
}

// log2_approx, cl2.cl:1498-1505
inline float log2_approx(float val) {
    int x = (int)as_uint(val);
    float log_2 = (float)(((x >> 23) & 255) - 128);
    x &= ~(255 << 23);
    x += 127 << 23;
    float v = as_float((uint32_t)x);
    log_2 += ((-0.3358287811f) * v + 2.0f) * v - 0.65871759316667f;
    return log_2;
}

// texture_filter_diff, cl2.cl:1511-1573
inline f4 texture_filter_diff(f2 vt, f2 vtdiff, int tid2, uint32_t mip_start, const orc_ctx* c) {
    int nv = (int)c->nums[tid2];
    int slice = nv >> 16;
    int tsize = (int)c->sizes[slice];
    f2 vtm = texture_mod(vt);
    f2 vs = vtdiff * (float)tsize;
    float worst = sqrtf(vs.x * vs.x + vs.y * vs.y);
    float worst_id_frac = log2_approx(worst);
    worst_id_frac = cl_max(worst_id_frac, 0.f);
    float mip_lower = floorf(worst_id_frac);
    mip_lower = cl_clamp(mip_lower, 0.f, (float)MIP_LEVELS);
    float fmd = worst_id_frac - mip_lower;
    int tid_lower = mip_lower == 0 ? tid2 : (int)(mip_lower - 1 + (float)mip_start + (float)(tid2 * MIP_LEVELS));
    int tid_higher = (int)(cl_clamp(mip_lower, 0.f, MIP_LEVELS - 1.f) + (float)mip_start + (float)(tid2 * MIP_LEVELS));
    int lower_nv = (int)c->nums[tid_lower], higher_nv = (int)c->nums[tid_higher];
    int slice_lower = lower_nv >> 16, slice_higher = higher_nv >> 16;
    int which_lower = lower_nv & 0x0000FFFF, which_higher = higher_nv & 0x0000FFFF;
    float size_lower = (float)c->sizes[slice_lower], size_higher = (float)c->sizes[slice_higher];
    f4 col1 = bilinear_pre(vtm * size_lower, which_lower, slice_lower, size_lower, c);
    f4 col2 = bilinear_pre(vtm * size_higher, which_higher, slice_higher, size_higher, c);
    f4 final_col = col1 + (col2 - col1) * fmd;    // mix()
    const float i255 = 1.f / 255.f;
    return final_col * i255;
}

// get_barycentric, cl2.cl:5372-5384
inline void get_barycentric(f3 p, f3 a, f3 b, f3 cc, float* u, float* v, float* w) {
    f3 v0 = b - a, v1 = cc - a, v2 = p - a;
    float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    *v = (d11 * d20 - d01 * d21) / denom;
    *w = (d00 * d21 - d01 * d20) / denom;
    *u = 1.0f - *v - *w;
}

// gamma_transform_approx / gamma_inverse_approx, cl2.cl:5763-5782 (per channel)
inline float gamma_fwd(float s) { return 0.012522878f * s + 0.682171111f * s * s + 0.305306011f * s * s * s; }
inline float gamma_inv(float c) {
    float S1 = sqrtf(c), S2 = sqrtf(S1), S3 = sqrtf(S2);
    return 0.585122381f * S1 + 0.783140355f * S2 - 0.368262736f * S3;
}

// generate_ssao, cl2.cl:2194-2260
inline float generate_ssao(int sx, int sy, const uint32_t* depth_buffer, const orc_ctx* c) {
    const int W = c->W, H = c->H;
    uint32_t seed1 = wang_hash((uint32_t)sx + (uint32_t)W * (uint32_t)H * (uint32_t)sy);   // int arithmetic, two's-complement wrap (q12)
    uint32_t seed2 = rand_xorshift(seed1);
    float foffset = (float)seed2 / U32MAXF;
    float depth = ((float)depth_buffer[sy * W + sx] / U32MAXF) * DEPTH_FAR;
    float rad = c->cfg.ssao_rad;
    rad += foffset / 2.f;
    float world_rad = rad * c->fov / depth;
    const int samples = 2;
    f2 fspos = {(float)sx, (float)sy};
    float acc = 0.f;
    for (int y = -samples; y <= samples; y++)
        for (int x = -samples; x <= samples; x++) {
            f2 offset = {(float)x * world_rad, (float)y * world_rad};
            offset = {roundf(offset.x), roundf(offset.y)};
            f2 world = fspos + offset;
            world.x = cl_clamp(world.x, 1.f, (float)W - 2.f);
            world.y = cl_clamp(world.y, 1.f, (float)H - 2.f);
            float d2 = ((float)depth_buffer[((int)world.y) * W + (int)world.x] / U32MAXF) * DEPTH_FAR;
            for (float z = -samples; z <= samples; z += 1.f)
                if (d2 > depth + z) acc += 1.f;
        }
    acc /= powf(samples * 2.f + 1.f, 3.f);
    return 1.f - (1.f - acc) / c->cfg.ssao_div;
}

// bilinear_interpolate, cl2.cl:2144-2158
inline float bilinear_interpolate(f2 coord, const float values[4]) {
    float mx = coord.x - 0.5f, my = coord.y - 0.5f;
    f2 uvratio = {mx - floorf(mx), my - floorf(my)};
    f2 buvr = {1.0f - uvratio.x, 1.0f - uvratio.y};
    return (values[0] * buvr.x + values[1] * uvratio.x) * buvr.y + (values[2] * buvr.x + values[3] * uvratio.x) * uvratio.y;
}

// generate_hard_occlusion, cl2.cl:2536-2701 (SMOOTH_SHADOWS branch)
inline float generate_hard_occlusion(f3 lpos, f3 normal, f3 position_to_light, const uint32_t* light_depth_buffer, int which_cubeface,
                                     f3 back_rotated, int shnum, const orc_ctx* c, const face_tab& ft) {
    const int L = c->L;
    const float Lf = (float)L;
    position_to_light = fast_normalize3(position_to_light);
    f3 local_pos = rot(back_rotated, lpos, ft.r[which_cubeface]);
    f3 pp = depth_project_singular(local_pos, Lf, Lf, Lf / 2.0f);
    float dpth = pp.z;
    const uint32_t* ldepth_map = &light_depth_buffer[(size_t)(which_cubeface + shnum * 6) * L * L];
    pp.x = cl_clamp(pp.x, 3.f, Lf - 4.f);
    pp.y = cl_clamp(pp.y, 3.f, Lf - 4.f);
    int ipx = (int)pp.x, ipy = (int)pp.y;
    float acos_res = rational_acos(cl_clamp(dot3(normal, position_to_light), 0.05f, 0.95f));
    float bias = c->cfg.shadow_bias * tanf(acos_res);
    bias = cl_clamp(bias, 0.1f * c->cfg.shadow_bias, powf(c->cfg.shadow_bias, c->cfg.shadow_exp));
    float shadow = 0.f;
    int conditions[16];
    for (int y = -1; y <= 2; y++)
        for (int x = -1; x <= 2; x++) {
            float ldp1 = ((float)ldepth_map[(ipy + y) * L + ipx + x] / U32MAXF) * DEPTH_FAR;
            conditions[(y + 1) * 4 + x + 1] = dpth > ldp1 + bias ? 1 : 0;
        }
    for (int y = -1; y <= 1; y++)
        for (int x = -1; x <= 1; x++) {
            float vals[4];
            vals[0] = (float)conditions[(y + 1) * 4 + x + 1];
            vals[1] = (float)conditions[(y + 1) * 4 + x + 2];
            vals[2] = (float)conditions[(y + 2) * 4 + x + 1];
            vals[3] = (float)conditions[(y + 2) * 4 + x + 2];
            shadow += bilinear_interpolate({pp.x + 0.5f + (float)x, pp.y + 0.5f + (float)y}, vals);
        }
    shadow /= 9.f;
    return shadow;
}

// float_to_short / encode_normal, cl2.cl:5588-5628
inline uint16_t to_ushort_sat(float v) { if (!(v > 0.f)) return 0; if (v >= 65535.f) return 65535; return (uint16_t)v; }
inline void encode_normal(f3 val, uint16_t out[2]) {
    float len_sq = val.x * val.x + val.y * val.y;
    if (len_sq < 0.0001f) val.x = 0.01f;
    float l = sqrtf(val.x * val.x + val.y * val.y);
    float k = sqrtf(cl_max(val.z * 0.5f + 0.5f, 0.f));
    f2 r = {(val.x / l) * k, (val.y / l) * k};
    out[0] = to_ushort_sat(((r.x + 1) / 2) * 65536 - 1);
    out[1] = to_ushort_sat(((r.y + 1) / 2) * 65536 - 1);
}

inline f4 get_vertex_col(const rr_vertex& v) {   // cl2.cl:5676-5686
    f4 rgba = {(float)(v.vertex_col >> 24), (float)((v.vertex_col >> 16) & 0xFF), (float)((v.vertex_col >> 8) & 0xFF), (float)(v.vertex_col & 0xFF)};
    return rgba / 255.f;
}

// ---- kernel3, cl2.cl:5795-6408 -------------------------------------------------------------------------------------
void kernel3(orc_ctx* c, const float c_pos4[4], const rotsc& crot, const float clear[4]) {
    const int W = c->W, H = c->H;
    const float FOV = c->fov;
    const uint32_t* depth_buffer = c->depth[c->cur].data();
    uint32_t* to_clear = c->depth[c->cur ^ 1].data();
    const f3 camera_pos = v3(c_pos4);
    const f3 zero3 = {0, 0, 0};
    const face_tab ft = make_face_tab();
    const bool linear = c->cfg.test_linear && c->cfg.use_linear_rendering;
    const int y0 = c->cfg.band_y1 > c->cfg.band_y0 ? c->cfg.band_y0 : 0;
    const int y1 = c->cfg.band_y1 > c->cfg.band_y0 ? c->cfg.band_y1 : H;
#pragma omp parallel for schedule(dynamic, 4) num_threads(c->threads)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            to_clear[px] = 0xFFFFFFFFu;                              // 5820
            if (y < y0 || y >= y1) continue;
            const uint32_t ft_depth = depth_buffer[px];
            float* out = &c->colour[px * 4];
            if (ft_depth == 0xFFFFFFFFu || c->ids[px] == 0u) {       // 5835-5862; unresolved id (never written by kernel2): same treatment
                for (int k = 0; k < 4; k++) out[k] = clear[k];
                continue;
            }
            uint32_t idv = c->ids[px] - 1u;
            uint32_t tri_global = c->frags[(size_t)idv * FRAG_MUL + 0];
            uint32_t ctri = c->frags[(size_t)idv * FRAG_MUL + 2];
            float rconst = as_float(c->frags[(size_t)idv * FRAG_MUL + 3]);
            int o_id = (int)c->frags[(size_t)idv * FRAG_MUL + 4];
            const rr_triangle& T = c->tris[tri_global];
            f3 p1 = v3(T.vertices[0].pos), p2 = v3(T.vertices[1].pos), p3 = v3(T.vertices[2].pos);
            f2 vt1 = {T.vertices[0].vt[0], T.vertices[0].vt[1]}, vt2 = {T.vertices[1].vt[0], T.vertices[1].vt[1]},
               vt3 = {T.vertices[2].vt[0], T.vertices[2].vt[1]};
            f3 n1 = v3(T.vertices[0].normal), n2 = v3(T.vertices[1].normal), n3 = v3(T.vertices[2].normal);
            const rr_obj_desc& G = c->objs[o_id];
            const f4 Gq = {G.world_rot_quat[0], G.world_rot_quat[1], G.world_rot_quat[2], G.world_rot_quat[3]};
            const f3 Gpos = v3(G.world_pos);
            p1 = p1 * G.scale; p2 = p2 * G.scale; p3 = p3 * G.scale;
            float ldepth = ((float)ft_depth / U32MAXF) * DEPTH_FAR;                 // 5897
            float actual_depth = ldepth;
            f3 local_position = {(((float)x - W / 2.0f) * actual_depth / FOV), (((float)y - H / 2.0f) * actual_depth / FOV), actual_depth};
            f3 global_position = back_rot(local_position, zero3, crot);
            global_position = global_position + camera_pos;
            f3 object_local = global_position - Gpos;
            object_local = back_rot_quat(object_local, Gq);
            float l1, l2, l3;
            get_barycentric(object_local, p1, p2, p3, &l1, &l2, &l3);
            f2 vt = mad2(vt1, l1, mad2(vt2, l2, vt3 * l3));
            f3 normal = mad3(n1, l1, mad3(n2, l2, n3 * l3));
            normal = rot_quat(normal, Gq);
            bool has_colour_already = false;
            f4 vertex_col = {0, 0, 0, 0};
            if (T.vertices[0].vertex_col != 0) {
                has_colour_already = true;
                vertex_col = mad4(get_vertex_col(T.vertices[0]), l1, mad4(get_vertex_col(T.vertices[1]), l2, get_vertex_col(T.vertices[2]) * l3));
            }
            f3 tris_proj[3];
            for (int k = 0; k < 3; k++) tris_proj[k] = {c->cutdown[(size_t)ctri * 3 + k].x, c->cutdown[(size_t)ctri * 3 + k].y, c->cutdown[(size_t)ctri * 3 + k].z};

            // get_vtdiff, cl2.cl:5691-5760
            f2 vtdiff;
            {
                float fx = (float)x, fy = (float)y;
                f3 xpv = {roundf(tris_proj[0].x), roundf(tris_proj[1].x), roundf(tris_proj[2].x)};
                f3 ypv = {roundf(tris_proj[0].y), roundf(tris_proj[1].y), roundf(tris_proj[2].y)};
                f3 depths = {1.0f / tris_proj[0].z, 1.0f / tris_proj[1].z, 1.0f / tris_proj[2].z};
                float DA, DB, DC;
                interpolate_get_const(depths, xpv, ypv, rconst, &DA, &DB, &DC);
                float dmx = fmaf(DA, fx + 1, fmaf(DB, fy, DC));
                float dmy = fmaf(DA, fx, fmaf(DB, fy + 1, DC));
                f3 lmx = {(fx + 1 - W / 2.f) / FOV, (fy - H / 2.f) / FOV, 1};
                f3 lmy = {(fx - W / 2.f) / FOV, (fy + 1 - H / 2.f) / FOV, 1};
                lmx = lmx / dmx;
                lmy = lmy / dmy;
                f3 gmx = back_rot_quat(back_rot(lmx, zero3, crot) + camera_pos - Gpos, Gq);
                f3 gmy = back_rot_quat(back_rot(lmy, zero3, crot) + camera_pos - Gpos, Gq);
                float lx1, lx2, lx3, ly1, ly2, ly3;
                get_barycentric(gmx, p1, p2, p3, &lx1, &lx2, &lx3);
                get_barycentric(gmy, p1, p2, p3, &ly1, &ly2, &ly3);
                f2 vtx = mad2(vt1, lx1, mad2(vt2, lx2, vt3 * lx3));
                f2 vty = mad2(vt1, ly1, mad2(vt2, ly2, vt3 * ly3));
                f2 vdx = vtx - vt, vdy = vty - vt;
                vdx = {fabsf(vdx.x), fabsf(vdx.y)};
                vdy = {fabsf(vdy.x), fabsf(vdy.y)};
                const float mip_bias = 1.f / c->cfg.mip_bias;
                vtdiff = f2{vdx.x + vdy.x, vdx.y + vdy.y} * mip_bias;
            }
            f4 col;
            if (!has_colour_already) col = texture_filter_diff(vt, vtdiff, (int)G.tid, c->mipmap_start, c);
            else col = vertex_col;
            if (linear) { col.x = gamma_fwd(col.x); col.y = gamma_fwd(col.y); col.z = gamma_fwd(col.z); }   // 5960-5963 (w kept)

            uint32_t seed1 = wang_hash((uint32_t)x + (uint32_t)y * (uint32_t)W * (uint32_t)H);   // 5965, wraps mod 2^32
            uint32_t seed2 = rand_xorshift(seed1), seed3 = rand_xorshift(seed2), seed4 = rand_xorshift(seed3);
            f3 rseed = {(float)seed2 / U32MAXF, (float)seed3 / U32MAXF, (float)seed4 / U32MAXF};
            rseed = {(rseed.x - 0.5f) * 2, (rseed.y - 0.5f) * 2, (rseed.z - 0.5f) * 2};

            int shnum = 0, static_num = 0;
            int num_lights = (int)c->lights.size();
            f3 diffuse_sum = {0, 0, 0}, specular_sum = {0, 0, 0};
            f3 l2p = camera_pos - global_position;
            l2p = fast_normalize3(l2p);
            int feature_flag = G.feature_flag;
            bool is_two_sided = (feature_flag & RR_FEATURE_TWO_SIDED) > 0;
            bool receives_dynamic_shadows = !((feature_flag & RR_FEATURE_NO_DYNAMIC_SHADOWS) > 0);
            int is_front = backface_cull_expanded(tris_proj[0], tris_proj[1], tris_proj[2]);
            int flip_normals = !is_front && is_two_sided == 1;
            if (flip_normals) normal = -normal;
            float ssao = c->cfg.no_ssao ? 1.f : generate_ssao(x, y, depth_buffer, c);
            normal = fast_normalize3(normal);
            f3 lighting_normal = normal + rseed / 100.f;
            lighting_normal = fast_normalize3(lighting_normal);
            float ambient = c->cfg.ambient;
            if (linear) ambient = gamma_fwd(c->cfg.ambient);

            for (int i = 0; i < num_lights; i++) {                                       // 6115-6278
                const rr_light& l = c->lights[i];
                const f3 lpos = v3(l.pos);
                f3 point_to_light = lpos - global_position;
                float occlusion = 1;
                if (l.shadow && l.is_static) {
                    int which_cubeface = ret_cubeface(global_position, lpos);
                    occlusion = 1.f - generate_hard_occlusion(lpos, normal, point_to_light, c->shadow_static.data(), which_cubeface,
                                                              global_position, static_num, c, ft);
                    static_num++;
                }
                float distance = fast_length3(point_to_light);
                float illumination = l.brightness / powf((distance / l.radius) + 1, 2.f);
                const float cutoff = 0.1f;
                illumination -= cutoff;
                illumination *= 1.f / (1.f - cutoff);
                if (illumination <= 0) continue;
                f3 light_col = {l.col[0], l.col[1], l.col[2]};
                if (linear) light_col = {gamma_fwd(light_col.x), gamma_fwd(light_col.y), gamma_fwd(light_col.z)};
                if (l.shadow && receives_dynamic_shadows) {
                    int which_cubeface = ret_cubeface(global_position, lpos);
                    float dyn_occlusion = 1.f - generate_hard_occlusion(lpos, normal, point_to_light, c->shadow_dyn.data(), which_cubeface,
                                                                        global_position, shnum, c, ft);
                    occlusion = cl_min(occlusion, dyn_occlusion);
                    shnum++;
                }
                point_to_light = fast_normalize3(point_to_light);
                float light = dot3(point_to_light, lighting_normal);
                light *= occlusion;
                light = cl_max(light, 0.f);
                float diffuse = (1.0f - ambient) * light;
                diffuse_sum = diffuse_sum + light_col * ((diffuse + ambient) * l.diffuse * G.diffuse * illumination);
                f3 Hh = fast_normalize3(l2p + point_to_light);
                f3 N = normal;
                const float kS = 0.4f;
                float ndh = cl_max(0.f, dot3(N, Hh));
                float ndv = cl_max(0.f, dot3(N, l2p));
                float vdh = cl_max(0.f, dot3(l2p, Hh));
                float ndl = cl_max(0.f, dot3(N, point_to_light));
                const float F0 = 0.4f;
                float fresnel = F0 + (1 - F0) * powf((1.f - vdh), 5.f);
                float rough = cl_clamp(1.f - G.specular, 0.001f, 10.f);
                const float gauss_constant = 0.8346f;
                float alpha = rational_acos(ndh);
                float microfacet = gauss_constant * expf(-alpha * alpha / (rough * rough));
                float sv = 2 * ndh / vdh;
                float c1 = sv * ndv, c2 = sv * ndl;
                float geometric = cl_min(cl_min(1.f, c1), c2);
                float spec = (fresnel * microfacet * geometric) / (CL_M_PI * ndv);
                specular_sum = specular_sum + light_col * (spec * kS * illumination) * G.spec_mult;
                specular_sum = {cl_max(specular_sum.x, 0.f), cl_max(specular_sum.y, 0.f), cl_max(specular_sum.z, 0.f)};
                specular_sum = specular_sum * occlusion;
            }
            specular_sum = specular_sum * ssao;
            diffuse_sum = diffuse_sum * ssao;
            const float reflected_surface_colour = 0.7f;
            f3 colclamp = f3{col.x, col.y, col.z} + f3{0, 0, 0} + specular_sum * reflected_surface_colour;
            f3 final_col = {fmaf(colclamp.x, diffuse_sum.x, specular_sum.x * (1.f - reflected_surface_colour)),
                            fmaf(colclamp.y, diffuse_sum.y, specular_sum.y * (1.f - reflected_surface_colour)),
                            fmaf(colclamp.z, diffuse_sum.z, specular_sum.z * (1.f - reflected_surface_colour))};
            if (linear) final_col = {gamma_inv(final_col.x), gamma_inv(final_col.y), gamma_inv(final_col.z)};
            final_col = {cl_clamp(final_col.x, 0.f, 1.f), cl_clamp(final_col.y, 0.f, 1.f), cl_clamp(final_col.z, 0.f, 1.f)};
            out[0] = final_col.x; out[1] = final_col.y; out[2] = final_col.z; out[3] = col.w;
            encode_normal(normal, &c->normals[px * 2]);
        }
    }
    // headless target quantiser (q15): q = (uint8)(clamp(c,0,1)*255 + 0.5f)
#pragma omp parallel for schedule(static) num_threads(c->threads)
    for (int64_t i = 0; i < (int64_t)W * H * 4; i++) {
        int yy = (int)((i / 4) / W);
        if (yy < y0 || yy >= y1) continue;
        c->rgba8[i] = (uint8_t)(cl_clamp(c->colour[i], 0.f, 1.f) * 255.f + 0.5f);
    }
}

// ---- do_pseudo_aa, cl2.cl:6437-6657 (post pass on the G-buffer) ---------------------------------
// short_to_float / decode_normal, cl2.cl:5598-5647
inline f3 decode_normal(const uint16_t* s) {
    f2 val = {(float)s[0], (float)s[1]};
    val = {val.x / 65535.f, val.y / 65535.f};
    val = {val.x * 2.f, val.y * 2.f};
    val = {val.x - 1.f, val.y - 1.f};
    f3 ret;
    const float d = val.x * val.x + val.y * val.y;
    ret.z = d * 2.f - 1.f;
    const float l = sqrtf(d);
    const float k = sqrtf(cl_max(1.f - ret.z * ret.z, 0.f));
    ret.x = (val.x / l) * k;
    ret.y = (val.y / l) * k;
    return ret;
}

// The reference runs this in place (screen == in_screen == gl_screen[1], engine.cpp:1854-1856), so a pixel may read
// neighbours the pass has already replaced: the outcome depends on scheduling. Canonical here: every read sees the frame
// kernel3 produced. Input is the headless RGBA8 target (c / 255.f), output goes through the same quantiser. The id /
// object lookups, avg_depth and avg_normal of the reference only feed code under REDUCED_AA / #if 0 and are not restated.
void pseudo_aa(orc_ctx* c) {
    const int W = c->W, H = c->H;
    const uint32_t* depth_buffer = c->depth[c->cur].data();
    const std::vector<uint8_t> in = c->rgba8;
    const float AA_angle_degrees = 20.f;
    const float aa_arg = AA_angle_degrees * 2 * CL_M_PI / 360.f;
    const float AA_angle_cosrad = (float)cos((double)aa_arg);          // cos() pinned: double on the host, rounded to float
    const float depth_bound = 100;
#pragma omp parallel for schedule(dynamic, 4) num_threads(c->threads)
    for (int y = 1; y < H - 1; y++) {
        for (int x = 1; x < W - 1; x++) {
            const size_t px = (size_t)y * W + x;
            f3 my_normal = fast_normalize3(decode_normal(&c->normals[px * 2]));
            const uint32_t my_depth_raw = depth_buffer[px];
            if (my_depth_raw == 0xFFFFFFFFu) continue;
            const float my_depth = ((float)my_depth_raw / U32MAXF) * DEPTH_FAR;             // idcalc
            int num_x[2] = {0, 0}, num_y[2] = {0, 0}, num_corner[2] = {0, 0};
            f3 my_accum[2] = {{0, 0, 0}, {0, 0, 0}}, their_accum[2] = {{0, 0, 0}, {0, 0, 0}};
            float my_samples[2] = {0, 0}, their_samples[2] = {0, 0};
            for (int j = -1; j < 2; j++) {
                for (int i = -1; i < 2; i++) {
                    if (i == 0 && j == 0) continue;
                    const size_t q = (size_t)(y + j) * W + x + i;
                    const float depth = ((float)depth_buffer[q] / U32MAXF) * DEPTH_FAR;
                    const f3 found_normal = decode_normal(&c->normals[q * 2]);
                    const f3 val = {(float)in[q * 4] / 255.f, (float)in[q * 4 + 1] / 255.f, (float)in[q * 4 + 2] / 255.f};
                    const int tests[2] = {fabsf(depth - my_depth) > depth_bound, dot3(my_normal, found_normal) < AA_angle_cosrad};
                    for (int kk = 0; kk < 2; kk++) {
                        if (tests[kk]) {
                            if (i == j || i == -j) num_corner[kk]++;
                            else if (i == 1 || i == -1) num_x[kk]++;
                            else if (j == 1 || j == -1) num_y[kk]++;
                            their_accum[kk] = their_accum[kk] + val;
                            their_samples[kk] += 1.f;
                        } else {
                            my_accum[kk] = my_accum[kk] + val;
                            my_samples[kk] += 1.f;
                        }
                    }
                }
            }
            for (int kk = 0; kk < 2; kk++) {
                if (num_x[kk] == 1 && num_y[kk] == 1 && num_corner[kk] >= 1) {
                    const float wm = 0.65f, wt = 0.35f;
                    const f3 ma = my_accum[kk] / my_samples[kk], ta = their_accum[kk] / their_samples[kk];
                    const f3 accum = ma * wm + ta * wt;
                    const float o[4] = {accum.x, accum.y, accum.z, 1.f};
                    for (int k = 0; k < 4; k++) c->rgba8[px * 4 + k] = (uint8_t)(cl_clamp(o[k], 0.f, 1.f) * 255.f + 0.5f);
                    break;
                }
            }
        }
    }
}

// ---- CLK_FILTER_LINEAR read of the colour target (OpenCL 1.2 §8.2, unnormalised coordinates) -------------------------
// (u, v) -> i0 = floor(u - 0.5), a = frac(u - 0.5), likewise j0 / b; T = texel / 255;
// result = (1-a)(1-b) T[i0,j0] + a(1-b) T[i0+1,j0] + (1-a)b T[i0,j0+1] + ab T[i0+1,j0+1].
// Taps are clamped to the image: that is CLK_ADDRESS_CLAMP_TO_EDGE (godrays); do_motion_blur's CLK_ADDRESS_NONE reads are
// in range except for the +1 tap on the last column / row, which the specification leaves undefined — clamped here.
// The reference's image is float (GL RGBA16 / half); the headless target is RGBA8, so the taps are the quantised colours.
inline f4 sample_linear_rgba8(const std::vector<uint8_t>& img, int W, int H, float u, float v) {
    const float fu = u - 0.5f, fv = v - 0.5f;
    const float i0f = floorf(fu), j0f = floorf(fv);
    const float a = fu - i0f, b = fv - j0f;
    const int i0 = (int)cl_clamp(i0f, 0.f, (float)(W - 1)), i1 = (int)cl_clamp(i0f + 1.f, 0.f, (float)(W - 1));
    const int j0 = (int)cl_clamp(j0f, 0.f, (float)(H - 1)), j1 = (int)cl_clamp(j0f + 1.f, 0.f, (float)(H - 1));
    auto T = [&](int i, int j) { const uint8_t* t = &img[((size_t)j * W + i) * 4]; return f4{t[0] / 255.f, t[1] / 255.f, t[2] / 255.f, t[3] / 255.f}; };
    const f4 t00 = T(i0, j0), t10 = T(i1, j0), t01 = T(i0, j1), t11 = T(i1, j1);
    return (t00 * (1.f - a) + t10 * a) * (1.f - b) + (t01 * (1.f - a) + t11 * a) * b;
}
inline void store_rgba8(std::vector<uint8_t>& img, size_t px, f4 col) {
    const float o[4] = {col.x, col.y, col.z, col.w};
    for (int k = 0; k < 4; k++) img[px * 4 + k] = (uint8_t)(cl_clamp(o[k], 0.f, 1.f) * 255.f + 0.5f);
}

// ---- do_motion_blur, cl2.cl:6714-6860 ------------------------------------------------------------------------------------
// in_screen = gl_screen[1], back_screen = gl_screen[0] (engine.cpp:1520-1521): kernel3 wrote the frame to both, so a pixel this
// pass does not write keeps the frame's colour. Covered pixels whose id did not resolve are treated like uncovered ones
// (the reference would follow a stale id, q7). The kernel also advances the motion history of every object it sees
// (old_world_pos_1/2 ping-pong on frame_id parity, racing identical stores): applied after the pass here.
void motion_blur(orc_ctx* c, float strength, float camera_contribution) {
    const int W = c->W, H = c->H;
    const float Wf = (float)W, Hf = (float)H, fov = c->fov;
    const uint32_t* depth_buffer = c->depth[c->cur].data();
    const std::vector<uint8_t> in = c->rgba8;
    const rotsc crot = make_rotsc(c->cam_rot[0], c->cam_rot[1], c->cam_rot[2]), crot_old = make_rotsc(c->cam_rot_old[0], c->cam_rot_old[1], c->cam_rot_old[2]);
    const f3 cpos = v3(c->cam_pos), cpos_old = v3(c->cam_pos_old);
    const uint32_t frame_id = c->frame_id;
    std::vector<uint8_t> seen(c->objs.size(), 0);
#pragma omp parallel for schedule(dynamic, 4) num_threads(c->threads)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            const uint32_t dbuf_val = depth_buffer[px];
            if (dbuf_val == 0xFFFFFFFFu) continue;
            const uint32_t idv = c->ids[px];
            if (idv == 0u || idv > c->n_frags) continue;
            const uint32_t o_id = c->frags[(size_t)(idv - 1u) * FRAG_MUL + 4];
            if (o_id >= c->objs.size()) continue;
            const rr_obj_desc& G = c->objs[o_id];
            const float actual_depth = ((float)dbuf_val / U32MAXF) * DEPTH_FAR;
            const f3 local_position = {((x - Wf / 2.0f) * actual_depth / fov), ((y - Hf / 2.0f) * actual_depth / fov), actual_depth};
            f3 global_position = back_rot(local_position, {0, 0, 0}, crot);
            global_position = global_position + cpos;
            f3 object_local = global_position - v3(G.world_pos);
            object_local = back_rot_quat(object_local, f4{G.world_rot_quat[0], G.world_rot_quat[1], G.world_rot_quat[2], G.world_rot_quat[3]});
            const float* owp = (frame_id & 1) == 0 ? G.old_world_pos_1 : G.old_world_pos_2;
            const float* owq = (frame_id & 1) == 0 ? G.old_world_rot_quat_1 : G.old_world_rot_quat_2;
            seen[o_id] = 1;
            f3 last_frame_pos = rot_quat(object_local, f4{owq[0], owq[1], owq[2], owq[3]});
            last_frame_pos = last_frame_pos + v3(owp);
            f3 last_frame_no_camera = rot(last_frame_pos, cpos, crot);
            last_frame_pos = rot(last_frame_pos, cpos_old, crot_old);
            last_frame_no_camera = depth_project_singular(last_frame_no_camera, Wf, Hf, fov);
            last_frame_pos = depth_project_singular(last_frame_pos, Wf, Hf, fov);
            if (last_frame_pos.z < (float)c->cfg.depth_icutoff) {
                store_rgba8(c->rgba8, px, sample_linear_rgba8(in, W, H, (float)x + 0.5f, (float)y + 0.5f));
                continue;
            }
            const f2 current_screen_pos = {(float)x, (float)y};
            f2 to_me_vector = current_screen_pos - f2{last_frame_pos.x, last_frame_pos.y};
            const f2 to_me_nocamera = current_screen_pos - f2{last_frame_no_camera.x, last_frame_no_camera.y};
            to_me_vector = to_me_vector * camera_contribution + to_me_nocamera * (1.f - camera_contribution);
            to_me_vector = to_me_vector * strength;
            int n = (int)(cl_max(fabsf(to_me_vector.x), fabsf(to_me_vector.y)) + 1);
            const int bound = 50;
            if (n > bound) {
                to_me_vector = {to_me_vector.x / (float)n, to_me_vector.y / (float)n};
                to_me_vector = to_me_vector * (float)bound;
                n = bound;
            }
            f2 diff = {0, 0};
            if (n != 0) diff = {to_me_vector.x / (float)n, to_me_vector.y / (float)n};
            f2 current = current_screen_pos - f2{to_me_vector.x / 2.f, to_me_vector.y / 2.f};
            f4 accum = {0, 0, 0, 0};
            float fcount = 0;
            for (int i = 0; i < n; i++, current = current + diff) {
                if (current.x < 0 || current.x >= Wf || current.y < 0 || current.y >= Hf) continue;
                const float w = 1;
                const f4 col = sample_linear_rgba8(in, W, H, current.x + 0.5f, current.y + 0.5f);
                accum = accum + col * w;
                fcount += w;
            }
            if (fcount != 0) accum = accum / fcount;
            store_rgba8(c->rgba8, px, accum);
        }
    }
    for (size_t o = 0; o < c->objs.size(); o++) {
        if (!seen[o]) continue;
        rr_obj_desc& G = c->objs[o];
        if ((frame_id & 1) == 0) { memcpy(G.old_world_pos_2, G.world_pos, 12); memcpy(G.old_world_rot_quat_2, G.world_rot_quat, 16); }
        else { memcpy(G.old_world_pos_1, G.world_pos, 12); memcpy(G.old_world_rot_quat_1, G.world_rot_quat, 16); }
    }
}

// ---- screenspace_godrays, cl2.cl:1792-1917 -------------------------------------------------------------------------------
// The reference runs it in place (screen_in == screen_out == gl_screen[0], engine.cpp:1471-1476); canonical here as for
// do_pseudo_aa: every read sees the frame as it was before the pass.
void godrays(orc_ctx* c) {
    const int W = c->W, H = c->H;
    const float Wf = (float)W, Hf = (float)H, fov = c->fov;
    const uint32_t* depth_buffer = c->depth[c->cur].data();
    const std::vector<uint8_t> in = c->rgba8;
    const rotsc crot = make_rotsc(c->cam_rot[0], c->cam_rot[1], c->cam_rot[2]);
    const f3 cpos = v3(c->cam_pos);
#pragma omp parallel for schedule(dynamic, 4) num_threads(c->threads)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const size_t px = (size_t)y * W + x;
            const float samples = 80.f;
            const uint32_t my_depth = depth_buffer[px];
            f4 my_col = sample_linear_rgba8(in, W, H, (float)x - 0.25f, (float)y - 0.25f);
            const float decay_factor = 0.97f, weight = 0.01f, max_length = 400.f;
            for (size_t i = 0; i < c->lights.size(); i++) {
                const rr_light& l = c->lights[i];
                const float ray_intensity = l.godray_intensity;
                if (ray_intensity <= 0) continue;
                float idecay = 1.f;
                f3 iter_col = {0, 0, 0};
                f3 slpos = rot(v3(l.pos), cpos, crot);
                slpos = depth_project_singular(slpos, Wf, Hf, fov);
                f3 current_pos = {(float)x, (float)y, ((float)my_depth / U32MAXF) * DEPTH_FAR};
                f3 destination_pos = slpos;
                const f3 original = current_pos;
                if (slpos.z < 0) destination_pos = (current_pos - destination_pos) + current_pos;
                const float vx = fabsf(current_pos.x - destination_pos.x), vy = fabsf(current_pos.y - destination_pos.y);
                const float mnum = vx > vy ? vx : vy;
                f3 dir = (destination_pos - current_pos) / mnum;
                dir = dir * (max_length / samples);
                const f3 col = v3(l.col);
                for (int j = 0; (float)j < mnum && (float)j < samples; j++) {
                    if (current_pos.x < 0 || current_pos.y < 0 || current_pos.x >= Wf - 1 || current_pos.y >= Hf - 1) continue;
                    const uint32_t cdepth = depth_buffer[(size_t)((int)current_pos.y) * W + (int)current_pos.x];
                    const float fdepth = ((float)cdepth / U32MAXF) * DEPTH_FAR;
                    f3 val = {0, 0, 0};
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
            const float exposure = 0.99f;
            store_rgba8(c->rgba8, px, my_col * exposure);
        }
    }
}

}  // namespace

// =====================================================================================================================
// C ABI — same shape as include/rr.h with an `orc_` prefix so the parity tests can drive both sides identically.
// =====================================================================================================================
extern "C" {

const char* orc_last_error(void) { return g_err; }

float orc_fov_const_from_hfov(float hfov_deg, float screenwidth) {   // engine.cpp:119-133, 474-477
    float fov_radians = (hfov_deg / 360.f) * 2 * (float)M_PI;     // float * double M_PI in the reference -> see below
    // engine.cpp:121 is `(horizontal_fov_degrees / 360.f) * 2 * M_PI` assigned to a float: the product is done in double.
    double fr = ((double)(hfov_deg / 360.f) * 2) * M_PI;
    fov_radians = (float)fr;
    float triangle_angle = fov_radians / 2;
    float fov_constant = (float)((double)(screenwidth / 2) / tan((double)triangle_angle));
    char buf[64];
    snprintf(buf, sizeof buf, "%f", fov_constant);                  // std::to_string(float) == "%f"
    return strtof(buf, nullptr);
}

orc_ctx* orc_create(const rr_config* cfg, int threads) {
    if (!cfg || cfg->width <= 0 || cfg->height <= 0) { snprintf(g_err, sizeof g_err, "bad config"); return nullptr; }
    orc_ctx* c = new orc_ctx();
    c->cfg = *cfg;
    c->W = cfg->width; c->H = cfg->height; c->L = cfg->light_dim;
    c->fov = cfg->fov_const > 0 ? cfg->fov_const : orc_fov_const_from_hfov(cfg->hfov_deg, (float)cfg->width);
#ifdef _OPENMP
    c->threads = threads > 0 ? threads : omp_get_max_threads();
#else
    c->threads = 1;
#endif
    size_t P = (size_t)c->W * c->H;
    c->depth[0].assign(P, 0xFFFFFFFFu);
    c->depth[1].assign(P, 0xFFFFFFFFu);
    c->ids.assign(P, 0u);
    c->rgba8.assign(P * 4, 0);
    c->colour.assign(P * 4, 0.f);
    c->normals.assign(P * 2, 0);
    memset(&c->tm, 0, sizeof c->tm);
    return c;
}
void orc_destroy(orc_ctx* c) { delete c; }
int orc_threads(orc_ctx* c) { return c->threads; }

int orc_scene_alloc(orc_ctx* c, uint32_t n_tris, uint32_t n_objs) { c->tris.assign(n_tris, rr_triangle{}); c->objs.assign(n_objs, rr_obj_desc{}); return RR_OK; }
// object_context::build(async) + flip_buffers: the checker rebuilds synchronously into a second scene and swaps on commit
int orc_scene_build_begin(orc_ctx* c, uint32_t n_tris, uint32_t n_objs) { c->back_tris.assign(n_tris, rr_triangle{}); c->back_objs.assign(n_objs, rr_obj_desc{}); return RR_OK; }
int orc_scene_build_write_tris(orc_ctx* c, uint32_t first, uint32_t count, const rr_triangle* t) {
    if ((size_t)first + count > c->back_tris.size()) return RR_ERR_INVALID;
    memcpy(&c->back_tris[first], t, (size_t)count * sizeof(rr_triangle)); return RR_OK;
}
int orc_scene_build_write_objs(orc_ctx* c, uint32_t first, uint32_t count, const rr_obj_desc* o) {
    if ((size_t)first + count > c->back_objs.size()) return RR_ERR_INVALID;
    memcpy(&c->back_objs[first], o, (size_t)count * sizeof(rr_obj_desc)); return RR_OK;
}
int orc_scene_build_ready(orc_ctx*) { return 1; }
int orc_scene_build_commit(orc_ctx* c) { c->tris.swap(c->back_tris); c->objs.swap(c->back_objs); return RR_OK; }
int orc_scene_write_tris(orc_ctx* c, uint32_t first, uint32_t count, const rr_triangle* t) {
    if ((size_t)first + count > c->tris.size()) return RR_ERR_INVALID;
    memcpy(&c->tris[first], t, (size_t)count * sizeof(rr_triangle)); return RR_OK;
}
int orc_scene_write_objs(orc_ctx* c, uint32_t first, uint32_t count, const rr_obj_desc* o) {
    if ((size_t)first + count > c->objs.size()) return RR_ERR_INVALID;
    memcpy(&c->objs[first], o, (size_t)count * sizeof(rr_obj_desc)); return RR_OK;
}
int orc_scene_patch_obj(orc_ctx* c, uint32_t obj_id, uint32_t off, uint32_t n, const void* src) {
    if (obj_id >= c->objs.size() || off + n > sizeof(rr_obj_desc)) return RR_ERR_INVALID;
    memcpy((char*)&c->objs[obj_id] + off, src, n); return RR_OK;
}

int orc_atlas_alloc(orc_ctx* c, uint32_t n_slices, const uint32_t* nums, uint32_t n_nums, const uint32_t* sizes, uint32_t n_sizes, uint32_t mipmap_start) {
    uint32_t s = std::max(n_slices, 2u);      // clamped_array_len, texture_context.cpp:441
    c->atlas.assign((size_t)s * ATLAS_DIM * ATLAS_DIM * 4, 0);
    c->nums.assign(nums, nums + n_nums);
    c->sizes.assign(sizes, sizes + n_sizes);
    c->mipmap_start = mipmap_start;
    return RR_OK;
}
int orc_atlas_upload(orc_ctx* c, uint32_t gpu_id, const uint8_t* rgba, uint32_t w, uint32_t h, int flip) {
    if (gpu_id >= c->nums.size()) return RR_ERR_INVALID;
    update_gpu_tex(c, gpu_id, rgba, (int)w, (int)h, flip);
    // texture::update_gpu_mipmaps, texture.cpp:465-493
    mip_pass(c, gpu_id, gpu_id * MIP_LEVELS + c->mipmap_start + 0, (int)w, (int)h);
    for (uint32_t i = 0; i < MIP_LEVELS - 1; i++) {
        uint32_t proper = gpu_id * MIP_LEVELS + c->mipmap_start + i;
        mip_pass(c, proper, proper + 1, (int)w, (int)h);
    }
    return RR_OK;
}
// the same uploads for a set of textures (texture_context.cpp:478-517 is a serial loop over them)
int orc_atlas_upload_batch(orc_ctx* c, uint32_t n, const uint32_t* gpu_ids, const uint8_t* const* rgba, const uint32_t* w, const uint32_t* h, int flip) {
    for (uint32_t i = 0; i < n; i++) { int r = orc_atlas_upload(c, gpu_ids[i], rgba[i], w[i], h[i], flip); if (r) return r; }
    return RR_OK;
}
// update_gpu_tex_colour, cl2.cl:955-984 (texture::update_gpu_texture_col, texture.cpp:445-463); global size = the texture's image size
int orc_atlas_fill_colour(orc_ctx* c, uint32_t tex_id, const float col[4], uint32_t gw, uint32_t gh) {
    if (tex_id >= c->nums.size()) return RR_ERR_INVALID;
    const int slice = (int)(c->nums[tex_id] >> 16);
    const float width = (float)c->sizes[slice];
    const uint32_t ucol[4] = {sat_u32(col[0]), sat_u32(col[1]), sat_u32(col[2]), sat_u32(col[3])};      // convert_uint4, pinned saturating
    for (int y = 0; y < (int)gh; y++)
        for (int x = 0; x < (int)gw; x++) {
            if ((float)x >= width || (float)y >= width) continue;
            write_tex_array(ucol, {(float)x, (float)y}, tex_id, c);
            for (int i = 0; i < MIP_LEVELS; i++) {
                const uint32_t mtexid = tex_id * MIP_LEVELS + c->mipmap_start + (uint32_t)i;
                const float nwidth = (float)c->sizes[c->nums[mtexid] >> 16];
                write_tex_array(ucol, {((float)x / width) * nwidth, ((float)y / width) * nwidth}, mtexid, c);
            }
        }
    return RR_OK;
}
// generate_from_raw, cl2.cl:1006-1031 (texture::update_gpu_texture_mono, texture.cpp:554-584)
int orc_atlas_upload_mono(orc_ctx* c, uint32_t tex_id, const uint8_t* raw, uint32_t len, uint32_t w, uint32_t h, int /*flip*/) {
    if (tex_id >= c->nums.size() || h == 0) return RR_ERR_INVALID;
    const int stride = (int)(len / h);
    const int width = (int)c->sizes[c->nums[tex_id] >> 16];
    for (int y = 0; y < (int)h; y++)
        for (int x = 0; x < (int)w; x++) {
            if (x >= width || y >= width) continue;
            const uint32_t v = raw[(size_t)y * stride + x];
            const uint32_t val[4] = {v, v, v, v};
            write_tex_array(val, {(float)x, (float)y}, tex_id, c);
        }
    return RR_OK;
}
int orc_atlas_write_raw(orc_ctx* c, const uint8_t* a, size_t n) { if (n > c->atlas.size()) return RR_ERR_INVALID; memcpy(c->atlas.data(), a, n); return RR_OK; }
int orc_atlas_read_raw(orc_ctx* c, uint8_t* d, size_t n) { if (n > c->atlas.size()) return RR_ERR_INVALID; memcpy(d, c->atlas.data(), n); return RR_OK; }

int orc_lights_write(orc_ctx* c, const rr_light* l, uint32_t n) {
    c->lights.assign(l, l + n);
    c->n_shadow = 0; c->n_static = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (l[i].shadow == 1) c->n_shadow++;
        if (l[i].shadow && l[i].is_static) c->n_static++;
    }
    size_t slab = (size_t)6 * c->L * c->L;
    c->shadow_dyn.assign(std::max<size_t>(slab * c->n_shadow, 1), 0xFFFFFFFFu);
    c->shadow_static.assign(std::max<size_t>(slab * c->n_static, 1), 0xFFFFFFFFu);
    return RR_OK;
}

// engine::generate_realtime_shadowing, engine.cpp:1601-1790
int orc_frame_shadows(orc_ctx* c, int static_dirty) {
    double t0 = now_ms();
    size_t slab = (size_t)6 * c->L * c->L;
    c->shadow_samples = 0; c->tm.n_shadow_fragments = 0;
    if (!c->lights.empty()) {
        std::fill(c->shadow_dyn.begin(), c->shadow_dyn.end(), 0xFFFFFFFFu);                 // engine.cpp:1615
        if (static_dirty) std::fill(c->shadow_static.begin(), c->shadow_static.end(), 0xFFFFFFFFu);   // 1620-1625
    }
    uint32_t nn = 0, kk = 0;
    for (size_t i = 0; i < c->lights.size(); i++) {
        const rr_light& l = c->lights[i];
        if (l.shadow == 1) {
            shadow_pass(c, l.pos, 0, &c->shadow_dyn[slab * nn], nn * 6);
            nn++;
        }
        if (l.shadow && l.is_static && static_dirty) {
            shadow_pass(c, l.pos, 1, &c->shadow_static[slab * kk], kk * 6);
            kk++;
        }
    }
    c->tm.shadow_depth_ms = (float)(now_ms() - t0);
    return RR_OK;
}

// engine::draw_bulk_objs_n -> render_tris, engine.cpp:1794-2025 (converged steady state: all fragments processed, q6)
int orc_frame_draw(orc_ctx* c, const float c_pos[4], const float c_rot[4], const float clear_rgba[4]) {
    if (c->tris.empty()) return RR_OK;                                  // engine.cpp:1806
    rotsc crot = make_rotsc(c_rot[0], c_rot[1], c_rot[2]);
    double t0 = now_ms();
    std::fill(c->ids.begin(), c->ids.end(), 0u);
    prearrange(c, c_pos, crot);
    double t1 = now_ms();
    kernel1(c);
    double t2 = now_ms();
    kernel2(c);
    double t3 = now_ms();
    kernel3(c, c_pos, crot, clear_rgba);
    double t4 = now_ms();
    c->tm.setup_ms = (float)(t1 - t0); c->tm.depth_ms = (float)(t2 - t1); c->tm.id_ms = (float)(t3 - t2);
    c->tm.shade_ms = (float)(t4 - t3); c->tm.frame_ms = (float)(t4 - t0);
    c->tm.n_cutdown = c->n_cut; c->tm.n_fragments = c->n_frags;
    memcpy(c->cam_pos, c_pos, 16); memcpy(c->cam_rot, c_rot, 16);
    c->frame_id++;                                                      // engine.cpp:2024
    return RR_OK;
}

// stage-by-stage entry points so tests can compare intermediates and bench can time stages
int orc_stage_setup(orc_ctx* c, const float c_pos[4], const float c_rot[4]) {
    rotsc crot = make_rotsc(c_rot[0], c_rot[1], c_rot[2]);
    std::fill(c->ids.begin(), c->ids.end(), 0u);
    prearrange(c, c_pos, crot); return RR_OK;
}
int orc_stage_depth(orc_ctx* c) { kernel1(c); return RR_OK; }
int orc_stage_ids(orc_ctx* c) { kernel2(c); return RR_OK; }

// engine::do_pseudo_aa, engine.cpp:1513-1516: after orc_frame_draw, before orc_swap_buffers
int orc_post_pseudo_aa(orc_ctx* c) { pseudo_aa(c); return RR_OK; }

// engine::do_motion_blur, engine.cpp:1518-1538 / engine::draw_godrays, engine.cpp:1463-1482: after orc_frame_draw, before orc_swap_buffers
int orc_post_motion_blur(orc_ctx* c, float strength, float camera_contribution) { motion_blur(c, strength, camera_contribution); return RR_OK; }
int orc_post_godrays(orc_ctx* c) { godrays(c); return RR_OK; }
int orc_scene_read_objs(orc_ctx* c, uint32_t first, uint32_t count, rr_obj_desc* dst) {
    if ((uint64_t)first + count > c->objs.size()) return RR_ERR_INVALID;
    if (count) memcpy(dst, &c->objs[first], (size_t)count * sizeof(rr_obj_desc));
    return RR_OK;
}

int orc_swap_buffers(orc_ctx* c) {                                      // object_context_data::swap_buffers, object_context.cpp:17-25
    c->cur ^= 1;                                                         // depth_buffer.flip()
    memcpy(c->cam_pos_old, c->cam_pos, 16); memcpy(c->cam_rot_old, c->cam_rot, 16);
    return RR_OK;
}
int orc_sync(orc_ctx*) { return RR_OK; }

// After orc_frame_draw and BEFORE orc_swap_buffers these return the frame just drawn.
int orc_read_depth(orc_ctx* c, uint32_t* d) { memcpy(d, c->depth[c->cur].data(), c->depth[0].size() * 4); return RR_OK; }
int orc_read_ids(orc_ctx* c, uint32_t* d) {
    for (size_t i = 0; i < c->ids.size(); i++) d[i] = c->ids[i] ? c->ids[i] - 1u : 0u;
    return RR_OK;
}
int orc_read_rgba8(orc_ctx* c, uint8_t* d) { memcpy(d, c->rgba8.data(), c->rgba8.size()); return RR_OK; }
int orc_read_colour_f32(orc_ctx* c, float* d) { memcpy(d, c->colour.data(), c->colour.size() * 4); return RR_OK; }
int orc_read_normals(orc_ctx* c, uint16_t* d) { memcpy(d, c->normals.data(), c->normals.size() * 2); return RR_OK; }
int orc_read_shadow(orc_ctx* c, int is_static, uint32_t slab, uint32_t* d) {
    size_t n = (size_t)6 * c->L * c->L;
    const std::vector<uint32_t>& b = is_static ? c->shadow_static : c->shadow_dyn;
    if ((slab + 1) * n > b.size()) return RR_ERR_INVALID;
    memcpy(d, &b[slab * n], n * 4); return RR_OK;
}
int orc_write_shadow(orc_ctx* c, int is_static, uint32_t slab, const uint32_t* s) {
    size_t n = (size_t)6 * c->L * c->L;
    std::vector<uint32_t>& b = is_static ? c->shadow_static : c->shadow_dyn;
    if ((slab + 1) * n > b.size()) return RR_ERR_INVALID;
    memcpy(&b[slab * n], s, n * 4); return RR_OK;
}
int orc_read_fragments(orc_ctx* c, uint32_t* d, uint32_t max_records, uint32_t* n) {
    uint32_t k = std::min(max_records, c->n_frags);
    if (d) memcpy(d, c->frags.data(), (size_t)k * FRAG_MUL * 4);
    if (n) *n = c->n_frags;
    return RR_OK;
}
int orc_read_cutdown(orc_ctx* c, float* d, uint32_t max_tris, uint32_t* n) {
    uint32_t k = std::min(max_tris, c->n_cut);
    if (d) memcpy(d, c->cutdown.data(), (size_t)k * 48);
    if (n) *n = c->n_cut;
    return RR_OK;
}
int orc_get_timings(orc_ctx* c, rr_timings* t) { *t = c->tm; return RR_OK; }
uint64_t orc_depth_samples(orc_ctx* c) { return c->depth_samples; }
uint64_t orc_shadow_samples(orc_ctx* c) { return c->shadow_samples; }
uint64_t orc_saturation_events(orc_ctx* c) { return c->sat_events; }

// ---- small pieces exported for unit tests against the reference's host mirrors -----------------------
void orc_unit_rot(const float p[3], const float cpos[3], const float crot[3], float out[3]) {
    f3 r = rot(v3(p), v3(cpos), make_rotsc(crot[0], crot[1], crot[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_unit_back_rot(const float p[3], const float cpos[3], const float crot[3], float out[3]) {
    f3 r = back_rot(v3(p), v3(cpos), make_rotsc(crot[0], crot[1], crot[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_unit_rot_quat(const float p[3], const float q[4], float out[3]) {
    f3 r = rot_quat(v3(p), f4{q[0], q[1], q[2], q[3]}); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_unit_back_rot_quat(const float p[3], const float q[4], float out[3]) {
    f3 r = back_rot_quat(v3(p), f4{q[0], q[1], q[2], q[3]}); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int orc_unit_point_in_tri(float px, float py, const float t[6]) { return point_in_tri({px, py}, {t[0], t[1]}, {t[2], t[3]}, {t[4], t[5]}); }
int orc_unit_cubeface(const float p[3], const float l[3]) { return ret_cubeface(v3(p), v3(l)); }
uint32_t orc_unit_wang_hash(uint32_t s) { return wang_hash(s); }
uint32_t orc_unit_xorshift(uint32_t s) { return rand_xorshift(s); }
float orc_unit_log2_approx(float v) { return log2_approx(v); }
void orc_unit_texture_mod(const float in[2], float out[2]) { f2 r = texture_mod({in[0], in[1]}); out[0] = r.x; out[1] = r.y; }
float orc_unit_rational_acos(float x) { return rational_acos(x); }
void orc_unit_encode_normal(const float n[3], uint16_t out[2]) { encode_normal(v3(n), out); }
// literal pixel walk: returns the visited (x,y) list of one fragment chunk
int orc_unit_scan(const float mm[4], int op_size, uint32_t distance, int32_t* xy_out, int max_out) {
    int n = 0;
    scan_fragment(mm, op_size, distance, [&](float x, float y) { if (n < max_out) { xy_out[2 * n] = (int)x; xy_out[2 * n + 1] = (int)y; } n++; });
    return n;
}
// clip + project one triangle (generate_new_triangles + depth_project): out = up to 2x3x3 floats, returns num
int orc_unit_clip_project(const float pr[9], int icut, float w, float h, float fov, float out[18]) {
    f3 p[3] = {v3(pr), v3(pr + 3), v3(pr + 6)};
    f3 t[2][3]; int num = 0;
    generate_new_triangles(p, icut, &num, t);
    for (int i = 0; i < num; i++) {
        f3 o[3]; depth_project(t[i], w, h, fov, o);
        for (int k = 0; k < 3; k++) { out[i * 9 + k * 3] = o[k].x; out[i * 9 + k * 3 + 1] = o[k].y; out[i * 9 + k * 3 + 2] = o[k].z; }
    }
    return num;
}

}  // extern "C"

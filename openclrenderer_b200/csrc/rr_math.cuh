// rr_math.cuh — device-side arithmetic of the raster path, written for sm_100a.
//
// The translation units that include this header are compiled with -fmad=false: a product followed by a sum is two
// roundings, and every place where the reference writes mad() is an explicit fmaf() (one rounding). Division and
// sqrt are IEEE (no -use_fast_math). This is the arithmetic pinned in DESIGN.md §3; the reference lines each function
// answers to are cited as cl2.cl:a-b.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rr {

#define RR_DEPTH_FAR 350000.0f       /* cl2.cl:17 */
#define RR_U32MAXF 4294967296.0f     /* (float)UINT_MAX, cl2.cl:19 */
#define RR_INV_U32MAXF 2.3283064365386963e-10f   /* 2^-32: x / 2^32 == x * 2^-32 bit for bit */
#define RR_PI_F 3.1415927f           /* cl2.cl:14 */
#define RR_OP_SIZE 500               /* cl2.cl:4247 */
#define RR_OP_SIZE_LIGHT 300         /* cl2.cl:4249 */
#define RR_FRAG_WORDS 5              /* FRAGMENT_ID_MUL cl2.cl:4252 */
#define RR_SFRAG_WORDS 4             /* FIDM1 cl2.cl:4411 */
#define RR_BUF_ERROR 20u             /* cl2.cl:5387 */
#define RR_MIP_LEVELS 4
#define RR_ATLAS_DIM 2048

struct RotSC { float sx, sy, sz, cx, cy, cz; };   // native_sin / native_cos of an euler triple, computed on the host

__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 operator/(float4 a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// a / b for b > 0 where a is often exactly zero: +-0 / b == +-0, and skipping the division avoids the IEEE slow path that
// a zero numerator triggers (bit-identical result)
__device__ __forceinline__ float div_pos(float a, float b) { return a == 0.f ? a : a / b; }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 normalize3(float3 a) { return a / sqrtf(dot3(a, a)); }      // fast_normalize, pinned
__device__ __forceinline__ float4 normalize4(float4 a) { return a / sqrtf(dot4(a, a)); }
__device__ __forceinline__ float3 mad3(float3 a, float b, float3 c) { return make_float3(fmaf(a.x, b, c.x), fmaf(a.y, b, c.y), fmaf(a.z, b, c.z)); }
__device__ __forceinline__ float2 mad2(float2 a, float b, float2 c) { return make_float2(fmaf(a.x, b, c.x), fmaf(a.y, b, c.y)); }
__device__ __forceinline__ float4 mad4(float4 a, float b, float4 c) {
    return make_float4(fmaf(a.x, b, c.x), fmaf(a.y, b, c.y), fmaf(a.z, b, c.z), fmaf(a.w, b, c.w));
}
__device__ __forceinline__ float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }

// atomic_min / atomic_max whose result nobody reads, as a reduction: red.global.{min,max}.u32 is fire-and-forget at the L2 (no
// return path). Written out because ptxas kept ATOMG with a discarded destination in the setup kernel (ncu: lts op_atom).
__device__ __forceinline__ void red_min_u32(uint32_t* p, uint32_t v) { asm volatile("red.global.min.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_max_u32(uint32_t* p, uint32_t v) { asm volatile("red.global.max.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// saturating float -> uint32 (NaN -> 0, negative -> 0, >= 2^32 -> UINT_MAX): cvt.rzi.u32.f32 does exactly this
__device__ __forceinline__ uint32_t sat_u32(float f) { return __float2uint_rz(f); }

// a / b for a divisor that is nearly always a power of two (tile size / 2048): multiplying by the exact reciprocal 2^-e gives the
// correctly rounded quotient, bit for bit, without the IEEE division sequence (whose slow path a zero numerator would take);
// any other divisor goes through the real division, kept out of line so that it is not evaluated speculatively
__device__ __noinline__ float div_general(float a, float b) { return a / b; }
__device__ __forceinline__ float div_pow2(float a, float b) {
    const uint32_t bb = __float_as_uint(b);
    // normal power of two whose reciprocal is normal as well
    if ((bb & 0x807FFFFFu) == 0u && bb >= 0x01000000u && bb <= 0x7E000000u) return a * __uint_as_float(0x7F000000u - bb);
    return div_general(a, b);
}

// cl2.cl:220-271
__device__ __forceinline__ float3 rot(float3 point, float3 c_pos, const RotSC& r) {
    float3 rel = point - c_pos;
    float t = fmaf(r.sz, rel.y, r.cz * rel.x);
    float u = fmaf(r.cy, rel.z, r.sy * t);
    float v = fmaf(r.cz, rel.y, -(r.sz * rel.x));
    return make_float3(fmaf(r.cy, t, -(r.sy * rel.z)), fmaf(r.sx, u, r.cx * v), fmaf(r.cx, u, -(r.sx * v)));
}

// cl2.cl:275-348 (factorisation at 325-345)
__device__ __forceinline__ float3 back_rot(float3 point, float3 c_pos, const RotSC& r) {
    float3 rel = point - c_pos;
    float3 ret;
    ret.x = r.cz * (fmaf(r.cy, rel.x, fmaf(r.sx, r.sy * rel.y, r.cx * r.sy * rel.z))) + r.sz * (r.sx * rel.z - r.cx * rel.y);
    ret.y = fmaf(r.sz, r.cy * rel.x, fmaf(fmaf(r.cx, r.cz, r.sx * r.sy * r.sz), rel.y, (fmaf(-r.sx, r.cz, r.cx * r.sy * r.sz) * rel.z)));
    ret.z = fmaf(-r.sy, rel.x, r.cy * fmaf(r.sx, rel.y, r.cx * rel.z));
    return ret;
}

// cl2.cl:350-357
__device__ __forceinline__ float3 rot_quat(float3 point, float4 quat) {
    quat = normalize4(quat);
    float3 q = xyz(quat);
    float3 t = 2.f * cross3(q, point);
    return point + quat.w * t + cross3(q, t);
}
// rot_quat for a quaternion that is already normalised by the caller (same arithmetic, hoisted)
__device__ __forceinline__ float3 rot_quat_n(float3 point, float4 nquat) {
    float3 q = xyz(nquat);
    float3 t = 2.f * cross3(q, point);
    return point + nquat.w * t + cross3(q, t);
}

// cl2.cl:359-370
__device__ __forceinline__ float4 back_quat(float4 quat) {          // the normalised conjugate rot_quat() will use
    float4 conj = make_float4(-quat.x, -quat.y, -quat.z, quat.w);
    float len_sq = dot4(conj, conj);
    return normalize4(conj / len_sq);
}

// cl2.cl:408-411
// 1.0f / d for a d that is often exactly +-0 (rounded vertices of a sub-pixel triangle are collinear): 1 / +-0 == +-inf, and
// skipping the division avoids the IEEE slow path a zero divisor takes (bit-identical result)
__device__ __forceinline__ float recip_or_inf(float d) { return d == 0.f ? __int_as_float(0x7f800000u | (__float_as_uint(d) & 0x80000000u)) : 1.0f / d; }

__device__ __forceinline__ float calc_rconstant_v(float3 x, float3 y) {
    return recip_or_inf(x.y * y.z + x.x * (y.y - y.z) - x.z * y.y + (x.z - x.y) * y.x);
}

// cl2.cl:413-418
__device__ __forceinline__ void interpolate_get_const(float3 f, float3 x, float3 y, float rconstant, float& A, float& B, float& C) {
    A = ((f.y * y.z + f.x * (y.y - y.z) - f.z * y.y + (f.z - f.y) * y.x) * rconstant);
    B = (-(f.y * x.z + f.x * (x.y - x.z) - f.z * x.y + (f.z - f.y) * x.x) * rconstant);
    C = f.x - A * x.x - B * y.x;
}

// cl2.cl:420-441 / 443-459: bbox [round(min)-1, round(max)] clamped to the viewport. mm = (min_x, max_x, min_y, max_y)
__device__ __forceinline__ float4 calc_min_max(float3 xr, float3 yr, float width, float height) {
    float4 mm;
    mm.x = fminf(fminf(xr.x, xr.y), xr.z) - 1.f;
    mm.y = fmaxf(fmaxf(xr.x, xr.y), xr.z);
    mm.z = fminf(fminf(yr.x, yr.y), yr.z) - 1.f;
    mm.w = fmaxf(fmaxf(yr.x, yr.y), yr.z);
    mm.x = clampf(mm.x, 0.0f, width - 1.f);
    mm.y = clampf(mm.y, 0.0f, width - 1.f);
    mm.z = clampf(mm.z, 0.0f, height - 1.f);
    mm.w = clampf(mm.w, 0.0f, height - 1.f);
    return mm;
}

// cl2.cl:491-494
__device__ __forceinline__ bool front_facing(float3 p0, float3 p1, float3 p2) { return cross3(p1 - p0, p2 - p0).z < 0.f; }

// cl2.cl:535-544
__device__ __forceinline__ float3 project(float3 r, float half_w, float half_h, float fovc) {
    float k = fovc / r.z;
    return make_float3(fmaf(r.x, k, half_w), fmaf(r.y, k, half_h), r.z);
}

// cl2.cl:577-663 — near-plane clip of one camera-space triangle into 0/1/2 triangles (register-only formulation:
// the reference's index arrays become selects so nothing is spilled to local memory)
__device__ __forceinline__ float3 sel3(int k, float3 a, float3 b, float3 c) { return k == 0 ? a : (k == 1 ? b : c); }

__device__ __forceinline__ int clip_near(float3 q0, float3 q1, float3 q2, float icut, float3& a0, float3& a1, float3& a2, float3& b0, float3& b1, float3& b2) {
    const bool h0 = q0.z <= icut || q0.z > RR_DEPTH_FAR, h1 = q1.z <= icut || q1.z > RR_DEPTH_FAR, h2 = q2.z <= icut || q2.z > RR_DEPTH_FAR;
    const int n_behind = (int)h0 + (int)h1 + (int)h2;
    if (n_behind == 0) { a0 = q0; a1 = q1; a2 = q2; return 1; }
    if (n_behind > 2) return 0;
    int g1, g2, g3;
    if (n_behind == 1) {
        const int id = h0 ? 0 : (h1 ? 1 : 2);                     // ids_behind[0]
        g1 = id;
        g2 = (id + 1) >= 3 ? id - 2 : id + 1;
        g3 = (id + 2) >= 3 ? id - 1 : id + 2;
    } else {
        g2 = h0 ? 0 : 1;                                          // ids_behind[0]
        g3 = h2 ? 2 : 1;                                          // ids_behind[1]
        g1 = !h0 ? 0 : (!h1 ? 1 : 2);                             // id_valid
    }
    const float3 P1 = sel3(g1, q0, q1, q2), P2 = sel3(g2, q0, q1, q2), P3 = sel3(g3, q0, q1, q2);
    const float3 p1 = P2 + ((icut - P2.z) * (P1 - P2)) / (P1.z - P2.z);
    const float3 p2 = P3 + ((icut - P3.z) * (P1 - P3)) / (P1.z - P3.z);
    if (n_behind == 1) {
        a0 = p1; a1 = P2; a2 = P3;
        b0 = p1; b1 = P3; b2 = p2;
        return 2;
    }
    // two behind: slot ids_behind[0] <- p1, slot ids_behind[1] <- p2, slot id_valid <- the valid vertex
    a0 = (0 == g2) ? p1 : ((0 == g3) ? p2 : P1);
    a1 = (1 == g2) ? p1 : ((1 == g3) ? p2 : P1);
    a2 = (2 == g2) ? p1 : ((2 == g3) ? p2 : P1);
    return 1;
}

// clip_near for callers that take the (at most two) output triangles one at a time: returns how many triangles the clip
// produces and writes triangle number `which` (0 or 1) to a0..a2 when it exists. Same arithmetic as clip_near.
__device__ __forceinline__ int clip_near_one(float3 q0, float3 q1, float3 q2, float icut, bool which, float3& a0, float3& a1, float3& a2) {
    const bool h0 = q0.z <= icut || q0.z > RR_DEPTH_FAR, h1 = q1.z <= icut || q1.z > RR_DEPTH_FAR, h2 = q2.z <= icut || q2.z > RR_DEPTH_FAR;
    if (!(h0 || h1 || h2)) { a0 = q0; a1 = q1; a2 = q2; return 1; }
    const int n_behind = (int)h0 + (int)h1 + (int)h2;
    if (n_behind > 2) return 0;
    int g1, g2, g3;
    if (n_behind == 1) {
        const int id = h0 ? 0 : (h1 ? 1 : 2);
        g1 = id;
        g2 = (id + 1) >= 3 ? id - 2 : id + 1;
        g3 = (id + 2) >= 3 ? id - 1 : id + 2;
    } else {
        g2 = h0 ? 0 : 1;
        g3 = h2 ? 2 : 1;
        g1 = !h0 ? 0 : (!h1 ? 1 : 2);
    }
    const float3 P1 = sel3(g1, q0, q1, q2), P2 = sel3(g2, q0, q1, q2), P3 = sel3(g3, q0, q1, q2);
    const float3 p1 = P2 + ((icut - P2.z) * (P1 - P2)) / (P1.z - P2.z);
    const float3 p2 = P3 + ((icut - P3.z) * (P1 - P3)) / (P1.z - P3.z);
    if (n_behind == 1) {
        if (!which) { a0 = p1; a1 = P2; a2 = P3; }
        else { a0 = p1; a1 = P3; a2 = p2; }
        return 2;
    }
    a0 = (0 == g2) ? p1 : ((0 == g3) ? p2 : P1);
    a1 = (1 == g2) ? p1 : ((1 == g3) ? p2 : P1);
    a2 = (2 == g2) ? p1 : ((2 == g3) ? p2 : P1);
    return 1;
}

// cl2.cl:4798-4807
__device__ __forceinline__ bool point_in_tri(float px, float py, float p0x, float p0y, float p1x, float p1y, float p2x, float p2y) {
    float A = 0.5f * (-p1y * p2x + p0y * (-p1x + p2x) + p0x * (p1y - p2y) + p1x * p2y);
    float sign = A < 0 ? -1.f : 1.f;
    float s = (p0y * p2x - p0x * p2y + (p2y - p0y) * px + (p0x - p2x) * py) * sign;
    float t = (p0x * p1y - p0y * p1x + (p0y - p1y) * px + (p1x - p0x) * py) * sign;
    return s > -0.0001f && t > -0.0001f && (s + t) < 2.0001f * A * sign;
}

// cl2.cl:1745-1790
__device__ __forceinline__ int ret_cubeface(float3 point, float3 light) {
    float3 rel = point - light;
    float ax = fabsf(rel.x), ay = fabsf(rel.y), az = fabsf(rel.z);
    if (ax >= ay && ax >= az) return rel.x < 0 ? 4 : 5;
    if (ay > ax && ay >= az) return rel.y < 0 ? 1 : 3;
    if (az > ax && az > ay) { if (rel.z < 0) return 2; }
    return 0;
}

// cl2.cl:1919-1936
__device__ __forceinline__ uint32_t wang_hash(uint32_t seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
__device__ __forceinline__ uint32_t rand_xorshift(uint32_t s) { s ^= (s << 13); s ^= (s >> 17); s ^= (s << 5); return s; }

// cl2.cl:2469-2477
__device__ __forceinline__ float rational_acos(float x) {
    const float a = -0.939115566365855f, b = 0.9217841528914573f, c = -1.2845906244690837f, d = 0.295624144969963174f;
    const float x2 = x * x;
    const float num = a * x + b * x * x * x, den = 1.f + c * x * x + d * (x2 * x2);      // pow(x, 4)
    return RR_PI_F / 2.f + ((num == 0.f && den > 0.f) ? num : num / den);                // +-0 / positive == +-0 without the slow path
}

// ---- colour-only arithmetic ------------------------------------------------------------------------------------------
// kernel3's lighting sums feed nothing but the RGBA8 colour, which is compared at +-1 LSB (the shipped reference itself builds
// them with -cl-fast-relaxed-math, FP_CONTRACT ON and native_* / fast_* calls). Everything that feeds a DISCRETE decision —
// texel and shadow-texel addresses, mip level, depth compares, the shadow bias, the illumination cut-off, the stored normal —
// keeps the pinned IEEE arithmetic; the smooth terms below use the SFU approximations (<= 2 ulp) and fused multiply-adds.
// Define RR_SHADE_EXACT to build the fully pinned variant for A/B runs.
#ifdef RR_SHADE_EXACT
__device__ __forceinline__ float3 normalize3_c(float3 a) { return normalize3(a); }
__device__ __forceinline__ float div_c(float a, float b) { return a / b; }
__device__ __forceinline__ float exp_c(float x) { return expf(x); }
__device__ __forceinline__ float sqrt_c(float x) { return sqrtf(x); }
#else
__device__ __forceinline__ float3 normalize3_c(float3 a) { const float r = rsqrtf(fmaf(a.x, a.x, fmaf(a.y, a.y, a.z * a.z))); return make_float3(a.x * r, a.y * r, a.z * r); }   // 0 -> NaN like 0 / 0
__device__ __forceinline__ float div_c(float a, float b) { return __fdividef(a, b); }       // 0 / 0 -> NaN, x / 0 -> inf (q17 semantics kept)
__device__ __forceinline__ float exp_c(float x) { return __expf(x); }
__device__ __forceinline__ float sqrt_c(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
__device__ __forceinline__ float rational_acos_c(float x) {           // rational_acos for the specular lobe
    const float a = -0.939115566365855f, b = 0.9217841528914573f, c = -1.2845906244690837f, d = 0.295624144969963174f;
    const float x2 = x * x;
    return RR_PI_F / 2.f + div_c(a * x + b * x * x * x, 1.f + c * x * x + d * (x2 * x2));
}
__device__ __forceinline__ float gamma_inv_c(float c) {
    const float S1 = sqrt_c(c), S2 = sqrt_c(S1), S3 = sqrt_c(S2);
    return 0.585122381f * S1 + 0.783140355f * S2 - 0.368262736f * S3;
}

// cl2.cl:5372-5384
__device__ __forceinline__ void get_barycentric(float3 p, float3 a, float3 b, float3 c, float& u, float& v, float& w) {
    float3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    v = (d11 * d20 - d01 * d21) / denom;
    w = (d00 * d21 - d01 * d20) / denom;
    u = 1.0f - v - w;
}

// cl2.cl:5763-5782, per channel
__device__ __forceinline__ float gamma_fwd(float s) { return 0.012522878f * s + 0.682171111f * s * s + 0.305306011f * s * s * s; }
__device__ __forceinline__ float gamma_inv(float c) {
    float S1 = sqrtf(c), S2 = sqrtf(S1), S3 = sqrtf(S2);
    return 0.585122381f * S1 + 0.783140355f * S2 - 0.368262736f * S3;
}

// cl2.cl:1457-1468
__device__ __forceinline__ float texture_mod1(float v) {
    v = v >= 1 ? 1.0f - (v - floorf(v)) : v;
    v = v < 0 ? 1.0f + fabsf(v) - fabsf(floorf(v)) : v;
    return v;
}

// cl2.cl:1498-1505
__device__ __forceinline__ float log2_approx(float val) {
    int x = __float_as_int(val);
    float log_2 = (float)(((x >> 23) & 255) - 128);
    x &= ~(255 << 23);
    x += 127 << 23;
    float v = __int_as_float(x);
    log_2 += ((-0.3358287811f) * v + 2.0f) * v - 0.65871759316667f;
    return log_2;
}

// cl2.cl:2144-2158
__device__ __forceinline__ float bilinear_interpolate(float cx, float cy, float v0, float v1, float v2, float v3) {
    float mx = cx - 0.5f, my = cy - 0.5f;
    float ux = mx - floorf(mx), uy = my - floorf(my);
    float bx = 1.0f - ux, by = 1.0f - uy;
    return (v0 * bx + v1 * ux) * by + (v2 * bx + v3 * ux) * uy;
}

// The geometry a fragment chunk needs: rounded vertices, depth-plane coefficients, bbox. cl2.cl:5011-5057
struct FragGeom {
    float3 xr, yr;       // rounded x / y of the three vertices
    float A, B, C;       // plane of 1/(z/far)
    float4 mm;           // (min_x, max_x, min_y, max_y)
};

__device__ __forceinline__ FragGeom frag_geom(float3 p0, float3 p1, float3 p2, float rconst, float width, float height) {
    FragGeom g;
    g.xr = make_float3(roundf(p0.x), roundf(p1.x), roundf(p2.x));
    g.yr = make_float3(roundf(p0.y), roundf(p1.y), roundf(p2.y));
    g.mm = calc_min_max(g.xr, g.yr, width, height);
    float3 d = make_float3(p0.z / RR_DEPTH_FAR, p1.z / RR_DEPTH_FAR, p2.z / RR_DEPTH_FAR);     // dcalc
    d = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);                                     // native_recip
    interpolate_get_const(d, g.xr, g.yr, rconst, g.A, g.B, g.C);
    return g;
}

// The reference's pixel walk (cl2.cl:5042-5095 == 5184-5227 == 5447-5508), replayed verbatim for ONE chunk.
// f(x, y) is called for every pixel the state machine tests. Used for small chunks (few slots: one thread each);
// large chunks go through the closed form below so their slots can be spread over many threads.
template <class F>
__device__ __forceinline__ void scan_chunk(const float4 mm, int op_size, uint32_t distance, F&& f) {
    int width = (int)(mm.y - mm.x);
    if (width <= 0) return;
    int pixel_along = op_size * (int)distance;
    int pcount = -1;
    float x = (float)(pixel_along % width) + mm.x - 1.f;
    float y = floorf((float)(pixel_along + pcount) / (float)width) + mm.z;
    float iwidth = 1.f / (float)width;
    float rwm = (float)((pixel_along + pcount) % width);
    const float fw = (float)width;
    while (pcount < op_size) {
        pcount++;
        x += 1.f;
        rwm += 1.f;
        if (rwm >= fw) rwm = 0.f;
        float ty = y;
        y = floorf(fmaf((float)(pixel_along + pcount), iwidth, mm.z));
        x = (y != ty) ? rwm + mm.x : x;
        if (y >= mm.w) break;
        if (x >= mm.y) continue;
        f(x, y);
    }
}

// ---- closed form of the walk ------------------------------------------------------------------------------------------
// The walk visits linear indices k = k0 .. k0+op of the row-major box, with a float row counter
//     y_k = floor(fma((float)k, 1.f/width, min_y))                  (cl2.cl:5076)
// that can lag or lead the true row by a step at exact row boundaries, and a column that restarts only when y_k changes.
// Because y_k is monotone in k, the walk has a closed form (property-tested against the literal state machine in
// tests/test_oracle_units.py::test_scan_closed_form_equals_literal_walk):
//     x_k = min_x + (j mod width) + (k - j),   j = max(k0, first index whose row counter equals y_k)
// and the walk of a triangle ends at k_end = first k with y_k >= max_y. This is what lets pixel slots be spread evenly
// over threads instead of being replayed sequentially per fragment.
__device__ __forceinline__ float walk_row(int k, float iw, float min_y) { return floorf(fmaf((float)k, iw, min_y)); }

__device__ __forceinline__ int walk_end(int width, int rows, float iw, float min_y, float max_y) {
    int k = width * rows;
    while (k > 0 && walk_row(k - 1, iw, min_y) >= max_y) k--;
    while (walk_row(k, iw, min_y) < max_y) k++;
    return k;
}

// pixel of slot k in the chunk starting at k0; returns false when the walk skips it (x beyond the box)
__device__ __forceinline__ bool walk_pixel(int k, int k0, int width, float iw, const float4 mm, float& x, float& y) {
    y = walk_row(k, iw, mm.z);
    int c = (int)(y - mm.z) * width;
    while (c > k0 && walk_row(c - 1, iw, mm.z) == y) c--;
    while (walk_row(c, iw, mm.z) < y) c++;
    const int j = max(c, k0);
    x = mm.x + (float)(j % width) + (float)(k - j);
    return x < mm.y;
}

}  // namespace rr

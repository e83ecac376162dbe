// gorder_math.cuh — device arithmetic shared by the kernels.
//
// Every function states the reference arithmetic it reproduces (file:line under the reference
// tree).  Where the reference's f32 expression has a visible rounding signature (the min-image
// fold, the wrap, the dot-product order) the device code performs the SAME f32 operations in the
// SAME order with contraction disabled (__fmul_rn/__fadd_rn), so that bond vectors, midpoints and
// reconstructed hydrogens are bit-identical to the reference and only the last step
// (cos(acos(c)) -> c) differs by <= 2 ulp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gorder {

struct f3 { float x, y, z; };

__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ float comp(const f3 &v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// x % L for |x| < 2L without calling fmodf: the subtraction is exact (Sterbenz), so the result is
// bit-identical to Rust's `%` (fmodf).  Falls back to fmodf for far-away images.
__device__ __forceinline__ float fmod_near(float x, float L) {
    float ax = fabsf(x);
    if (ax < L) return x;
    if (ax < 2.0f * L) return x - copysignf(L, x);
    return fmodf(x, L);
}

// Slow path of the fold (bond longer than half a box, atoms outside the box, degenerate box):
// the literal expression, out of line so that the hot loop stays small in the instruction cache.
__device__ __noinline__ float min_image_slow(float d, float L, float half) {
    if (!(L > 0.0f)) return d;
    float t = __fadd_rn(d, half);
    t = fmod_near(t, L);      // exact for |t| < 2L (one subtraction), fmodf beyond: the same bits either way
    float u = __fadd_rn(t, L);
    u = fmod_near(u, L);
    return __fsub_rn(u, half);
}

// groan_rs Vector3D::vector_to per component (call site src/analysis/pbc.rs:378-385):
//   (((d + L/2) % L) + L) % L - L/2     in f32, every operation rounded.
// The roundings of this expression are part of the reference's results (oracle/gorder_oracle.c
// min_image, pinned by cgorder.rs:188-241 and uaorder.rs:1113-1200), so they are reproduced.
// Fast path: for 0 <= t = fl(d + L/2) and u = fl(t + L) < 2L both `%` are exact
// (t % L = t, u % L = u - L), so the result is fl(fl(u - L) - L/2) with identical bits.
__device__ __forceinline__ float min_image(float d, float L, float half) {
    const float t = __fadd_rn(d, half);
    const float u = __fadd_rn(t, L);
    if (t >= 0.0f && u < __fadd_rn(L, L)) return __fsub_rn(__fsub_rn(u, L), half);
    return min_image_slow(d, L, half);
}

// Same result with ONE compare on the hot path: |d| <= guard (guard = 0.99 L/2, precomputed per
// frame) implies 0 <= t and u < 2L, i.e. the exact fast path; everything else takes the literal path.
__device__ __forceinline__ float min_image_g(float d, float L, float half, float guard) {
    if (fabsf(d) <= guard) return __fsub_rn(__fsub_rn(__fadd_rn(__fadd_rn(d, half), L), L), half);
    return min_image_slow(d, L, half);
}

// Vector3D::wrap per component (call site pbc.rs:388-390): c % L, + L if negative.
__device__ __forceinline__ float wrap1(float c, float L) {
    float w = fmod_near(c, L);
    if (w < 0.0f) w = __fadd_rn(w, L);
    return w;
}

struct Box {
    float L[3];
    float half[3];
};

template <bool PBC>
__device__ __forceinline__ f3 vector_to(const f3 &p1, const f3 &p2, const Box &b) {
    f3 d = mk3(__fsub_rn(p2.x, p1.x), __fsub_rn(p2.y, p1.y), __fsub_rn(p2.z, p1.z));
    if (PBC) {   // a degenerate dimension (L <= 0) falls into the slow path, which returns d
        d.x = min_image(d.x, b.L[0], b.half[0]);
        d.y = min_image(d.y, b.L[1], b.half[1]);
        d.z = min_image(d.z, b.L[2], b.half[2]);
    }
    return d;
}

template <bool PBC>
__device__ __forceinline__ f3 wrap_point(const f3 &p, const Box &b) {
    if (!PBC) return p;  // NoPBC::wrap is a no-op (pbc.rs:185)
    return mk3(b.L[0] > 0.0f ? wrap1(p.x, b.L[0]) : p.x, b.L[1] > 0.0f ? wrap1(p.y, b.L[1]) : p.y,
               b.L[2] > 0.0f ? wrap1(p.z, b.L[2]) : p.z);
}

// ---- order-free group centre (oracle/gorder_oracle.c group_center) ------------------------------
// groan's group_get_center folds its sums sequentially in f32, which no parallel machine reproduces.  Engine and
// oracle share one order-free definition instead: every term (cos, sin, minimum-image displacement, coordinate) is
// computed in f32 and added as the integer  rint(term * 2^24)  -- exact and associative, so any partition of the
// group over threads, CTAs or devices gives the same bits -- and the circular-mean estimate, which only seeds the
// refinement, evaluates its transcendentals with fixed polynomials in IEEE operations + fmaf (no MUFU, no libm).
constexpr float kCenterScale = 16777216.0f;   // 2^24

// sin / cos of 2 pi u (u in turns): Taylor about the nearest quarter turn, |t| <= 1/8, <= 3e-8
__device__ __forceinline__ void sincos_turns(float u, float &sn, float &cs) {
    const float r = __fsub_rn(u, rintf(u));          // exact, [-0.5, 0.5]
    const float j = rintf(__fmul_rn(4.0f, r));       // -2 .. 2
    const float t = fmaf(j, -0.25f, r);              // exact
    const float z = __fmul_rn(t, t);
    float sp = fmaf(z, 0x1.507834p+5f, -0x1.32d2ccp+6f);
    sp = fmaf(z, sp, 0x1.466bc6p+6f); sp = fmaf(z, sp, -0x1.4abbcep+5f); sp = fmaf(z, sp, 0x1.921fb6p+2f);
    sp = __fmul_rn(sp, t);
    float cp = fmaf(z, 0x1.e1f506p+5f, -0x1.55d3c8p+6f);
    cp = fmaf(z, cp, 0x1.03c1f0p+6f); cp = fmaf(z, cp, -0x1.3bd3ccp+4f); cp = fmaf(z, cp, 1.0f);
    const int q = (int)j & 3;
    sn = q == 0 ? sp : (q == 1 ? cp : (q == 2 ? -sp : -cp));
    cs = q == 0 ? cp : (q == 1 ? -sp : (q == 2 ? -cp : sp));
}

// atan2(y, x) / 2 pi in [-0.5, 0.5] (minimax of atan(a) / (2 pi a) in a^2, <= 2e-8 turns)
__device__ __forceinline__ float atan2_turns(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    const float a = mx > 0.0f ? __fdiv_rn(mn, mx) : 0.0f;
    const float z = __fmul_rn(a, a);
    float p = fmaf(z, -0x1.931afcp-11f, 0x1.0236b0p-8f);
    p = fmaf(z, p, -0x1.3a0c64p-7f); p = fmaf(z, p, 0x1.03ebd4p-6f); p = fmaf(z, p, -0x1.6e1bfcp-6f);
    p = fmaf(z, p, 0x1.046a90p-5f); p = fmaf(z, p, -0x1.b295eep-5f); p = fmaf(z, p, 0x1.45f306p-3f);
    float r = __fmul_rn(p, a);
    if (ay > ax) r = __fsub_rn(0.25f, r);
    if (x < 0.0f) r = __fsub_rn(0.5f, r);
    if (y < 0.0f) r = -r;
    return r;
}

// term -> fixed point; `bad` is raised for NaN / Inf / out-of-range terms (the centre is then NaN)
__device__ __forceinline__ long long center_q(float term, bool &bad) {
    if (!(fabsf(term) < 1073741824.0f)) { bad = true; return 0; }
    return __float2ll_rn(__fmul_rn(term, kCenterScale));
}
__device__ __forceinline__ float center_mean(long long sum, int n) { return (float)(((double)sum * (1.0 / 16777216.0)) / (double)n); }
// circular-mean estimate from the fixed-point sums of cos / sin
__device__ __forceinline__ float center_estimate(long long sc, long long ss, float L) {
    return __fmul_rn(L, __fadd_rn(atan2_turns(-(float)ss, -(float)sc), 0.5f));
}

// nalgebra dot for 3-vectors: (x0*y0 + x1*y1) + x2*y2, no FMA.
__device__ __forceinline__ float dot_ref(const f3 &a, const f3 &b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ float norm_ref(const f3 &a) { return __fsqrt_rn(dot_ref(a, a)); }
__device__ __forceinline__ f3 cross_ref(const f3 &a, const f3 &b) {
    return mk3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
               __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// Vector3D::to_unit / Unit::new_normalize: component / norm (IEEE division).
__device__ __forceinline__ f3 unit_ref(const f3 &a) {
    float n = norm_ref(a);
    return mk3(__fdiv_rn(a.x, n), __fdiv_rn(a.y, n), __fdiv_rn(a.z, n));
}

// calc_sch (src/analysis/mod.rs:76-82):  S = 1.5 cos^2(theta) - 0.5 with
// theta = acos(clamp(v.n / (|v||n|))) (nalgebra Vector::angle).  cos(acos(c)) == c up to 2 ulp, so
// the device evaluates S from c directly:  c = v.n / (|v||n|),  S = (1.5 c) c - 0.5  (same
// operation order as the reference for the last step).  angle() returns 0 when a norm is 0 -> S = 1.
__device__ __forceinline__ float calc_sch(const f3 &v, const f3 &n) {
    float prod = dot_ref(v, n);
    float n1 = norm_ref(v), n2 = norm_ref(n);
    float c = __fdiv_rn(prod, __fmul_rn(n1, n2));
    c = fminf(1.0f, fmaxf(-1.0f, c));
    if (n1 == 0.0f || n2 == 0.0f) c = 1.0f;
    return __fsub_rn(__fmul_rn(__fmul_rn(1.5f, c), c), 0.5f);
}

// Static axis normal (0,0,1)-like: v.n == v[axis] exactly and |n| == 1, so the general expression
// reduces to c = v_axis / |v| with identical roundings.
__device__ __forceinline__ float calc_sch_axis(const f3 &v, float v_axis) {
    float n1 = norm_ref(v);
    float c = __fdiv_rn(v_axis, n1);
    c = fminf(1.0f, fmaxf(-1.0f, c));
    if (n1 == 0.0f) c = 1.0f;
    return __fsub_rn(__fmul_rn(__fmul_rn(1.5f, c), c), 0.5f);
}

// Streaming variants used by the accumulation kernels: c = v.n * rsqrt(|v|^2 |n|^2) (MUFU.RSQ,
// <= 2 ulp) instead of sqrt + IEEE division.  cos(acos(c)) of the reference is itself only defined
// to ~2 ulp of c, so nothing is lost: |dS| <= 4e-7 per sample either way.
__device__ __forceinline__ float calc_sch_fast(const f3 &v, const f3 &n) {
    const float prod = fmaf(v.z, n.z, fmaf(v.y, n.y, v.x * n.x));
    const float n1 = fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x)), n2 = fmaf(n.z, n.z, fmaf(n.y, n.y, n.x * n.x));
    const float den = n1 * n2;
    float c = prod * rsqrtf(den);
    c = fminf(1.0f, fmaxf(-1.0f, c));
    if (den == 0.0f) c = 1.0f;      // angle() returns 0 when a norm is 0 -> S = 1
    return fmaf(1.5f * c, c, -0.5f);   // NaN coordinates propagate (den != den -> c NaN -> S NaN)
}
// one MUFU.RSQ (rsqrtf() would add denormal scaling; |v|^2 of a bond is never denormal, and 0 is handled)
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float calc_sch_axis_fast(const f3 &v, float v_axis) {
    const float n1 = fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x));
    const float c = v_axis * rsqrt_ftz(n1);
    float c2 = fminf(c * c, 1.0f);   // clamp(c, -1, 1)^2
    if (n1 == 0.0f) c2 = 1.0f;       // angle() returns 0 when a norm is 0 -> S = 1
    return fmaf(1.5f, c2, -0.5f);    // NaN coordinates propagate (fminf would hide them: re-poison below)
}

// OrderValue::from(f32) in one FMUL + F2I: round-to-nearest of fl(S * 1e6).  fl() moves the product
// by <= 0.03 units of 1e-6, far below the +-0.3 unit uncertainty S already carries.
__device__ __forceinline__ int order_value_fast(float s) { return __float2int_rn(s * 1000000.0f); }

// OrderValue::from(f32) (src/analysis/order.rs:21-26): (value as f64 * 1e6).round() as i64.
// S is f32 and 1e6 = 2^6 * 15625, so the f64 product is exact; round() is half away from zero.
__device__ __forceinline__ int order_value(float s) {
    double x = (double)s * 1000000.0;
    return __double2int_rz(x + copysign(0.5, x));
}

// nalgebra Rotation3::from_axis_angle (Rodrigues) applied as R * v with rows summed left to
// right (uaorder.rs:958-1037); sin/cos of the angle are passed in.
__device__ __forceinline__ f3 rotate_axis(const f3 &v, const f3 &u, float s, float c) {
    float sqx = __fmul_rn(u.x, u.x), sqy = __fmul_rn(u.y, u.y), sqz = __fmul_rn(u.z, u.z);
    float omc = __fsub_rn(1.0f, c);
    float xy = __fmul_rn(__fmul_rn(u.x, u.y), omc), xz = __fmul_rn(__fmul_rn(u.x, u.z), omc),
          yz = __fmul_rn(__fmul_rn(u.y, u.z), omc);
    float xs = __fmul_rn(u.x, s), ys = __fmul_rn(u.y, s), zs = __fmul_rn(u.z, s);
    float m00 = __fadd_rn(sqx, __fmul_rn(__fsub_rn(1.0f, sqx), c)), m01 = __fsub_rn(xy, zs), m02 = __fadd_rn(xz, ys);
    float m10 = __fadd_rn(xy, zs), m11 = __fadd_rn(sqy, __fmul_rn(__fsub_rn(1.0f, sqy), c)), m12 = __fsub_rn(yz, xs);
    float m20 = __fsub_rn(xz, ys), m21 = __fadd_rn(yz, xs), m22 = __fadd_rn(sqz, __fmul_rn(__fsub_rn(1.0f, sqz), c));
    return mk3(__fadd_rn(__fadd_rn(__fmul_rn(m00, v.x), __fmul_rn(m01, v.y)), __fmul_rn(m02, v.z)),
               __fadd_rn(__fadd_rn(__fmul_rn(m10, v.x), __fmul_rn(m11, v.y)), __fmul_rn(m12, v.z)),
               __fadd_rn(__fadd_rn(__fmul_rn(m20, v.x), __fmul_rn(m21, v.y)), __fmul_rn(m22, v.z)));
}

// Vector3D::shift(direction, length): p += length * unit(direction).
__device__ __forceinline__ f3 shift_ref(const f3 &p, const f3 &dir, float len) {
    f3 u = unit_ref(dir);
    return mk3(__fadd_rn(p.x, __fmul_rn(u.x, len)), __fadd_rn(p.y, __fmul_rn(u.y, len)), __fadd_rn(p.z, __fmul_rn(u.z, len)));
}

}  // namespace gorder

/*
 * gorder_oracle.c — CPU restatement of gorder's per-frame order-parameter engine.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path (gorder_b200/, include/) may import,
 * link or call this file; it exists to check the CUDA engine (tests/, __graft_entry__.smoke())
 * and to be timed as the CPU baseline (bench.py cpu_baseline / --impl reference).
 *
 * The reference (VachaLab/gorder v1.4.1) is Rust and cannot be built in this image (no cargo /
 * rustc), so this is a restatement in plain C of the algorithm, function by
 * function, with the reference file:line each function follows.  Arithmetic that lives in the
 * un-vendored crates groan_rs 0.11.2 / nalgebra 0.34.0 / statistical 1.0.0 (Cargo.lock:684,965,1907)
 * is restated from their published behaviour and anchored on the reference's own call sites and
 * known-answer tests (tests/test_oracle_pins.py lists every pin).
 *
 * All floating-point work is f32 exactly as in the reference (compile with -ffp-contract=off:
 * rustc never contracts a*b+c into an FMA); sums are the reference's fixed-point i64.
 *
 * It consumes the same GorderSetup / GorderResults structs as the CUDA library so that the
 * parity tests feed both engines byte-identical inputs.
 */
#define _GNU_SOURCE
#include "../include/gorder_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_F 3.14159265358979323846f /* std::f32::consts::PI */

typedef struct { float x, y, z; } vec3;

/* ------------------------------------------------------------------------------------------
 * groan_rs Vector3D / SimBox arithmetic (un-vendored)
 * ---------------------------------------------------------------------------------------- */

static inline vec3 v3(float x, float y, float z) { vec3 v = {x, y, z}; return v; }
static inline vec3 vadd(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 vsub(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 vdivs(vec3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline float get_c(vec3 v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

/* nalgebra dot for 3-vectors: (x0*y0 + x1*y1) + x2*y2, no FMA. */
static inline float vdot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float vnorm(vec3 a) { return sqrtf(vdot(a, a)); }
static inline vec3 vcross(vec3 a, vec3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* Vector3D::to_unit / nalgebra Unit::new_normalize: component / norm. */
static inline vec3 vunit(vec3 a) { float n = vnorm(a); return vdivs(a, n); }

/* Variant switches, kept only so that tests/test_oracle_pins.py can show that the alternatives do
 * NOT reproduce the reference's goldens.  Frozen values: min-image 1, wrap 0 — with them all 18
 * hydrogen coordinates of uaorder.rs:1113-1200 are reproduced to 0 ulp and the single-frame sums of
 * cgorder.rs:188-241 to the goldens' printed precision. */
int gorder_oracle_variant_minimage = 1; /* 0: fold loop, 1: (((d+h)%L)+L)%L-h */
int gorder_oracle_variant_wrap = 0;     /* 0: c%L, +L if negative, 1: ((c%L)+L)%L */

/* Vector3D::vector_to -> per-component minimum image (call site pbc.rs:378-385).
 * groan_rs folds with the "positive modulo" idiom  (((d + L/2) % L) + L) % L - L/2  in f32; the
 * roundings of that expression are what the reference comment at pbc.rs:379-383 refers to
 * ("introduces minor numerical errors compared to the NoPBC version") and they are visible in the
 * single-frame goldens cgorder.rs:188-241 / aaorder.rs:226-350: with an exact fold the sums are off
 * by up to 6e-5, with this expression they agree to the goldens' printed precision. */
static inline float min_image(float d, float L) {
    float half = L / 2.0f;
    if (!(L > 0.0f)) return d; /* degenerate dimension: nothing to fold */
    if (gorder_oracle_variant_minimage == 0) {
        while (d > half) d -= L;
        while (d < -half) d += L;
        return d;
    }
    return fmodf(fmodf(d + half, L) + L, L) - half;
}

/* PBC3D::vector_to (pbc.rs:378-385) / NoPBC::vector_to (pbc.rs:180-182). */
static inline vec3 vector_to(vec3 p1, vec3 p2, const float *box, int pbc) {
    vec3 d = vsub(p2, p1);
    if (pbc) { d.x = min_image(d.x, box[0]); d.y = min_image(d.y, box[1]); d.z = min_image(d.z, box[2]); }
    return d;
}

/* Vector3D::wrap (call site pbc.rs:388-390): coordinate into [0, L). */
static inline float wrap1(float c, float L) {
    if (!(L > 0.0f)) return c;
    if (gorder_oracle_variant_wrap == 0) {
        float w = fmodf(c, L);
        if (w < 0.0f) w += L;
        return w;
    }
    return fmodf(fmodf(c, L) + L, L);
}
static inline vec3 wrap_point(vec3 p, const float *box, int pbc) {
    if (!pbc) return p; /* NoPBC::wrap is a no-op (pbc.rs:185) */
    return v3(wrap1(p.x, box[0]), wrap1(p.y, box[1]), wrap1(p.z, box[2]));
}

/* Vector3D::distance(other, Dimension::X|Y|Z, box): signed minimum-image self - other
 * (call sites pbc.rs:354-356, leaflets.rs:725, :788). distance_naive: plain difference. */
static inline float distance_1d(vec3 a, vec3 b, int axis, const float *box, int pbc) {
    float d = get_c(a, axis) - get_c(b, axis);
    return pbc ? min_image(d, box[axis]) : d;
}

/* Vector3D::angle -> nalgebra Vector::angle: acos(clamp(a.b/(|a||b|), -1, 1)), 0 if a norm is 0. */
static inline float vangle(vec3 a, vec3 b) {
    float prod = vdot(a, b), n1 = vnorm(a), n2 = vnorm(b);
    if (n1 == 0.0f || n2 == 0.0f) return 0.0f;
    float c = prod / (n1 * n2);
    if (c < -1.0f) c = -1.0f;
    if (c > 1.0f) c = 1.0f;
    return acosf(c);
}

/* calc_sch (analysis/mod.rs:76-82). */
float gorder_oracle_calc_sch(const float *vec, const float *normal) {
    float angle = vangle(v3(vec[0], vec[1], vec[2]), v3(normal[0], normal[1], normal[2]));
    float c = cosf(angle);
    return (1.5f * c * c) - 0.5f;
}
static inline float calc_sch(vec3 v, vec3 n) {
    float c = cosf(vangle(v, n));
    return (1.5f * c * c) - 0.5f;
}

/* OrderValue::from(f32) (order.rs:21-26): (value as f64 * 1e6).round() as i64. */
static inline int64_t order_value(float s) { return (int64_t)round((double)s * 1000000.0); }
int64_t gorder_oracle_order_value(float s) { return order_value(s); }

void gorder_oracle_vector_to(const float *p1, const float *p2, const float *box, int pbc, float *out) {
    vec3 d = vector_to(v3(p1[0], p1[1], p1[2]), v3(p2[0], p2[1], p2[2]), box, pbc);
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}

/* nalgebra Rotation3::from_axis_angle (Rodrigues) applied as R * v; rows summed left to right. */
static vec3 rotate_axis_angle(vec3 v, vec3 u, float angle) {
    if (angle == 0.0f) return v;
    float sqx = u.x * u.x, sqy = u.y * u.y, sqz = u.z * u.z;
    float s = sinf(angle), c = cosf(angle), omc = 1.0f - c;
    float m00 = sqx + (1.0f - sqx) * c, m01 = u.x * u.y * omc - u.z * s, m02 = u.x * u.z * omc + u.y * s;
    float m10 = u.x * u.y * omc + u.z * s, m11 = sqy + (1.0f - sqy) * c, m12 = u.y * u.z * omc - u.x * s;
    float m20 = u.x * u.z * omc - u.y * s, m21 = u.y * u.z * omc + u.x * s, m22 = sqz + (1.0f - sqz) * c;
    return v3((m00 * v.x + m01 * v.y) + m02 * v.z, (m10 * v.x + m11 * v.y) + m12 * v.z,
              (m20 * v.x + m21 * v.y) + m22 * v.z);
}

/* Vector3D::shift(direction, length): p += length * unit(direction). */
static inline vec3 shift(vec3 p, vec3 dir, float len) {
    vec3 u = vunit(dir);
    return v3(p.x + u.x * len, p.y + u.y * len, p.z + u.z * len);
}

/* ------------------------------------------------------------------------------------------
 * United-atom hydrogen reconstruction (uaorder.rs:35-41, :947-1104)
 * ---------------------------------------------------------------------------------------- */
#define TETRAHEDRAL_ANGLE 1.910633f
#define TETRAHEDRAL_ANGLE_HALF 0.9553165f
#define BOND_LENGTH 0.109f
#define CH3_ANGLE 2.0943952f

/* returns number of hydrogens written to out[3] */
static int predict_hydrogens(int kind, vec3 t, vec3 h1, vec3 h2, vec3 h3, const float *box, int pbc, vec3 *out) {
    switch (kind) {
    case GORDER_UA_CH3: { /* predict_hydrogens_ch3, uaorder.rs:947-981 */
        vec3 th1 = vector_to(t, h1, box, pbc), th2 = vector_to(t, h2, box, pbc);
        vec3 axis = vunit(vcross(th2, th1));
        vec3 hv1 = rotate_axis_angle(th1, axis, TETRAHEDRAL_ANGLE);
        out[0] = wrap_point(shift(t, hv1, BOND_LENGTH), box, pbc);
        vec3 nth1 = vunit(th1);
        out[1] = wrap_point(shift(t, rotate_axis_angle(hv1, nth1, CH3_ANGLE), BOND_LENGTH), box, pbc);
        out[2] = wrap_point(shift(t, rotate_axis_angle(hv1, nth1, -CH3_ANGLE), BOND_LENGTH), box, pbc);
        return 3;
    }
    case GORDER_UA_CH2: { /* predict_hydrogens_ch2, uaorder.rs:985-1020 */
        vec3 th1 = vunit(vector_to(t, h1, box, pbc)), th2 = vunit(vector_to(t, h2, box, pbc));
        vec3 plane_normal = vcross(th2, th1);
        vec3 rot_axis = vunit(vsub(th1, th2));
        vec3 rot_vec = vcross(plane_normal, rot_axis);
        vec3 ax = vunit(rot_axis); /* Unit::new_normalize of an already-unit vector */
        out[0] = wrap_point(shift(t, rotate_axis_angle(rot_vec, ax, TETRAHEDRAL_ANGLE_HALF), BOND_LENGTH), box, pbc);
        out[1] = wrap_point(shift(t, rotate_axis_angle(rot_vec, ax, -TETRAHEDRAL_ANGLE_HALF), BOND_LENGTH), box, pbc);
        return 2;
    }
    case GORDER_UA_CH1_UNSAT: { /* predict_hydrogen_ch1_unsaturated, uaorder.rs:1024-1045 */
        vec3 th1 = vector_to(t, h1, box, pbc), th2 = vector_to(t, h2, box, pbc);
        float gamma = vangle(th1, th2);
        vec3 axis = vunit(vcross(th1, th2));
        vec3 hv = rotate_axis_angle(th2, axis, PI_F - (gamma / 2.0f));
        out[0] = wrap_point(shift(t, hv, BOND_LENGTH), box, pbc);
        return 1;
    }
    case GORDER_UA_CH1_SAT: { /* predict_hydrogen_ch1_saturated, uaorder.rs:1087-1104 */
        vec3 a = vunit(vector_to(t, h1, box, pbc)), b = vunit(vector_to(t, h2, box, pbc)),
             c = vunit(vector_to(t, h3, box, pbc));
        vec3 s = vadd(vadd(a, b), c);
        vec3 hv = v3(-s.x, -s.y, -s.z);
        out[0] = wrap_point(shift(t, hv, BOND_LENGTH), box, pbc);
        return 1;
    }
    }
    return 0;
}

int gorder_oracle_predict_hydrogens(int kind, const float *t, const float *h1, const float *h2, const float *h3,
                                    const float *box, int pbc, float *out) {
    vec3 o[3];
    vec3 hh3 = h3 ? v3(h3[0], h3[1], h3[2]) : v3(0, 0, 0);
    int n = predict_hydrogens(kind, v3(t[0], t[1], t[2]), v3(h1[0], h1[1], h1[2]), v3(h2[0], h2[1], h2[2]), hh3,
                              box, pbc, o);
    for (int i = 0; i < n; i++) { out[3 * i] = o[i].x; out[3 * i + 1] = o[i].y; out[3 * i + 2] = o[i].z; }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Centre of geometry of a group (System::group_get_center / _naive; pbc.rs:101,270)
 * groan_rs: "refined Bai-Breen" (CHANGELOG.md v1.0.0): circular-mean estimate per axis, then the
 * mean of minimum-image displacements around the estimate, wrapped into the box.
 *
 * groan folds these sums sequentially in f32.  A sequential f32 fold over 1e4 - 1e6 atoms cannot be
 * reproduced by a parallel machine, and its own rounding noise (~1e-7 L sqrt(N)) is larger than
 * anything below, so the engine AND this oracle use one ORDER-FREE definition:
 *   - every term (cos, sin, minimum-image displacement, coordinate) is computed in f32 and added
 *     as the integer  llrint(term * 2^24)  (exact, associative: any order gives the same bits);
 *   - the circular-mean estimate only seeds the refinement (the centre is independent of it up to
 *     rounding), so its transcendental functions are evaluated by fixed polynomials in IEEE
 *     operations and fmaf (sincos_turns / atan2_turns: <= 3e-8 / 2e-8 turns) instead of libm, whose
 *     results no device reproduces bit for bit.
 * variant 1 (gorder_oracle_variant_center) is the previous sequential-f32 / libm evaluation: the
 * pins show that both agree to an ulp of the centre and give identical leaflets on every fixture.
 * ---------------------------------------------------------------------------------------- */
int gorder_oracle_variant_center = 0; /* 0: order-free fixed point + polynomial seed; 1: sequential f32 + libm */

#define CENTER_SCALE 16777216.0f /* 2^24 */

/* sin / cos of 2 pi u (u in turns), IEEE ops + fmaf only: Taylor about the nearest quarter turn, |t| <= 1/8 */
static inline void sincos_turns(float u, float *sn, float *cs) {
    const float r = u - rintf(u);          /* exact, [-0.5, 0.5] */
    const float j = rintf(4.0f * r);       /* -2 .. 2 */
    const float t = fmaf(j, -0.25f, r);    /* exact */
    const float z = t * t;
    float sp = fmaf(z, 0x1.507834p+5f, -0x1.32d2ccp+6f);
    sp = fmaf(z, sp, 0x1.466bc6p+6f); sp = fmaf(z, sp, -0x1.4abbcep+5f); sp = fmaf(z, sp, 0x1.921fb6p+2f);
    sp = sp * t;
    float cp = fmaf(z, 0x1.e1f506p+5f, -0x1.55d3c8p+6f);
    cp = fmaf(z, cp, 0x1.03c1f0p+6f); cp = fmaf(z, cp, -0x1.3bd3ccp+4f); cp = fmaf(z, cp, 1.0f);
    const int q = (int)j & 3;
    *sn = q == 0 ? sp : (q == 1 ? cp : (q == 2 ? -sp : -cp));
    *cs = q == 0 ? cp : (q == 1 ? -sp : (q == 2 ? -cp : sp));
}

/* atan2(y, x) / 2 pi in [-0.5, 0.5], IEEE ops + fmaf only (minimax of atan(a) / (2 pi a) in a^2, <= 2e-8 turns) */
static inline float atan2_turns(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    const float a = mx > 0.0f ? mn / mx : 0.0f;
    const float z = a * a;
    float p = fmaf(z, -0x1.931afcp-11f, 0x1.0236b0p-8f);
    p = fmaf(z, p, -0x1.3a0c64p-7f); p = fmaf(z, p, 0x1.03ebd4p-6f); p = fmaf(z, p, -0x1.6e1bfcp-6f);
    p = fmaf(z, p, 0x1.046a90p-5f); p = fmaf(z, p, -0x1.b295eep-5f); p = fmaf(z, p, 0x1.45f306p-3f);
    float r = p * a;
    if (ay > ax) r = 0.25f - r;
    if (x < 0.0f) r = 0.5f - r;
    if (y < 0.0f) r = -r;
    return r;
}

/* term -> fixed point; *bad is raised for NaN / Inf / out-of-range terms (the centre is then NaN) */
static inline int64_t center_q(float term, int *bad) {
    if (!(fabsf(term) < 1073741824.0f)) { *bad = 1; return 0; }
    return llrintf(term * CENTER_SCALE);
}
static inline float center_mean(int64_t sum, int n) { return (float)(((double)sum * (1.0 / 16777216.0)) / (double)n); }

static vec3 group_center_sequential(const float *xyz, const int32_t *idx, int n, const float *box, int pbc) {
    if (!pbc) {
        vec3 s = v3(0, 0, 0);
        for (int i = 0; i < n; i++) { const float *p = xyz + 3 * (size_t)idx[i]; s = vadd(s, v3(p[0], p[1], p[2])); }
        return vdivs(s, (float)n);
    }
    float c0[3];
    for (int a = 0; a < 3; a++) {
        float sc = 0.0f, ss = 0.0f, L = box[a];
        float scale = 2.0f * PI_F / L;
        for (int i = 0; i < n; i++) {
            float th = xyz[3 * (size_t)idx[i] + a] * scale;
            sc += cosf(th);
            ss += sinf(th);
        }
        float th = atan2f(-ss, -sc) + PI_F;
        c0[a] = L * th / (2.0f * PI_F);
    }
    vec3 est = v3(c0[0], c0[1], c0[2]);
    vec3 s = v3(0, 0, 0);
    for (int i = 0; i < n; i++) {
        const float *p = xyz + 3 * (size_t)idx[i];
        s = vadd(s, vector_to(est, v3(p[0], p[1], p[2]), box, 1));
    }
    vec3 c = vadd(est, vdivs(s, (float)n));
    return wrap_point(c, box, 1);
}

static vec3 group_center(const float *xyz, const int32_t *idx, int n, const float *box, int pbc) {
    if (n <= 0) { float q = nanf(""); return v3(q, q, q); }
    if (gorder_oracle_variant_center == 1) return group_center_sequential(xyz, idx, n, box, pbc);
    float c[3];
    for (int a = 0; a < 3; a++) {
        int bad = 0;
        const float L = box ? box[a] : 0.0f;
        if (!pbc) {
            int64_t sx = 0;
            for (int i = 0; i < n; i++) sx += center_q(xyz[3 * (size_t)idx[i] + a], &bad);
            c[a] = bad ? nanf("") : center_mean(sx, n);
            continue;
        }
        const float inv = 1.0f / L;
        int64_t sc = 0, ss = 0, sd = 0;
        for (int i = 0; i < n; i++) {
            float sn, cs;
            sincos_turns(xyz[3 * (size_t)idx[i] + a] * inv, &sn, &cs);
            sc += center_q(cs, &bad); ss += center_q(sn, &bad);
        }
        const float est = L * (atan2_turns(-(float)ss, -(float)sc) + 0.5f);
        for (int i = 0; i < n; i++) sd += center_q(min_image(xyz[3 * (size_t)idx[i] + a] - est, L), &bad);
        if (bad || !(est == est)) { c[a] = nanf(""); continue; }
        c[a] = wrap1(est + center_mean(sd, n), L);
    }
    return v3(c[0], c[1], c[2]);
}

void gorder_oracle_group_center(const float *xyz, const int32_t *idx, int n, const float *box, int pbc, float *out) {
    vec3 c = group_center(xyz, idx, n, box, pbc);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

/* ------------------------------------------------------------------------------------------
 * Geometry selection (geometry.rs:181-210, :328-357, :422-451, :507-514; groan shapes)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int kind, invert, axis;
    vec3 origin;     /* cuboid: lower corner; cylinder: centre of base; sphere: centre */
    float len[3];    /* cuboid extents */
    float radius, height;
} Shape;

static int is_inf_span(float lo, float hi) { return isinf(lo) && lo < 0 && isinf(hi) && hi > 0; }

/* construct_shape for the three selections, given this frame's reference point. */
static Shape construct_shape(const GorderSetup *s, vec3 ref, const float *box) {
    Shape sh;
    memset(&sh, 0, sizeof(sh));
    sh.kind = s->geom_kind; sh.invert = s->geom_invert; sh.axis = s->geom_axis;
    int pbc = s->handle_pbc;
    float p[3] = {ref.x, ref.y, ref.z};
    if (s->geom_kind == GORDER_GEOM_CUBOID) { /* geometry.rs:328-357 */
        for (int a = 0; a < 3; a++) {
            float lo = s->geom_dims[2 * a], hi = s->geom_dims[2 * a + 1];
            if (is_inf_span(lo, hi)) { /* get_infinite_span: pbc.rs:393-396 / :188-191 */
                p[a] = pbc ? 0.0f : -3.40282347e+38f;
                sh.len[a] = INFINITY;
            } else { p[a] += lo; sh.len[a] = hi - lo; }
        }
    } else if (s->geom_kind == GORDER_GEOM_CYLINDER) { /* geometry.rs:422-451 */
        float lo = s->geom_dims[1], hi = s->geom_dims[2];
        sh.radius = s->geom_dims[0];
        if (is_inf_span(lo, hi)) { p[s->geom_axis] = pbc ? 0.0f : -3.40282347e+38f; sh.height = INFINITY; }
        else { p[s->geom_axis] += lo; sh.height = hi - lo; }
    } else if (s->geom_kind == GORDER_GEOM_SPHERE) { /* geometry.rs:507-514 */
        sh.radius = s->geom_dims[0];
    }
    {   /* a fixed reference point: the shape was built once, with the structure file's box (geometry.rs:297-312) */
        const float *wb = box;
        if (s->geom_ref_kind == GORDER_GEOMREF_POINT && (s->structure_box[0] != 0.0f || s->structure_box[1] != 0.0f || s->structure_box[2] != 0.0f))
            wb = s->structure_box;
        sh.origin = wrap_point(v3(p[0], p[1], p[2]), wb, pbc);
    }
    return sh;
}

/* Shape::inside (PBC) / NaiveShape::inside_naive, XOR invert (geometry.rs:181-190). */
static int shape_inside(const Shape *sh, vec3 pt, const float *box, int pbc) {
    int in = 1;
    float o[3] = {sh->origin.x, sh->origin.y, sh->origin.z}, q[3] = {pt.x, pt.y, pt.z};
    switch (sh->kind) {
    case GORDER_GEOM_NONE: return 1;
    case GORDER_GEOM_CUBOID:
        for (int a = 0; a < 3 && in; a++) {
            if (pbc) { float d = wrap1(q[a] - o[a], box[a]); in = (d <= sh->len[a]); }
            else in = (q[a] >= o[a]) && (q[a] <= o[a] + sh->len[a]);
        }
        break;
    case GORDER_GEOM_CYLINDER: {
        int ax = sh->axis;
        float r2 = 0.0f;
        for (int a = 0; a < 3; a++) {
            if (a == ax) continue;
            float d = q[a] - o[a];
            if (pbc) d = min_image(d, box[a]);
            r2 += d * d;
        }
        in = sqrtf(r2) < sh->radius;
        if (in) {
            if (pbc) { float d = wrap1(q[ax] - o[ax], box[ax]); in = (d <= sh->height); }
            else { float d = q[ax] - o[ax]; in = (d >= 0.0f) && (d <= sh->height); }
        }
        break;
    }
    case GORDER_GEOM_SPHERE: {
        vec3 d = vector_to(sh->origin, pt, box, pbc);
        in = vnorm(d) < sh->radius;
        break;
    }
    }
    return in ^ (sh->invert ? 1 : 0);
}

int gorder_oracle_shape_inside(const GorderSetup *s, const float *ref, const float *pt, const float *box) {
    Shape sh = construct_shape(s, v3(ref[0], ref[1], ref[2]), box);
    return shape_inside(&sh, v3(pt[0], pt[1], pt[2]), box, s->handle_pbc);
}
void gorder_oracle_shape_origin(const GorderSetup *s, const float *ref, const float *box, float *out) {
    Shape sh = construct_shape(s, v3(ref[0], ref[1], ref[2]), box);
    out[0] = sh.origin.x; out[1] = sh.origin.y; out[2] = sh.origin.z;
    out[3] = sh.len[0]; out[4] = sh.len[1]; out[5] = sh.len[2]; out[6] = sh.radius; out[7] = sh.height;
}

/* ------------------------------------------------------------------------------------------
 * PCA normal (normal.rs:421-458): smallest right-singular vector of the demeaned N x 3 cloud ==
 * eigenvector of the smallest eigenvalue of the 3x3 scatter matrix (sign is implementation-defined
 * in nalgebra; S is sign-invariant).  Centroid and demeaning in f32 as the reference; the 3x3
 * eigenproblem by cyclic Jacobi in f64.
 * ---------------------------------------------------------------------------------------- */
static void jacobi_eig3(double a[3][3], double v[3][3], double w[3]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) v[i][j] = (i == j);
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (fabs(a[p][q]) < 1e-300) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++) w[i] = a[i][i];
}

/* ------------------------------------------------------------------------------------------
 * nalgebra 0.34 SVD::new(data, true, true) of the demeaned N x 3 cloud (normal.rs:443), restated: the reference takes
 * the LAST ROW of V^T, sign included (its exported normals are compared signed, tests/common/mod.rs:84-87), so the
 * conventions of nalgebra's algorithm are part of the result:
 *   - the matrix is divided by its largest |entry|;
 *   - Bidiagonal::new: Householder reflections with the axis  x + sign(x0) |x| e1  (normalised twice), applied as
 *     sign * H with sign = signum(-sign(x0) |x|), alternately to columns and rows; v_t() replays the row reflections
 *     on the identity, from the last to the first, with the sign of the stored (signed) off-diagonal;
 *   - the diagonal and off-diagonal handed to the iteration are the MODULI of the stored values;
 *   - implicit-shift QR sweeps (Wilkinson shift of the trailing 2 x 2 of B^T B) by Givens rotations
 *     (GivensRotation::cancel_y: c = |x| / r, s = -y / (sign(x) r)), V^T rotated by the right-hand rotation; off-diagonal
 *     entries are zeroed when <= 5 eps (|d_i| + |d_i+1|); a remaining 2 x 2 block by compute_2x2_uptrig_svd;
 *   - rows sorted by decreasing singular value (the rows keep their sign).
 * Pinned by the reference's own signed vectors: the 274 normals of normal.rs:664-963 (pcpepg.tpr, P atoms, r = 2 nm) are
 * reproduced with their signs (tests/test_oracle_pins.py), and so is tests/files/ua_normals.yaml.
 * Zero diagonal entries (singular value <= 5 eps of the largest entry: collinear clouds) are not chased as nalgebra does;
 * the sweep then simply stops at 64 iterations.
 * ---------------------------------------------------------------------------------------- */
static inline float signumf_rust(float x) { return signbit(x) ? -1.0f : 1.0f; }

/* nalgebra's dot (base/blas.rs): chunks of eight products, res += (p0 + p4); (p1 + p5); (p2 + p6); (p3 + p7); then the tail */
static float dot8(const float *x, int sx, const float *y, int sy, int n) {
    float res = 0.0f;
    int i = 0;
    for (; n - i >= 8; i += 8) {
        float p[8];
        for (int k = 0; k < 8; k++) p[k] = x[(i + k) * sx] * y[(i + k) * sy];
        res += p[0] + p[4]; res += p[1] + p[5]; res += p[2] + p[6]; res += p[3] + p[7];
    }
    for (; i < n; i++) res += x[i * sx] * y[i * sy];
    return res;
}

/* householder::reflection_axis_mut on a strided vector; returns the (signed) reflection norm, *nz = axis is usable */
static float reflection_axis(float *col, int n, int stride, int *nz) {
    float sq = dot8(col, stride, col, stride, n);
    const float norm = sqrtf(sq);
    const float modulus = fabsf(col[0]), sign = signumf_rust(col[0]);
    const float signed_norm = sign * norm;
    const float factor = (sq + modulus * norm) * 2.0f;
    col[0] += signed_norm;
    if (factor != 0.0f) {
        const float f = sqrtf(factor);
        for (int i = 0; i < n; i++) col[i * stride] /= f;
        const float nn = sqrtf(dot8(col, stride, col, stride, n));
        for (int i = 0; i < n; i++) col[i * stride] /= nn;
        *nz = 1;
        return -signed_norm;
    }
    *nz = 0;
    return signed_norm;
}

typedef struct { float c, s; } givens_t;
static int givens_cancel_y(float x, float y, givens_t *g, float *r) {
    if (y == 0.0f) return 0;
    const float mod0 = fabsf(x), sign0 = signumf_rust(x);
    const float denom = sqrtf(mod0 * mod0 + y * y);
    g->c = mod0 / denom; g->s = -y / (sign0 * denom);
    *r = sign0 * denom;
    return 1;
}
static void givens_new(float c, float s, givens_t *g, float *norm) {
    const float mod0 = fabsf(c), sign0 = signumf_rust(c);
    const float denom = sqrtf(mod0 * mod0 + s * s);
    if (denom > 0.0f) { *norm = sign0 * denom; g->c = mod0 / denom; g->s = s / *norm; }
    else { g->c = 1.0f; g->s = 0.0f; *norm = 0.0f; }
}
/* rows k, k + 1 of a 3-column matrix: rhs = R rhs with R = [[c, -s], [s, c]] */
static void givens_rotate_rows3(givens_t g, float *r0, float *r1) {
    for (int j = 0; j < 3; j++) { const float a = r0[j], b = r1[j]; r0[j] = a * g.c - g.s * b; r1[j] = g.s * a + b * g.c; }
}

static void svd_delimit(float *d, float *o, int end, float eps, int *start_out, int *end_out) {
    int n = end;
    while (n > 0) {
        const int m = n - 1;
        if (o[m] == 0.0f || fabsf(o[m]) <= eps * (fabsf(d[n]) + fabsf(d[m]))) o[m] = 0.0f;
        else break;
        n--;
    }
    if (n == 0) { *start_out = 0; *end_out = 0; return; }
    int ns = n - 1;
    while (ns > 0) {
        const int m = ns - 1;
        if (fabsf(o[m]) <= eps * (fabsf(d[ns]) + fabsf(d[m]))) { o[m] = 0.0f; break; }
        ns--;
    }
    *start_out = ns; *end_out = n;
}

/* a: [n][3] row-major, overwritten.  out: last row of V^T of SVD::new(a). */
static void nalgebra_svd_last_row(float *a, int n, float out[3]) {
    float amax = 0.0f;
    for (int i = 0; i < 3 * n; i++) amax = fmaxf(amax, fabsf(a[i]));
    if (amax != 0.0f) for (int i = 0; i < 3 * n; i++) a[i] /= amax;
    float ds[3], os[2];   /* stored (signed) diagonal / off-diagonal of Bidiagonal::new */
    int nz;
    for (int ite = 0; ite < 2; ite++) {
        /* clear_column_unchecked(matrix, ite, 0, None) */
        float *axis = a + 3 * ite + ite;   /* rows ite.., column ite, stride 3 */
        const int len = n - ite;
        float rn = reflection_axis(axis, len, 3, &nz);
        if (nz) {
            const float sign = signumf_rust(rn);
            for (int j = ite + 1; j < 3; j++) {
                float *col = a + 3 * ite + j;
                const float factor = dot8(axis, 3, col, 3, len) * (sign * -2.0f);
                for (int i = 0; i < len; i++) col[3 * i] = factor * axis[3 * i] + sign * col[3 * i];
            }
        }
        ds[ite] = rn;
        /* clear_row_unchecked(matrix, axis_packed, work, ite, 1) */
        float rax[2];
        const int rl = 3 - (ite + 1);
        for (int j = 0; j < rl; j++) rax[j] = a[3 * ite + ite + 1 + j];
        rn = reflection_axis(rax, rl, 1, &nz);
        if (nz) {
            const float sign = signumf_rust(rn);
            for (int i = ite + 1; i < n; i++) {
                float *row = a + 3 * i + ite + 1;
                float w = 0.0f;
                for (int j = 0; j < rl; j++) w += row[j] * rax[j];
                const float f = w * (sign * -2.0f);
                for (int j = 0; j < rl; j++) row[j] = sign * row[j] + f * rax[j];
            }
        }
        for (int j = 0; j < rl; j++) a[3 * ite + ite + 1 + j] = rax[j];
        os[ite] = rn;
    }
    ds[2] = reflection_axis(a + 3 * 2 + 2, n - 2, 3, &nz);
    /* v_t(): the row reflections replayed on the identity, i = 1, 0 */
    float vt[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 1; i >= 0; i--) {
        const int rl = 3 - (i + 1);
        const float *axis = a + 3 * i + i + 1;
        float sq = 0.0f;
        for (int j = 0; j < rl; j++) sq += axis[j] * axis[j];
        if (sq == 0.0f) continue;
        const float sign = signumf_rust(os[i]);
        for (int r = i; r < 3; r++) {
            float w = 0.0f;
            for (int j = 0; j < rl; j++) w += vt[r][i + 1 + j] * axis[j];
            const float f = w * (sign * -2.0f);
            for (int j = 0; j < rl; j++) vt[r][i + 1 + j] = sign * vt[r][i + 1 + j] + f * axis[j];
        }
    }
    float d[3] = {fabsf(ds[0]), fabsf(ds[1]), fabsf(ds[2])}, o[2] = {fabsf(os[0]), fabsf(os[1])};
    const float eps = 1.1920929e-07f * 5.0f;
    int start, end;
    svd_delimit(d, o, 2, eps, &start, &end);
    for (int niter = 0; end != start && niter < 64; niter++) {
        const int subdim = end - start + 1;
        if (subdim > 2) {
            const int m = end - 1, nn = end;
            const float dm = d[m], dn = d[nn], fm = o[m];
            const float tmm = dm * dm + o[m - 1] * o[m - 1], tmn = dm * fm, tnn = dn * dn + fm * fm;
            float shift = tnn;
            const float sq = tmn * tmn;
            if (sq != 0.0f) {   /* symmetric_eigen::wilkinson_shift */
                const float dd = (tmm - tnn) * 0.5f;
                shift = tnn - sq / (dd + signumf_rust(dd) * sqrtf(dd * dd + sq));
            }
            float vx = d[start] * d[start] - shift, vy = d[start] * o[start];
            for (int k = start; k < nn; k++) {
                const float m12 = (k == nn - 1) ? 0.0f : o[k + 1];
                float s00 = d[k], s01 = o[k], s02 = 0.0f, s10 = 0.0f, s11 = d[k + 1], s12 = m12;
                givens_t r1, r2;
                float norm1, norm2;
                if (!givens_cancel_y(vx, vy, &r1, &norm1)) break;
                {   /* rot1.inverse().rotate_rows(columns 0..2): lhs = lhs * [[c, s], [-s, c]] */
                    const float c = r1.c, sn = -r1.s;
                    float a0 = s00, b0 = s01; s00 = a0 * c + sn * b0; s01 = -sn * a0 + b0 * c;
                    a0 = s10; b0 = s11; s10 = a0 * c + sn * b0; s11 = -sn * a0 + b0 * c;
                }
                if (k > start) o[k - 1] = norm1;
                if (!givens_cancel_y(s00, s10, &r2, &norm2)) { r2.c = 1.0f; r2.s = 0.0f; norm2 = s00; }
                {   /* rot2.rotate(columns 1..3) */
                    float a0 = s01, b0 = s11; s01 = a0 * r2.c - r2.s * b0; s11 = r2.s * a0 + b0 * r2.c;
                    a0 = s02; b0 = s12; s02 = a0 * r2.c - r2.s * b0; s12 = r2.s * a0 + b0 * r2.c;
                }
                s00 = norm2;
                givens_rotate_rows3(r1, vt[k], vt[k + 1]);
                d[k] = s00; d[k + 1] = s11; o[k] = s01;
                if (k != nn - 1) o[k + 1] = s12;
                vx = s01; vy = s02;
            }
        } else if (subdim == 2) {   /* compute_2x2_uptrig_svd */
            const float m11 = d[start], m12 = o[start], m22 = d[start + 1];
            const float denom = hypotf(m11 + m22, m12) + hypotf(m11 - m22, m12);
            float v1 = m11 * m22 * 2.0f / denom, v2 = 0.5f * denom;
            givens_t csv, csu;
            float sgn_v, sgn_u;
            /* nalgebra builds the rotation from (m11 m12, v1^2 - m11^2), which is free of cancellation when |m11| >= |m22| (v1
             * is then the singular value next to m22).  QR sweeps may leave the block the other way round; the difference is
             * then evaluated from the characteristic equation, v1^2 - m11^2 = -m11^2 m12^2 / (m12^2 + m22^2 - v1^2): the same
             * number in exact arithmetic (a borderline deflation decides whether the reference gets here at all). */
            const float diff = fabsf(m11) >= fabsf(m22) ? v1 * v1 - m11 * m11 : -(m11 * m11 * m12 * m12) / (m12 * m12 + m22 * m22 - v1 * v1);
            givens_new(m11 * m12, diff, &csv, &sgn_v);
            v1 *= sgn_v; v2 *= sgn_v;
            const float cu = (m11 * csv.c + m12 * csv.s) / v1, su = (m22 * csv.s) / v1;
            givens_new(cu, su, &csu, &sgn_u);
            v1 *= sgn_u; v2 *= sgn_u;
            d[start] = v1; d[start + 1] = v2; o[start] = 0.0f;
            givens_t inv = {csv.c, -csv.s};
            givens_rotate_rows3(inv, vt[start], vt[start + 1]);
            end -= 1;
        }
        svd_delimit(d, o, end, eps, &start, &end);
    }
    int k = 0;   /* sort_by_singular_values: the last row belongs to the smallest singular value (stable) */
    for (int i = 1; i < 3; i++) if (fabsf(d[i]) <= fabsf(d[k])) k = i;
    out[0] = vt[k][0]; out[1] = vt[k][1]; out[2] = vt[k][2];
}

/* membrane_normal_from_cloud (normal.rs:421-458).  returns 0 ok, GORDER_ERR_DYNAMIC_NORMAL_POINTS if n < 3 */
int gorder_oracle_variant_normal = 0;   /* 0: nalgebra SVD restatement (signed); 1: eigenvector of the scatter matrix (sign arbitrary) */
static int normal_from_cloud(const vec3 *pts, int n, vec3 *out) {
    if (n < 3) return GORDER_ERR_DYNAMIC_NORMAL_POINTS;
    vec3 c = v3(0, 0, 0);
    for (int i = 0; i < n; i++) c = vadd(c, pts[i]);
    c = vdivs(c, (float)n);
    if (gorder_oracle_variant_normal == 0) {
        float *a = (float *)malloc(sizeof(float) * 3 * (size_t)n);
        for (int i = 0; i < n; i++) { a[3 * i] = pts[i].x - c.x; a[3 * i + 1] = pts[i].y - c.y; a[3 * i + 2] = pts[i].z - c.z; }
        float r[3];
        nalgebra_svd_last_row(a, n, r);
        free(a);
        *out = vunit(v3(r[0], r[1], r[2]));
        return 0;
    }
    double a[3][3] = {{0}};
    for (int i = 0; i < n; i++) {
        float d[3] = {pts[i].x - c.x, pts[i].y - c.y, pts[i].z - c.z};
        for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) a[r][q] += (double)d[r] * (double)d[q];
    }
    double v[3][3], w[3];
    jacobi_eig3(a, v, w);
    int k = 0;
    if (w[1] < w[k]) k = 1;
    if (w[2] < w[k]) k = 2;
    vec3 nrm = v3((float)v[0][k], (float)v[1][k], (float)v[2][k]);
    *out = vunit(nrm);
    return 0;
}

int gorder_oracle_normal_from_cloud(const float *pts, int n, float *out) {
    vec3 *p = (vec3 *)malloc(sizeof(vec3) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) p[i] = v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    vec3 o = v3(0, 0, 0);
    int rc = normal_from_cloud(p, n, &o);
    free(p);
    out[0] = o.x; out[1] = o.y; out[2] = o.z;
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Engine state
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n_mol, n_bonds, n_ua, head_rel, n_methyls, normal_head_rel;
    int32_t *mol_base, *bond_rel, *ua_kind, *ua_rel, *methyl_rel;
    int n_manual_leaflet_frames, n_manual_normal_frames;
    uint8_t *manual_leaflets;
    float *manual_normals;
    int slot0;     /* first order slot of this type */
    int n_orders;  /* order slots of this type */
    int mol0;      /* offset of this type in per-molecule tables */
    int32_t *ua_slot; /* [n_ua] first slot of each carbon type (relative to slot0) */
} OMolType;

typedef struct GorderOracle {
    GorderSetup s;
    OMolType *mt;
    int32_t *normal_heads, *membrane, *geom_ref;
    int n_slots, n_mol_total;
    int map_nx, map_ny;
    int64_t n_bins;
    /* accumulators */
    int64_t *sum; uint64_t *cnt;            /* [n_slots][3] */
    int64_t *map_sum; uint64_t *map_cnt;    /* [n_slots][3][n_bins] */
    /* timewise rows, in submit order */
    int64_t n_frames, cap_frames;
    int64_t *tw_sum; uint64_t *tw_cnt; int64_t *tw_frame;
    /* leaflets */
    uint8_t *cur_leaflets; int64_t cur_leaflet_frame; int have_leaflets;
    int64_t n_leaf_frames, cap_leaf_frames; uint8_t *leaf_rows; int64_t *leaf_frame;
    /* normals storage */
    float *normals; /* [n_frames][n_mol_total][3] when collect_normals */
    /* error */
    int err; int64_t err_detail;
} GorderOracle;

static void *dup_mem(const void *p, size_t n) {
    if (!p || !n) return NULL;
    void *q = malloc(n);
    memcpy(q, p, n);
    return q;
}

static int ua_hydrogens(int kind) { return kind == GORDER_UA_CH3 ? 3 : (kind == GORDER_UA_CH2 ? 2 : 1); }

/* GridMap::new node count per axis: round(span / bin) + 1 (node-centred). */
static int grid_nodes(float lo, float hi, float bin) { return (int)roundf((hi - lo) / bin) + 1; }

int gorder_oracle_create(const GorderSetup *s, GorderOracle **out) {
    if (!s || !out || s->abi_version != GORDER_ABI_VERSION) return GORDER_ERR_INVALID_ARGUMENT;
    GorderOracle *o = (GorderOracle *)calloc(1, sizeof(GorderOracle));
    o->s = *s;
    o->mt = (OMolType *)calloc(s->n_moltypes > 0 ? s->n_moltypes : 1, sizeof(OMolType));
    int slot = 0, mol = 0;
    for (int t = 0; t < s->n_moltypes; t++) {
        const GorderMolType *m = &s->moltypes[t];
        OMolType *q = &o->mt[t];
        q->n_mol = m->n_molecules; q->n_bonds = m->n_bond_types; q->n_ua = m->n_ua_atoms;
        q->head_rel = m->head_rel; q->n_methyls = m->n_methyls; q->normal_head_rel = m->normal_head_rel;
        q->mol_base = (int32_t *)dup_mem(m->mol_base, sizeof(int32_t) * m->n_molecules);
        q->bond_rel = (int32_t *)dup_mem(m->bond_rel, sizeof(int32_t) * 2 * m->n_bond_types);
        q->ua_kind = (int32_t *)dup_mem(m->ua_kind, sizeof(int32_t) * m->n_ua_atoms);
        q->ua_rel = (int32_t *)dup_mem(m->ua_rel, sizeof(int32_t) * 4 * m->n_ua_atoms);
        q->methyl_rel = (int32_t *)dup_mem(m->methyl_rel, sizeof(int32_t) * m->n_methyls);
        q->n_manual_leaflet_frames = m->n_manual_leaflet_frames;
        q->manual_leaflets = (uint8_t *)dup_mem(m->manual_leaflets, (size_t)m->n_manual_leaflet_frames * m->n_molecules);
        q->n_manual_normal_frames = m->n_manual_normal_frames;
        q->manual_normals = (float *)dup_mem(m->manual_normals, sizeof(float) * 3 * (size_t)m->n_manual_normal_frames * m->n_molecules);
        q->slot0 = slot; q->mol0 = mol;
        if (s->kind == GORDER_KIND_UA) {
            q->ua_slot = (int32_t *)calloc(m->n_ua_atoms > 0 ? m->n_ua_atoms : 1, sizeof(int32_t));
            int k = 0;
            for (int i = 0; i < m->n_ua_atoms; i++) { q->ua_slot[i] = k; k += ua_hydrogens(m->ua_kind[i]); }
            q->n_orders = k;
        } else q->n_orders = m->n_bond_types;
        slot += q->n_orders; mol += m->n_molecules;
    }
    o->n_slots = slot; o->n_mol_total = mol;
    o->normal_heads = (int32_t *)dup_mem(s->normal_heads, sizeof(int32_t) * s->n_normal_heads);
    o->membrane = (int32_t *)dup_mem(s->membrane, sizeof(int32_t) * s->n_membrane);
    o->geom_ref = (int32_t *)dup_mem(s->geom_ref, sizeof(int32_t) * s->n_geom_ref);
    o->s.moltypes = NULL; o->s.normal_heads = NULL; o->s.membrane = NULL; o->s.geom_ref = NULL;
    size_t na = (size_t)(o->n_slots > 0 ? o->n_slots : 1) * 3;
    o->sum = (int64_t *)calloc(na, sizeof(int64_t));
    o->cnt = (uint64_t *)calloc(na, sizeof(uint64_t));
    if (s->map_enabled) {
        /* Map::new (ordermap.rs:40-96): GridMapError::InvalidGridTile -> BinTooLarge */
        if (s->map_bin[0] > s->map_span_x[1] - s->map_span_x[0] || s->map_bin[1] > s->map_span_y[1] - s->map_span_y[0]) {
            free(o); return GORDER_ERR_ORDERMAP_BIN_TOO_LARGE;
        }
        o->map_nx = grid_nodes(s->map_span_x[0], s->map_span_x[1], s->map_bin[0]);
        o->map_ny = grid_nodes(s->map_span_y[0], s->map_span_y[1], s->map_bin[1]);
        o->n_bins = (int64_t)o->map_nx * o->map_ny;
        o->map_sum = (int64_t *)calloc(na * o->n_bins, sizeof(int64_t));
        o->map_cnt = (uint64_t *)calloc(na * o->n_bins, sizeof(uint64_t));
    }
    o->cur_leaflets = (uint8_t *)calloc(o->n_mol_total > 0 ? o->n_mol_total : 1, 1);
    o->cur_leaflet_frame = -1;
    *out = o;
    return GORDER_OK;
}

void gorder_oracle_destroy(GorderOracle *o) {
    if (!o) return;
    for (int t = 0; t < o->s.n_moltypes; t++) {
        OMolType *q = &o->mt[t];
        free(q->mol_base); free(q->bond_rel); free(q->ua_kind); free(q->ua_rel); free(q->methyl_rel);
        free(q->manual_leaflets); free(q->manual_normals); free(q->ua_slot);
    }
    free(o->mt); free(o->normal_heads); free(o->membrane); free(o->geom_ref);
    free(o->sum); free(o->cnt); free(o->map_sum); free(o->map_cnt);
    free(o->tw_sum); free(o->tw_cnt); free(o->tw_frame);
    free(o->cur_leaflets); free(o->leaf_rows); free(o->leaf_frame); free(o->normals);
    free(o);
}

/* should_assign (leaflets.rs:435-441) */
static int should_assign(const GorderSetup *s, int64_t frame) {
    if (s->leaflet_mode == GORDER_LEAFLET_NONE) return 0;
    if (s->leaflet_freq_kind == GORDER_FREQ_ONCE) return frame == 0;
    return (frame % (s->leaflet_freq > 0 ? s->leaflet_freq : 1)) == 0;
}
/* frame whose assignment a frame uses (get_assigned_leaflet, leaflets.rs:1438-1473) */
static int64_t assignment_frame(const GorderSetup *s, int64_t frame) {
    if (s->leaflet_freq_kind == GORDER_FREQ_ONCE) return 0;
    int64_t n = s->leaflet_freq > 0 ? s->leaflet_freq : 1;
    return (frame / n) * n;
}

static inline vec3 atom_pos(const float *xyz, int idx) { const float *p = xyz + 3 * (size_t)idx; return v3(p[0], p[1], p[2]); }
static inline int pos_undefined(vec3 p) { return isnan(p.x) || isnan(p.y) || isnan(p.z); }

typedef struct { int err; int64_t detail; } FrameErr;
#define FAIL(e, code, d) do { if (!(e)->err) { (e)->err = (code); (e)->detail = (d); } } while (0)

/* PBC3D::calc_local_membrane_centers (pbc.rs:273-318) for one head: centre of the membrane atoms
 * inside an infinite cylinder of `radius` around the head, oriented along the leaflet axis. */
static vec3 local_membrane_center(const GorderOracle *o, const float *xyz, const float *box, vec3 head, int *n_in, int32_t *scratch) {
    const GorderSetup *s = &o->s;
    int ax = s->leaflet_axis, pbc = s->handle_pbc, n = 0;
    for (int i = 0; i < s->n_membrane; i++) {
        vec3 p = atom_pos(xyz, o->membrane[i]);
        float r2 = 0.0f;
        for (int a = 0; a < 3; a++) {
            if (a == ax) continue;
            float d = get_c(p, a) - get_c(head, a);
            if (pbc) d = min_image(d, box[a]);
            r2 += d * d;
        }
        if (sqrtf(r2) < s->leaflet_radius) scratch[n++] = o->membrane[i];
    }
    *n_in = n;
    return group_center(xyz, scratch, n, box, pbc);
}

/* ---- spherical clustering: two-component 1-D Gaussian mixture of the head - vesicle-centre distances ----
 * (spherical_clustering.rs:22-272; sums are the reference's sequential f32 folds, ln / exp are libm's) */
#define GMM_MAX_ITERATIONS 50
#define GMM_TOLERANCE 1e-4f

static inline float gmm_log_gaussian(float x, float mean, float variance) {   /* :103-107 */
    float diff = x - mean;
    return -0.5f * (logf(2.0f * 3.14159274101257324f) + logf(variance) + diff * diff / variance);
}
static inline float gmm_log_sum_exp(float a, float b) {                       /* :110-114 */
    float m = a > b ? a : b;
    return m + logf(expf(a - m) + expf(b - m));
}
static int cmp_float(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }

/* fit_gmm_1d_two_components (:138-237) with initialize_params (:116-136).  params: weight_a, mean_a, var_a, mean_b, var_b. */
static float gmm_fit(const float *data, int n, float *resp_a, float *params) {
    float n_f = (float)n;
    const float var_floor = 1e-6f, weight_floor = 1e-4f;
    float *sorted = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
    memcpy(sorted, data, sizeof(float) * n);
    qsort(sorted, n, sizeof(float), cmp_float);
    float mean_a = sorted[n / 4], mean_b = sorted[(3 * n) / 4];
    free(sorted);
    float gsum = 0.0f;
    for (int i = 0; i < n; i++) gsum += data[i];
    float gmean = gsum / n_f, gdev = 0.0f;
    for (int i = 0; i < n; i++) { float d = data[i] - gmean; gdev += d * d; }
    float gvar = gdev / (n_f - 1.0f);
    if (!isfinite(gvar) || gvar <= 0.0f) gvar = 1.0f;
    float weight_a = 0.5f, var_a = gvar > var_floor ? gvar : var_floor, var_b = var_a;
    for (int i = 0; i < n; i++) resp_a[i] = 0.5f;
    float prev_avg_ll = -INFINITY;
    for (int it = 0; it < GMM_MAX_ITERATIONS; it++) {
        float loglik_sum = 0.0f;
        float log_weight_a = logf(weight_a), log_weight_b = logf(1.0f - weight_a);
        for (int i = 0; i < n; i++) {
            float x = data[i];
            float ja = log_weight_a + gmm_log_gaussian(x, mean_a, var_a), jb = log_weight_b + gmm_log_gaussian(x, mean_b, var_b);
            float log_px = gmm_log_sum_exp(ja, jb);
            loglik_sum += log_px;
            resp_a[i] = expf(ja - log_px);
        }
        float avg_ll = loglik_sum / n_f;
        if (fabsf(avg_ll - prev_avg_ll) < GMM_TOLERANCE) { prev_avg_ll = avg_ll; break; }
        prev_avg_ll = avg_ll;
        float sum_a = 0.0f;
        for (int i = 0; i < n; i++) sum_a += resp_a[i];
        float sum_b = n_f - sum_a;
        sum_a = sum_a > 1e-6f ? sum_a : 1e-6f;
        sum_b = sum_b > 1e-6f ? sum_b : 1e-6f;
        weight_a = sum_a / n_f;
        if (weight_a < weight_floor) weight_a = weight_floor;
        if (weight_a > 1.0f - weight_floor) weight_a = 1.0f - weight_floor;
        float ma = 0.0f, mb = 0.0f;
        for (int i = 0; i < n; i++) { ma += resp_a[i] * data[i]; mb += (1.0f - resp_a[i]) * data[i]; }
        mean_a = ma / sum_a; mean_b = mb / sum_b;
        float va = 0.0f, vb = 0.0f;
        for (int i = 0; i < n; i++) {
            float da = data[i] - mean_a, db = data[i] - mean_b;
            va += resp_a[i] * da * da; vb += (1.0f - resp_a[i]) * db * db;
        }
        var_a = va / sum_a; var_b = vb / sum_b;
        if (var_a < var_floor) var_a = var_floor;
        if (var_b < var_floor) var_b = var_floor;
    }
    if (params) { params[0] = weight_a; params[1] = mean_a; params[2] = var_a; params[3] = mean_b; params[4] = var_b; }
    return prev_avg_ll;
}

/* Clusters::from_responsibilities (:239-272): r < 0.5 -> cluster 1; the cluster farther from the centre is the upper (outer) one. */
static void gmm_clusters(const float *resp_a, const float *dist, int n, uint8_t *upper) {
    float s1 = 0.0f, s2 = 0.0f;
    int n1 = 0, n2 = 0;
    for (int i = 0; i < n; i++) {
        if (resp_a[i] < 0.5f) { n1++; s1 += dist[i]; } else { n2++; s2 += dist[i]; }
    }
    int first_is_upper = (s1 / (float)n1) > (s2 / (float)n2);
    for (int i = 0; i < n; i++) upper[i] = ((resp_a[i] < 0.5f) == first_is_upper) ? 1 : 0;
}

float gorder_oracle_gmm_fit(const float *data, int n, float *resp_a, float *params) { return gmm_fit(data, n, resp_a, params); }
void gorder_oracle_gmm_clusters(const float *resp_a, const float *dist, int n, uint8_t *upper) { gmm_clusters(resp_a, dist, n, upper); }

/* SystemSphericalClusterClassification::cluster (:42-76): distances of the ClusterHeads group (GorderSetup.membrane in this
 * mode) from its PBC-aware centre, mixture fit, clusters.  upper[i] = 1 for the outer leaflet. */
static void spherical_cluster(const GorderOracle *o, const float *xyz, const float *box, uint8_t *upper) {
    const GorderSetup *s = &o->s;
    int n = s->n_membrane, pbc = s->handle_pbc;
    vec3 center = group_center(xyz, o->membrane, n, box, pbc);
    float *dist = (float *)malloc(sizeof(float) * (n > 0 ? n : 1)), *resp = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) dist[i] = vnorm(vector_to(center, atom_pos(xyz, o->membrane[i]), box, pbc));   /* atom - centre */
    gmm_fit(dist, n, resp, NULL);
    gmm_clusters(resp, dist, n, upper);
    free(dist); free(resp);
}

/* Assign every molecule of every type (AssignedLeaflets::assign_lipids, leaflets.rs:1406-1435;
 * classifiers: Global :571-624 + common_identify_leaflet :711-732, Local :630-707,
 * Individual :736-811, Manual :816-874), then maybe_flip (:68-73).  Output GORDER_UPPER/LOWER. */
static void assign_leaflets(const GorderOracle *o, const float *xyz, const float *box, int64_t frame, uint8_t *out, FrameErr *fe) {
    const GorderSetup *s = &o->s;
    int pbc = s->handle_pbc, ax = s->leaflet_axis;
    vec3 center = v3(0, 0, 0);
    if (s->leaflet_mode == GORDER_LEAFLET_GLOBAL) { /* SystemLeafletClassification::run, leaflets.rs:171-205 */
        center = group_center(xyz, o->membrane, s->n_membrane, box, pbc);
        if (pos_undefined(center)) { FAIL(fe, GORDER_ERR_INVALID_GLOBAL_CENTER, 0); return; }
    }
    uint8_t *cl_upper = NULL;
    int32_t *cl_pos = NULL;
    if (s->leaflet_mode == GORDER_LEAFLET_SPHERICAL) {   /* leaflets.rs:201-204, 1351-1367 */
        cl_upper = (uint8_t *)malloc(s->n_membrane > 0 ? s->n_membrane : 1);
        cl_pos = (int32_t *)malloc(sizeof(int32_t) * (s->n_atoms > 0 ? s->n_atoms : 1));
        for (int i = 0; i < s->n_atoms; i++) cl_pos[i] = -1;
        for (int i = 0; i < s->n_membrane; i++) cl_pos[o->membrane[i]] = i;
        spherical_cluster(o, xyz, box, cl_upper);
    }
    int32_t *scratch = NULL;
    if (s->leaflet_mode == GORDER_LEAFLET_LOCAL) scratch = (int32_t *)malloc(sizeof(int32_t) * (s->n_membrane > 0 ? s->n_membrane : 1));
    for (int t = 0; t < s->n_moltypes; t++) {
        const OMolType *q = &o->mt[t];
        for (int m = 0; m < q->n_mol; m++) {
            int upper = 1;
            switch (s->leaflet_mode) {
            case GORDER_LEAFLET_GLOBAL: {
                vec3 head = atom_pos(xyz, q->mol_base[m] + q->head_rel);
                if (pos_undefined(head)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, q->mol_base[m] + q->head_rel); break; }
                upper = distance_1d(head, center, ax, box, pbc) >= 0.0f;
                break;
            }
            case GORDER_LEAFLET_LOCAL: {
                int hi = q->mol_base[m] + q->head_rel, n_in = 0;
                vec3 head = atom_pos(xyz, hi);
                if (pos_undefined(head)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, hi); break; }
                vec3 c = local_membrane_center(o, xyz, box, head, &n_in, scratch);
                if (pos_undefined(c)) { FAIL(fe, GORDER_ERR_INVALID_LOCAL_CENTER, hi); break; }
                upper = distance_1d(head, c, ax, box, pbc) >= 0.0f;
                break;
            }
            case GORDER_LEAFLET_INDIVIDUAL: {
                vec3 head = atom_pos(xyz, q->mol_base[m] + q->head_rel);
                if (pos_undefined(head)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, q->mol_base[m] + q->head_rel); break; }
                float total = 0.0f;
                for (int k = 0; k < q->n_methyls; k++) {
                    vec3 me = atom_pos(xyz, q->mol_base[m] + q->methyl_rel[k]);
                    if (pos_undefined(me)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, q->mol_base[m] + q->methyl_rel[k]); break; }
                    total += distance_1d(head, me, ax, box, pbc);
                }
                upper = total >= 0.0f;
                break;
            }
            case GORDER_LEAFLET_SPHERICAL: {
                int at = cl_pos[q->mol_base[m] + q->head_rel];
                if (at < 0) { FAIL(fe, GORDER_ERR_INVALID_ARGUMENT, q->mol_base[m] + q->head_rel); break; }   /* the reference panics */
                upper = cl_upper[at];
                break;
            }
            case GORDER_LEAFLET_MANUAL: {
                int64_t row = s->leaflet_freq_kind == GORDER_FREQ_ONCE ? 0 : frame / (s->leaflet_freq > 0 ? s->leaflet_freq : 1);
                if (row >= q->n_manual_leaflet_frames) { FAIL(fe, GORDER_ERR_MANUAL_LEAFLET_FRAME, frame); break; }
                upper = q->manual_leaflets[(size_t)row * q->n_mol + m] == GORDER_UPPER;
                break;
            }
            }
            if (s->leaflet_flip) upper ^= 1;
            out[q->mol0 + m] = upper ? GORDER_UPPER : GORDER_LOWER;
        }
    }
    free(scratch); free(cl_upper); free(cl_pos);
}

/* Cell grid over the NormalHeads group (groan CellGrid::new(system, "NormalHeads", radius), pbc.rs:327-333):
 * cell edge >= radius, the 3x3x3 block around the reference is scanned and filtered by the exact
 * sphere test, so the result equals the brute-force scan except for the order in which the f32
 * centroid is summed.  Used only for large head groups (the brute force is O(n^2) per frame). */
#define ORACLE_CELL_MIN_HEADS 4096
typedef struct { int n[3]; int *start; int *sorted; } CellGridO;

static void cellgrid_build(const GorderOracle *o, const float *xyz, const float *box, CellGridO *g) {
    const GorderSetup *s = &o->s;
    int nh = s->n_normal_heads;
    for (int k = 0; k < 3; k++) { int q = (int)floorf(box[k] / s->dynamic_radius); g->n[k] = q < 1 ? 1 : (q > 64 ? 64 : q); }
    int nc = g->n[0] * g->n[1] * g->n[2];
    g->start = (int *)calloc(nc + 1, sizeof(int));
    g->sorted = (int *)malloc(sizeof(int) * (nh > 0 ? nh : 1));
    int *cell = (int *)malloc(sizeof(int) * (nh > 0 ? nh : 1));
    for (int i = 0; i < nh; i++) {
        vec3 p = atom_pos(xyz, o->normal_heads[i]);
        int c[3];
        float q[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; k++) {
            float w = wrap1(q[k], box[k]);
            int ci = (int)(w * (float)g->n[k] / box[k]);
            c[k] = ci < 0 ? 0 : (ci >= g->n[k] ? g->n[k] - 1 : ci);
            if (isnan(q[k])) c[k] = 0;
        }
        cell[i] = (c[0] * g->n[1] + c[1]) * g->n[2] + c[2];
        g->start[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; c++) g->start[c + 1] += g->start[c];
    int *cur = (int *)malloc(sizeof(int) * (nc > 0 ? nc : 1));
    memcpy(cur, g->start, sizeof(int) * nc);
    for (int i = 0; i < nh; i++) g->sorted[cur[cell[i]]++] = i;
    free(cur); free(cell);
}
static void cellgrid_free(CellGridO *g) { free(g->start); free(g->sorted); g->start = NULL; g->sorted = NULL; }

/* PBC3D::get_heads_cloud (pbc.rs:321-351) / NoPBC (:142-161) + membrane_normal_from_cloud. */
static int dynamic_normal(const GorderOracle *o, const float *xyz, const float *box, int head_idx, vec3 *out, FrameErr *fe, vec3 *cloud,
                          const CellGridO *g) {
    const GorderSetup *s = &o->s;
    vec3 ref = atom_pos(xyz, head_idx);
    if (pos_undefined(ref)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, head_idx); return 1; }
    int n = 0;
    if (g && g->start) {
        int c0[3];
        float q[3] = {ref.x, ref.y, ref.z};
        for (int k = 0; k < 3; k++) {
            int ci = (int)(wrap1(q[k], box[k]) * (float)g->n[k] / box[k]);
            c0[k] = ci < 0 ? 0 : (ci >= g->n[k] ? g->n[k] - 1 : ci);
        }
        int lo[3], hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = g->n[k] >= 3 ? -1 : 0; hi[k] = g->n[k] >= 3 ? 1 : g->n[k] - 1; }
        for (int dx = lo[0]; dx <= hi[0]; dx++)
            for (int dy = lo[1]; dy <= hi[1]; dy++)
                for (int dz = lo[2]; dz <= hi[2]; dz++) {
                    int cx = g->n[0] >= 3 ? (c0[0] + dx + g->n[0]) % g->n[0] : dx;
                    int cy = g->n[1] >= 3 ? (c0[1] + dy + g->n[1]) % g->n[1] : dy;
                    int cz = g->n[2] >= 3 ? (c0[2] + dz + g->n[2]) % g->n[2] : dz;
                    int c = (cx * g->n[1] + cy) * g->n[2] + cz;
                    for (int k = g->start[c]; k < g->start[c + 1]; k++) {
                        int idx = o->normal_heads[g->sorted[k]];
                        vec3 p = atom_pos(xyz, idx);
                        vec3 d = vector_to(ref, p, box, 1);
                        if (vnorm(d) < s->dynamic_radius) cloud[n++] = vadd(ref, d);
                    }
                }
    } else {
        for (int i = 0; i < s->n_normal_heads; i++) {
            vec3 p = atom_pos(xyz, o->normal_heads[i]);
            vec3 d = vector_to(ref, p, box, s->handle_pbc);
            if (vnorm(d) < s->dynamic_radius) {
                if (pos_undefined(p)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, o->normal_heads[i]); return 1; }
                cloud[n++] = s->handle_pbc ? vadd(ref, d) : p;
            }
        }
    }
    int rc = normal_from_cloud(cloud, n, out);
    if (rc) { FAIL(fe, rc, n); return 1; }
    return 0;
}

/* Map::add_order (ordermap.rs:100-113): project, nearest node, drop if outside. */
static inline int64_t map_bin(const GorderOracle *o, vec3 pos) {
    const GorderSetup *s = &o->s;
    float x, y;
    switch (s->map_plane) { /* Plane::projection2plane, input/ordermap.rs:44-50 */
    case GORDER_PLANE_XY: x = pos.x; y = pos.y; break;
    case GORDER_PLANE_XZ: x = pos.x; y = pos.z; break;
    default: x = pos.z; y = pos.y; break;
    }
    /* groan GridMap::get_mut_at: nearest node = round((x - min) / bin), half away from zero.  Pinned by the AA map fixtures
     * (tests/files/ordermaps/, bin 0.1 nm on a 0.01 nm lattice: bond midpoints sit exactly on bin edges); floor(v + 0.5) and
     * round-half-even do not reproduce them. */
    float fx = roundf((x - s->map_span_x[0]) / s->map_bin[0]);
    float fy = roundf((y - s->map_span_y[0]) / s->map_bin[1]);
    if (!(fx >= 0.0f) || !(fy >= 0.0f) || fx >= (float)o->map_nx || fy >= (float)o->map_ny) return -1;
    return (int64_t)fx * o->map_ny + (int64_t)fy; /* x-major (x slow, y fast) */
}

/* BondLike::add_order (bond.rs:184-215). fsum/fcnt: this frame's [n_slots][3] accumulators. */
static inline void add_order(GorderOracle *o, int slot, float sch, vec3 pos, int leaflet /* -1 none */, int64_t *fsum, uint64_t *fcnt) {
    int64_t q = order_value(sch);
    fsum[slot * 3 + GORDER_TOTAL] += q; fcnt[slot * 3 + GORDER_TOTAL] += 1;
    int which = leaflet < 0 ? -1 : (leaflet == GORDER_UPPER ? GORDER_ACC_UPPER : GORDER_ACC_LOWER);
    if (which >= 0) { fsum[slot * 3 + which] += q; fcnt[slot * 3 + which] += 1; }
    if (o->s.map_enabled) {
        int64_t b = map_bin(o, pos);
        if (b >= 0) {
            size_t i0 = ((size_t)slot * 3 + GORDER_TOTAL) * o->n_bins + b;
#pragma omp atomic
            o->map_sum[i0] += q;
#pragma omp atomic
            o->map_cnt[i0] += 1;
            if (which >= 0) {
                size_t i1 = ((size_t)slot * 3 + which) * o->n_bins + b;
#pragma omp atomic
                o->map_sum[i1] += q;
#pragma omp atomic
                o->map_cnt[i1] += 1;
            }
        }
    }
}

/* analyze_frame (common.rs:201-235) -> MoleculeTypes::analyze_frame (molecule.rs:54-95) for one
 * frame whose leaflet table is already known.  normals_out: [n_mol_total][3] or NULL. */
static void analyze_one_frame(GorderOracle *o, const float *xyz, const float *box, int64_t frame, const uint8_t *leaflets,
                              int64_t *fsum, uint64_t *fcnt, float *normals_out, FrameErr *fe) {
    const GorderSetup *s = &o->s;
    int pbc = s->handle_pbc;
    if (pbc && box[0] == 0.0f && box[1] == 0.0f && box[2] == 0.0f) { FAIL(fe, GORDER_ERR_ZERO_BOX, frame); return; }

    /* geometry.init_new_frame -> init_reference (geometry.rs:192-210) */
    Shape shape;
    memset(&shape, 0, sizeof(shape));
    if (s->geom_kind != GORDER_GEOM_NONE) {
        vec3 ref = v3(s->geom_ref_point[0], s->geom_ref_point[1], s->geom_ref_point[2]);
        if (s->geom_ref_kind == GORDER_GEOMREF_SELECTION) ref = group_center(xyz, o->geom_ref, s->n_geom_ref, box, pbc);
        else if (s->geom_ref_kind == GORDER_GEOMREF_BOX_CENTER) ref = v3(box[0] / 2.0f, box[1] / 2.0f, box[2] / 2.0f);
        shape = construct_shape(s, ref, box);
    }
    vec3 static_normal = v3(s->normal_axis == 0, s->normal_axis == 1, s->normal_axis == 2);
    vec3 *cloud = NULL;
    CellGridO grid = {{0, 0, 0}, NULL, NULL};
    if (s->normal_mode == GORDER_NORMAL_DYNAMIC) {
        cloud = (vec3 *)malloc(sizeof(vec3) * (s->n_normal_heads > 0 ? s->n_normal_heads : 1));
        if (pbc && s->n_normal_heads >= ORACLE_CELL_MIN_HEADS) cellgrid_build(o, xyz, box, &grid);
    }
    if (normals_out) for (int i = 0; i < o->n_mol_total * 3; i++) normals_out[i] = nanf("");

    for (int t = 0; t < s->n_moltypes && !fe->err; t++) {
        const OMolType *q = &o->mt[t];
        /* per-frame OnceCell cache of dynamic normals (normal.rs:160-175) */
        vec3 *ncache = NULL; uint8_t *nhave = NULL;
        if (s->normal_mode == GORDER_NORMAL_DYNAMIC) {
            ncache = (vec3 *)malloc(sizeof(vec3) * (q->n_mol > 0 ? q->n_mol : 1));
            nhave = (uint8_t *)calloc(q->n_mol > 0 ? q->n_mol : 1, 1);
        }
#define GET_NORMAL(m, nvar)                                                                                         \
    do {                                                                                                            \
        if (s->normal_mode == GORDER_NORMAL_STATIC) nvar = static_normal;                                           \
        else if (s->normal_mode == GORDER_NORMAL_DYNAMIC) {                                                         \
            if (!nhave[m]) {                                                                                        \
                if (dynamic_normal(o, xyz, box, q->mol_base[m] + q->normal_head_rel, &ncache[m], fe, cloud, &grid)) break; \
                nhave[m] = 1;                                                                                       \
            }                                                                                                       \
            nvar = ncache[m];                                                                                       \
        } else { /* ManualMembraneNormal::get_normal, normal.rs:266-298 */                                          \
            int64_t row = frame / (s->step > 0 ? s->step : 1);                                                      \
            if (row >= q->n_manual_normal_frames) { FAIL(fe, GORDER_ERR_MANUAL_NORMAL_FRAME, frame); break; }       \
            const float *mn = q->manual_normals + 3 * ((size_t)row * q->n_mol + m);                                 \
            nvar = v3(mn[0], mn[1], mn[2]);                                                                         \
        }                                                                                                           \
    } while (0)

        if (s->kind != GORDER_KIND_UA) {
            /* OrderBonds::analyze_frame (bond.rs:103-117) -> BondType::analyze_frame (bond.rs:396-446) */
            for (int b = 0; b < q->n_bonds && !fe->err; b++) {
                int r1 = q->bond_rel[2 * b], r2 = q->bond_rel[2 * b + 1];
                for (int m = 0; m < q->n_mol; m++) {
                    int i1 = q->mol_base[m] + r1, i2 = q->mol_base[m] + r2;
                    vec3 p1 = atom_pos(xyz, i1), p2 = atom_pos(xyz, i2);
                    if (pos_undefined(p1)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, i1); break; }
                    if (pos_undefined(p2)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, i2); break; }
                    vec3 vec = vector_to(p1, p2, box, pbc);
                    vec3 bond_pos = vadd(p1, vdivs(vec, 2.0f));
                    if (!shape_inside(&shape, bond_pos, box, pbc)) continue;
                    vec3 normal = static_normal;
                    GET_NORMAL(m, normal);
                    if (fe->err) break;
                    float sch = calc_sch(vec, normal);
                    add_order(o, q->slot0 + b, sch, bond_pos, leaflets ? leaflets[q->mol0 + m] : -1, fsum, fcnt);
                }
            }
        } else {
            /* UAOrderAtoms::analyze_frame (uatom.rs:73-88) -> UAAtom::analyze_frame (uaorder.rs:400-437) */
            for (int a = 0; a < q->n_ua && !fe->err; a++) {
                const int32_t *rel = q->ua_rel + 4 * a;
                int kind = q->ua_kind[a];
                for (int m = 0; m < q->n_mol; m++) {
                    vec3 normal = static_normal;
                    GET_NORMAL(m, normal); /* the normal is requested BEFORE the geometry test (uaorder.rs:412-413) */
                    if (fe->err) break;
                    int base = q->mol_base[m];
                    vec3 tp = atom_pos(xyz, base + rel[0]), h1 = atom_pos(xyz, base + rel[1]), h2 = atom_pos(xyz, base + rel[2]);
                    vec3 h3 = rel[3] >= 0 ? atom_pos(xyz, base + rel[3]) : v3(0, 0, 0);
                    if (pos_undefined(h1)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, base + rel[1]); break; }
                    if (pos_undefined(tp)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, base + rel[0]); break; }
                    if (pos_undefined(h2)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, base + rel[2]); break; }
                    if (rel[3] >= 0 && pos_undefined(h3)) { FAIL(fe, GORDER_ERR_UNDEFINED_POSITION, base + rel[3]); break; }
                    vec3 hyd[3];
                    int nh = predict_hydrogens(kind, tp, h1, h2, h3, box, pbc, hyd);
                    for (int i = 0; i < nh; i++) { /* calculate_sch, uaorder.rs:375-397 */
                        vec3 vec = vector_to(tp, hyd[i], box, pbc);
                        vec3 bond_pos = vadd(hyd[i], vdivs(vec, 2.0f)); /* sic: hydrogen + vec/2 (uaorder.rs:385) */
                        if (!shape_inside(&shape, bond_pos, box, pbc)) continue;
                        float sch = calc_sch(vec, normal);
                        add_order(o, q->slot0 + q->ua_slot[a] + i, sch, bond_pos, leaflets ? leaflets[q->mol0 + m] : -1, fsum, fcnt);
                    }
                }
            }
        }
        /* store_normals (normal.rs:211-227) */
        if (normals_out && ncache)
            for (int m = 0; m < q->n_mol; m++)
                if (nhave[m]) {
                    float *d = normals_out + 3 * (size_t)(q->mol0 + m);
                    d[0] = ncache[m].x; d[1] = ncache[m].y; d[2] = ncache[m].z;
                }
        free(ncache); free(nhave);
#undef GET_NORMAL
    }
    free(cloud);
    cellgrid_free(&grid);
}

/* Analyse a batch of frames; n_threads > 1 mimics the reference's interleaved frame ownership
 * (thread t owns frames t, t+N, ...; topology/mod.rs:139-144,274-277) with private accumulators
 * merged by integer addition (topology/mod.rs:236-272). */
int gorder_oracle_analyze(GorderOracle *o, const float *xyz, const float *box, const int64_t *frame_index, int n_frames, int n_threads) {
    const GorderSetup *s = &o->s;
    if (o->err) return o->err;
    if (n_frames <= 0) return GORDER_OK;
    size_t fstride = (size_t)s->n_atoms * 3;
    size_t na = (size_t)o->n_slots * 3;
    int use_leaflets = s->leaflet_mode != GORDER_LEAFLET_NONE;

    /* grow timewise / normals storage */
    if (o->n_frames + n_frames > o->cap_frames) {
        int64_t cap = (o->n_frames + n_frames) * 2;
        o->tw_frame = (int64_t *)realloc(o->tw_frame, sizeof(int64_t) * cap);
        if (s->timewise) {
            o->tw_sum = (int64_t *)realloc(o->tw_sum, sizeof(int64_t) * cap * (na ? na : 1));
            o->tw_cnt = (uint64_t *)realloc(o->tw_cnt, sizeof(uint64_t) * cap * (na ? na : 1));
        }
        if (s->collect_normals && s->normal_mode == GORDER_NORMAL_DYNAMIC)
            o->normals = (float *)realloc(o->normals, sizeof(float) * 3 * (size_t)cap * (o->n_mol_total ? o->n_mol_total : 1));
        o->cap_frames = cap;
    }

    /* phase A: assignment frames of this batch */
    int n_assign = 0;
    int *assign_row = (int *)malloc(sizeof(int) * n_frames); /* row used by frame f: -1 carry-over */
    uint8_t *rows = NULL;
    FrameErr *ferr = (FrameErr *)calloc(n_frames, sizeof(FrameErr));
    if (use_leaflets) {
        int *assign_of = (int *)malloc(sizeof(int) * n_frames);
        for (int f = 0; f < n_frames; f++) assign_of[f] = should_assign(s, frame_index[f]) ? n_assign++ : -1;
        rows = (uint8_t *)malloc((size_t)(n_assign > 0 ? n_assign : 1) * (o->n_mol_total ? o->n_mol_total : 1));
#pragma omp parallel for schedule(static, 1) num_threads(n_threads > 0 ? n_threads : 1)
        for (int f = 0; f < n_frames; f++)
            if (assign_of[f] >= 0)
                assign_leaflets(o, xyz + f * fstride, box + 3 * f, frame_index[f], rows + (size_t)assign_of[f] * o->n_mol_total, &ferr[f]);
        int last = -1;
        for (int f = 0; f < n_frames; f++) {
            if (assign_of[f] >= 0) last = assign_of[f];
            assign_row[f] = last;
            if (last < 0) { /* needs the carried-over table */
                if (!o->have_leaflets || o->cur_leaflet_frame != assignment_frame(s, frame_index[f]))
                    FAIL(&ferr[f], GORDER_ERR_LEAFLET_FRAME_UNAVAILABLE, frame_index[f]);
            }
        }
        /* collect for export (shared storage doubles as export buffer, leaflets.rs:536-548) */
        if (s->collect_leaflets && n_assign > 0) {
            if (o->n_leaf_frames + n_assign > o->cap_leaf_frames) {
                int64_t cap = (o->n_leaf_frames + n_assign) * 2;
                o->leaf_rows = (uint8_t *)realloc(o->leaf_rows, (size_t)cap * (o->n_mol_total ? o->n_mol_total : 1));
                o->leaf_frame = (int64_t *)realloc(o->leaf_frame, sizeof(int64_t) * cap);
                o->cap_leaf_frames = cap;
            }
            for (int f = 0; f < n_frames; f++)
                if (assign_of[f] >= 0) {
                    memcpy(o->leaf_rows + (size_t)o->n_leaf_frames * o->n_mol_total, rows + (size_t)assign_of[f] * o->n_mol_total, o->n_mol_total);
                    o->leaf_frame[o->n_leaf_frames++] = frame_index[f];
                }
        }
        free(assign_of);
    }

    /* phase B: accumulate */
    int nt = n_threads > 0 ? n_threads : 1;
    int64_t *psum = (int64_t *)calloc((size_t)nt * (na ? na : 1), sizeof(int64_t));
    uint64_t *pcnt = (uint64_t *)calloc((size_t)nt * (na ? na : 1), sizeof(uint64_t));
#pragma omp parallel num_threads(nt)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        int64_t *fsum = (int64_t *)malloc(sizeof(int64_t) * (na ? na : 1));
        uint64_t *fcnt = (uint64_t *)malloc(sizeof(uint64_t) * (na ? na : 1));
#pragma omp for schedule(static, 1)
        for (int f = 0; f < n_frames; f++) {
            memset(fsum, 0, sizeof(int64_t) * na);
            memset(fcnt, 0, sizeof(uint64_t) * na);
            const uint8_t *lf = NULL;
            if (use_leaflets) lf = assign_row[f] >= 0 ? rows + (size_t)assign_row[f] * o->n_mol_total : o->cur_leaflets;
            float *nout = (s->collect_normals && s->normal_mode == GORDER_NORMAL_DYNAMIC)
                              ? o->normals + 3 * (size_t)(o->n_frames + f) * o->n_mol_total : NULL;
            if (!ferr[f].err) analyze_one_frame(o, xyz + f * fstride, box + 3 * f, frame_index[f], lf, fsum, fcnt, nout, &ferr[f]);
            if (!ferr[f].err) {
                for (size_t i = 0; i < na; i++) { psum[tid * na + i] += fsum[i]; pcnt[tid * na + i] += fcnt[i]; }
                if (s->timewise) {
                    memcpy(o->tw_sum + (size_t)(o->n_frames + f) * na, fsum, sizeof(int64_t) * na);
                    memcpy(o->tw_cnt + (size_t)(o->n_frames + f) * na, fcnt, sizeof(uint64_t) * na);
                }
            }
        }
        free(fsum); free(fcnt);
    }
    for (int f = 0; f < n_frames && !o->err; f++)
        if (ferr[f].err) { o->err = ferr[f].err; o->err_detail = ferr[f].detail; }
    if (!o->err) {
        for (int t = 0; t < nt; t++)
            for (size_t i = 0; i < na; i++) { o->sum[i] += psum[t * na + i]; o->cnt[i] += pcnt[t * na + i]; }
        for (int f = 0; f < n_frames; f++) o->tw_frame[o->n_frames + f] = frame_index[f];
        o->n_frames += n_frames;
        if (use_leaflets && n_assign > 0) {
            memcpy(o->cur_leaflets, rows + (size_t)(n_assign - 1) * o->n_mol_total, o->n_mol_total);
            for (int f = n_frames - 1; f >= 0; f--)
                if (should_assign(s, frame_index[f])) { o->cur_leaflet_frame = frame_index[f]; break; }
            o->have_leaflets = 1;
        }
    }
    free(psum); free(pcnt); free(rows); free(assign_row); free(ferr);
    return o->err;
}

int gorder_oracle_set_leaflets(GorderOracle *o, const uint8_t *table, int64_t frame_index) {
    memcpy(o->cur_leaflets, table, o->n_mol_total);
    o->cur_leaflet_frame = frame_index; o->have_leaflets = 1;
    return GORDER_OK;
}

int64_t gorder_oracle_error_detail(GorderOracle *o) { return o->err_detail; }

int gorder_oracle_result_sizes(GorderOracle *o, GorderResults *r) {
    r->n_slots = o->n_slots; r->n_frames = o->n_frames; r->n_map_bins = o->n_bins;
    r->map_nx = o->map_nx; r->map_ny = o->map_ny; r->n_leaflet_frames = o->n_leaf_frames;
    r->n_molecules_total = o->n_mol_total;
    return GORDER_OK;
}

int gorder_oracle_finish(GorderOracle *o, GorderResults *r) {
    if (o->err) return o->err;
    gorder_oracle_result_sizes(o, r);
    size_t na = (size_t)o->n_slots * 3;
    if (r->sum) memcpy(r->sum, o->sum, sizeof(int64_t) * na);
    if (r->count) memcpy(r->count, o->cnt, sizeof(uint64_t) * na);
    if (o->s.timewise) {
        if (r->tw_sum) memcpy(r->tw_sum, o->tw_sum, sizeof(int64_t) * na * o->n_frames);
        if (r->tw_count) memcpy(r->tw_count, o->tw_cnt, sizeof(uint64_t) * na * o->n_frames);
    }
    if (r->tw_frame_index) memcpy(r->tw_frame_index, o->tw_frame, sizeof(int64_t) * o->n_frames);
    if (o->s.map_enabled) {
        if (r->map_sum) memcpy(r->map_sum, o->map_sum, sizeof(int64_t) * na * o->n_bins);
        if (r->map_count) memcpy(r->map_count, o->map_cnt, sizeof(uint64_t) * na * o->n_bins);
    }
    if (r->leaflets && o->n_leaf_frames) memcpy(r->leaflets, o->leaf_rows, (size_t)o->n_leaf_frames * o->n_mol_total);
    if (r->leaflet_frame_index && o->n_leaf_frames) memcpy(r->leaflet_frame_index, o->leaf_frame, sizeof(int64_t) * o->n_leaf_frames);
    if (r->normals && o->normals) memcpy(r->normals, o->normals, sizeof(float) * 3 * (size_t)o->n_frames * o->n_mol_total);
    return GORDER_OK;
}

/* ------------------------------------------------------------------------------------------
 * Post-processing that the reference applies to the accumulators (the step after the path;
 * needed to compare against the reference's YAML fixtures).
 * ---------------------------------------------------------------------------------------- */

/* AnalysisOrder::calc_order (order.rs:97-107): integer division (order.rs:34-41), /1e6 in f64 -> f32 */
float gorder_oracle_calc_order(int64_t sum, uint64_t n, uint64_t min_samples) {
    if (n < min_samples || n == 0) return nanf("");
    return (float)((double)(sum / (int64_t)n) / 1000000.0);
}

/* TimeWiseData::estimate_error (timewise.rs:191-231) with statistical::standard_deviation(v, None)
 * (sample standard deviation, divisor n-1, f32). Returns NaN-coded: -1 => None. */
int gorder_oracle_estimate_error(const int64_t *order, const uint64_t *n_samples, int64_t n_frames, int n_blocks, float *out) {
    if (n_frames <= 0) return -1;
    int64_t block_size = n_frames / n_blocks;
    int64_t *bo = (int64_t *)calloc(n_blocks, sizeof(int64_t));
    uint64_t *bs = (uint64_t *)calloc(n_blocks, sizeof(uint64_t));
    for (int64_t i = 0; i < n_frames; i++) {
        int64_t id = block_size > 0 ? i / block_size : n_blocks; /* block_size 0: reference divides by zero (panics) */
        if (id < n_blocks) { bo[id] += order[i]; bs[id] += n_samples[i]; }
    }
    float *vals = (float *)malloc(sizeof(float) * n_blocks);
    int ok = 1;
    for (int b = 0; b < n_blocks; b++) {
        if (bs[b] > 0) vals[b] = (float)((double)(bo[b] / (int64_t)bs[b]) / 1000000.0);
        else { ok = 0; break; }
    }
    if (!ok) *out = nanf("");
    else {
        /* statistical 1.0.0: mean = sum / n; variance = sum((x-mean)^2) / (n-1); sqrt */
        float sum = 0.0f;
        for (int b = 0; b < n_blocks; b++) sum += vals[b];
        float mean = sum / (float)n_blocks, v = 0.0f;
        for (int b = 0; b < n_blocks; b++) { float d = vals[b] - mean; v += d * d; } /* statistical: sum of squared deviations */
        *out = sqrtf(v / (float)(n_blocks - 1));
    }
    free(bo); free(bs); free(vals);
    return 0;
}

/* TimeWiseData::prefix_average (timewise.rs:259-274) */
void gorder_oracle_prefix_average(const int64_t *order, const uint64_t *n_samples, int64_t n_frames, float *out) {
    int64_t s = 0; uint64_t n = 0;
    for (int64_t i = 0; i < n_frames; i++) {
        s += order[i]; n += n_samples[i];
        out[i] = n == 0 ? nanf("") : (float)((double)(s / (int64_t)n) / 1000000.0);
    }
}

const char *gorder_oracle_version(void) { return "gorder-oracle 0.1 (restates gorder v1.4.1)"; }

// gorder_kernels.cuh — sm_100a kernels of the per-frame order-parameter engine.
//
//   relayout_kernel         host AoS frames -> device planes (+ undefined-position check)
//   frame_setup_kernel      box, geometry shape of every frame      (geometry.rs:192-210, :328-514)
//   group_center_*          PBC-aware centre of a group             (pbc.rs:270; groan group_get_center)
//   leaflet_assign_kernel   Global / Individual / Manual / Local    (leaflets.rs:571-874)
//   dynamic_normal_kernel   per-lipid PCA normal                    (normal.rs:160-199, :421-458)
//   bond_order_kernel       K1: AA / CG bond engine                 (topology/bond.rs:396-446, :184-215)
//   ua_order_kernel         K2: united-atom engine                  (uaorder.rs:375-437, :947-1104)
//   fold_kernel             per-frame accumulators -> running totals (order.rs:160-188)
//
// All of them are HBM/L2-bound integer/f32 streaming kernels: no tensor cores (nothing on this path
// is a dense contraction).
#pragma once
#include <math_constants.h>

#include "gorder_engine.cuh"
#include "gorder_math.cuh"

namespace gorder {

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void raise_error(const DeviceView &v, int code, long long detail) {
    if (atomicCAS(v.err, 0, code) == 0) *v.err_detail = detail;
}

// DRAM -> L2 prefetch of a contiguous run (one instruction, one thread); 16-byte aligned address, size a multiple of 16
__device__ __forceinline__ void l2_prefetch_run(const float *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// L2 eviction-priority hint (createpolicy + ld.global.L2::cache_hint): the centre passes want the axis planes to
// survive from pass 0 to pass 1 (evict_last).
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_hint4(const float *p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}

template <int N> struct Vec;
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float *p) { v[0] = __ldg(p); }
};
template <> struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float *p) { float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y; }
};
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float *p) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};

// float offset (inside a frame) of molecule m's slot in the x plane of the atom at in-tile offset 0
__device__ __forceinline__ size_t mol_offset(const TypeDesc &td, int m) {
    return (size_t)td.plane_base + (size_t)(m / td.tile) * td.tile_stride + (m % td.tile);
}

__device__ __forceinline__ Box load_box(const FrameAux &a) {
    Box b;
    b.L[0] = a.L[0]; b.L[1] = a.L[1]; b.L[2] = a.L[2];
    b.half[0] = a.half[0]; b.half[1] = a.half[1]; b.half[2] = a.half[2];
    return b;
}

// sqrt_rn(r2) < radius, with the IEEE square root only where it decides: sqrt is monotone and radius^2 is within one rounding
// of the square, so two parts in a million away from radius^2 the comparison of the squares gives the same answer.
__device__ __forceinline__ bool radius_less(float r2, float radius) {
    const float R2 = radius * radius;
    if (r2 < R2 * 0.999998f) return true;
    if (r2 > R2 * 1.000002f) return false;
    return __fsqrt_rn(r2) < radius;
}

// Shape::inside / inside_naive XOR invert (geometry.rs:181-190; groan Rectangular / Cylinder / Sphere).
template <bool PBC>
__device__ __forceinline__ bool shape_inside(const ShapeParams &sp, const FrameAux &a, const Box &bx, const f3 &pt) {
    bool in = true;
    const float o[3] = {a.shape_origin[0], a.shape_origin[1], a.shape_origin[2]};
    const float q[3] = {pt.x, pt.y, pt.z};
    if (sp.kind == GORDER_GEOM_CUBOID) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (PBC) {
                float d = __fsub_rn(q[k], o[k]);
                if (bx.L[k] > 0.0f) d = wrap1(d, bx.L[k]);
                in = in && (d <= a.shape_len[k]);
            } else in = in && (q[k] >= o[k]) && (q[k] <= __fadd_rn(o[k], a.shape_len[k]));
        }
    } else if (sp.kind == GORDER_GEOM_CYLINDER) {
        float r2 = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (k == sp.axis) continue;
            float d = __fsub_rn(q[k], o[k]);
            if (PBC && bx.L[k] > 0.0f) d = min_image(d, bx.L[k], bx.half[k]);
            r2 = __fadd_rn(r2, __fmul_rn(d, d));
        }
        in = radius_less(r2, a.shape_radius);
        float d = __fsub_rn(q[sp.axis], o[sp.axis]);
        if (PBC) {
            // an infinite span: every wrapped distance lies below it (NaN coordinates are caught by the accumulator's check)
            if (a.shape_height < CUDART_INF_F) {
                if (bx.L[sp.axis] > 0.0f) d = wrap1(d, bx.L[sp.axis]);
                in = in && (d <= a.shape_height);
            }
        } else in = in && (d >= 0.0f) && (d <= a.shape_height);
    } else if (sp.kind == GORDER_GEOM_SPHERE) {
        f3 c = mk3(o[0], o[1], o[2]);
        f3 d = vector_to<PBC>(c, pt, bx);
        in = radius_less(dot_ref(d, d), a.shape_radius);
    }
    return in != (sp.invert != 0);
}

// Map::add_order bin lookup (ordermap.rs:100-113): nearest node, -1 if outside.
// groan GridMap: nearest node = round((x - min) / bin), half away from zero (pinned by the reference's AA map fixtures,
// where bond midpoints sit exactly on bin edges; floor(v + 0.5) does not reproduce them).
// The IEEE division and roundf are only needed next to a bin edge: t = (x - min) * (1 / bin) is within 2 ulp of the quotient,
// so unless t lies within that distance of a half-integer its nearest integer (one FRND) IS the reference's node.
__device__ __forceinline__ float map_node(float d, float bin, float inv_bin) {
    const float t = __fmul_rn(d, inv_bin), r = rintf(t);
    if (fabsf(__fsub_rn(t, r)) < 0.5f - fmaf(fabsf(t), 1e-6f, 1e-6f)) return r;
    return roundf(__fdiv_rn(d, bin));
}
__device__ __forceinline__ long long map_bin(const MapParams &mp, const f3 &pos) {
    float x, y;
    if (mp.plane == GORDER_PLANE_XY) { x = pos.x; y = pos.y; }
    else if (mp.plane == GORDER_PLANE_XZ) { x = pos.x; y = pos.z; }
    else { x = pos.z; y = pos.y; }   // sic: YZ projects to (z, y), input/ordermap.rs:48
    const float fx = map_node(__fsub_rn(x, mp.x0), mp.binx, mp.inv_binx);
    const float fy = map_node(__fsub_rn(y, mp.y0), mp.biny, mp.inv_biny);
    if (!(fx >= 0.0f) || !(fy >= 0.0f) || fx >= (float)mp.nx || fy >= (float)mp.ny) return -1;
    return (long long)((int)fx * mp.ny + (int)fy);   // n_bins < 2^31 (checked at create)
}

// ---------------------------------------------------------------------------------------------
// relayout: [F][n_atoms][3] AoS -> planes.  One thread per (frame, slot).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relayout_kernel(DeviceView v, const float *__restrict__ xyz, float *__restrict__ planes,
                                                       const int *__restrict__ slot_off, const int *__restrict__ slot_cs, int n_atoms) {
    const int f = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_atoms) return;
    const int off = slot_off[s];
    if (off < 0) return;
    const int cs = slot_cs[s];
    const float *src = xyz + ((size_t)f * n_atoms + s) * 3;
    float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
    // AnalysisError::UndefinedPosition: a NaN coordinate marks an atom without position
    if (x != x || y != y || z != z) raise_error(v, GORDER_ERR_UNDEFINED_POSITION, s);
    float *dst = planes + (size_t)f * v.frame_floats + off;
    dst[0] = x; dst[cs] = y; dst[2 * (size_t)cs] = z;
}

// ---------------------------------------------------------------------------------------------
// frame setup: box -> FrameAux.L/half, zero-box check (common.rs:186-198), geometry shape.
// One thread per frame.  `stage` 0: box; 1: shape (after the optional reference-centre reduction).
// ---------------------------------------------------------------------------------------------
template <bool PBC>
__device__ void construct_shape(const DeviceView &v, FrameAux &a) {
    const ShapeParams &sp = v.shape;
    if (sp.kind == GORDER_GEOM_NONE) return;
    Box bx = load_box(a);
    float p[3];
    if (sp.ref_kind == GORDER_GEOMREF_POINT) { p[0] = sp.ref_point[0]; p[1] = sp.ref_point[1]; p[2] = sp.ref_point[2]; }
    else if (sp.ref_kind == GORDER_GEOMREF_BOX_CENTER) { p[0] = a.L[0] / 2.0f; p[1] = a.L[1] / 2.0f; p[2] = a.L[2] / 2.0f; }  // pbc.rs:399-405
    else { p[0] = a.center[0]; p[1] = a.center[1]; p[2] = a.center[2]; }
    const float inf_pos = PBC ? 0.0f : -3.40282347e+38f;   // get_infinite_span: pbc.rs:393-396 / :188-191
    a.shape_len[0] = a.shape_len[1] = a.shape_len[2] = 0.0f;
    a.shape_radius = 0.0f; a.shape_height = 0.0f;
    if (sp.kind == GORDER_GEOM_CUBOID) {   // geometry.rs:328-357
        for (int k = 0; k < 3; k++) {
            float lo = sp.dims[2 * k], hi = sp.dims[2 * k + 1];
            if (isinf(lo) && lo < 0 && isinf(hi) && hi > 0) { p[k] = inf_pos; a.shape_len[k] = CUDART_INF_F; }
            else { p[k] = __fadd_rn(p[k], lo); a.shape_len[k] = __fsub_rn(hi, lo); }
        }
    } else if (sp.kind == GORDER_GEOM_CYLINDER) {   // geometry.rs:422-451
        float lo = sp.dims[1], hi = sp.dims[2];
        a.shape_radius = sp.dims[0];
        if (isinf(lo) && lo < 0 && isinf(hi) && hi > 0) { p[sp.axis] = inf_pos; a.shape_height = CUDART_INF_F; }
        else { p[sp.axis] = __fadd_rn(p[sp.axis], lo); a.shape_height = __fsub_rn(hi, lo); }
    } else {   // sphere, geometry.rs:507-514
        a.shape_radius = sp.dims[0];
    }
    // a fixed reference point: the reference built the shape once, with the structure file's box (geometry.rs:297-312)
    if (sp.ref_kind == GORDER_GEOMREF_POINT && (sp.structure_box[0] != 0.0f || sp.structure_box[1] != 0.0f || sp.structure_box[2] != 0.0f)) {
#pragma unroll
        for (int k = 0; k < 3; k++) { bx.L[k] = sp.structure_box[k]; bx.half[k] = sp.structure_box[k] / 2.0f; }
    }
    f3 o = wrap_point<PBC>(mk3(p[0], p[1], p[2]), bx);
    a.shape_origin[0] = o.x; a.shape_origin[1] = o.y; a.shape_origin[2] = o.z;
}

__global__ void frame_setup_kernel(DeviceView v, FrameAux *aux, const float *__restrict__ box, int n_frames, int stage) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    FrameAux &a = aux[f];
    if (stage == 0) {
        if (v.handle_pbc) {
            float bx = box[3 * f], by = box[3 * f + 1], bz = box[3 * f + 2];
            if (bx == 0.0f && by == 0.0f && bz == 0.0f) raise_error(v, GORDER_ERR_ZERO_BOX, a.frame_index);
            a.L[0] = bx; a.L[1] = by; a.L[2] = bz;
            a.half[0] = bx / 2.0f; a.half[1] = by / 2.0f; a.half[2] = bz / 2.0f;
        } else {
            a.L[0] = a.L[1] = a.L[2] = 0.0f;
            a.half[0] = a.half[1] = a.half[2] = 0.0f;
        }
        a.guard[0] = 0.99f * a.half[0]; a.guard[1] = 0.99f * a.half[1]; a.guard[2] = 0.99f * a.half[2];
    } else {
        if (v.handle_pbc) construct_shape<true>(v, a);
        else construct_shape<false>(v, a);
    }
}

// ---------------------------------------------------------------------------------------------
// group centre along one axis (groan group_get_center: refined Bai-Breen; _naive: mean), in the order-free
// fixed-point arithmetic of gorder_math.cuh (bit-identical to oracle/gorder_oracle.c group_center):
//   pass 0: sum q(cos 2 pi x / L), q(sin 2 pi x / L)                 -> estimate
//   pass 1: sum q(min_image(x - estimate))                           -> centre = wrap(estimate + mean)
// The group is given as runs of contiguous floats of the native frame (the axis component of its
// atoms: whole planes for typical selections), so the reads are coalesced and carry no index
// traffic.  The sums are integers: every CTA writes its partials, the last CTA of a frame (ticket) adds
// them and finishes the pass -- no separate launch; the result does not depend on the partition.
// frame_list[i]: index in the batch of the i-th frame that needs the centre.
// out: est[3 * i + axis] (pass 0) / center[3 * i + axis] (pass 1; pass 0 when !pbc).
// ---------------------------------------------------------------------------------------------
constexpr int kCenterBlocks = 64;

// sums of one virtual block `vb` of center_axis_kernel (256 threads), result in thread 0 after the reduction
__device__ __forceinline__ void center_block_sums(const Seg *__restrict__ segs, int n_segs, int vb, int n_blocks, const float *__restrict__ fr,
                                                  bool pbc, int pass, float inv_l, float e, float L, float half, long long (&s_red)[3][8],
                                                  long long &t0, long long &t1, bool &tbad) {
    long long a0 = 0, a1 = 0;
    bool bad = false;
    const unsigned long long pol = l2_policy_evict_last();
    const float guard = 0.99f * half;
    auto add = [&](float p) {
        if (!pbc) a0 += center_q(p, bad);
        else if (pass == 0) {
            float sn, cs;
            sincos_turns(__fmul_rn(p, inv_l), sn, cs);
            a0 += center_q(cs, bad); a1 += center_q(sn, bad);
        } else a0 += center_q(min_image_g(__fsub_rn(p, e), L, half, guard), bad);
    };
    for (int sg = vb; sg < n_segs; sg += n_blocks) {
        const Seg sgm = segs[sg];
        const float *src = fr + sgm.off;
        if ((((size_t)src) & 15) == 0) {
            const int n4 = sgm.len >> 2;
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            int i = threadIdx.x;
            for (; i + 3 * (int)blockDim.x < n4; i += 4 * blockDim.x) {
                const float4 q0 = ldg_hint4(src + 4 * (size_t)i, pol), q1 = ldg_hint4(src + 4 * (size_t)(i + blockDim.x), pol),
                             q2 = ldg_hint4(src + 4 * (size_t)(i + 2 * blockDim.x), pol), q3 = ldg_hint4(src + 4 * (size_t)(i + 3 * blockDim.x), pol);
                add(q0.x); add(q0.y); add(q0.z); add(q0.w); add(q1.x); add(q1.y); add(q1.z); add(q1.w);
                add(q2.x); add(q2.y); add(q2.z); add(q2.w); add(q3.x); add(q3.y); add(q3.z); add(q3.w);
            }
            for (; i < n4; i += blockDim.x) { const float4 q = __ldg(s4 + i); add(q.x); add(q.y); add(q.z); add(q.w); }
            for (int k = (n4 << 2) + threadIdx.x; k < sgm.len; k += blockDim.x) add(__ldg(src + k));
        } else {
            for (int i = threadIdx.x; i < sgm.len; i += blockDim.x) add(__ldg(src + i));
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long x0 = a0, x1 = a1;
    for (int o = 16; o > 0; o >>= 1) { x0 += __shfl_down_sync(0xffffffffu, x0, o); x1 += __shfl_down_sync(0xffffffffu, x1, o); }
    const bool wbad = __any_sync(0xffffffffu, bad);
    __syncthreads();   // s_red may still be read from the previous call
    if (lane == 0) { s_red[0][warp] = x0; s_red[1][warp] = x1; s_red[2][warp] = wbad ? 1 : 0; }
    __syncthreads();
    t0 = 0; t1 = 0; tbad = false;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; w++) { t0 += s_red[0][w]; t1 += s_red[1][w]; tbad = tbad || s_red[2][w] != 0; }
}

// ticket word of a (frame, pass): low half = CTAs that have published, high half = CTAs that saw a NaN / Inf term
__global__ void __launch_bounds__(256) center_axis_kernel(DeviceView v, const Seg *__restrict__ segs, int n_segs, int n_group, int axis,
                                                          const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                          const int *__restrict__ frame_list, float *__restrict__ est,
                                                          float *__restrict__ center, long long *__restrict__ partial,
                                                          unsigned *__restrict__ ticket, int pass) {
    const int fi = blockIdx.y, f = frame_list[fi];
    const float *fr = planes + (size_t)f * v.frame_floats;
    const float L = aux[f].L[axis], half = aux[f].half[axis];
    const bool pbc = v.handle_pbc != 0;
    const float inv_l = pbc ? __fdiv_rn(1.0f, L) : 0.0f;
    const float e = (pbc && pass == 1) ? est[3 * fi + axis] : 0.0f;
    __shared__ long long s_red[3][8];
    __shared__ unsigned s_ticket;
    long long b0, b1;
    bool bbad;
    center_block_sums(segs, n_segs, blockIdx.x, gridDim.x, fr, pbc, pass, inv_l, e, L, half, s_red, b0, b1, bbad);
    if (threadIdx.x == 0) {
        long long *pp = partial + ((size_t)fi * gridDim.x + blockIdx.x) * 2;
        pp[0] = b0; pp[1] = b1;
        __threadfence();
        s_ticket = atomicAdd(&ticket[fi], 1u | (bbad ? 0x10000u : 0u)) + (bbad ? 0x10000u : 0u);
    }
    __syncthreads();
    if ((s_ticket & 0xffffu) != gridDim.x - 1 || threadIdx.x != 0) return;
    __threadfence();
    const bool bad = (s_ticket >> 16) != 0;
    long long t0 = 0, t1 = 0;
    for (unsigned b = 0; b < gridDim.x; b++) {
        const volatile long long *pp = partial + ((size_t)fi * gridDim.x + b) * 2;
        t0 += pp[0]; t1 += pp[1];
    }
    ticket[fi] = 0;
    if (!pbc) center[3 * fi + axis] = (n_group > 0 && !bad) ? center_mean(t0, n_group) : CUDART_NAN_F;
    else if (pass == 0) est[3 * fi + axis] = (n_group > 0 && !bad) ? center_estimate(t0, t1, L) : CUDART_NAN_F;
    else {
        const float c = __fadd_rn(e, center_mean(t0, n_group));
        center[3 * fi + axis] = (n_group > 0 && !bad && e == e) ? ((L > 0.0f) ? wrap1(c, L) : c) : CUDART_NAN_F;
    }
}

// copy centres into FrameAux.center (geometry reference)
__global__ void store_center_kernel(FrameAux *aux, const int *frame_list, int n_list, const float *center) {
    int fi = blockIdx.x * blockDim.x + threadIdx.x;
    if (fi >= n_list) return;
    FrameAux &a = aux[frame_list[fi]];
    a.center[0] = center[3 * fi]; a.center[1] = center[3 * fi + 1]; a.center[2] = center[3 * fi + 2];
}

// ---------------------------------------------------------------------------------------------
// leaflet assignment (AssignedLeaflets::assign_lipids, leaflets.rs:1406-1435).
// grid (ceil(n_molpad / 256), n_assign): one thread per (padded molecule, assignment frame).
// Row written: rows[(1 + ai) * n_molpad + molpad]   (row 0 = table carried over from the last batch).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float distance_1d(float a, float b, float L, float half, bool pbc) {
    float d = __fsub_rn(a, b);
    return (pbc && L > 0.0f) ? min_image(d, L, half) : d;
}

// 2-D cell list over the membrane atoms for the Local method (the reference builds a CellGrid with cell edge >= radius
// and visits the 3 x 3 columns around the head, pbc.rs:280-303): columns along the leaflet axis, cells over the two
// lateral axes.  lcell_sorted holds (lateral 0, lateral 1, axis coordinate, -) per atom in cell order.
constexpr int kLCellMaxDim = 256;
__device__ __forceinline__ void lcell_dims(const FrameAux &a, float radius, int ax, int (&n)[2]) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float L = a.L[(ax + 1 + k) % 3];
        n[k] = min(max((int)floorf(L / radius), 1), kLCellMaxDim);
    }
}

__global__ void __launch_bounds__(256) leaflet_assign_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                             const int *__restrict__ frame_list, const float *__restrict__ center,
                                                             const int *__restrict__ molpad_type, unsigned char *__restrict__ rows,
                                                             const int *__restrict__ lcell_start = nullptr, const float4 *__restrict__ lcell_sorted = nullptr,
                                                             int lcells_cap = 0) {
    const int ai = blockIdx.y, f = frame_list[ai];
    const int mp = blockIdx.x * blockDim.x + threadIdx.x;
    if (mp >= v.n_molpad) return;
    const int t = molpad_type[mp];
    if (t < 0) return;
    const TypeDesc &td = v.types[t];   // fields are read individually (read-only cache), not copied
    const int m = mp - __ldg(&td.molpad0);
    unsigned char out = GORDER_UPPER;
    if (m < __ldg(&td.n_mol)) {
        const FrameAux &a = aux[f];
        const int ax = v.leaflet_axis;
        const bool pbc = v.handle_pbc != 0;
        const float L = a.L[ax], half = a.half[ax];
        const float *fr = planes + (size_t)f * v.frame_floats + mol_offset(td, m);
        const int cst = td.cstride;   // component stride
        bool upper = true;
        if (v.leaflet_mode == GORDER_LEAFLET_GLOBAL) {   // leaflets.rs:571-624, :711-732
            float c = center[3 * ai + ax];
            if (c != c) raise_error(v, GORDER_ERR_INVALID_GLOBAL_CENTER, a.frame_index);
            float head = fr[td.head_off + ax * cst];
            upper = distance_1d(head, c, L, half, pbc) >= 0.0f;
        } else if (v.leaflet_mode == GORDER_LEAFLET_INDIVIDUAL) {   // leaflets.rs:736-811
            float head = fr[td.head_off + ax * cst];
            float total = 0.0f;
            for (int k = 0; k < td.n_methyls; k++) {
                float me = fr[v.methyl_offs[td.methyl_off + k] + ax * cst];
                total = __fadd_rn(total, distance_1d(head, me, L, half, pbc));
            }
            upper = total >= 0.0f;
        } else if (v.leaflet_mode == GORDER_LEAFLET_MANUAL) {   // leaflets.rs:816-874
            long long row = v.leaflet_freq_kind == GORDER_FREQ_ONCE ? 0 : a.frame_index / (v.leaflet_freq > 0 ? v.leaflet_freq : 1);
            if (row >= td.n_manual_leaf) { raise_error(v, GORDER_ERR_MANUAL_LEAFLET_FRAME, a.frame_index); }
            else upper = v.manual_leaflets[td.manual_leaf_off + row * td.n_mol + m] == GORDER_UPPER;
        } else if (v.leaflet_mode == GORDER_LEAFLET_LOCAL && lcell_start) {   // leaflets.rs:630-707, pbc.rs:273-318 with the cell list
            const int a0 = (ax + 1) % 3, a1 = (ax + 2) % 3;
            const float h0 = fr[td.head_off + a0 * cst], h1 = fr[td.head_off + a1 * cst], hax = fr[td.head_off + ax * cst];
            int n[2];
            lcell_dims(a, v.leaflet_radius, ax, n);
            const float L0 = a.L[a0], L1 = a.L[a1], hf0 = a.half[a0], hf1 = a.half[a1];
            const int c0 = min(max((int)(wrap1(h0, L0) * (float)n[0] / L0), 0), n[0] - 1), c1 = min(max((int)(wrap1(h1, L1) * (float)n[1] / L1), 0), n[1] - 1);
            const int *st = lcell_start + (size_t)ai * (lcells_cap + 1);
            const float4 *srt = lcell_sorted + (size_t)ai * v.membrane.n;
            const int lo0 = n[0] >= 3 ? -1 : 0, hi0 = n[0] >= 3 ? 1 : n[0] - 1, lo1 = n[1] >= 3 ? -1 : 0, hi1 = n[1] >= 3 ? 1 : n[1] - 1;
            // centre of the atoms inside the cylinder: the order-free sums of center_axis_kernel (gorder_math.cuh)
            const float inv_l = __fdiv_rn(1.0f, L);
            long long sc = 0, ss = 0, sn = 0;
            bool bad = false;
            int cnt = 0;
            float est = 0.0f;
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 1) {
                    if (cnt == 0) break;
                    est = center_estimate(sc, ss, L);
                }
                for (int d0 = lo0; d0 <= hi0; d0++) {
                    const int x0 = n[0] >= 3 ? (c0 + d0 + n[0]) % n[0] : d0;
                    for (int d1 = lo1; d1 <= hi1; d1++) {
                        const int x1 = n[1] >= 3 ? (c1 + d1 + n[1]) % n[1] : d1;
                        const int cell = x0 * n[1] + x1;
                        for (int k = st[cell]; k < st[cell + 1]; k++) {
                            const float4 q = __ldg(srt + k);
                            const float e0 = min_image(__fsub_rn(q.x, h0), L0, hf0), e1 = min_image(__fsub_rn(q.y, h1), L1, hf1);
                            // r2 in the oracle's component order (ascending axis index)
                            const float r2 = a0 < a1 ? __fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)) : __fadd_rn(__fmul_rn(e1, e1), __fmul_rn(e0, e0));
                            if (!(__fsqrt_rn(r2) < v.leaflet_radius)) continue;
                            if (pass == 0) { cnt++; float sv, cv; sincos_turns(__fmul_rn(q.z, inv_l), sv, cv); sc += center_q(cv, bad); ss += center_q(sv, bad); }
                            else sn += center_q(min_image(__fsub_rn(q.z, est), L, half), bad);
                        }
                    }
                }
            }
            float c = CUDART_NAN_F;
            if (cnt > 0 && !bad && est == est) c = wrap1(__fadd_rn(est, center_mean(sn, cnt)), L);
            if (c != c) raise_error(v, GORDER_ERR_INVALID_LOCAL_CENTER, ((long long)t << 32) | (unsigned)m);
            upper = distance_1d(hax, c, L, half, true) >= 0.0f;
        } else if (v.leaflet_mode == GORDER_LEAFLET_LOCAL) {   // leaflets.rs:630-707, pbc.rs:273-318
            // centre of the membrane atoms inside an infinite cylinder around the head (brute force)
            f3 head = mk3(fr[td.head_off], fr[td.head_off + cst], fr[td.head_off + 2 * cst]);
            const float *frame0 = planes + (size_t)f * v.frame_floats;
            const float inv_l = pbc ? __fdiv_rn(1.0f, L) : 0.0f;
            long long sc = 0, ss = 0, sn = 0;
            bool bad = false;
            int cnt = 0;
            float est = 0.0f;
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 1) {
                    if (!pbc || cnt == 0) break;
                    est = center_estimate(sc, ss, L);
                }
                for (int i = 0; i < v.membrane.n; i++) {
                    const int off = v.membrane.off[i], cs = v.membrane.cs[i];
                    float p[3] = {frame0[off], frame0[off + cs], frame0[off + 2 * (size_t)cs]};
                    float r2 = 0.0f;
                    for (int k = 0; k < 3; k++) {
                        if (k == ax) continue;
                        float d = __fsub_rn(p[k], comp(head, k));
                        if (pbc && a.L[k] > 0.0f) d = min_image(d, a.L[k], a.half[k]);
                        r2 = __fadd_rn(r2, __fmul_rn(d, d));
                    }
                    if (!(__fsqrt_rn(r2) < v.leaflet_radius)) continue;
                    if (pass == 0) {
                        cnt++;
                        if (pbc) { float sv, cv; sincos_turns(__fmul_rn(p[ax], inv_l), sv, cv); sc += center_q(cv, bad); ss += center_q(sv, bad); }
                        else sn += center_q(p[ax], bad);
                    } else sn += center_q((L > 0.0f) ? min_image(__fsub_rn(p[ax], est), L, half) : __fsub_rn(p[ax], est), bad);
                }
            }
            float c = CUDART_NAN_F;
            if (cnt > 0 && !bad && est == est) c = pbc ? ((L > 0.0f) ? wrap1(__fadd_rn(est, center_mean(sn, cnt)), L) : __fadd_rn(est, center_mean(sn, cnt))) : center_mean(sn, cnt);
            if (c != c) raise_error(v, GORDER_ERR_INVALID_LOCAL_CENTER, ((long long)t << 32) | (unsigned)m);
            upper = distance_1d(comp(head, ax), c, L, half, pbc) >= 0.0f;
        }
        if (v.leaflet_flip) upper = !upper;   // maybe_flip, leaflets.rs:68-73
        out = upper ? GORDER_UPPER : GORDER_LOWER;
    }
    rows[(size_t)(1 + ai) * v.n_molpad + mp] = out;
}

// ---------------------------------------------------------------------------------------------
// dynamic membrane normals (normal.rs:160-199, pbc.rs:321-351, normal.rs:421-458).
// One thread per (padded molecule, frame): heads within `radius` of this lipid's head (minimum
// image, re-imaged next to the reference), centroid, 3x3 scatter matrix, eigenvector of the
// smallest eigenvalue by cyclic Jacobi, all in registers.  Brute-force neighbour scan (round 1).
// Output planes normals[(f*3 + c) * n_molpad + molpad]; NaN + n_points when < 3 points.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void jacobi_smallest(float a00, float a01, float a02, float a11, float a12, float a22, f3 &out) {
    float A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    // Cyclic Jacobi converges quadratically: once the off-diagonal part is at f32 resolution of the diagonal, one more
    // sweep finishes the eigenvectors.  (Twelve unconditional sweeps, unrolled, were 22 % of the normals kernel.)
#pragma unroll 1
    for (int sweep = 0; sweep < 12; sweep++) {
        const float off = fabsf(A[0][1]) + fabsf(A[0][2]) + fabsf(A[1][2]);
        if (off < 1e-30f) break;
        const bool last = off <= 3e-7f * (fabsf(A[0][0]) + fabsf(A[1][1]) + fabsf(A[2][2]));
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                if (fabsf(A[p][q]) < 1e-37f) continue;
                float theta = (A[q][q] - A[p][p]) / (2.0f * A[p][q]);
                float t = copysignf(1.0f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
                float c = rsqrtf(t * t + 1.0f), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; k++) { float x = A[k][p], y = A[k][q]; A[k][p] = c * x - s * y; A[k][q] = s * x + c * y; }
#pragma unroll
                for (int k = 0; k < 3; k++) { float x = A[p][k], y = A[q][k]; A[p][k] = c * x - s * y; A[q][k] = s * x + c * y; }
#pragma unroll
                for (int k = 0; k < 3; k++) { float x = V[k][p], y = V[k][q]; V[k][p] = c * x - s * y; V[k][q] = s * x + c * y; }
            }
        if (last) break;
    }
    int k = 0;
    if (A[1][1] < A[k][k]) k = 1;
    if (A[2][2] < A[k][k]) k = 2;
    f3 n = mk3(V[0][k], V[1][k], V[2][k]);
    out = unit_ref(n);
}

// ---- the SIGN of the reference's normal ----------------------------------------------------------------------------------
// membrane_normal_from_cloud (normal.rs:421-458) returns the last row of V^T of nalgebra's SVD of the demeaned N x 3 cloud,
// sign included (exported normals are compared signed: tests/common/mod.rs:84-87).  nalgebra's bidiagonalisation applies its
// Householder reflections as sign * H, which makes the bidiagonal matrix and V^T functions of A^T A alone (the first row after
// the first step is (|a0|, a0.a1 / |a0|, a0.a2 / |a0|) whatever the order or the signs of the rows), so running the SAME
// algorithm -- reflection axes  x + sign(x0)|x| e1,  v_t() replayed on the identity, Wilkinson-shift QR sweeps with
// GivensRotation::cancel_y, the 2 x 2 closing step, rows ordered by singular value -- on the 3 x 3 Cholesky factor of the
// scatter matrix gives the same last row.  The device takes the DIRECTION from the Jacobi eigenvector (f64 moments) and only
// the sign from this replay (oracle/gorder_oracle.c nalgebra_svd_last_row is the N x 3 restatement, pinned on the reference's
// 274 signed vectors; tests/test_gpu_parity.py compares signed normals).
__device__ __forceinline__ float signum_rust(float x) { return signbit(x) ? -1.0f : 1.0f; }
__device__ __forceinline__ float refl_axis(float *col, int n, bool &nz) {   // householder::reflection_axis_mut, n <= 3
    float sq = 0.0f;
    for (int i = 0; i < n; i++) sq += col[i] * col[i];
    const float norm = sqrtf(sq), modulus = fabsf(col[0]), signed_norm = signum_rust(col[0]) * norm;
    const float factor = (sq + modulus * norm) * 2.0f;
    col[0] += signed_norm;
    if (factor != 0.0f) {
        const float f = sqrtf(factor);
        float s2 = 0.0f;
        for (int i = 0; i < n; i++) { col[i] /= f; s2 += col[i] * col[i]; }
        const float nn = sqrtf(s2);
        for (int i = 0; i < n; i++) col[i] /= nn;
        nz = true;
        return -signed_norm;
    }
    nz = false;
    return signed_norm;
}
__device__ __forceinline__ bool givens_cancel_y(float x, float y, float &c, float &s, float &r) {
    if (y == 0.0f) return false;
    const float mod0 = fabsf(x), sign0 = signum_rust(x), denom = sqrtf(mod0 * mod0 + y * y);
    c = mod0 / denom; s = -y / (sign0 * denom); r = sign0 * denom;
    return true;
}
__device__ __forceinline__ void givens_new(float cin, float sin_, float &c, float &s, float &norm) {
    const float mod0 = fabsf(cin), sign0 = signum_rust(cin), denom = sqrtf(mod0 * mod0 + sin_ * sin_);
    if (denom > 0.0f) { norm = sign0 * denom; c = mod0 / denom; s = sin_ / norm; }
    else { c = 1.0f; s = 0.0f; norm = 0.0f; }
}
__device__ __forceinline__ void svd3_delimit(float *d, float *o, int end, float eps, int &start_out, int &end_out) {
    int n = end;
    while (n > 0) {
        const int m = n - 1;
        if (o[m] == 0.0f || fabsf(o[m]) <= eps * (fabsf(d[n]) + fabsf(d[m]))) o[m] = 0.0f;
        else break;
        n--;
    }
    if (n == 0) { start_out = 0; end_out = 0; return; }
    int ns = n - 1;
    while (ns > 0) {
        const int m = ns - 1;
        if (fabsf(o[m]) <= eps * (fabsf(d[ns]) + fabsf(d[m]))) { o[m] = 0.0f; break; }
        ns--;
    }
    start_out = ns; end_out = n;
}
// last row of V^T of nalgebra's SVD::new applied to the upper-triangular 3 x 3 matrix a (row-major, overwritten)
__device__ __noinline__ void nalgebra_svd3_last_row(float (&a)[3][3], f3 &out) {
    float amax = 0.0f;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) amax = fmaxf(amax, fabsf(a[i][j]));
    if (amax != 0.0f) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] /= amax;
    float ds[3], os[2];
    bool nz;
    for (int ite = 0; ite < 2; ite++) {
        const int len = 3 - ite;
        float axis[3];
        for (int i = 0; i < len; i++) axis[i] = a[ite + i][ite];
        float rn = refl_axis(axis, len, nz);
        for (int i = 0; i < len; i++) a[ite + i][ite] = axis[i];
        if (nz) {
            const float sign = signum_rust(rn);
            for (int j = ite + 1; j < 3; j++) {
                float dot = 0.0f;
                for (int i = 0; i < len; i++) dot += axis[i] * a[ite + i][j];
                const float factor = dot * (sign * -2.0f);
                for (int i = 0; i < len; i++) a[ite + i][j] = factor * axis[i] + sign * a[ite + i][j];
            }
        }
        ds[ite] = rn;
        const int rl = 2 - ite;
        float rax[2];
        for (int j = 0; j < rl; j++) rax[j] = a[ite][ite + 1 + j];
        rn = refl_axis(rax, rl, nz);
        if (nz) {
            const float sign = signum_rust(rn);
            for (int i = ite + 1; i < 3; i++) {
                float w = 0.0f;
                for (int j = 0; j < rl; j++) w += a[i][ite + 1 + j] * rax[j];
                const float f = w * (sign * -2.0f);
                for (int j = 0; j < rl; j++) a[i][ite + 1 + j] = sign * a[i][ite + 1 + j] + f * rax[j];
            }
        }
        for (int j = 0; j < rl; j++) a[ite][ite + 1 + j] = rax[j];
        os[ite] = rn;
    }
    {
        float axis[1] = {a[2][2]};
        ds[2] = refl_axis(axis, 1, nz);
    }
    float vt[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 1; i >= 0; i--) {
        const int rl = 2 - i;
        float sq = 0.0f;
        for (int j = 0; j < rl; j++) sq += a[i][i + 1 + j] * a[i][i + 1 + j];
        if (sq == 0.0f) continue;
        const float sign = signum_rust(os[i]);
        for (int r = i; r < 3; r++) {
            float w = 0.0f;
            for (int j = 0; j < rl; j++) w += vt[r][i + 1 + j] * a[i][i + 1 + j];
            const float f = w * (sign * -2.0f);
            for (int j = 0; j < rl; j++) vt[r][i + 1 + j] = sign * vt[r][i + 1 + j] + f * a[i][i + 1 + j];
        }
    }
    float d[3] = {fabsf(ds[0]), fabsf(ds[1]), fabsf(ds[2])}, o[2] = {fabsf(os[0]), fabsf(os[1])};
    const float eps = 1.1920929e-07f * 5.0f;
    int start, end;
    svd3_delimit(d, o, 2, eps, start, end);
    for (int niter = 0; end != start && niter < 64; niter++) {
        if (end - start + 1 > 2) {
            const int m = end - 1, nn = end;
            const float dm = d[m], dn = d[nn], fm = o[m];
            const float tmm = dm * dm + o[m - 1] * o[m - 1], tmn = dm * fm, tnn = dn * dn + fm * fm;
            float shift = tnn;
            const float sq = tmn * tmn;
            if (sq != 0.0f) {
                const float dd = (tmm - tnn) * 0.5f;
                shift = tnn - sq / (dd + signum_rust(dd) * sqrtf(dd * dd + sq));
            }
            float vx = d[start] * d[start] - shift, vy = d[start] * o[start];
            for (int k = start; k < nn; k++) {
                const float m12 = (k == nn - 1) ? 0.0f : o[k + 1];
                float s00 = d[k], s01 = o[k], s02 = 0.0f, s10 = 0.0f, s11 = d[k + 1], s12 = m12;
                float c1, s1, norm1, c2, s2, norm2;
                if (!givens_cancel_y(vx, vy, c1, s1, norm1)) break;
                {
                    const float sn = -s1;
                    float a0 = s00, b0 = s01; s00 = a0 * c1 + sn * b0; s01 = -sn * a0 + b0 * c1;
                    a0 = s10; b0 = s11; s10 = a0 * c1 + sn * b0; s11 = -sn * a0 + b0 * c1;
                }
                if (k > start) o[k - 1] = norm1;
                if (!givens_cancel_y(s00, s10, c2, s2, norm2)) { c2 = 1.0f; s2 = 0.0f; norm2 = s00; }
                {
                    float a0 = s01, b0 = s11; s01 = a0 * c2 - s2 * b0; s11 = s2 * a0 + b0 * c2;
                    a0 = s02; b0 = s12; s02 = a0 * c2 - s2 * b0; s12 = s2 * a0 + b0 * c2;
                }
                s00 = norm2;
                for (int j = 0; j < 3; j++) { const float p = vt[k][j], q = vt[k + 1][j]; vt[k][j] = p * c1 - s1 * q; vt[k + 1][j] = s1 * p + q * c1; }
                d[k] = s00; d[k + 1] = s11; o[k] = s01;
                if (k != nn - 1) o[k + 1] = s12;
                vx = s01; vy = s02;
            }
        } else {
            const float m11 = d[start], m12 = o[start], m22 = d[start + 1];
            const float denom = hypotf(m11 + m22, m12) + hypotf(m11 - m22, m12);
            float v1 = m11 * m22 * 2.0f / denom, v2 = 0.5f * denom;
            float cv, sv, sgn_v, cu_, su_, sgn_u;
            const float diff = fabsf(m11) >= fabsf(m22) ? v1 * v1 - m11 * m11 : -(m11 * m11 * m12 * m12) / (m12 * m12 + m22 * m22 - v1 * v1);
            givens_new(m11 * m12, diff, cv, sv, sgn_v);
            v1 *= sgn_v; v2 *= sgn_v;
            givens_new((m11 * cv + m12 * sv) / v1, (m22 * sv) / v1, cu_, su_, sgn_u);
            v1 *= sgn_u; v2 *= sgn_u;
            d[start] = v1; d[start + 1] = v2; o[start] = 0.0f;
            const float si = -sv;
            for (int j = 0; j < 3; j++) { const float p = vt[start][j], q = vt[start + 1][j]; vt[start][j] = p * cv - si * q; vt[start + 1][j] = si * p + q * cv; }
            end -= 1;
        }
        svd3_delimit(d, o, end, eps, start, end);
    }
    int k = 0;
    for (int i = 1; i < 3; i++) if (fabsf(d[i]) <= fabsf(d[k])) k = i;
    out = mk3(vt[k][0], vt[k][1], vt[k][2]);
}

// Normal of a cloud from its scatter matrix (sum of outer products of the demeaned positions): direction by Jacobi, sign as
// the reference's SVD gives it.
// `signed_`: only exported normals need the reference's sign (S is invariant under n -> -n).
__device__ __forceinline__ void pca_normal(double a00, double a01, double a02, double a11, double a12, double a22, bool signed_, f3 &out) {
    f3 n;
    jacobi_smallest((float)a00, (float)a01, (float)a02, (float)a11, (float)a12, (float)a22, n);
    // Cholesky factor (upper) of the scatter matrix = R of the cloud's QR factorisation with a positive diagonal
    const double r00 = sqrt(fmax(a00, 0.0));
    if (signed_ && r00 > 0.0) {
        const double r01 = a01 / r00, r02 = a02 / r00;
        const double r11 = sqrt(fmax(a11 - r01 * r01, 0.0));
        if (r11 > 0.0) {
            const double r12 = (a12 - r01 * r02) / r11;
            const double r22 = sqrt(fmax(a22 - r02 * r02 - r12 * r12, 0.0));
            float R[3][3] = {{(float)r00, (float)r01, (float)r02}, {0.0f, (float)r11, (float)r12}, {0.0f, 0.0f, (float)r22}};
            f3 ref;
            nalgebra_svd3_last_row(R, ref);
            if (ref.x * n.x + ref.y * n.y + ref.z * n.z < 0.0f) n = mk3(-n.x, -n.y, -n.z);
        }
    }
    out = n;
}

__global__ void __launch_bounds__(128) dynamic_normal_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                             const int *__restrict__ molpad_type, float *__restrict__ normals,
                                                             int *__restrict__ normal_npoints) {
    const int f = blockIdx.y;
    const int mp = blockIdx.x * blockDim.x + threadIdx.x;
    if (mp >= v.n_molpad) return;
    const int t = molpad_type[mp];
    float *nx = normals + ((size_t)f * 3) * v.n_molpad + mp;
    if (t < 0) { nx[0] = CUDART_NAN_F; nx[v.n_molpad] = CUDART_NAN_F; nx[2 * (size_t)v.n_molpad] = CUDART_NAN_F; return; }
    const TypeDesc &td = v.types[t];
    const int m = mp - td.molpad0;
    if (m >= td.n_mol || td.nhead_off < 0) { nx[0] = CUDART_NAN_F; nx[v.n_molpad] = CUDART_NAN_F; nx[2 * (size_t)v.n_molpad] = CUDART_NAN_F; return; }
    const FrameAux &a = aux[f];
    const Box bx = load_box(a);
    const float *frame0 = planes + (size_t)f * v.frame_floats;
    const float *fr = frame0 + mol_offset(td, m);
    const f3 ref = mk3(fr[td.nhead_off], fr[td.nhead_off + td.cstride], fr[td.nhead_off + 2 * td.cstride]);
    const bool pbc = v.handle_pbc != 0;
    // pass 1: centroid of the cloud (f32 running sum in group order, as the reference's fold)
    int n = 0;
    f3 sum = mk3(0, 0, 0);
    for (int i = 0; i < v.normal_heads.n; i++) {
        const int off = v.normal_heads.off[i], cs = v.normal_heads.cs[i];
        f3 p = mk3(frame0[off], frame0[off + cs], frame0[off + 2 * (size_t)cs]);
        f3 d = pbc ? vector_to<true>(ref, p, bx) : vector_to<false>(ref, p, bx);
        if (norm_ref(d) < v.dynamic_radius) {
            f3 q = pbc ? mk3(__fadd_rn(ref.x, d.x), __fadd_rn(ref.y, d.y), __fadd_rn(ref.z, d.z)) : p;
            sum = mk3(__fadd_rn(sum.x, q.x), __fadd_rn(sum.y, q.y), __fadd_rn(sum.z, q.z));
            n++;
        }
    }
    normal_npoints[(size_t)f * v.n_molpad + mp] = n;
    if (n < 3) { nx[0] = CUDART_NAN_F; nx[v.n_molpad] = CUDART_NAN_F; nx[2 * (size_t)v.n_molpad] = CUDART_NAN_F; return; }
    const float fn = (float)n;
    const f3 c = mk3(__fdiv_rn(sum.x, fn), __fdiv_rn(sum.y, fn), __fdiv_rn(sum.z, fn));
    // pass 2: scatter matrix of the demeaned cloud
    float a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    for (int i = 0; i < v.normal_heads.n; i++) {
        const int off = v.normal_heads.off[i], cs = v.normal_heads.cs[i];
        f3 p = mk3(frame0[off], frame0[off + cs], frame0[off + 2 * (size_t)cs]);
        f3 d = pbc ? vector_to<true>(ref, p, bx) : vector_to<false>(ref, p, bx);
        if (norm_ref(d) < v.dynamic_radius) {
            f3 q = pbc ? mk3(__fadd_rn(ref.x, d.x), __fadd_rn(ref.y, d.y), __fadd_rn(ref.z, d.z)) : p;
            float dx = __fsub_rn(q.x, c.x), dy = __fsub_rn(q.y, c.y), dz = __fsub_rn(q.z, c.z);
            a00 += dx * dx; a01 += dx * dy; a02 += dx * dz; a11 += dy * dy; a12 += dy * dz; a22 += dz * dz;
        }
    }
    f3 nrm;
    pca_normal(a00, a01, a02, a11, a12, a22, v.collect_normals != 0, nrm);
    nx[0] = nrm.x; nx[v.n_molpad] = nrm.y; nx[2 * (size_t)v.n_molpad] = nrm.z;
}

// ---------------------------------------------------------------------------------------------
// K4 with a cell list (PBC): the reference builds a CellGrid over the NormalHeads group with cell
// edge >= radius and scans the 27 neighbouring cells (pbc.rs:327-339); so does this.  Per frame:
//   cell_count_kernel  wrap every head into the box, bin it, count                 (1 thread / head)
//   cell_scan_kernel   exclusive scan of the counts -> cell_start                  (1 CTA / frame)
//   cell_fill_kernel   scatter head ids into their cell's range
//   dynamic_normal_cell_kernel  per lipid: gather heads within `radius` from the 3x3x3 block,
//                      centroid + scatter matrix relative to the reference head (f64 sums: the
//                      gather order inside a cell is arbitrary, f64 makes the result independent of
//                      it to ~1e-13, i.e. bit-stable after rounding to f32), Jacobi smallest eigenvector.
// Grid: n_k = max(floor(L_k / radius), 1) cells along k (cell edge L_k / n_k >= radius), at most kCellBudget cells per frame.
// ---------------------------------------------------------------------------------------------
constexpr int kCellBudget = 1 << 18;   // cells per frame (count + start arrays: 2 MB per frame)

__device__ __forceinline__ void cell_dims(const FrameAux &a, float radius, int (&n)[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) n[k] = min(max((int)floorf(a.L[k] / radius), 1), 4096);
    // a very large box: shrink the grid uniformly (coarser cells are still correct, the exact distance test decides)
    long long prod = (long long)n[0] * n[1] * n[2];
    if (prod > kCellBudget) {
        const float sc = cbrtf((float)kCellBudget / (float)prod);
#pragma unroll
        for (int k = 0; k < 3; k++) n[k] = max((int)((float)n[k] * sc), 1);
        while ((long long)n[0] * n[1] * n[2] > kCellBudget) {
            const int big = (n[0] >= n[1] && n[0] >= n[2]) ? 0 : (n[1] >= n[2] ? 1 : 2);
            if (big == 0) n[0]--; else if (big == 1) n[1]--; else n[2]--;
        }
    }
}
__device__ __forceinline__ int cell_coord(float x, float L, int n) {
    const float w = (L > 0.0f) ? wrap1(x, L) : 0.0f;
    return min(max((int)(w * (float)n / L), 0), n - 1);
}

__global__ void __launch_bounds__(256) cell_count_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                         int *__restrict__ head_cell, int *__restrict__ cell_count, int cells_cap) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.normal_heads.n) return;
    const FrameAux &a = aux[f];
    int n[3];
    cell_dims(a, v.dynamic_radius, n);
    const float *fr = planes + (size_t)f * v.frame_floats;
    const int off = v.normal_heads.off[i], cs = v.normal_heads.cs[i];
    const int cx = cell_coord(fr[off], a.L[0], n[0]), cy = cell_coord(fr[off + cs], a.L[1], n[1]), cz = cell_coord(fr[off + 2 * (size_t)cs], a.L[2], n[2]);
    const int c = (cx * n[1] + cy) * n[2] + cz;
    head_cell[(size_t)f * v.normal_heads.n + i] = c;
    atomicAdd(&cell_count[(size_t)f * cells_cap + c], 1);
}

// Exclusive scan of the cell counts, two levels: a CTA owns kScanSpan consecutive cells of one frame (a vesicle in a big box
// has 2e5 cells, mostly empty: one CTA per frame walked them in 23 passes, 15 us per frame of pure latency).
constexpr int kScanPer = 8;                      // consecutive cells per thread
constexpr int kScanSpan = 1024 * kScanPer;       // cells per CTA
constexpr int kScanBlocks = (kCellBudget + kScanSpan - 1) / kScanSpan;   // CTAs per frame (32)

__global__ void __launch_bounds__(1024) cell_span_sum_kernel(DeviceView v, const FrameAux *__restrict__ aux, const int *__restrict__ cell_count,
                                                             int *__restrict__ span_sum, int cells_cap) {
    const int f = blockIdx.y;
    int n[3];
    cell_dims(aux[f], v.dynamic_radius, n);
    const int nc = n[0] * n[1] * n[2];
    const int i0 = blockIdx.x * kScanSpan + threadIdx.x * kScanPer;
    const int *cnt = cell_count + (size_t)f * cells_cap;
    int tot = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++) tot += i0 + j < nc ? cnt[i0 + j] : 0;
    __shared__ int s_warp[32];
    tot = __reduce_add_sync(0xffffffffu, tot);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int w = __reduce_add_sync(0xffffffffu, s_warp[threadIdx.x]);
        if (threadIdx.x == 0) span_sum[f * kScanBlocks + blockIdx.x] = w;
    }
}

__global__ void __launch_bounds__(1024) cell_scan_kernel(DeviceView v, const FrameAux *__restrict__ aux, int *__restrict__ cell_count,
                                                         int *__restrict__ cell_start, const int *__restrict__ span_sum, int cells_cap) {
    const int f = blockIdx.y;
    int n[3];
    cell_dims(aux[f], v.dynamic_radius, n);
    const int nc = n[0] * n[1] * n[2];
    if (blockIdx.x * kScanSpan >= nc) return;
    int *cnt = cell_count + (size_t)f * cells_cap, *st = cell_start + (size_t)f * (cells_cap + 1);
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {   // cells in the spans before this one
        const int w = __reduce_add_sync(0xffffffffu, lane < (int)blockIdx.x ? span_sum[f * kScanBlocks + lane] : 0);
        if (lane == 0) s_carry = w;
    }
    const int i0 = blockIdx.x * kScanSpan + threadIdx.x * kScanPer;
    int x[kScanPer], tot = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++) { x[j] = i0 + j < nc ? cnt[i0 + j] : 0; tot += x[j]; }
    int incl = tot;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        s_warp[lane] = w;
    }
    __syncthreads();
    int excl = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - tot;
#pragma unroll
    for (int j = 0; j < kScanPer; j++) {
        if (i0 + j < nc) { st[i0 + j] = excl; cnt[i0 + j] = 0; }   // the counts become the fill cursors
        excl += x[j];
    }
    if (i0 <= nc - 1 && nc - 1 < i0 + kScanPer) st[nc] = excl;      // the thread that owns the last cell closes the table
}

__global__ void __launch_bounds__(256) cell_fill_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux, const int *__restrict__ head_cell,
                                                        int *__restrict__ cell_count, const int *__restrict__ cell_start,
                                                        float4 *__restrict__ sorted_pos, int cells_cap) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.normal_heads.n) return;
    const int c = head_cell[(size_t)f * v.normal_heads.n + i];
    const int pos = cell_start[(size_t)f * (cells_cap + 1) + c] + atomicAdd(&cell_count[(size_t)f * cells_cap + c], 1);
    // the coordinates travel with the index: the gather loop reads ONE contiguous float4 per candidate
    const float *fr = planes + (size_t)f * v.frame_floats;
    const int off = v.normal_heads.off[i], cs = v.normal_heads.cs[i];
    const float x = fr[off], y = fr[off + cs], z = fr[off + 2 * (size_t)cs];
    // w: the head's index; complemented (negative) when the head lies outside [0, L): the gather loop then skips its
    // cell-based image shortcut for this candidate
    const FrameAux &a = aux[f];
    const bool inside = x >= 0.0f && x < a.L[0] && y >= 0.0f && y < a.L[1] && z >= 0.0f && z < a.L[2];
    sorted_pos[(size_t)f * v.normal_heads.n + pos] = make_float4(x, y, z, __int_as_float(inside ? i : ~i));
}

__global__ void __launch_bounds__(128) dynamic_normal_cell_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                                  const int *__restrict__ molpad_type, const int *__restrict__ cell_start,
                                                                  const float4 *__restrict__ sorted_pos, int cells_cap, float *__restrict__ normals,
                                                                  int *__restrict__ normal_npoints) {
    const int f = blockIdx.y;
    const int mp = blockIdx.x * blockDim.x + threadIdx.x;
    if (mp >= v.n_molpad) return;
    float *nx = normals + ((size_t)f * 3) * v.n_molpad + mp;
    const int t = molpad_type[mp];
    const TypeDesc &td = v.types[t];
    const int m = mp - td.molpad0;
    if (m >= td.n_mol || td.nhead_off < 0) { nx[0] = CUDART_NAN_F; nx[v.n_molpad] = CUDART_NAN_F; nx[2 * (size_t)v.n_molpad] = CUDART_NAN_F; return; }
    const FrameAux &a = aux[f];
    const Box bx = load_box(a);
    int n[3];
    cell_dims(a, v.dynamic_radius, n);
    const float *frame0 = planes + (size_t)f * v.frame_floats;
    const float *fr = frame0 + mol_offset(td, m);
    const f3 ref = mk3(fr[td.nhead_off], fr[td.nhead_off + td.cstride], fr[td.nhead_off + 2 * td.cstride]);
    const int c0[3] = {cell_coord(ref.x, a.L[0], n[0]), cell_coord(ref.y, a.L[1], n[1]), cell_coord(ref.z, a.L[2], n[2])};
    const int *st = cell_start + (size_t)f * (cells_cap + 1);
    const float4 *srt = sorted_pos + (size_t)f * v.normal_heads.n;
    const float g0 = 0.99f * a.half[0], g1 = 0.99f * a.half[1], g2 = 0.99f * a.half[2];
    const float r2max = v.dynamic_radius;
    // Cheap rejection: inside a cell of the 3 x 3 x 3 block the periodic image next to the reference is known from the cell
    // offset alone (cell edge >= radius, >= 3 cells per axis, both atoms inside the box), so most candidates cost
    // 3 subtractions + |r|^2; everything within radius + slack is decided by the exact fold and the exact `<`.
    const bool ref_in = ref.x >= 0.0f && ref.x < a.L[0] && ref.y >= 0.0f && ref.y < a.L[1] && ref.z >= 0.0f && ref.z < a.L[2];
    const bool quick = ref_in && n[0] >= 3 && n[1] >= 3 && n[2] >= 3;
    // slack: the shifted reference is rounded at the magnitude of the box (ulp(L) ~ 1.2e-7 L)
    const float r_hi = r2max + 1e-5f * r2max + 6e-7f * fmaxf(a.L[0], fmaxf(a.L[1], a.L[2]));
    const float r2hi = r_hi * r_hi;
    int cnt = 0;
    double sx = 0, sy = 0, sz = 0, xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    // with fewer than 3 cells along an axis the +-1 neighbours alias: visit every cell of that axis once
    const int lo0 = n[0] >= 3 ? -1 : 0, hi0 = n[0] >= 3 ? 1 : n[0] - 1;
    const int lo1 = n[1] >= 3 ? -1 : 0, hi1 = n[1] >= 3 ? 1 : n[1] - 1;
    const int lo2 = n[2] >= 3 ? -1 : 0, hi2 = n[2] >= 3 ? 1 : n[2] - 1;
    // (four-way unrolled candidate loads with merged z columns were measured 2x SLOWER: 85 registers and more
    //  divergent code; the simple loop below, one candidate at a time, is what the scheduler hides best)
    for (int dx = lo0; dx <= hi0; dx++) {
        const int cx = n[0] >= 3 ? (c0[0] + dx + n[0]) % n[0] : dx;
        // reference shifted into the candidate cell's image: r = q - refs is the nearest-image vector
        const float rsx = ref.x + (c0[0] + dx < 0 ? a.L[0] : (c0[0] + dx >= n[0] ? -a.L[0] : 0.0f));
        for (int dy = lo1; dy <= hi1; dy++) {
            const int cy = n[1] >= 3 ? (c0[1] + dy + n[1]) % n[1] : dy;
            const float rsy = ref.y + (c0[1] + dy < 0 ? a.L[1] : (c0[1] + dy >= n[1] ? -a.L[1] : 0.0f));
            for (int dz = lo2; dz <= hi2; dz++) {
                const int cz = n[2] >= 3 ? (c0[2] + dz + n[2]) % n[2] : dz;
                const float rsz = ref.z + (c0[2] + dz < 0 ? a.L[2] : (c0[2] + dz >= n[2] ? -a.L[2] : 0.0f));
                const int c = (cx * n[1] + cy) * n[2] + cz;
                for (int k = st[c]; k < st[c + 1]; k++) {
                    const float4 q = __ldg(srt + k);
                    if (quick && __float_as_int(q.w) >= 0) {
                        const float ex = q.x - rsx, ey = q.y - rsy, ez = q.z - rsz;
                        if (fmaf(ez, ez, fmaf(ey, ey, ex * ex)) > r2hi) continue;
                    }
                    f3 d;   // Vector3D::vector_to(reference, head): same fold as everywhere (single-compare fast path)
                    d.x = min_image_g(__fsub_rn(q.x, ref.x), a.L[0], a.half[0], g0);
                    d.y = min_image_g(__fsub_rn(q.y, ref.y), a.L[1], a.half[1], g1);
                    d.z = min_image_g(__fsub_rn(q.z, ref.z), a.L[2], a.half[2], g2);
                    if (norm_ref(d) < r2max) {
                        cnt++;
                        sx += d.x; sy += d.y; sz += d.z;
                        xx += (double)d.x * d.x; xy += (double)d.x * d.y; xz += (double)d.x * d.z;
                        yy += (double)d.y * d.y; yz += (double)d.y * d.z; zz += (double)d.z * d.z;
                    }
                }
            }
        }
    }
    normal_npoints[(size_t)f * v.n_molpad + mp] = cnt;
    if (cnt < 3) { nx[0] = CUDART_NAN_F; nx[v.n_molpad] = CUDART_NAN_F; nx[2 * (size_t)v.n_molpad] = CUDART_NAN_F; return; }
    const double inv = 1.0 / cnt, mx = sx * inv, my = sy * inv, mz = sz * inv;
    f3 nrm;
    pca_normal(xx - cnt * mx * mx, xy - cnt * mx * my, xz - cnt * mx * mz, yy - cnt * my * my, yz - cnt * my * mz, zz - cnt * mz * mz, v.collect_normals != 0, nrm);
    nx[0] = nrm.x; nx[v.n_molpad] = nrm.y; nx[2 * (size_t)v.n_molpad] = nrm.z;
}

// K4, sorted variant (default when every analysed lipid's head is a member of the NormalHeads group).
// dynamic_normal_cell_kernel above gives lane j the lipid j, so the lanes of a warp sit in unrelated cells: their 27 candidate
// loops have unrelated trip counts and the f64 accumulation runs whenever ANY lane accepts -- ncu: 17 of 32 lanes active in
// the candidate loop, 10 of 32 in the accepting branch, 67 % of the kernel's instructions there.  Here
//   * lane k takes the k-th head of the CELL-SORTED list (cell_fill_kernel): the lanes of a warp share a handful of home
//     cells, hence candidate ranges, and `sorted_pos[k]` is the lipid's own head (coalesced);
//   * the three z cells of a column are one contiguous run of the sorted list (cells are numbered z-fastest), split only at
//     the periodic boundary: 9 - 18 ranges instead of 27 cell look-ups;
//   * stage 1 only filters: a candidate that passes the cheap image-by-cell-offset bound (or cannot use it) has its index
//     appended to the lane's list in shared memory; stage 2 runs the exact fold, the exact `<` and the f64 moments over the
//     list, ~22 entries that nearly all pass, so the lanes stay together.
// Same arithmetic per accepted head as the kernel above (the sums are f64: independent of the order to ~1e-13).
constexpr int kNormCap = 48;   // list entries per lane (a full list is drained early)

struct NormAcc {
    int cnt;
    double sx, sy, sz, xx, xy, xz, yy, yz, zz;
};
__device__ __forceinline__ void normal_drain(const int *lst, int nl, const float4 *__restrict__ srt, const f3 &ref, const FrameAux &a,
                                             float g0, float g1, float g2, float radius, NormAcc &s) {
    for (int j = 0; j < nl; j++) {
        const float4 q = __ldg(srt + lst[j * 128]);
        f3 d;   // Vector3D::vector_to(reference, head)
        d.x = min_image_g(__fsub_rn(q.x, ref.x), a.L[0], a.half[0], g0);
        d.y = min_image_g(__fsub_rn(q.y, ref.y), a.L[1], a.half[1], g1);
        d.z = min_image_g(__fsub_rn(q.z, ref.z), a.L[2], a.half[2], g2);
        if (radius_less(dot_ref(d, d), radius)) {   // norm_ref(d) < radius, the IEEE square root only next to the sphere
            s.cnt++;
            s.sx += d.x; s.sy += d.y; s.sz += d.z;
            s.xx += (double)d.x * d.x; s.xy += (double)d.x * d.y; s.xz += (double)d.x * d.z;
            s.yy += (double)d.y * d.y; s.yz += (double)d.y * d.z; s.zz += (double)d.z * d.z;
        }
    }
}

__global__ void __launch_bounds__(128) dynamic_normal_sorted_kernel(DeviceView v, const FrameAux *__restrict__ aux, const int *__restrict__ head_molpad,
                                                                    const int *__restrict__ cell_start, const float4 *__restrict__ sorted_pos,
                                                                    int cells_cap, float *__restrict__ normals, int *__restrict__ normal_npoints) {
    __shared__ int s_list[kNormCap * 128];
    const int f = blockIdx.y, nh = v.normal_heads.n;
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= nh) return;
    const float4 *srt = sorted_pos + (size_t)f * nh;
    const float4 me = __ldg(srt + k);
    const int wi = __float_as_int(me.w);
    const int mp = head_molpad[wi >= 0 ? wi : ~wi];
    if (mp < 0) return;   // a head that only takes part in the clouds of others
    const FrameAux &a = aux[f];
    int n[3];
    cell_dims(a, v.dynamic_radius, n);
    const f3 ref = mk3(me.x, me.y, me.z);
    const int c0[3] = {cell_coord(ref.x, a.L[0], n[0]), cell_coord(ref.y, a.L[1], n[1]), cell_coord(ref.z, a.L[2], n[2])};
    const int *st = cell_start + (size_t)f * (cells_cap + 1);
    const float g0 = 0.99f * a.half[0], g1 = 0.99f * a.half[1], g2 = 0.99f * a.half[2];
    const float radius = v.dynamic_radius;
    const bool quick = wi >= 0 && n[0] >= 3 && n[1] >= 3 && n[2] >= 3;   // wi >= 0: the reference lies inside [0, L)
    const float r_hi = radius + 1e-5f * radius + 6e-7f * fmaxf(a.L[0], fmaxf(a.L[1], a.L[2]));
    const float r2hi = r_hi * r_hi;
    NormAcc acc{};
    int *lst = s_list + threadIdx.x;
    int nl = 0;
    const int lo0 = n[0] >= 3 ? -1 : 0, hi0 = n[0] >= 3 ? 1 : n[0] - 1;
    const int lo1 = n[1] >= 3 ? -1 : 0, hi1 = n[1] >= 3 ? 1 : n[1] - 1;
    // z runs of a column: the cells inside [0, n2) as one run, a wrapped cell below / above as a run of its own
    const bool z3 = n[2] >= 3;
    const int za = z3 ? max(c0[2] - 1, 0) : 0, zb = z3 ? min(c0[2] + 1, n[2] - 1) : n[2] - 1;
    const int zw = !z3 ? -1 : (c0[2] == 0 ? n[2] - 1 : (c0[2] == n[2] - 1 ? 0 : -1));   // the wrapped cell, if any
    const float zshift = c0[2] == 0 ? a.L[2] : -a.L[2];
    for (int dx = lo0; dx <= hi0; dx++) {
        int cx = dx;
        float rsx = ref.x;
        if (n[0] >= 3) {   // neighbour cell and the image of the reference next to it (no division: |dx| <= 1)
            cx = c0[0] + dx;
            if (cx < 0) { cx += n[0]; rsx += a.L[0]; } else if (cx >= n[0]) { cx -= n[0]; rsx -= a.L[0]; }
        }
        for (int dy = lo1; dy <= hi1; dy++) {
            int cy = dy;
            float rsy = ref.y;
            if (n[1] >= 3) {
                cy = c0[1] + dy;
                if (cy < 0) { cy += n[1]; rsy += a.L[1]; } else if (cy >= n[1]) { cy -= n[1]; rsy -= a.L[1]; }
            }
            const int cb = (cx * n[1] + cy) * n[2];
            for (int part = 0; part < (zw >= 0 ? 2 : 1); part++) {
                const int kb = part == 0 ? st[cb + za] : st[cb + zw], ke = part == 0 ? st[cb + zb + 1] : st[cb + zw + 1];
                const float rsz = part == 0 ? ref.z : ref.z + zshift;
                const float4 *qp = srt + kb;
                for (int kk = kb; kk < ke; kk++, qp++) {
                    const float4 q = __ldg(qp);
                    // no branch on the outcome: the slot behind the list is always free, the index is written there and
                    // kept by advancing the count (the accepting branch ran with a quarter of the lanes)
                    const float ex = q.x - rsx, ey = q.y - rsy, ez = q.z - rsz;
                    const bool far = fmaf(ez, ez, fmaf(ey, ey, ex * ex)) > r2hi;
                    lst[nl * 128] = kk;
                    nl += (far && quick && __float_as_int(q.w) >= 0) ? 0 : 1;
                    if (nl == kNormCap) { normal_drain(lst, nl, srt, ref, a, g0, g1, g2, radius, acc); nl = 0; }
                }
            }
        }
    }
    normal_drain(lst, nl, srt, ref, a, g0, g1, g2, radius, acc);
    float *nx = normals + ((size_t)f * 3) * v.n_molpad + mp;
    const int cnt = acc.cnt;
    normal_npoints[(size_t)f * v.n_molpad + mp] = cnt;
    if (cnt < 3) { nx[0] = CUDART_NAN_F; nx[v.n_molpad] = CUDART_NAN_F; nx[2 * (size_t)v.n_molpad] = CUDART_NAN_F; return; }
    const double inv = 1.0 / cnt, mx = acc.sx * inv, my = acc.sy * inv, mz = acc.sz * inv;
    f3 nrm;
    pca_normal(acc.xx - cnt * mx * mx, acc.xy - cnt * mx * my, acc.xz - cnt * mx * mz, acc.yy - cnt * my * my, acc.yz - cnt * my * mz,
               acc.zz - cnt * mz * mz, v.collect_normals != 0, nrm);
    nx[0] = nrm.x; nx[v.n_molpad] = nrm.y; nx[2 * (size_t)v.n_molpad] = nrm.z;
}

// 2-D cell list of the membrane atoms (Local leaflets): count / scan / fill over the assignment frames of a batch.
__global__ void __launch_bounds__(256) lcell_count_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                          const int *__restrict__ frame_list, int *__restrict__ atom_cell, int *__restrict__ cell_count,
                                                          int cells_cap) {
    const int ai = blockIdx.y, f = frame_list[ai], i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.membrane.n) return;
    const FrameAux &a = aux[f];
    const int ax = v.leaflet_axis, a0 = (ax + 1) % 3, a1 = (ax + 2) % 3;
    int n[2];
    lcell_dims(a, v.leaflet_radius, ax, n);
    const float *fr = planes + (size_t)f * v.frame_floats;
    const int off = v.membrane.off[i];
    const size_t cs = (size_t)v.membrane.cs[i];
    const float p0 = fr[off + a0 * cs], p1 = fr[off + a1 * cs];
    const int c0 = min(max((int)(wrap1(p0, a.L[a0]) * (float)n[0] / a.L[a0]), 0), n[0] - 1);
    const int c1 = min(max((int)(wrap1(p1, a.L[a1]) * (float)n[1] / a.L[a1]), 0), n[1] - 1);
    const int c = c0 * n[1] + c1;
    atom_cell[(size_t)ai * v.membrane.n + i] = c;
    atomicAdd(&cell_count[(size_t)ai * cells_cap + c], 1);
}

__global__ void __launch_bounds__(1024) lcell_scan_kernel(DeviceView v, const FrameAux *__restrict__ aux, const int *__restrict__ frame_list,
                                                          int *__restrict__ cell_count, int *__restrict__ cell_start, int cells_cap) {
    const int ai = blockIdx.x;
    int n[2];
    lcell_dims(aux[frame_list[ai]], v.leaflet_radius, v.leaflet_axis, n);
    const int nc = n[0] * n[1];
    int *cnt = cell_count + (size_t)ai * cells_cap, *st = cell_start + (size_t)ai * (cells_cap + 1);
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nc; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int x = i < nc ? cnt[i] : 0;
        int incl = x;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int excl = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - x;
        if (i < nc) { st[i] = excl; cnt[i] = 0; }   // the counts become the fill cursors
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) st[nc] = s_carry;
}

__global__ void __launch_bounds__(256) lcell_fill_kernel(DeviceView v, const float *__restrict__ planes, const int *__restrict__ frame_list,
                                                         const int *__restrict__ atom_cell, int *__restrict__ cell_count,
                                                         const int *__restrict__ cell_start, float4 *__restrict__ sorted_pos, int cells_cap) {
    const int ai = blockIdx.y, f = frame_list[ai], i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.membrane.n) return;
    const int c = atom_cell[(size_t)ai * v.membrane.n + i];
    const int pos = cell_start[(size_t)ai * (cells_cap + 1) + c] + atomicAdd(&cell_count[(size_t)ai * cells_cap + c], 1);
    const int ax = v.leaflet_axis, a0 = (ax + 1) % 3, a1 = (ax + 2) % 3;
    const float *fr = planes + (size_t)f * v.frame_floats;
    const int off = v.membrane.off[i];
    const size_t cs = (size_t)v.membrane.cs[i];
    sorted_pos[(size_t)ai * v.membrane.n + pos] = make_float4(fr[off + a0 * cs], fr[off + a1 * cs], fr[off + ax * cs], 0.0f);
}

// manual membrane normals (ManualMembraneNormal::get_normal, normal.rs:266-298): copy the row of
// this frame into the per-frame normal planes; NaN when the frame is not available (the error is
// raised by the accumulation kernel when the normal is actually used).
__global__ void __launch_bounds__(256) manual_normal_kernel(DeviceView v, const FrameAux *__restrict__ aux, const int *__restrict__ molpad_type,
                                                            float *__restrict__ normals) {
    const int f = blockIdx.y;
    const int mp = blockIdx.x * blockDim.x + threadIdx.x;
    if (mp >= v.n_molpad) return;
    float *nx = normals + ((size_t)f * 3) * v.n_molpad + mp;
    float x = CUDART_NAN_F, y = CUDART_NAN_F, z = CUDART_NAN_F;
    const int t = molpad_type[mp];
    if (t >= 0) {
        const TypeDesc &td = v.types[t];
        const int m = mp - td.molpad0, row = aux[f].manual_norm_row;
        if (m < td.n_mol && td.manual_norm_off >= 0 && row < td.n_manual_norm) {
            const float *p = v.manual_normals + td.manual_norm_off + 3 * ((size_t)row * td.n_mol + m);
            x = p[0]; y = p[1]; z = p[2];
        }
    }
    nx[0] = x; nx[v.n_molpad] = y; nx[2 * (size_t)v.n_molpad] = z;
}

__global__ void __launch_bounds__(256) mask_normals_kernel(int n_molpad, const unsigned char *__restrict__ used, float *__restrict__ normals) {
    const int f = blockIdx.y;
    const int mp = blockIdx.x * blockDim.x + threadIdx.x;
    if (mp >= n_molpad) return;
    if (!used[(size_t)f * n_molpad + mp]) {
        float *nx = normals + ((size_t)f * 3) * n_molpad + mp;
        nx[0] = CUDART_NAN_F; nx[n_molpad] = CUDART_NAN_F; nx[2 * (size_t)n_molpad] = CUDART_NAN_F;
    }
}

// ---------------------------------------------------------------------------------------------
// accumulation helpers shared by K1 and K2
// ---------------------------------------------------------------------------------------------
struct AccumOut {
    // Global leaflets assigned on every frame: the accumulation kernel classifies its own molecules
    // from the head coordinate and the frame's membrane centre (common_identify_leaflet,
    // leaflets.rs:711-732) instead of reading a table; it writes the table rows only for export.
    const float *inline_center;     // [F][3] centre of frame f, or nullptr (table path)
    unsigned char *leaf_out;        // rows [(1 + f)][n_molpad] when the tables are collected, else nullptr
    long long *bsum;                // [rows][n_slots][3]
    unsigned long long *bcnt;       // [rows][n_slots][3]
    long long *map_sum;             // [n_slots][3][n_bins]
    unsigned long long *map_cnt;
    unsigned char *normal_used;     // [F][n_molpad] or nullptr
    // speculative Global leaflets (SPEC): classify against the provisional centre *spec_ref while the
    // kernel sums the membrane's displacements from it; the last CTA of a frame derives the frame's true
    // centre and flags the frame for spec_repair_kernel when a head could change sides.
    const float *spec_ref;          // [1] provisional centre along the leaflet axis
    float *spec_ref_next;           // [1] provisional centre of the next batch (centre of this batch's last frame)
    double *spec_sum;               // [F][n_chunks][2] sum d, sum d^2 (d = min-image displacement from *spec_ref)
    float *spec_mm;                 // [F][n_chunks][4] max |d|, min |d_head|, max |d_head|, NaN seen
    unsigned *spec_ticket;          // [F] CTAs of the frame that have published their partials (self-resetting)
    const double *spec_left_sum;    // [F][spec_left_parts][2] the same sums over the membrane atoms no bond loads
    const float *spec_left_mm;      // [F][spec_left_parts][2] max |d|, NaN seen            (spec_leftover_kernel)
    int spec_left_parts;
    float *spec_center;             // [F] centre derived from the partials
    unsigned char *spec_flag;       // [F] 1 = the frame needs the exact centre (repair)
    unsigned *spec_nflag;           // [1] flagged frames of this batch
    int n_membrane;
};

// number of int accumulators per order slot in shared memory
template <bool LEAF, bool EXTRA> struct AccLayout { static constexpr int N = LEAF ? (EXTRA ? 3 : 2) : (EXTRA ? 2 : 1); };
//  !LEAF: [sum_total, (cnt_total)]      LEAF: [sum_up, sum_lo, (cnt_up | cnt_lo << 16)]

template <bool LEAF, bool EXTRA>
__device__ __forceinline__ void warp_commit(int *s_acc, int lane, int su, int sl, int cu, int cl) {
    // fixed-order hardware tree (REDUX) -> deterministic; one plain shared store per value
    const unsigned full = 0xffffffffu;
    if (LEAF) {
        int a = __reduce_add_sync(full, su), b = __reduce_add_sync(full, sl);
        if (EXTRA) {
            int c = __reduce_add_sync(full, cu | (cl << 16));
            if (lane == 0) { s_acc[0] = a; s_acc[1] = b; s_acc[2] = c; }
        } else if (lane == 0) { s_acc[0] = a; s_acc[1] = b; }
    } else {
        int a = __reduce_add_sync(full, su + sl);
        if (EXTRA) {
            int c = __reduce_add_sync(full, cu + cl);
            if (lane == 0) { s_acc[0] = a; s_acc[1] = c; }
        } else if (lane == 0) s_acc[0] = a;
    }
}

// CTA epilogue: add the per-warp partials of every order slot to this frame's accumulators.
template <bool LEAF, bool EXTRA, int WARPS = kWarps>
__device__ __forceinline__ void cta_flush(const DeviceView &v, const AccumOut &o, const int *s_acc, int n_orders, int slot0, int tw_row,
                                          int cnt_total, int cnt_up) {
    constexpr int NA = AccLayout<LEAF, EXTRA>::N;
    for (int i = threadIdx.x; i < n_orders; i += blockDim.x) {
        long long acc[NA];
#pragma unroll
        for (int k = 0; k < NA; k++) acc[k] = 0;
        int c_up = 0, c_lo = 0;
        for (int w = 0; w < WARPS; w++) {
            const int *p = s_acc + ((size_t)w * n_orders + i) * NA;
            if (LEAF) {
                acc[0] += p[0]; acc[1] += p[1];
                if (EXTRA) { c_up += p[2] & 0xffff; c_lo += (p[2] >> 16) & 0xffff; }
            } else {
                acc[0] += p[0];
                if (EXTRA) c_up += p[1];
            }
        }
        const size_t base = ((size_t)tw_row * v.n_slots + slot0 + i) * 3;
        if (LEAF) {
            if (!EXTRA) { c_up = cnt_up; c_lo = cnt_total - cnt_up; }
            if (c_up) { atomicAdd((unsigned long long *)&o.bsum[base + GORDER_ACC_UPPER], (unsigned long long)acc[0]); atomicAdd(&o.bcnt[base + GORDER_ACC_UPPER], (unsigned long long)c_up); }
            if (c_lo) { atomicAdd((unsigned long long *)&o.bsum[base + GORDER_ACC_LOWER], (unsigned long long)acc[1]); atomicAdd(&o.bcnt[base + GORDER_ACC_LOWER], (unsigned long long)c_lo); }
        } else {
            if (!EXTRA) c_up = cnt_total;
            if (c_up) { atomicAdd((unsigned long long *)&o.bsum[base + GORDER_TOTAL], (unsigned long long)acc[0]); atomicAdd(&o.bcnt[base + GORDER_TOTAL], (unsigned long long)c_up); }
        }
    }
}

template <bool LEAF>
__device__ __forceinline__ void map_add(const DeviceView &v, const AccumOut &o, int slot, const f3 &pos, int q, bool upper) {
    long long b = map_bin(v.map, pos);
    if (b < 0) return;
    // with leaflets the total map is upper + lower (derived when results are fetched)
    const int which = LEAF ? (upper ? GORDER_ACC_UPPER : GORDER_ACC_LOWER) : GORDER_TOTAL;
    const size_t i = ((size_t)slot * 3 + which) * v.map.n_bins + b;
    atomicAdd((unsigned long long *)&o.map_sum[i], (unsigned long long)(long long)q);
    atomicAdd(&o.map_cnt[i], 1ull);
}

// ---------------------------------------------------------------------------------------------
// SPEC epilogue (thread 0 of a CTA): publish the CTA's partial sums of the membrane's displacements
// from the provisional centre `sref`; the last CTA of the frame adds them in chunk order
// (deterministic) and decides whether the speculative leaflets are provably the exact ones.
//
// group_get_center = wrap(est + mean(min_image(z - est))), est = circular mean of the membrane.  With
// d_i = min_image(z_i - sref) this equals wrap(sref + mean d) (up to f32 rounding) as long as every atom
// keeps its periodic image when seen from est instead of sref:  |est - sref| < L/2 - max|d_i|.
// est is not computed; it is bounded from the moments: with t_i = 2 pi d_i / L, T = max|t_i|,
//     sum cos t_i >= N - S2/2,  |sum sin t_i| <= |S1| + T S2 / 6     (S1 = sum t_i, S2 = sum t_i^2)
// so |est - sref| <= atan2(|S1| + T S2/6, N - S2/2) L / 2 pi whenever N - S2/2 > 0.
// The leaflets are those of the exact centre if no head lies within |delta| (+ margin) of sref or of the
// far cut sref + L/2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void spec_publish(const AccumOut &o, int f, float sref, float L, float half, double ds, double dq, float dabs,
                                             float hmin, float hmax, bool bad) {
    const size_t pi = (size_t)f * gridDim.x + blockIdx.x;
    o.spec_sum[2 * pi] = ds; o.spec_sum[2 * pi + 1] = dq;
    reinterpret_cast<float4 *>(o.spec_mm)[pi] = make_float4(dabs, hmin, hmax, bad ? 1.0f : 0.0f);
    __threadfence();
    if (atomicAdd(&o.spec_ticket[f], 1u) != gridDim.x - 1) return;
    __threadfence();
    o.spec_ticket[f] = 0;
    double tot = 0.0, tot2 = 0.0;
    float dmax = 0.0f, hmn = CUDART_INF_F, hmx = 0.0f;
    bool nan = false;
    for (unsigned c = 0; c < gridDim.x; c++) {
        const size_t qi = (size_t)f * gridDim.x + c;
        tot += __ldcg(o.spec_sum + 2 * qi); tot2 += __ldcg(o.spec_sum + 2 * qi + 1);
        const float4 m4 = __ldcg(reinterpret_cast<const float4 *>(o.spec_mm) + qi);
        nan = nan || m4.w != 0.0f;
        dmax = fmaxf(dmax, m4.x); hmn = fminf(hmn, m4.y); hmx = fmaxf(hmx, m4.z);
    }
    for (int c = 0; c < o.spec_left_parts; c++) {   // written by spec_leftover_kernel, earlier in the stream
        const size_t qi = (size_t)f * o.spec_left_parts + c;
        tot += __ldcg(o.spec_left_sum + 2 * qi); tot2 += __ldcg(o.spec_left_sum + 2 * qi + 1);
        const float2 m2 = __ldcg(reinterpret_cast<const float2 *>(o.spec_left_mm) + qi);
        nan = nan || m2.y != 0.0f || m2.x != m2.x;
        dmax = fmaxf(dmax, m2.x);
    }
    const double n = (double)o.n_membrane;
    const float delta = (float)(tot / n);
    const float margin = 1e-4f + 1e-3f * L;
    const double ts = 6.283185307179586 / (double)L;
    const double s1 = fabs(tot) * ts, s2 = tot2 * ts * ts, x_lo = n - 0.5 * s2, y_hi = s1 + (double)dmax * ts * s2 / 6.0;
    const float est_bound = x_lo > 0.0 ? (float)(atan2(y_hi, x_lo) / ts) : CUDART_INF_F;
    const bool ok = !nan && o.n_membrane > 0 && L > 0.0f && (delta == delta) && est_bound + margin < half - dmax &&
                    fabsf(delta) + margin < hmn && hmx + fabsf(delta) + margin < half;
    const float c = wrap1(__fadd_rn(sref, delta), L);
    o.spec_center[f] = ok ? c : CUDART_NAN_F;
    o.spec_flag[f] = ok ? 0 : 1;
    if (!ok) atomicAdd(o.spec_nflag, 1u);
    if (f == (int)gridDim.y - 1) *o.spec_ref_next = ok ? c : sref;
}

// SPEC side pass, grid (<= kSpecLeftBlocks, F): the membrane atoms that no bond of the bond kernel loads (AA: the
// head-group atoms outside the analysed bonds, lipids without analysed bonds).  One read of their leaflet-axis
// component gives the sums of spec_add over them; spec_publish adds these partials to the bond kernel's.
constexpr int kSpecLeftBlocks = 32;
__global__ void __launch_bounds__(256) spec_leftover_kernel(DeviceView v, const Seg *__restrict__ segs, int n_segs, const float *__restrict__ planes,
                                                            const FrameAux *__restrict__ aux, const float *__restrict__ ref,
                                                            double *__restrict__ left_sum, float *__restrict__ left_mm) {
    const int f = blockIdx.y;
    const float *fr = planes + (size_t)f * v.frame_floats;
    const float L = aux[f].L[v.leaflet_axis];
    const float sref = __ldg(ref);
    const float invL = L > 0.0f ? __frcp_rn(L) : 0.0f;
    double ds = 0.0, dq = 0.0;
    float dabs = 0.0f;
    for (int sg = blockIdx.x; sg < n_segs; sg += gridDim.x) {
        const Seg sgm = segs[sg];
        float a = 0.0f, q = 0.0f;   // <= 16 terms per thread and run
        for (int i = threadIdx.x; i < sgm.len; i += blockDim.x) {
            const float t = __ldg(fr + sgm.off + i) - sref;
            const float k = __fadd_rn(fmaf(t, invL, 12582912.0f), -12582912.0f);
            const float d = fmaf(-L, k, t);
            a += d; q = fmaf(d, d, q); dabs = fmaxf(dabs, fabsf(d));
        }
        ds += (double)a; dq += (double)q;
    }
    float bad = (ds != ds) ? 1.0f : 0.0f;   // NaN / Inf coordinate
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, ofs); dq += __shfl_xor_sync(0xffffffffu, dq, ofs);
        dabs = fmaxf(dabs, __shfl_xor_sync(0xffffffffu, dabs, ofs)); bad = fmaxf(bad, __shfl_xor_sync(0xffffffffu, bad, ofs));
    }
    __shared__ double s_d[2][8];
    __shared__ float s_m[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_d[0][warp] = ds; s_d[1][warp] = dq; s_m[0][warp] = dabs; s_m[1][warp] = bad; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    ds = 0.0; dq = 0.0; dabs = 0.0f; bad = 0.0f;
    for (int w = 0; w < 8; w++) { ds += s_d[0][w]; dq += s_d[1][w]; dabs = fmaxf(dabs, s_m[0][w]); bad = fmaxf(bad, s_m[1][w]); }
    const size_t pi = (size_t)f * gridDim.x + blockIdx.x;
    left_sum[2 * pi] = ds; left_sum[2 * pi + 1] = dq;
    left_mm[2 * pi] = dabs; left_mm[2 * pi + 1] = bad;
}

// ---------------------------------------------------------------------------------------------
// K1: bond engine (topology/bond.rs:396-446 + :184-215).
//   grid (n_chunks, F); a lane owns MPT consecutive molecules of one molecule type and walks the
//   type's bond table (staged in shared memory); every plane read is a fully-used 128 B * MPT line.
// template: MPT molecules per thread; PBC; NVEC per-molecule normal vector (dynamic / manual)
//           instead of a static axis; LEAF per-leaflet accumulation; EXTRA geometry filter / maps.
// ---------------------------------------------------------------------------------------------
//           SPEC speculative Global leaflets: no centre pre-pass (see AccumOut::spec_*).
template <int MPT, bool PBC, bool NVEC, bool LEAF, bool EXTRA, bool SPEC = false, int BLOCK = kBlock>
__device__ __forceinline__ void bond_order_body(const DeviceView &v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                const unsigned char *__restrict__ leaf_rows, const float *__restrict__ normals,
                                                const int *__restrict__ normal_npoints, const AccumOut &o) {
    constexpr int NA = AccLayout<LEAF, EXTRA>::N;
    constexpr int WARPS = BLOCK / 32;   // the CTA has BLOCK threads (256, or fewer when K1f hands over the partial tile of a small system)
    // Static normal without geometry / maps: S depends only on (d_axis, |d|^2), so the kernel reads the
    // components in the order (axis+1, axis+2, axis) and never selects a component at run time.
    constexpr bool PERMUTE = !NVEC && !EXTRA;
    extern __shared__ int smem[];
    const Chunk ch = v.chunks[blockIdx.x];
    const TypeDesc td = v.types[ch.type];
    const int f = blockIdx.y;
    const FrameAux &ax = aux[f];
    const int nb = td.n_items;
    BondItem *s_bonds = reinterpret_cast<BondItem *>(smem);
    int *s_acc = smem + 2 * nb;                 // [WARPS][nb][NA]
    int *s_cnt = s_acc + WARPS * nb * NA;      // [2] valid, valid & upper
    for (int i = threadIdx.x; i < nb; i += BLOCK) s_bonds[i] = v.bonds[td.item_off + i];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = ch.first_mol + threadIdx.x * MPT;
    const bool active = m0 < td.mpad;
    const int mpad = td.cstride;   // component stride inside the CTA's tile
    const int c0 = PERMUTE ? (v.normal_axis + 1) % 3 : 0, c1 = PERMUTE ? (v.normal_axis + 2) % 3 : 1, c2 = PERMUTE ? v.normal_axis : 2;
    const int o0 = c0 * mpad, o1 = c1 * mpad, o2 = c2 * mpad;
    // box in the kernel's component order, with the fast-path guard of the fold
    const float L0 = ax.L[c0], L1 = ax.L[c1], L2 = ax.L[c2];
    const float h0 = ax.half[c0], h1 = ax.half[c1], h2 = ax.half[c2];
    const float g0 = 0.99f * h0, g1 = 0.99f * h1, g2 = 0.99f * h2;
    const Box bx = load_box(ax);
    const float *base_mol = planes + (size_t)f * v.frame_floats + mol_offset(td, m0);
    bool valid[MPT];
    int upmask[MPT];
    f3 nrm[MPT];
    int nvalid = 0, nup = 0;
    // SPEC: provisional centre, |head - centre| extremes of this thread's molecules
    const float sref = SPEC ? __ldg(o.spec_ref) : 0.0f;
    float hmin = CUDART_INF_F, hmax = 0.0f;
    bool hnan = false;
#pragma unroll
    for (int j = 0; j < MPT; j++) {
        valid[j] = active && (m0 + j < td.n_mol);
        bool up = false;
        if (LEAF && valid[j]) {
            if (SPEC) {
                const int la = v.leaflet_axis;
                const float hd = __ldg(base_mol + td.head_off + la * mpad + j);
                const float dh = distance_1d(hd, sref, ax.L[la], ax.half[la], PBC);
                up = dh >= 0.0f;
                hmin = fminf(hmin, fabsf(dh)); hmax = fmaxf(hmax, fabsf(dh));
                hnan = hnan || dh != dh;   // NaN head: the frame is flagged and the exact path reports it
                if (v.leaflet_flip) up = !up;
                if (o.leaf_out) o.leaf_out[(size_t)(1 + f) * v.n_molpad + td.molpad0 + m0 + j] = up ? GORDER_UPPER : GORDER_LOWER;
            } else if (o.inline_center) {
                const int la = v.leaflet_axis;
                const float c = o.inline_center[3 * f + la];
                if (c != c) raise_error(v, GORDER_ERR_INVALID_GLOBAL_CENTER, ax.frame_index);
                const float hd = __ldg(base_mol + td.head_off + la * mpad + j);
                up = distance_1d(hd, c, ax.L[la], ax.half[la], PBC) >= 0.0f;
                if (v.leaflet_flip) up = !up;
                if (o.leaf_out) o.leaf_out[(size_t)(1 + f) * v.n_molpad + td.molpad0 + m0 + j] = up ? GORDER_UPPER : GORDER_LOWER;
            } else up = leaf_rows[(size_t)ax.leaf_row * v.n_molpad + td.molpad0 + m0 + j] == GORDER_UPPER;
        }
        upmask[j] = up ? -1 : 0;
        if (NVEC) {
            nrm[j] = mk3(0.f, 0.f, 1.f);
            if (valid[j]) {
                const float *np = normals + ((size_t)f * 3) * v.n_molpad + td.molpad0 + m0 + j;
                nrm[j] = mk3(np[0], np[v.n_molpad], np[2 * (size_t)v.n_molpad]);
            }
        }
        nvalid += valid[j]; nup += valid[j] && up;
    }
    if (!EXTRA) {
        int a = __reduce_add_sync(0xffffffffu, nvalid), b = __reduce_add_sync(0xffffffffu, nup);
        if (lane == 0) { atomicAdd(&s_cnt[0], a); atomicAdd(&s_cnt[1], b); }
    }
    bool any_used[MPT];
#pragma unroll
    for (int j = 0; j < MPT; j++) any_used[j] = false;
    float nan_acc = 0.0f;   // NaN coordinates poison this accumulator (checked once after the loop)
    // SPEC: head extremes of the CTA -> shared (frees the registers for the loop)
    __shared__ float s_hmm[2][WARPS];
    __shared__ double s_dsum[2][WARPS];
    __shared__ float s_dabs[WARPS];
    if (SPEC) {
        float a = hnan ? CUDART_NAN_F : hmin, b = hmax;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            const float a2 = __shfl_xor_sync(0xffffffffu, a, ofs), b2 = __shfl_xor_sync(0xffffffffu, b, ofs);
            a = (a != a || a2 != a2) ? CUDART_NAN_F : fminf(a, a2); b = fmaxf(b, b2);
        }
        if (lane == 0) { s_hmm[0][warp] = a; s_hmm[1][warp] = b; }
    }
    // SPEC: displacement of the membrane atoms from the provisional centre, d = t - L rint(t / L)
    // (any periodic image), summed per thread in f32 (<= 2 MPT n_items terms), across threads in f64.
    const float sp_invL = (SPEC && L2 > 0.0f) ? __frcp_rn(L2) : 0.0f;
    float dsum = 0.0f, dsq = 0.0f, dabs = 0.0f;   // sum d, sum d^2, max |d|
    auto spec_add = [&](float z, bool ok) {
        const float t = z - sref;
        const float k = __fadd_rn(fmaf(t, sp_invL, 12582912.0f), -12582912.0f);
        float d = fmaf(-L2, k, t);
        d = ok ? d : 0.0f;
        dsum += d; dsq = fmaf(d, d, dsq); dabs = fmaxf(dabs, fabsf(d));
    };

    const float *base = base_mol;
    Vec<MPT> x1, y1, z1, x2, y2, z2;
#pragma unroll
    for (int j = 0; j < MPT; j++) x1.v[j] = y1.v[j] = z1.v[j] = x2.v[j] = y2.v[j] = z2.v[j] = 0.0f;
    // As in K1f: one lane per CTA asks L2 for the CTA's slices of the planes the bonds two iterations ahead will read (the
    // variants with geometry / maps / per-molecule normals stall on these loads: long scoreboard led their ncu profiles)
    constexpr int kAhead = 2;
    const float *tile0 = planes + (size_t)f * v.frame_floats + mol_offset(td, ch.first_mol);
    const unsigned slice_bytes = (unsigned)(BLOCK * MPT * sizeof(float));
    auto prefetch_bond = [&](int bb) {
        if (threadIdx.x != 0 || bb >= nb) return;
        const BondItem it = s_bonds[bb];
        if ((it.a_off & 3) == 0) {
            const float *pa = tile0 + (it.a_off & ~15);
            l2_prefetch_run(pa, slice_bytes); l2_prefetch_run(pa + mpad, slice_bytes); l2_prefetch_run(pa + 2 * mpad, slice_bytes);
        }
        const float *pb = tile0 + it.b_off;
        l2_prefetch_run(pb, slice_bytes); l2_prefetch_run(pb + mpad, slice_bytes); l2_prefetch_run(pb + 2 * mpad, slice_bytes);
    };
    if (EXTRA || NVEC) for (int bb = 0; bb < kAhead; bb++) prefetch_bond(bb);
    for (int b = 0; b < nb; b++) {
        if (EXTRA || NVEC) prefetch_bond(b + kAhead);
        {
            const BondItem bi = s_bonds[b];
            // low bits of a_off: 1 = first atom is the previous bond's first atom, 2 = ... second atom;
            // 4 / 8 = first / second atom is a membrane atom seen here for the first time (speculative centre)
            const int reuse = bi.a_off & 3, a_off = bi.a_off & ~15;
            if (reuse == 2) { x1 = x2; y1 = y2; z1 = z2; }
            if (active) {
                if (reuse == 0) { x1.load(base + a_off + o0); y1.load(base + a_off + o1); z1.load(base + a_off + o2); }
                x2.load(base + bi.b_off + o0); y2.load(base + bi.b_off + o1); z2.load(base + bi.b_off + o2);
            }
        }
        if (SPEC) {
            const int cf = s_bonds[b].a_off;
            if (cf & 4) {
#pragma unroll
                for (int j = 0; j < MPT; j++) spec_add(z1.v[j], valid[j]);
            }
            if (cf & 8) {
#pragma unroll
                for (int j = 0; j < MPT; j++) spec_add(z2.v[j], valid[j]);
            }
        }
        int st = 0, su = 0, ct = 0, cu = 0;   // total / upper (lower = total - upper)
        // bond vectors of the thread's MPT molecules.  The fold's exact fast path
        //   fl(fl(fl(fl(d + L/2) + L) - L) - L/2)
        // is evaluated unconditionally; ONE predicate per iteration sends the (rare) warp that holds a bond
        // outside the guard |d| <= 0.99 L/2 to the literal, out-of-line expression.
        f3 dv[MPT];
        bool slow = false;
#pragma unroll
        for (int j = 0; j < MPT; j++) {
            const float rx = __fsub_rn(x2.v[j], x1.v[j]), ry = __fsub_rn(y2.v[j], y1.v[j]), rz = __fsub_rn(z2.v[j], z1.v[j]);
            if (PBC) {
                slow = slow | (valid[j] & ((fabsf(rx) > g0) | (fabsf(ry) > g1) | (fabsf(rz) > g2)));
                dv[j] = mk3(__fsub_rn(__fsub_rn(__fadd_rn(__fadd_rn(rx, h0), L0), L0), h0),
                            __fsub_rn(__fsub_rn(__fadd_rn(__fadd_rn(ry, h1), L1), L1), h1),
                            __fsub_rn(__fsub_rn(__fadd_rn(__fadd_rn(rz, h2), L2), L2), h2));
            } else dv[j] = mk3(rx, ry, rz);
        }
        if (PBC && slow) {
#pragma unroll
            for (int j = 0; j < MPT; j++) {
                const float rx = __fsub_rn(x2.v[j], x1.v[j]), ry = __fsub_rn(y2.v[j], y1.v[j]), rz = __fsub_rn(z2.v[j], z1.v[j]);
                if (fabsf(rx) > g0) dv[j].x = min_image_slow(rx, L0, h0);
                if (fabsf(ry) > g1) dv[j].y = min_image_slow(ry, L1, h1);
                if (fabsf(rz) > g2) dv[j].z = min_image_slow(rz, L2, h2);
            }
        }
#pragma unroll
        for (int j = 0; j < MPT; j++) {
            const f3 d = dv[j];
            bool use = valid[j];
            f3 mid;
            if (EXTRA) {
                // bond_pos = pos1 + vec / 2 (bond.rs:422)
                mid = mk3(__fadd_rn(x1.v[j], d.x * 0.5f), __fadd_rn(y1.v[j], d.y * 0.5f), __fadd_rn(z1.v[j], d.z * 0.5f));
                if (use && v.shape.kind != GORDER_GEOM_NONE) use = shape_inside<PBC>(v.shape, ax, bx, mid);
                any_used[j] = any_used[j] || use;
            }
            float s;
            if (NVEC) {
                if (use && nrm[j].x != nrm[j].x) {   // normal could not be computed (normal.rs:424)
                    int npts = normal_npoints ? normal_npoints[(size_t)f * v.n_molpad + td.molpad0 + m0 + j] : 0;
                    raise_error(v, v.normal_mode == GORDER_NORMAL_DYNAMIC ? GORDER_ERR_DYNAMIC_NORMAL_POINTS : GORDER_ERR_MANUAL_NORMAL_FRAME, npts);
                    use = false;
                }
                s = calc_sch_fast(d, nrm[j]);
            } else s = calc_sch_axis_fast(d, PERMUTE ? d.z : comp(d, v.normal_axis));
            s = use ? s : 0.0f;
            nan_acc = fmaf(use ? (d.x + d.y + d.z) : 0.0f, 0.0f, nan_acc);   // NaN / Inf coordinates poison the accumulator
            const int q = order_value_fast(s);
            st += q; su += q & upmask[j];
            if (EXTRA) {
                ct += use; cu += use & (upmask[j] & 1);
                if (use && v.map.enabled) map_add<LEAF>(v, o, td.slot0 + b, mid, q, upmask[j] != 0);
            }
        }
        if (LEAF) warp_commit<LEAF, EXTRA>(s_acc + ((size_t)warp * nb + b) * NA, lane, su, st - su, cu, ct - cu);
        else warp_commit<LEAF, EXTRA>(s_acc + ((size_t)warp * nb + b) * NA, lane, st, 0, ct, 0);
    }
    if (nan_acc != nan_acc)   // AnalysisError::UndefinedPosition: a NaN coordinate reached the engine
        raise_error(v, GORDER_ERR_UNDEFINED_POSITION, ((long long)ch.type << 48) | (unsigned)m0);
    if (EXTRA && o.normal_used) {
#pragma unroll
        for (int j = 0; j < MPT; j++)
            if (valid[j] && any_used[j]) o.normal_used[(size_t)f * v.n_molpad + td.molpad0 + m0 + j] = 1;
    }
    if (SPEC) {
        double ds = (double)dsum, dq = (double)dsq;
        float a = dabs;
        if (dsum != dsum) a = dsum;   // NaN / Inf coordinate: poison the extent so that the frame is flagged
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            ds += __shfl_xor_sync(0xffffffffu, ds, ofs); dq += __shfl_xor_sync(0xffffffffu, dq, ofs);
            const float a2 = __shfl_xor_sync(0xffffffffu, a, ofs);
            a = (a != a || a2 != a2) ? CUDART_NAN_F : fmaxf(a, a2);
        }
        if (lane == 0) { s_dsum[0][warp] = ds; s_dsum[1][warp] = dq; s_dabs[warp] = a; }
    }
    __syncthreads();
    cta_flush<LEAF, EXTRA, WARPS>(v, o, s_acc, nb, td.slot0, ax.tw_row, s_cnt[0], s_cnt[1]);
    if (SPEC && threadIdx.x == 0) {
        double ds = 0.0, dq = 0.0;
        float p0 = 0.0f, p1 = CUDART_INF_F, p2 = 0.0f;
        bool bad = false;
        for (int w = 0; w < WARPS; w++) {
            ds += s_dsum[0][w]; dq += s_dsum[1][w];
            bad = bad || s_dabs[w] != s_dabs[w] || s_hmm[0][w] != s_hmm[0][w];
            p0 = fmaxf(p0, s_dabs[w]); p1 = fminf(p1, s_hmm[0][w]); p2 = fmaxf(p2, s_hmm[1][w]);
        }
        spec_publish(o, f, sref, L2, h2, ds, dq, p0, p1, p2, bad);
    }
}

// Resident CTAs per SM the variants are compiled for (measured on B200, round 2).  Geometry / maps: 4 (64 registers, a few
// spills) -- S-AA-large launches 512 CTAs, which 4 per SM hold in ONE wave (0.77 -> 0.55 ms per 128 frames; at 2 or 3 per SM
// the second wave is a tail).  Per-molecule normals without geometry: 3 (0.056 -> 0.046 ms per 8 frames of S-CG).
#ifndef GORDER_EXTRA_MINB
#define GORDER_EXTRA_MINB 4
#endif
#ifndef GORDER_NVEC_MINB
#define GORDER_NVEC_MINB 3
#endif
template <int MPT, bool PBC, bool NVEC, bool LEAF, bool EXTRA, bool SPEC = false>
__global__ void __launch_bounds__(kBlock, EXTRA ? GORDER_EXTRA_MINB : (NVEC ? GORDER_NVEC_MINB : 4)) bond_order_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                            const unsigned char *__restrict__ leaf_rows, const float *__restrict__ normals,
                                                            const int *__restrict__ normal_npoints, AccumOut o) {
    bond_order_body<MPT, PBC, NVEC, LEAF, EXTRA, SPEC>(v, planes, aux, leaf_rows, normals, normal_npoints, o);
}

// ---------------------------------------------------------------------------------------------
// spec_repair_kernel: frames whose speculative leaflets are not provably those of the exact centre.
// grid (n_chunks, R); CTA (x, y) visits the flagged frames y, y + R, ...: it recomputes the exact centre
// with the arithmetic of center_axis_kernel (same partial sums in the same order; every CTA of the column
// does so redundantly -- this is the rare path), re-classifies the molecules of chunk x and moves the
// samples of every molecule that changes sides from one leaflet's accumulators to the other's.
// ---------------------------------------------------------------------------------------------
struct RepairParams {
    const Seg *segs;
    int n_segs, n_blocks, n_group;
    const float *ref;            // provisional centre the batch was classified with
    const unsigned char *flag;   // [F]
    const unsigned *nflag;       // [1]
    float *center;               // [F] exact centre of repaired frames
    unsigned *host_counters;     // mapped pinned: [0] frames speculated, [1] frames repaired
    unsigned char *leaf_out;
};

template <int MPT>
__global__ void __launch_bounds__(kBlock) spec_repair_kernel(DeviceView v, RepairParams rp, const float *__restrict__ planes,
                                                             const FrameAux *__restrict__ aux, AccumOut o, int n_frames) {
    const unsigned nflag = *rp.nflag;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        volatile unsigned *hc = rp.host_counters;
        hc[0] = hc[0] + (unsigned)n_frames; hc[1] = hc[1] + nflag;
        __threadfence_system();
    }
    if (nflag == 0) return;
    __shared__ long long s_red[3][8];
    __shared__ float s_c;
    const Chunk ch = v.chunks[blockIdx.x];
    const TypeDesc td = v.types[ch.type];
    const int axis = v.leaflet_axis;
    const float sref = *rp.ref;
    for (int f = blockIdx.y; f < n_frames; f += gridDim.y) {
        if (!rp.flag[f]) continue;   // block-uniform
        const FrameAux &ax = aux[f];
        const float L = ax.L[axis], half = ax.half[axis];
        const float *fr = planes + (size_t)f * v.frame_floats;
        // ---- exact centre (the sums of center_axis_kernel: integers, any partition gives the same bits) ----
        const float inv_l = __fdiv_rn(1.0f, L);
        long long T0 = 0, T1 = 0, t0, t1;
        bool bad = false, tb;
        for (int vb = 0; vb < rp.n_blocks; vb++) { center_block_sums(rp.segs, rp.n_segs, vb, rp.n_blocks, fr, true, 0, inv_l, 0.0f, L, half, s_red, t0, t1, tb); T0 += t0; T1 += t1; bad = bad || tb; }
        if (threadIdx.x == 0) s_c = (rp.n_group > 0 && !bad) ? center_estimate(T0, T1, L) : CUDART_NAN_F;
        __syncthreads();
        const float e = s_c;
        T0 = 0;
        for (int vb = 0; vb < rp.n_blocks; vb++) { center_block_sums(rp.segs, rp.n_segs, vb, rp.n_blocks, fr, true, 1, inv_l, e, L, half, s_red, t0, t1, tb); T0 += t0; bad = bad || tb; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const float c = __fadd_rn(e, center_mean(T0, rp.n_group));
            s_c = (rp.n_group > 0 && !bad && e == e) ? ((L > 0.0f) ? wrap1(c, L) : c) : CUDART_NAN_F;
        }
        __syncthreads();
        const float center = s_c;
        if (center != center) { if (threadIdx.x == 0) raise_error(v, GORDER_ERR_INVALID_GLOBAL_CENTER, ax.frame_index); continue; }
        if (blockIdx.x == 0 && threadIdx.x == 0) rp.center[f] = center;
        // ---- molecules of this chunk that change sides ----
        const int mpad = td.cstride;
        // the bond kernel's component order (PERMUTE) so that every sample rounds exactly as it did there
        const int c0 = (v.normal_axis + 1) % 3, c1 = (v.normal_axis + 2) % 3, c2 = v.normal_axis;
#pragma unroll 1
        for (int j = 0; j < MPT; j++) {
            const int m = ch.first_mol + threadIdx.x * MPT + j;
            if (m >= td.n_mol) continue;
            const float *base = fr + mol_offset(td, m);
            const float hd = __ldg(base + td.head_off + axis * mpad);
            bool up_spec = distance_1d(hd, sref, L, half, true) >= 0.0f, up = distance_1d(hd, center, L, half, true) >= 0.0f;
            if (v.leaflet_flip) { up_spec = !up_spec; up = !up; }
            if (rp.leaf_out) rp.leaf_out[(size_t)(1 + f) * v.n_molpad + td.molpad0 + m] = up ? GORDER_UPPER : GORDER_LOWER;
            if (up == up_spec) continue;
            const int to = up ? GORDER_ACC_UPPER : GORDER_ACC_LOWER, from = up ? GORDER_ACC_LOWER : GORDER_ACC_UPPER;
            for (int b = 0; b < td.n_items; b++) {
                const BondItem bi = v.bonds[td.item_off + b];
                const int a_off = bi.a_off & ~15;
                const float *pa = base + a_off, *pb = base + bi.b_off;
                const f3 d = mk3(min_image(__fsub_rn(__ldg(pb + c0 * mpad), __ldg(pa + c0 * mpad)), ax.L[c0], ax.half[c0]),
                                 min_image(__fsub_rn(__ldg(pb + c1 * mpad), __ldg(pa + c1 * mpad)), ax.L[c1], ax.half[c1]),
                                 min_image(__fsub_rn(__ldg(pb + c2 * mpad), __ldg(pa + c2 * mpad)), ax.L[c2], ax.half[c2]));
                const long long q = order_value_fast(calc_sch_axis_fast(d, d.z));
                const size_t rb = ((size_t)ax.tw_row * v.n_slots + td.slot0 + b) * 3;
                atomicAdd((unsigned long long *)&o.bsum[rb + to], (unsigned long long)q);
                atomicAdd((unsigned long long *)&o.bsum[rb + from], (unsigned long long)(-q));
                atomicAdd(&o.bcnt[rb + to], 1ull);
                atomicAdd(&o.bcnt[rb + from], ~0ull);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: united-atom engine (uaorder.rs:400-437): rebuild 1-3 hydrogens per carbon from 2-3 helper
// heavy atoms, then the same S / accumulate tail as K1.  One molecule per lane (the hydrogen
// construction is register-heavy), carbon-type table in shared memory.
// ---------------------------------------------------------------------------------------------
template <bool PBC>
__device__ __forceinline__ int predict_hydrogens(const DeviceView &v, int kind, const f3 &t, const f3 &h1, const f3 &h2, const f3 &h3,
                                                 const Box &bx, f3 (&out)[3]) {
    const float BOND_LENGTH = 0.109f;   // uaorder.rs:39
    if (kind == GORDER_UA_CH3) {   // predict_hydrogens_ch3, uaorder.rs:947-981
        f3 th1 = vector_to<PBC>(t, h1, bx), th2 = vector_to<PBC>(t, h2, bx);
        f3 axis = unit_ref(cross_ref(th2, th1));
        f3 hv1 = rotate_axis(th1, axis, v.tet_s, v.tet_c);
        out[0] = wrap_point<PBC>(shift_ref(t, hv1, BOND_LENGTH), bx);
        f3 nth1 = unit_ref(th1);
        out[1] = wrap_point<PBC>(shift_ref(t, rotate_axis(hv1, nth1, v.ch3_s, v.ch3_c), BOND_LENGTH), bx);
        out[2] = wrap_point<PBC>(shift_ref(t, rotate_axis(hv1, nth1, -v.ch3_s, v.ch3_c), BOND_LENGTH), bx);
        return 3;
    } else if (kind == GORDER_UA_CH2) {   // predict_hydrogens_ch2, uaorder.rs:985-1020
        f3 th1 = unit_ref(vector_to<PBC>(t, h1, bx)), th2 = unit_ref(vector_to<PBC>(t, h2, bx));
        f3 plane_normal = cross_ref(th2, th1);
        f3 rot_axis = unit_ref(mk3(__fsub_rn(th1.x, th2.x), __fsub_rn(th1.y, th2.y), __fsub_rn(th1.z, th2.z)));
        f3 rot_vec = cross_ref(plane_normal, rot_axis);
        f3 ax = unit_ref(rot_axis);
        out[0] = wrap_point<PBC>(shift_ref(t, rotate_axis(rot_vec, ax, v.tet_half_s, v.tet_half_c), BOND_LENGTH), bx);
        out[1] = wrap_point<PBC>(shift_ref(t, rotate_axis(rot_vec, ax, -v.tet_half_s, v.tet_half_c), BOND_LENGTH), bx);
        return 2;
    } else if (kind == GORDER_UA_CH1_UNSAT) {   // predict_hydrogen_ch1_unsaturated, uaorder.rs:1024-1045
        f3 th1 = vector_to<PBC>(t, h1, bx), th2 = vector_to<PBC>(t, h2, bx);
        // gamma = angle(th1, th2); rotation by pi - gamma/2:  sin = cos(gamma/2), cos = -sin(gamma/2)... evaluated
        // with the half-angle identities from cos(gamma) (no acos on the device)
        float n1 = norm_ref(th1), n2 = norm_ref(th2);
        float cg = __fdiv_rn(dot_ref(th1, th2), __fmul_rn(n1, n2));
        cg = fminf(1.0f, fmaxf(-1.0f, cg));
        if (n1 == 0.0f || n2 == 0.0f) cg = 1.0f;
        float ch = sqrtf(fmaxf(0.0f, 0.5f * (1.0f + cg)));   // cos(gamma/2)
        float sh = sqrtf(fmaxf(0.0f, 0.5f * (1.0f - cg)));   // sin(gamma/2)
        f3 axis = unit_ref(cross_ref(th1, th2));
        // sin(pi - g/2) = sin(g/2), cos(pi - g/2) = -cos(g/2)
        f3 hv = rotate_axis(th2, axis, sh, -ch);
        out[0] = wrap_point<PBC>(shift_ref(t, hv, BOND_LENGTH), bx);
        return 1;
    } else {   // predict_hydrogen_ch1_saturated, uaorder.rs:1087-1104
        f3 a = unit_ref(vector_to<PBC>(t, h1, bx)), b = unit_ref(vector_to<PBC>(t, h2, bx)), c = unit_ref(vector_to<PBC>(t, h3, bx));
        f3 s = mk3(__fadd_rn(__fadd_rn(a.x, b.x), c.x), __fadd_rn(__fadd_rn(a.y, b.y), c.y), __fadd_rn(__fadd_rn(a.z, b.z), c.z));
        out[0] = wrap_point<PBC>(shift_ref(t, mk3(-s.x, -s.y, -s.z), BOND_LENGTH), bx);
        return 1;
    }
}

// Streaming variant of the hydrogen construction (no geometry / maps): the same constructions with every
// normalisation done as x * rsqrt(|x|^2) (MUFU + one Newton step, <= 1 ulp) instead of IEEE sqrt + three IEEE
// divisions, rotations of vectors perpendicular to their axis as  v cos + (u x v) sin,  FMA contraction allowed.
// The directions agree with the exact path to ~2e-7; the hydrogen is still placed as fl(t + fl(0.109 u)) and the bond
// vector folded with the reference's expression, so that the dominant rounding (positions at box magnitude) is the
// reference's own.  Per sample |dS| <~ 1e-6 (zero mean); bit-exact hydrogens stay available (GORDER_UA_EXACT=1, and
// always with geometry selections / order maps, where a sample's position decides a count).
__device__ __forceinline__ float rsq_newton(float x) {
    const float r = rsqrt_ftz(x);
    return r * fmaf(-0.5f * x, r * r, 1.5f);
}
__device__ __forceinline__ f3 cross_f(const f3 &a, const f3 &b) {
    return mk3(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ float dot_f(const f3 &a, const f3 &b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ f3 scale_f(const f3 &a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 unit_f(const f3 &a) { return scale_f(a, rsq_newton(dot_f(a, a))); }
__device__ __forceinline__ f3 lin2(const f3 &a, float ca, const f3 &b, float cb) {
    return mk3(fmaf(a.x, ca, b.x * cb), fmaf(a.y, ca, b.y * cb), fmaf(a.z, ca, b.z * cb));
}

template <bool PBC>
__device__ __forceinline__ int predict_directions_fast(const DeviceView &v, int kind, const f3 &t, const f3 &h1, const f3 &h2, const f3 &h3,
                                                       const Box &bx, f3 (&u)[3]) {
    if (kind == GORDER_UA_CH2) {   // uaorder.rs:985-1020
        const f3 a = unit_f(vector_to<PBC>(t, h1, bx)), b = unit_f(vector_to<PBC>(t, h2, bx));
        const f3 pn = cross_f(b, a);
        const f3 ra = unit_f(mk3(a.x - b.x, a.y - b.y, a.z - b.z));
        const f3 rv = cross_f(pn, ra), w = cross_f(ra, rv);   // rv is perpendicular to ra: R v = v cos + (ra x v) sin
        const float inv = rsq_newton(dot_f(rv, rv));
        u[0] = lin2(rv, v.tet_half_c * inv, w, v.tet_half_s * inv);
        u[1] = lin2(rv, v.tet_half_c * inv, w, -v.tet_half_s * inv);
        return 2;
    } else if (kind == GORDER_UA_CH3) {   // uaorder.rs:947-981
        const f3 th1 = vector_to<PBC>(t, h1, bx), th2 = vector_to<PBC>(t, h2, bx);
        const f3 ax = unit_f(cross_f(th2, th1));
        const f3 hv1 = lin2(th1, v.tet_c, cross_f(ax, th1), v.tet_s);
        u[0] = unit_f(hv1);
        const f3 n = unit_f(th1), nxu = cross_f(n, u[0]);
        const float nd = dot_f(n, u[0]) * (1.0f - v.ch3_c);
        const f3 base = mk3(fmaf(u[0].x, v.ch3_c, n.x * nd), fmaf(u[0].y, v.ch3_c, n.y * nd), fmaf(u[0].z, v.ch3_c, n.z * nd));
        u[1] = mk3(fmaf(nxu.x, v.ch3_s, base.x), fmaf(nxu.y, v.ch3_s, base.y), fmaf(nxu.z, v.ch3_s, base.z));
        u[2] = mk3(fmaf(nxu.x, -v.ch3_s, base.x), fmaf(nxu.y, -v.ch3_s, base.y), fmaf(nxu.z, -v.ch3_s, base.z));
        return 3;
    } else if (kind == GORDER_UA_CH1_UNSAT) {   // uaorder.rs:1024-1045
        const f3 th1 = vector_to<PBC>(t, h1, bx), th2 = vector_to<PBC>(t, h2, bx);
        const float nn = dot_f(th1, th1) * dot_f(th2, th2);
        float cg = dot_f(th1, th2) * rsq_newton(nn);
        cg = fminf(1.0f, fmaxf(-1.0f, cg));
        if (nn == 0.0f) cg = 1.0f;
        const float ch = sqrtf(fmaxf(0.0f, 0.5f * (1.0f + cg))), sh = sqrtf(fmaxf(0.0f, 0.5f * (1.0f - cg)));
        const f3 ax = unit_f(cross_f(th1, th2));
        u[0] = unit_f(lin2(th2, -ch, cross_f(ax, th2), sh));   // rotation by pi - gamma/2 about ax (perpendicular to th2)
        return 1;
    } else {   // uaorder.rs:1087-1104
        const f3 a = unit_f(vector_to<PBC>(t, h1, bx)), b = unit_f(vector_to<PBC>(t, h2, bx)), c = unit_f(vector_to<PBC>(t, h3, bx));
        u[0] = unit_f(mk3(-(a.x + b.x + c.x), -(a.y + b.y + c.y), -(a.z + b.z + c.z)));
        return 1;
    }
}

template <bool PBC, bool NVEC, bool LEAF, bool EXTRA>
__global__ void __launch_bounds__(kBlock) ua_order_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                          const unsigned char *__restrict__ leaf_rows, const float *__restrict__ normals,
                                                          const int *__restrict__ normal_npoints, AccumOut o) {
    constexpr int NA = AccLayout<LEAF, EXTRA>::N;
    extern __shared__ int smem[];
    const Chunk ch = v.chunks[blockIdx.x];
    const TypeDesc td = v.types[ch.type];
    const int f = blockIdx.y;
    const FrameAux &ax = aux[f];
    const int ni = td.n_items, no = td.n_orders;
    UAItem *s_items = reinterpret_cast<UAItem *>(smem);
    int *s_acc = smem + 8 * ni;                 // [kWarps][no][NA]
    int *s_cnt = s_acc + kWarps * no * NA;
    for (int i = threadIdx.x; i < ni; i += kBlock) s_items[i] = v.ua[td.item_off + i];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = ch.first_mol + threadIdx.x;
    const bool valid = m < td.n_mol;
    const Box bx = load_box(ax);
    bool up = false;
    if (LEAF && valid) up = leaf_rows[(size_t)ax.leaf_row * v.n_molpad + td.molpad0 + m] == GORDER_UPPER;
    f3 nrm = mk3(v.normal_axis == 0 ? 1.f : 0.f, v.normal_axis == 1 ? 1.f : 0.f, v.normal_axis == 2 ? 1.f : 0.f);
    bool normal_bad = false;
    if (NVEC && valid) {
        const float *np = normals + ((size_t)f * 3) * v.n_molpad + td.molpad0 + m;
        nrm = mk3(np[0], np[v.n_molpad], np[2 * (size_t)v.n_molpad]);
        // the normal is requested for every molecule, before the geometry test (uaorder.rs:412-413)
        if (nrm.x != nrm.x) {
            normal_bad = true;
            int npts = normal_npoints ? normal_npoints[(size_t)f * v.n_molpad + td.molpad0 + m] : 0;
            raise_error(v, v.normal_mode == GORDER_NORMAL_DYNAMIC ? GORDER_ERR_DYNAMIC_NORMAL_POINTS : GORDER_ERR_MANUAL_NORMAL_FRAME, npts);
        }
    }
    if (!EXTRA) {
        int a = __reduce_add_sync(0xffffffffu, (int)valid), b = __reduce_add_sync(0xffffffffu, (int)(valid && up));
        if (lane == 0) { atomicAdd(&s_cnt[0], a); atomicAdd(&s_cnt[1], b); }
    }
    const float *base = planes + (size_t)f * v.frame_floats + mol_offset(td, valid ? m : 0);
    const int mpad = td.cstride;
    for (int i = 0; i < ni; i++) {
        const UAItem it = s_items[i];
        const int nh = it.kind == GORDER_UA_CH3 ? 3 : (it.kind == GORDER_UA_CH2 ? 2 : 1);
        int q[3] = {0, 0, 0};
        bool in[3] = {false, false, false};
        if (valid && !normal_bad) {
            const f3 t = mk3(__ldg(base + it.t_off), __ldg(base + it.t_off + mpad), __ldg(base + it.t_off + 2 * mpad));
            const f3 h1 = mk3(__ldg(base + it.h1_off), __ldg(base + it.h1_off + mpad), __ldg(base + it.h1_off + 2 * mpad));
            const f3 h2 = mk3(__ldg(base + it.h2_off), __ldg(base + it.h2_off + mpad), __ldg(base + it.h2_off + 2 * mpad));
            f3 h3 = mk3(0, 0, 0);
            if (it.kind == GORDER_UA_CH1_SAT) h3 = mk3(__ldg(base + it.h3_off), __ldg(base + it.h3_off + mpad), __ldg(base + it.h3_off + 2 * mpad));
            f3 hyd[3];
            if (!EXTRA && !v.ua_exact) {
                f3 u[3];
                predict_directions_fast<PBC>(v, it.kind, t, h1, h2, h3, bx, u);
#pragma unroll
                for (int k = 0; k < 3; k++)   // Vector3D::shift: p + u * 0.109 (the wrap into the box is undone by the fold below)
                    hyd[k] = mk3(__fadd_rn(t.x, __fmul_rn(u[k].x, 0.109f)), __fadd_rn(t.y, __fmul_rn(u[k].y, 0.109f)), __fadd_rn(t.z, __fmul_rn(u[k].z, 0.109f)));
            } else predict_hydrogens<PBC>(v, it.kind, t, h1, h2, h3, bx, hyd);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (k >= nh) break;
                const f3 d = vector_to<PBC>(t, hyd[k], bx);   // calculate_sch, uaorder.rs:375-397
                f3 mid;
                if (EXTRA) {
                    // sic: bond_pos = hydrogen + vec / 2 (uaorder.rs:385)
                    mid = mk3(__fadd_rn(hyd[k].x, d.x * 0.5f), __fadd_rn(hyd[k].y, d.y * 0.5f), __fadd_rn(hyd[k].z, d.z * 0.5f));
                    if (v.shape.kind != GORDER_GEOM_NONE && !shape_inside<PBC>(v.shape, ax, bx, mid)) continue;
                }
                const float s = calc_sch_fast(d, nrm);
                const float chk = (d.x + d.y + d.z) * 0.0f;   // NaN / Inf coordinates (the clamp in calc_sch would hide them)
                if (chk != chk) {
                    raise_error(v, GORDER_ERR_UNDEFINED_POSITION, ((long long)ch.type << 48) | ((long long)i << 32) | (unsigned)m);
                    continue;
                }
                q[k] = order_value_fast(s);
                in[k] = true;
                if (EXTRA && v.map.enabled) map_add<LEAF>(v, o, td.slot0 + it.slot_rel + k, mid, q[k], up);
            }
        }
        for (int k = 0; k < nh; k++) {   // warp-uniform trip count
            int su = 0, sl = 0, cu = 0, cl = 0;
            if (in[k]) { if (LEAF && !up) { sl = q[k]; cl = 1; } else { su = q[k]; cu = 1; } }
            warp_commit<LEAF, EXTRA>(s_acc + ((size_t)warp * no + it.slot_rel + k) * NA, lane, su, sl, cu, cl);
        }
    }
    __syncthreads();
    cta_flush<LEAF, EXTRA>(v, o, s_acc, no, td.slot0, ax.tw_row, s_cnt[0], s_cnt[1]);
}

// ---------------------------------------------------------------------------------------------
// fold: add the per-frame accumulators of a batch into the running totals (order.rs:160-188).
// One thread per (slot, leaf); frames added in ascending order (fixed order, integer anyway).
// With leaflets the kernels only fill upper / lower; total = upper + lower is completed here
// (bond.rs:184-215 adds every sample to total and to exactly one leaflet).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fold_kernel(int n_slots, int n_rows, int leaf, long long *__restrict__ bsum,
                                                   unsigned long long *__restrict__ bcnt, long long *__restrict__ tot_sum,
                                                   unsigned long long *__restrict__ tot_cnt) {
    const int s = blockIdx.x;
    long long ts[3] = {0, 0, 0};
    unsigned long long tc[3] = {0, 0, 0};
    for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const size_t b = ((size_t)r * n_slots + s) * 3;
        if (leaf) {
            bsum[b + GORDER_TOTAL] = bsum[b + GORDER_ACC_UPPER] + bsum[b + GORDER_ACC_LOWER];
            bcnt[b + GORDER_TOTAL] = bcnt[b + GORDER_ACC_UPPER] + bcnt[b + GORDER_ACC_LOWER];
        }
#pragma unroll
        for (int k = 0; k < 3; k++) { ts[k] += bsum[b + k]; tc[k] += bcnt[b + k]; }
    }
    __shared__ long long s_s[3][4];
    __shared__ unsigned long long s_c[3][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) { ts[k] += __shfl_down_sync(0xffffffffu, ts[k], o); tc[k] += __shfl_down_sync(0xffffffffu, tc[k], o); }
        if (lane == 0) { s_s[k][warp] = ts[k]; s_c[k][warp] = tc[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        long long a = 0; unsigned long long c = 0;
        for (int w = 0; w < 4; w++) { a += s_s[k][w]; c += s_c[k][w]; }
        tot_sum[s * 3 + k] += a; tot_cnt[s * 3 + k] += c;
    }
}

}  // namespace gorder

// gorder_fast.cuh — K1f: the bond engine of the headline configurations, written for issue slots.
//
// Same arithmetic and results as bond_order_kernel<MPT, PBC=1, NVEC=0, LEAF, EXTRA=0, SPEC> (static normal,
// PBC, no geometry / maps: AAOrder / CGOrder "basic" and "leaflets" runs; topology/bond.rs:396-446,
// :184-215, analysis/mod.rs:76-82, order.rs:21-26, leaflets.rs:711-732), bit for bit.  The generic kernel
// spends ~82 thread-instructions per sample and is bound by the issue rate, not by HBM
// (profiles/README.md); this one needs about a third of that:
//   * packed f32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE-rn operations per issue slot):
//     a lane owns 2 NP consecutive molecules, the j-th pair lives in one 64-bit register pair, exactly as
//     the 128-bit plane loads deliver it;
//   * no register copies between bonds that share an atom: the two atom buffers swap ROLES (two
//     specialisations of the loop body, selected by a warp-uniform branch);
//   * one |r| <= guard test per component and thread (FMNMX3 over the lane's molecules);
//   * NaN / Inf coordinates are caught by an integer max over the bit patterns of |d|^2;
//   * warp sums (REDUX) stay in registers (lane b keeps bond b) and reach shared memory once per 32 bonds;
//   * every load address is one IMAD.WIDE from an opaque per-lane pointer;
//   * no validity masks: the (at most one) partial tile of a molecule type runs the generic body instead.
#pragma once
#include "gorder_kernels.cuh"

namespace gorder {

__device__ __forceinline__ float2 padd(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 psub(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; "
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

// DRAM -> L2 prefetch of a contiguous run (one instruction, one thread): the CTA's slice of a plane
__device__ __forceinline__ void l2_prefetch_bulk(const float *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// one atom of the lane's 2 NP molecules: component c of pair p
template <int NP> struct AtomBuf { float2 c[3][NP]; };

template <int NP>
__device__ __forceinline__ void load_atom(AtomBuf<NP> &a, const float *p0, const float *p1, const float *p2) {
    if (NP == 2) {
        const float4 x = __ldg(reinterpret_cast<const float4 *>(p0)), y = __ldg(reinterpret_cast<const float4 *>(p1)),
                     z = __ldg(reinterpret_cast<const float4 *>(p2));
        a.c[0][0] = make_float2(x.x, x.y); a.c[0][NP - 1] = make_float2(x.z, x.w);
        a.c[1][0] = make_float2(y.x, y.y); a.c[1][NP - 1] = make_float2(y.z, y.w);
        a.c[2][0] = make_float2(z.x, z.y); a.c[2][NP - 1] = make_float2(z.z, z.w);
    } else {
        a.c[0][0] = __ldg(reinterpret_cast<const float2 *>(p0));
        a.c[1][0] = __ldg(reinterpret_cast<const float2 *>(p1));
        a.c[2][0] = __ldg(reinterpret_cast<const float2 *>(p2));
    }
}

// Frame-invariant tables of a handle in constant memory: a CTA lives for ~15 us, the chain of dependent
// global loads chunk -> type -> bond table at its start (3 x L2 latency) was a quarter of its stall time.
constexpr int kFastSlots = 8, kFastTypes = 8, kFastBonds = 128;
struct FastTables {
    int n_types;
    int chunk0[kFastTypes + 1];   // first chunk of every molecule type
    TypeDesc types[kFastTypes];
    BondItem bonds[kFastBonds];
};
__constant__ FastTables c_fast[kFastSlots];

// per-thread state of the bond loop
template <int NP> struct FastState {
    unsigned imax;         // max bit pattern of |d|^2 seen so far (NaN / Inf detector)
    float2 dsum, dsq;      // SPEC: sum d, sum d^2 of the membrane atoms (two lanes of partial sums)
    float dabs;            // SPEC: max |d|
    int ru, rl;            // lane (b & 31) keeps the warp sums of bond b until they are flushed to shared memory
};

// everything that is uniform over the bond loop
struct FastConst {
    const float *tp;       // plane x of in-tile offset 0, first molecule of the lane
    int o1, o2;            // component strides (the normal axis comes last)
    float L0, L1, L2, h0, h1, h2, g0, g1, g2;
    float sref, inv_l2;    // SPEC: provisional centre, 1 / L along the leaflet axis
    int nb, lane;
    int *s_acc_warp;       // this warp's rows of the shared accumulators
};

// One bond of the lane's molecules: `first` holds the bond's first atom, `other` receives the second.
template <int NP, bool LEAF, bool SPEC>
__device__ __forceinline__ void fast_bond(AtomBuf<NP> &first, AtomBuf<NP> &other, const FastConst &k, const BondItem bi, int b,
                                          const bool (&up)[2 * NP], FastState<NP> &s) {
    constexpr int NA = LEAF ? 2 : 1;
    const int a_item = bi.a_off;
    if ((a_item & 3) == 0) {
        const float *pa = k.tp + (a_item & ~15);
        load_atom<NP>(first, pa, pa + k.o1, pa + k.o2);
    }
    {
        const float *pb = k.tp + bi.b_off;
        load_atom<NP>(other, pb, pb + k.o1, pb + k.o2);
    }
    if (SPEC && (a_item & 12)) {   // membrane atoms seen for the first time: displacement from the provisional centre
        auto add = [&](const AtomBuf<NP> &at) {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const float2 t = padd(at.c[2][p], bc2(-k.sref));
                const float2 q = padd(pfma(t, bc2(k.inv_l2), bc2(12582912.0f)), bc2(-12582912.0f));   // rint(t / L)
                const float2 d = pfma(q, bc2(-k.L2), t);
                s.dsum = padd(s.dsum, d); s.dsq = pfma(d, d, s.dsq);
                s.dabs = fmaxf(s.dabs, fmaxf(fabsf(d.x), fabsf(d.y)));
            }
        };
        if (a_item & 4) add(first);
        if (a_item & 8) add(other);
    }

    // bond vectors; the fold's exact fast path  fl(fl(fl(fl(r + L/2) + L) - L) - L/2)  for everybody, the literal
    // expression (out of line) for the rare lane that holds a component beyond the guard 0.99 L/2
    float2 d[3][NP];
    float m0x = 0.0f, m1x = 0.0f, m2x = 0.0f;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const float2 r0 = psub(other.c[0][p], first.c[0][p]), r1 = psub(other.c[1][p], first.c[1][p]), r2 = psub(other.c[2][p], first.c[2][p]);
        m0x = fmaxf(m0x, fmaxf(fabsf(r0.x), fabsf(r0.y)));
        m1x = fmaxf(m1x, fmaxf(fabsf(r1.x), fabsf(r1.y)));
        m2x = fmaxf(m2x, fmaxf(fabsf(r2.x), fabsf(r2.y)));
        d[0][p] = padd(padd(padd(padd(r0, bc2(k.h0)), bc2(k.L0)), bc2(-k.L0)), bc2(-k.h0));
        d[1][p] = padd(padd(padd(padd(r1, bc2(k.h1)), bc2(k.L1)), bc2(-k.L1)), bc2(-k.h1));
        d[2][p] = padd(padd(padd(padd(r2, bc2(k.h2)), bc2(k.L2)), bc2(-k.L2)), bc2(-k.h2));
    }
    if ((m0x > k.g0) | (m1x > k.g1) | (m2x > k.g2)) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const float2 r0 = psub(other.c[0][p], first.c[0][p]), r1 = psub(other.c[1][p], first.c[1][p]), r2 = psub(other.c[2][p], first.c[2][p]);
            if (fabsf(r0.x) > k.g0) d[0][p].x = min_image_slow(r0.x, k.L0, k.h0);
            if (fabsf(r0.y) > k.g0) d[0][p].y = min_image_slow(r0.y, k.L0, k.h0);
            if (fabsf(r1.x) > k.g1) d[1][p].x = min_image_slow(r1.x, k.L1, k.h1);
            if (fabsf(r1.y) > k.g1) d[1][p].y = min_image_slow(r1.y, k.L1, k.h1);
            if (fabsf(r2.x) > k.g2) d[2][p].x = min_image_slow(r2.x, k.L2, k.h2);
            if (fabsf(r2.y) > k.g2) d[2][p].y = min_image_slow(r2.y, k.L2, k.h2);
        }
    }
    // S = 1.5 c^2 - 0.5, c = d_axis rsqrt(|d|^2) (calc_sch_axis_fast); |d| = 0 -> c = 0 * inf = NaN -> min(NaN, 1) = 1 -> S = 1,
    // which is what Vector3D::angle's zero-norm rule gives.  q = round(S 1e6) (order_value_fast).
    int st = 0, su = 0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const float2 n1 = pfma(d[2][p], d[2][p], pfma(d[1][p], d[1][p], pmul(d[0][p], d[0][p])));
        s.imax = max(s.imax, max(__float_as_uint(n1.x), __float_as_uint(n1.y)));
        const float2 c = pmul(d[2][p], make_float2(rsqrt_ftz(n1.x), rsqrt_ftz(n1.y)));
        float2 c2 = pmul(c, c);
        c2.x = fminf(c2.x, 1.0f); c2.y = fminf(c2.y, 1.0f);
        const float2 sv = pmul(pfma(bc2(1.5f), c2, bc2(-0.5f)), bc2(1000000.0f));
        const int qa = __float2int_rn(sv.x), qb = __float2int_rn(sv.y);
        st += qa + qb;
        if (LEAF) su += (up[2 * p] ? qa : 0) + (up[2 * p + 1] ? qb : 0);
    }
    // fixed-order hardware tree (REDUX) -> deterministic
    const int wu = __reduce_add_sync(0xffffffffu, LEAF ? su : st);
    const int wl = LEAF ? __reduce_add_sync(0xffffffffu, st - su) : 0;
    if ((b & 31) == k.lane) { s.ru = wu; s.rl = wl; }
    if ((b & 31) == 31 || b == k.nb - 1) {
        if (k.lane <= (b & 31)) {
            int *p = k.s_acc_warp + ((b & ~31) + k.lane) * NA;
            p[0] = s.ru;
            if (LEAF) p[1] = s.rl;
        }
    }
}

// BLOCK threads x 2 NP molecules = the tile of the molecule type's layout (256, 512 or 1024 molecules): small systems
// (one tile of 256 molecules per frame, BASELINE configs[0]) run 64-thread CTAs with four molecules per lane, so that the
// batch's frames supply the parallelism and every lane still issues 128-bit loads and packed arithmetic.
template <int NP, bool LEAF, bool SPEC, int BLOCK = kBlock>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) bond_fast_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                              const unsigned char *__restrict__ leaf_rows, AccumOut o, int fast_slot) {
    constexpr int MPT = 2 * NP;
    constexpr int NA = LEAF ? 2 : 1;
    constexpr int WARPS = BLOCK / 32;
    extern __shared__ int smem[];
    __shared__ int s_nup[WARPS];
    __shared__ unsigned s_done;
    __shared__ float s_hmm[2][WARPS];
    __shared__ double s_dsum[2][WARPS];
    __shared__ float s_dabs[WARPS];
    Chunk ch;
    TypeDesc td;
    if (fast_slot >= 0) {   // tables in constant memory: no global load before the first plane load
        const FastTables &ft = c_fast[fast_slot];
        int t = 0;
        while (t + 1 < ft.n_types && (int)blockIdx.x >= ft.chunk0[t + 1]) t++;
        td = ft.types[t];
        ch.type = t; ch.first_mol = ((int)blockIdx.x - ft.chunk0[t]) * (BLOCK * MPT);
    } else {
        ch = v.chunks[blockIdx.x];
        td = v.types[ch.type];
    }
    if (ch.first_mol + BLOCK * MPT > td.n_mol) {   // the last tile of a molecule type (molecules past the end) takes the generic, masked code
        bond_order_body<MPT, true, false, LEAF, false, SPEC, BLOCK>(v, planes, aux, leaf_rows, nullptr, nullptr, o);
        return;
    }
    const int f = blockIdx.y;
    const FrameAux &ax = aux[f];
    const int nb = td.n_items;
    const int mpad = td.cstride;
    const int c0 = (v.normal_axis + 1) % 3, c1 = (v.normal_axis + 2) % 3, c2 = v.normal_axis;   // the normal axis comes last
    BondItem *s_bonds = reinterpret_cast<BondItem *>(smem);
    int *s_acc = smem + 2 * nb;                 // [WARPS][nb][NA]
    for (int i = threadIdx.x; i < nb; i += BLOCK) s_bonds[i] = fast_slot >= 0 ? c_fast[fast_slot].bonds[td.item_off + i] : v.bonds[td.item_off + i];
    if (threadIdx.x == 0) s_done = 0u;
    // The kernel is bound by memory latency (8 warps per scheduler, ~150 issue slots between a warp's loads): one
    // lane per CTA asks L2 for the CTA's slices (BLOCK * MPT floats per component: contiguous) of the planes that the
    // bonds kPrefetchAhead iterations later will read, so that the plane loads find their lines in L2.
    constexpr int kPrefetchAhead = 2;   // measured: 1 -> 0.78, 2 -> 0.80, 3 -> 0.78, 6 -> 0.77, 11 -> 0.75 of the HBM peak
    constexpr unsigned kSliceBytes = BLOCK * MPT * sizeof(float);
    const float *tile0 = planes + (size_t)f * v.frame_floats + mol_offset(td, ch.first_mol);
    if (threadIdx.x == 0 && LEAF && (SPEC || o.inline_center)) l2_prefetch_bulk(tile0 + td.head_off + v.leaflet_axis * mpad, kSliceBytes);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = ch.first_mol + threadIdx.x * MPT;
    FastConst k;
    k.o1 = (c1 - c0) * mpad; k.o2 = (c2 - c0) * mpad;
    auto prefetch_bond = [&](int bb) {
        if (threadIdx.x != 0 || bb >= nb) return;
        const BondItem it = s_bonds[bb];
        const float *t0 = tile0 + c0 * mpad;
        if ((it.a_off & 3) == 0) {
            const float *pa = t0 + (it.a_off & ~15);
            l2_prefetch_bulk(pa, kSliceBytes); l2_prefetch_bulk(pa + k.o1, kSliceBytes); l2_prefetch_bulk(pa + k.o2, kSliceBytes);
        }
        const float *pb = t0 + it.b_off;
        l2_prefetch_bulk(pb, kSliceBytes); l2_prefetch_bulk(pb + k.o1, kSliceBytes); l2_prefetch_bulk(pb + k.o2, kSliceBytes);
    };
    for (int bb = 0; bb < kPrefetchAhead; bb++) prefetch_bond(bb);
    k.L0 = ax.L[c0]; k.L1 = ax.L[c1]; k.L2 = ax.L[c2];
    k.h0 = ax.half[c0]; k.h1 = ax.half[c1]; k.h2 = ax.half[c2];
    k.g0 = ax.guard[c0]; k.g1 = ax.guard[c1]; k.g2 = ax.guard[c2];
    k.nb = nb; k.lane = lane; k.s_acc_warp = s_acc + (size_t)warp * nb * NA;
    const float *tp0 = planes + (size_t)f * v.frame_floats + mol_offset(td, m0);

    // ---- leaflets of the lane's molecules (leaflets.rs:711-732 inline, or the table) ----
    bool up[MPT];
    int nup = 0;
    k.sref = SPEC ? __ldg(o.spec_ref) : 0.0f;
    float hmin = CUDART_INF_F, hmax = 0.0f;
    bool hnan = false;
#pragma unroll
    for (int j = 0; j < MPT; j++) {
        up[j] = false;
        if (LEAF) {
            const int la = v.leaflet_axis;
            if (SPEC || o.inline_center) {
                const float cen = SPEC ? k.sref : o.inline_center[3 * f + la];
                if (!SPEC && cen != cen) raise_error(v, GORDER_ERR_INVALID_GLOBAL_CENTER, ax.frame_index);
                const float hd = __ldg(tp0 + td.head_off + la * mpad + j);
                const float dh = distance_1d(hd, cen, ax.L[la], ax.half[la], true);
                up[j] = dh >= 0.0f;
                if (SPEC) { hmin = fminf(hmin, fabsf(dh)); hmax = fmaxf(hmax, fabsf(dh)); hnan = hnan || dh != dh; }
                if (v.leaflet_flip) up[j] = !up[j];
                if (o.leaf_out) o.leaf_out[(size_t)(1 + f) * v.n_molpad + td.molpad0 + m0 + j] = up[j] ? GORDER_UPPER : GORDER_LOWER;
            } else up[j] = leaf_rows[(size_t)ax.leaf_row * v.n_molpad + td.molpad0 + m0 + j] == GORDER_UPPER;
        }
        nup += up[j];
    }
    {
        const int a = __reduce_add_sync(0xffffffffu, nup);
        if (lane == 0) s_nup[warp] = a;
    }
    if (SPEC) {
        float a = hnan ? CUDART_NAN_F : hmin, b = hmax;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            const float a2 = __shfl_xor_sync(0xffffffffu, a, ofs), b2 = __shfl_xor_sync(0xffffffffu, b, ofs);
            a = (a != a || a2 != a2) ? CUDART_NAN_F : fminf(a, a2); b = fmaxf(b, b2);
        }
        if (lane == 0) { s_hmm[0][warp] = a; s_hmm[1][warp] = b; }
    }

    // ---- bond loop: the role of the two atom buffers is encoded in the program counter ----
    FastState<NP> s;
    s.imax = 0u; s.dsum = s.dsq = make_float2(0.0f, 0.0f); s.dabs = 0.0f; s.ru = s.rl = 0;
    k.inv_l2 = (SPEC && k.L2 > 0.0f) ? __frcp_rn(k.L2) : 0.0f;
    const float *tp = tp0 + c0 * mpad;
    asm volatile("" : "+l"(tp));   // keep it a pointer: every load address is ONE IMAD.WIDE away
    k.tp = tp;
    AtomBuf<NP> A, B;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int p = 0; p < NP; p++) A.c[c][p] = B.c[c][p] = make_float2(0.0f, 0.0f);
    int b = 0;
    bool done = nb <= 0;
    while (!done) {
        do {   // the bond's first atom lives in A
            prefetch_bond(b + kPrefetchAhead);
            fast_bond<NP, LEAF, SPEC>(A, B, k, s_bonds[b], b, up, s);
            b++;
            done = b >= nb;
        } while (!done && (s_bonds[b].a_off & 3) != 2);
        if (done) break;
        do {   // ... in B (the previous bond's second atom became this bond's first)
            prefetch_bond(b + kPrefetchAhead);
            fast_bond<NP, LEAF, SPEC>(B, A, k, s_bonds[b], b, up, s);
            b++;
            done = b >= nb;
        } while (!done && (s_bonds[b].a_off & 3) != 2);
    }
    if (s.imax >= 0x7f800000u)   // AnalysisError::UndefinedPosition: a NaN / Inf coordinate reached the engine
        raise_error(v, GORDER_ERR_UNDEFINED_POSITION, ((long long)ch.type << 48) | (unsigned)m0);
    if (SPEC) {
        const float tsum = s.dsum.x + s.dsum.y;
        double ds = (double)s.dsum.x + (double)s.dsum.y, dq = (double)s.dsq.x + (double)s.dsq.y;
        float a = s.dabs;
        if (tsum != tsum) a = tsum;   // NaN / Inf coordinate: poison the extent so that the frame is flagged
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            ds += __shfl_xor_sync(0xffffffffu, ds, ofs); dq += __shfl_xor_sync(0xffffffffu, dq, ofs);
            const float a2 = __shfl_xor_sync(0xffffffffu, a, ofs);
            a = (a != a || a2 != a2) ? CUDART_NAN_F : fmaxf(a, a2);
        }
        if (lane == 0) { s_dsum[0][warp] = ds; s_dsum[1][warp] = dq; s_dabs[warp] = a; }
    }
    // No barrier at the end: the warp that finishes last adds the CTA's partials (in warp order: deterministic)
    // to the frame's accumulators; the others retire and free their slots for the next CTA.
    __syncwarp();
    __threadfence_block();
    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(&s_done, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != WARPS - 1) return;
    __threadfence_block();
    int cta_up = 0;
    for (int w = 0; w < WARPS; w++) cta_up += s_nup[w];
    const int cnt_total = BLOCK * MPT;
    for (int i = lane; i < nb; i += 32) {
        long long acc0 = 0, acc1 = 0;
        for (int w = 0; w < WARPS; w++) {
            const int *p = s_acc + ((size_t)w * nb + i) * NA;
            acc0 += p[0];
            if (LEAF) acc1 += p[1];
        }
        const size_t base = ((size_t)ax.tw_row * v.n_slots + td.slot0 + i) * 3;
        if (LEAF) {
            const int c_up = cta_up, c_lo = cnt_total - cta_up;
            if (c_up) { atomicAdd((unsigned long long *)&o.bsum[base + GORDER_ACC_UPPER], (unsigned long long)acc0); atomicAdd(&o.bcnt[base + GORDER_ACC_UPPER], (unsigned long long)c_up); }
            if (c_lo) { atomicAdd((unsigned long long *)&o.bsum[base + GORDER_ACC_LOWER], (unsigned long long)acc1); atomicAdd(&o.bcnt[base + GORDER_ACC_LOWER], (unsigned long long)c_lo); }
        } else {
            atomicAdd((unsigned long long *)&o.bsum[base + GORDER_TOTAL], (unsigned long long)acc0); atomicAdd(&o.bcnt[base + GORDER_TOTAL], (unsigned long long)cnt_total);
        }
    }
    if (SPEC && lane == 0) {
        double ds = 0.0, dq = 0.0;
        float p0 = 0.0f, p1 = CUDART_INF_F, p2 = 0.0f;
        bool bad = false;
        for (int w = 0; w < WARPS; w++) {
            ds += s_dsum[0][w]; dq += s_dsum[1][w];
            bad = bad || s_dabs[w] != s_dabs[w] || s_hmm[0][w] != s_hmm[0][w];
            p0 = fmaxf(p0, s_dabs[w]); p1 = fminf(p1, s_hmm[0][w]); p2 = fmaxf(p2, s_hmm[1][w]);
        }
        spec_publish(o, f, k.sref, k.L2, k.h2, ds, dq, p0, p1, p2, bad);
    }
}

}  // namespace gorder

// gorder_ua_fast.cuh — K2f: the united-atom engine of the headline UA configuration, written for issue slots.
//
// Same construction and results as ua_order_kernel<PBC=1, NVEC=0, LEAF, EXTRA=0> with the streaming hydrogen
// construction (predict_directions_fast: uaorder.rs:375-437, :947-1104; static normal, PBC, no geometry / maps), bit for
// bit.  That kernel is compute-bound (0.13 of the HBM peak, profiles/README.md): one molecule per lane, scalar f32,
// every carbon re-reading its three or four heavy atoms through L1.  This one:
//   * is persistent: one CTA per SM walks the (frame, tile) work items of a batch; a tile -- ONE contiguous region of
//     3 x used atoms x 256 molecules floats, the layout of DESIGN.md §3 -- is staged in shared memory with bulk asynchronous
//     copies (cp.async.bulk, completion on an mbarrier) while the previous tile is being worked on (two buffers), so the
//     bytes in flight of an SM are a whole tile, independent of registers and occupancy, and every heavy atom crosses
//     L2 -> SM exactly once however many carbons use it as a helper;
//   * a lane owns TWO molecules and computes on packed f32x2 (FADD2 / FMUL2 / FFMA2: two IEEE-rn operations per issue
//     slot), reading its pairs of coordinates with one conflict-free 64-bit shared load each;
//   * the 512 threads are 128 lane pairs x 4 carbon groups: the carbons of a molecule type are dealt to the quarters of the
//     CTA in turn, so a tile of 256 molecules keeps 16 warps busy;
//   * one minimum-image guard test per carbon (not per vector), none for the 0.109 nm hydrogen vectors;
//   * NaN / Inf coordinates are caught by an integer max over the bit patterns of |d|^2 (as in K1f).
// The kernel is bound by the FP32 pipe, not by HBM: the reference's construction costs ~79 f32 operations per sample
// (a third of them the four roundings of every minimum-image fold, which are part of the reference's results), i.e. at
// most 0.48 of the HBM peak at 100 % pipe utilisation (profiles/README.md).
#pragma once
#include "gorder_fast.cuh"

namespace gorder {

struct P3 { float2 x, y, z; };

__device__ __forceinline__ float2 pneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ P3 p3sub(const P3 &a, const P3 &b) { P3 r; r.x = psub(a.x, b.x); r.y = psub(a.y, b.y); r.z = psub(a.z, b.z); return r; }
// dot_f / cross_f / unit_f / lin2 of gorder_kernels.cuh on pairs: the same operations in the same order
__device__ __forceinline__ float2 p3dot(const P3 &a, const P3 &b) { return pfma(a.z, b.z, pfma(a.y, b.y, pmul(a.x, b.x))); }
__device__ __forceinline__ P3 p3cross(const P3 &a, const P3 &b) {
    P3 r;
    r.x = pfma(a.y, b.z, pmul(pneg(a.z), b.y)); r.y = pfma(a.z, b.x, pmul(pneg(a.x), b.z)); r.z = pfma(a.x, b.y, pmul(pneg(a.y), b.x));
    return r;
}
__device__ __forceinline__ P3 p3scale(const P3 &a, float2 s) { P3 r; r.x = pmul(a.x, s); r.y = pmul(a.y, s); r.z = pmul(a.z, s); return r; }
__device__ __forceinline__ float2 prsq_newton(float2 x) {   // rsq_newton on both halves
    const float2 r = make_float2(rsqrt_ftz(x.x), rsqrt_ftz(x.y));
    return pmul(r, pfma(pmul(bc2(-0.5f), x), pmul(r, r), bc2(1.5f)));
}
__device__ __forceinline__ P3 p3unit(const P3 &a) { return p3scale(a, prsq_newton(p3dot(a, a))); }
__device__ __forceinline__ P3 p3lin2(const P3 &a, float2 ca, const P3 &b, float2 cb) {
    P3 r; r.x = pfma(a.x, ca, pmul(b.x, cb)); r.y = pfma(a.y, ca, pmul(b.y, cb)); r.z = pfma(a.z, ca, pmul(b.z, cb)); return r;
}

// box constants of a frame (component order x, y, z)
struct UaBox { float L[3], h[3], g[3]; };

// Vector3D::vector_to on a pair of molecules.  The fold's exact fast path fl(fl(fl(fl(r + L/2) + L) - L) - L/2) is valid for
// |r| <= guard = 0.99 L/2; a half that holds a component beyond it (a bond through the periodic boundary) takes the literal
// expression (out of line).  p3fold evaluates the fast path, UaGuard collects max |r| per component over the vectors of one
// carbon so that ONE test per carbon decides whether any of them needs p3fold_fix.
__device__ __forceinline__ P3 p3fold(const P3 &r, const UaBox &b) {
    P3 d;
    d.x = padd(padd(padd(padd(r.x, bc2(b.h[0])), bc2(b.L[0])), bc2(-b.L[0])), bc2(-b.h[0]));
    d.y = padd(padd(padd(padd(r.y, bc2(b.h[1])), bc2(b.L[1])), bc2(-b.L[1])), bc2(-b.h[1]));
    d.z = padd(padd(padd(padd(r.z, bc2(b.h[2])), bc2(b.L[2])), bc2(-b.L[2])), bc2(-b.h[2]));
    return d;
}
struct UaGuard {
    float m0 = 0.0f, m1 = 0.0f, m2 = 0.0f;
    __device__ __forceinline__ void see(const P3 &r) {
        m0 = fmaxf(m0, fmaxf(fabsf(r.x.x), fabsf(r.x.y))); m1 = fmaxf(m1, fmaxf(fabsf(r.y.x), fabsf(r.y.y))); m2 = fmaxf(m2, fmaxf(fabsf(r.z.x), fabsf(r.z.y)));
    }
    __device__ __forceinline__ bool beyond(const UaBox &b) const { return (m0 > b.g[0]) | (m1 > b.g[1]) | (m2 > b.g[2]); }
};
__device__ __forceinline__ void p3fold_fix(const P3 &r, const UaBox &b, P3 &d) {
    if (fabsf(r.x.x) > b.g[0]) d.x.x = min_image_slow(r.x.x, b.L[0], b.h[0]);
    if (fabsf(r.x.y) > b.g[0]) d.x.y = min_image_slow(r.x.y, b.L[0], b.h[0]);
    if (fabsf(r.y.x) > b.g[1]) d.y.x = min_image_slow(r.y.x, b.L[1], b.h[1]);
    if (fabsf(r.y.y) > b.g[1]) d.y.y = min_image_slow(r.y.y, b.L[1], b.h[1]);
    if (fabsf(r.z.x) > b.g[2]) d.z.x = min_image_slow(r.z.x, b.L[2], b.h[2]);
    if (fabsf(r.z.y) > b.g[2]) d.z.y = min_image_slow(r.z.y, b.L[2], b.h[2]);
}
__device__ __forceinline__ P3 p3vector_to(const P3 &from, const P3 &to, const UaBox &b) {   // one vector, checked on its own
    const P3 r = p3sub(to, from);
    P3 d = p3fold(r, b);
    UaGuard g;
    g.see(r);
    if (g.beyond(b)) p3fold_fix(r, b, d);
    return d;
}

// predict_directions_fast (gorder_kernels.cuh) on a pair of molecules
__device__ __forceinline__ void predict_directions_pair(const DeviceView &v, int kind, const P3 &t, const P3 &h1, const P3 &h2, const P3 &h3,
                                                        const UaBox &bx, P3 (&u)[3]) {
    // the carbon's bond vectors to its helpers: fast fold for all, one guard test for the carbon
    const P3 r1 = p3sub(h1, t), r2 = p3sub(h2, t);
    P3 th1 = p3fold(r1, bx), th2 = p3fold(r2, bx), th3 = th1;
    UaGuard g;
    g.see(r1); g.see(r2);
    if (kind == GORDER_UA_CH1_SAT) {
        const P3 r3 = p3sub(h3, t);
        th3 = p3fold(r3, bx);
        g.see(r3);
        if (g.beyond(bx)) p3fold_fix(r3, bx, th3);
    }
    if (g.beyond(bx)) { p3fold_fix(r1, bx, th1); p3fold_fix(r2, bx, th2); }
    if (kind == GORDER_UA_CH2) {   // uaorder.rs:985-1020
        const P3 a = p3unit(th1), b = p3unit(th2);
        const P3 pn = p3cross(b, a);
        const P3 ra = p3unit(p3sub(a, b));
        const P3 rv = p3cross(pn, ra), w = p3cross(ra, rv);
        const float2 inv = prsq_newton(p3dot(rv, rv));
        const float2 cc = pmul(bc2(v.tet_half_c), inv), cs = pmul(bc2(v.tet_half_s), inv);
        u[0] = p3lin2(rv, cc, w, cs);
        u[1] = p3lin2(rv, cc, w, pneg(cs));
    } else if (kind == GORDER_UA_CH3) {   // uaorder.rs:947-981
        const P3 ax = p3unit(p3cross(th2, th1));
        const P3 hv1 = p3lin2(th1, bc2(v.tet_c), p3cross(ax, th1), bc2(v.tet_s));
        u[0] = p3unit(hv1);
        const P3 n = p3unit(th1), nxu = p3cross(n, u[0]);
        const float2 nd = pmul(p3dot(n, u[0]), bc2(1.0f - v.ch3_c));
        P3 base;
        base.x = pfma(u[0].x, bc2(v.ch3_c), pmul(n.x, nd)); base.y = pfma(u[0].y, bc2(v.ch3_c), pmul(n.y, nd)); base.z = pfma(u[0].z, bc2(v.ch3_c), pmul(n.z, nd));
        u[1].x = pfma(nxu.x, bc2(v.ch3_s), base.x); u[1].y = pfma(nxu.y, bc2(v.ch3_s), base.y); u[1].z = pfma(nxu.z, bc2(v.ch3_s), base.z);
        u[2].x = pfma(nxu.x, bc2(-v.ch3_s), base.x); u[2].y = pfma(nxu.y, bc2(-v.ch3_s), base.y); u[2].z = pfma(nxu.z, bc2(-v.ch3_s), base.z);
    } else if (kind == GORDER_UA_CH1_UNSAT) {   // uaorder.rs:1024-1045
        const float2 nn = pmul(p3dot(th1, th1), p3dot(th2, th2));
        float2 cg = pmul(p3dot(th1, th2), prsq_newton(nn));
        cg.x = fminf(1.0f, fmaxf(-1.0f, cg.x)); cg.y = fminf(1.0f, fmaxf(-1.0f, cg.y));
        if (nn.x == 0.0f) cg.x = 1.0f;
        if (nn.y == 0.0f) cg.y = 1.0f;
        const float2 ch = make_float2(sqrtf(fmaxf(0.0f, 0.5f * (1.0f + cg.x))), sqrtf(fmaxf(0.0f, 0.5f * (1.0f + cg.y))));
        const float2 sh = make_float2(sqrtf(fmaxf(0.0f, 0.5f * (1.0f - cg.x))), sqrtf(fmaxf(0.0f, 0.5f * (1.0f - cg.y))));
        const P3 ax = p3unit(p3cross(th1, th2));
        u[0] = p3unit(p3lin2(th2, pneg(ch), p3cross(ax, th2), sh));
    } else {   // uaorder.rs:1087-1104
        const P3 a = p3unit(th1), b = p3unit(th2), c = p3unit(th3);
        P3 s;
        s.x = pneg(padd(padd(a.x, b.x), c.x)); s.y = pneg(padd(padd(a.y, b.y), c.y)); s.z = pneg(padd(padd(a.z, b.z), c.z));
        u[0] = p3unit(s);
    }
}

// ---- bulk asynchronous copy global -> shared, completion on an mbarrier (TMA without a tensor map) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

constexpr int kUaTile = 256;        // molecules per tile (= TypeDesc::tile of a UA type)
constexpr int kUaGroups = 4;        // carbon groups: the carbons of a molecule type are dealt to the quarters of the CTA in turn
constexpr int kUaPairWarps = kUaTile / 2 / 32;   // warps that cover the tile's 128 lane pairs
constexpr int kUaThreads = 32 * kUaPairWarps * kUaGroups;   // 512

// Persistent: one CTA per SM walks the (frame, tile) work items of the batch, item w = blockIdx.x + k gridDim.x.  While
// the 16 warps work on tile k out of one shared-memory buffer, the bulk copy of tile k + 1 fills the other one, so the
// arithmetic never waits for HBM and the bytes in flight per SM are a whole tile (S-UA: 105 KB).  One __syncthreads
// per tile: the accumulators and the upper-leaflet counts are double-buffered like the tiles.
// dynamic shared memory: [NBUF tiles: max_tile floats each][items of all types][accumulators: 2 x kUaPairWarps x max_orders x NA ints]
template <bool LEAF>
__global__ void __launch_bounds__(kUaThreads, 1) ua_fast_kernel(DeviceView v, const float *__restrict__ planes, const FrameAux *__restrict__ aux,
                                                                const unsigned char *__restrict__ leaf_rows, AccumOut o, int n_frames, int max_tile_floats,
                                                                int n_items_total, int max_orders, int nbuf) {
    constexpr int NA = LEAF ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long s_bar[2];
    __shared__ int s_up[2][kUaPairWarps];
    float *s_tiles = reinterpret_cast<float *>(smem_raw);
    UAItem *s_items = reinterpret_cast<UAItem *>(s_tiles + (size_t)nbuf * max_tile_floats);
    int *s_acc_all = reinterpret_cast<int *>(s_items + n_items_total);   // [2][kUaPairWarps][max_orders][NA]
    const int acc_words = kUaPairWarps * max_orders * NA;
    const int n_work = n_frames * v.n_chunks;
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    for (int i = threadIdx.x; i < n_items_total; i += kUaThreads) s_items[i] = v.ua[i];
    __syncthreads();

    auto issue = [&](int w, int buf) {   // thread 0: the tile of work item w is one contiguous region of its frame
        const int f = w / v.n_chunks, c = w - f * v.n_chunks;
        const Chunk ch = v.chunks[c];
        const TypeDesc &td = v.types[ch.type];
        const unsigned bytes = (unsigned)__ldg(&td.tile_stride) * sizeof(float);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(planes + (size_t)f * v.frame_floats + mol_offset(td, ch.first_mol));
        unsigned char *dst = reinterpret_cast<unsigned char *>(s_tiles + (size_t)buf * max_tile_floats);
        mbar_expect_tx(&s_bar[buf], bytes);
        constexpr unsigned kPiece = 32768;
        for (unsigned off = 0; off < bytes; off += kPiece) bulk_g2s(dst + off, src + off, min(kPiece, bytes - off), &s_bar[buf]);
    };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp / kUaPairWarps, pw = warp % kUaPairWarps;   // carbon group, warp inside the group
    const int p = pw * 32 + lane;                                    // lane pair: molecules 2p, 2p + 1 of the tile
    const int na = v.normal_axis;
    // leaflets of the lane's two molecules in work item w (UPPER bits 0 / 1), fetched one tile ahead
    auto leaf_bits = [&](int w) -> int {
        if (!LEAF || w >= n_work) return 0;
        const int f = w / v.n_chunks, c = w - f * v.n_chunks;
        const Chunk ch = v.chunks[c];
        const TypeDesc &td = v.types[ch.type];
        const int m0 = ch.first_mol + 2 * p, n_mol = __ldg(&td.n_mol);
        const unsigned char *row = leaf_rows + (size_t)aux[f].leaf_row * v.n_molpad + __ldg(&td.molpad0) + m0;
        return ((m0 < n_mol && row[0] == GORDER_UPPER) ? 1 : 0) | ((m0 + 1 < n_mol && row[1] == GORDER_UPPER) ? 2 : 0);
    };
    int w = blockIdx.x;
    if (threadIdx.x == 0 && w < n_work) issue(w, 0);
    int leaf_next = leaf_bits(w);
    unsigned imax = 0u;
    long long bad_detail = 0;
    for (int k = 0; w < n_work; k++, w += gridDim.x) {
        const int buf = nbuf == 2 ? (k & 1) : 0, ab = k & 1;
        const int w_next = w + gridDim.x;
        // the other buffer was last read by tile k - 1, which every warp left before the barrier that closed it
        if (nbuf == 2 && threadIdx.x == 0 && w_next < n_work) issue(w_next, buf ^ 1);
        const int leaf_cur = leaf_next;
        leaf_next = leaf_bits(w_next);
        const int f = w / v.n_chunks, c = w - f * v.n_chunks;
        const Chunk ch = v.chunks[c];
        const TypeDesc td = v.types[ch.type];
        const FrameAux &ax = aux[f];
        const int ni = td.n_items, no = td.n_orders;
        const int m0 = ch.first_mol + 2 * p;
        const bool valid0 = m0 < td.n_mol, valid1 = m0 + 1 < td.n_mol;
        const bool up0 = (leaf_cur & 1) != 0, up1 = (leaf_cur & 2) != 0;
        if (LEAF && grp == 0) {
            const int b = __reduce_add_sync(0xffffffffu, (int)up0 + (int)up1);
            if (lane == 0) s_up[ab][pw] = b;
        }
        UaBox bx;
#pragma unroll
        for (int q = 0; q < 3; q++) { bx.L[q] = ax.L[q]; bx.h[q] = ax.half[q]; bx.g[q] = ax.guard[q]; }
        // |hydrogen - carbon| = 0.109 nm: its fold needs no guard test unless the box is smaller than a quarter of a nanometre
        const bool tiny_box = fminf(bx.g[0], fminf(bx.g[1], bx.g[2])) < 0.125f;
        const int mpad = td.cstride;
        const UAItem *items = s_items + td.item_off;
        int *s_acc = s_acc_all + (size_t)ab * acc_words;
        mbar_wait(&s_bar[buf], nbuf == 2 ? ((k >> 1) & 1) : (k & 1));   // tile landed

        const float *base = s_tiles + (size_t)buf * max_tile_floats + 2 * p;
        auto atom = [&](int off) {
            P3 a;
            a.x = *reinterpret_cast<const float2 *>(base + off); a.y = *reinterpret_cast<const float2 *>(base + off + mpad);
            a.z = *reinterpret_cast<const float2 *>(base + off + 2 * mpad);
            return a;
        };
        unsigned imax_tile = 0u;
        for (int i = grp; i < ni; i += kUaGroups) {
            const UAItem it = items[i];
            const int nh = it.kind == GORDER_UA_CH3 ? 3 : (it.kind == GORDER_UA_CH2 ? 2 : 1);
            const P3 t = atom(it.t_off), h1 = atom(it.h1_off), h2 = atom(it.h2_off);
            P3 h3 = t;
            if (it.kind == GORDER_UA_CH1_SAT) h3 = atom(it.h3_off);
            P3 u[3];
            predict_directions_pair(v, it.kind, t, h1, h2, h3, bx, u);
            for (int q = 0; q < nh; q++) {   // warp-uniform trip count
                // Vector3D::shift: hydrogen = t + u * 0.109 (its wrap into the box is undone by the fold), then calculate_sch (uaorder.rs:375-397)
                P3 hyd;
                hyd.x = padd(t.x, pmul(u[q].x, bc2(0.109f))); hyd.y = padd(t.y, pmul(u[q].y, bc2(0.109f))); hyd.z = padd(t.z, pmul(u[q].z, bc2(0.109f)));
                const P3 d = tiny_box ? p3vector_to(t, hyd, bx) : p3fold(p3sub(hyd, t), bx);
                const float2 n1 = p3dot(d, d);
                imax_tile = max(imax_tile, max(valid0 ? __float_as_uint(n1.x) : 0u, valid1 ? __float_as_uint(n1.y) : 0u));
                const float2 dax = na == 0 ? d.x : (na == 1 ? d.y : d.z);
                const float2 cth = pmul(dax, make_float2(rsqrt_ftz(n1.x), rsqrt_ftz(n1.y)));
                float2 c2 = pmul(cth, cth);
                c2.x = fminf(c2.x, 1.0f); c2.y = fminf(c2.y, 1.0f);   // |d| = 0: c = 0 * inf = NaN -> min(NaN, 1) = 1 -> S = 1 (angle()'s zero-norm rule)
                const float2 sv = pmul(pfma(bc2(1.5f), c2, bc2(-0.5f)), bc2(1000000.0f));
                const int qa = valid0 ? __float2int_rn(sv.x) : 0, qb = valid1 ? __float2int_rn(sv.y) : 0;
                const int st = qa + qb, su = (up0 ? qa : 0) + (up1 ? qb : 0);
                int *acc = s_acc + ((size_t)pw * no + it.slot_rel + q) * NA;
                const int wu = __reduce_add_sync(0xffffffffu, LEAF ? su : st);
                if (LEAF) {
                    const int wl = __reduce_add_sync(0xffffffffu, st - su);
                    if (lane == 0) { acc[0] = wu; acc[1] = wl; }
                } else if (lane == 0) acc[0] = wu;
            }
        }
        if (imax_tile >= 0x7f800000u && imax < 0x7f800000u) bad_detail = ((long long)ch.type << 48) | (unsigned)m0;
        imax = max(imax, imax_tile);
        __syncthreads();   // accumulators of tile k complete; its buffer may be refilled
        if (nbuf == 1 && threadIdx.x == 0 && w_next < n_work) issue(w_next, 0);
        const int cnt_total = min(kUaTile, td.n_mol - ch.first_mol);
        int cnt_up = 0;
        if (LEAF)
            for (int q = 0; q < kUaPairWarps; q++) cnt_up += s_up[ab][q];
        for (int i = threadIdx.x; i < no; i += kUaThreads) {
            long long acc0 = 0, acc1 = 0;
            for (int q = 0; q < kUaPairWarps; q++) {
                const int *r = s_acc + ((size_t)q * no + i) * NA;
                acc0 += r[0];
                if (LEAF) acc1 += r[1];
            }
            const size_t rb = ((size_t)ax.tw_row * v.n_slots + td.slot0 + i) * 3;
            if (LEAF) {
                const int c_lo = cnt_total - cnt_up;
                if (cnt_up) { atomicAdd((unsigned long long *)&o.bsum[rb + GORDER_ACC_UPPER], (unsigned long long)acc0); atomicAdd(&o.bcnt[rb + GORDER_ACC_UPPER], (unsigned long long)cnt_up); }
                if (c_lo) { atomicAdd((unsigned long long *)&o.bsum[rb + GORDER_ACC_LOWER], (unsigned long long)acc1); atomicAdd(&o.bcnt[rb + GORDER_ACC_LOWER], (unsigned long long)c_lo); }
            } else if (cnt_total) {
                atomicAdd((unsigned long long *)&o.bsum[rb + GORDER_TOTAL], (unsigned long long)acc0); atomicAdd(&o.bcnt[rb + GORDER_TOTAL], (unsigned long long)cnt_total);
            }
        }
    }
    if (imax >= 0x7f800000u)   // AnalysisError::UndefinedPosition: a NaN / Inf coordinate reached the engine
        raise_error(v, GORDER_ERR_UNDEFINED_POSITION, bad_detail);
}

}  // namespace gorder

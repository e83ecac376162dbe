// gorder_capi.cu — host side of the engine and its C ABI (include/gorder_b200.h).
//
// Replaces, behind the reference's frame loop (src/analysis/common.rs:283-339):
//   SystemTopology::new / clone per worker   -> gorder_gpu_create      (topology/mod.rs:70-118)
//   analyze_frame                            -> gorder_gpu_submit*     (common.rs:201-235)
//   ParallelTrajData::reduce + Add chain     -> gorder_gpu_finish      (topology/mod.rs:236-272)
// There is NO CPU fallback: every entry point fails with GORDER_ERR_NO_DEVICE / GORDER_ERR_CUDA
// when the device is not usable.
#include <emmintrin.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "gorder_kernels.cuh"
#include "gorder_fast.cuh"
#include "gorder_ua_fast.cuh"
#include "gorder_spherical.cuh"

using namespace gorder;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            h->set_error(e__ == cudaErrorMemoryAllocation ? GORDER_ERR_OUT_OF_MEMORY : GORDER_ERR_CUDA,   \
                         std::string(#call) + ": " + cudaGetErrorString(e__));                           \
            return h->err_code;                                                                          \
        }                                                                                                \
    } while (0)

// slots of c_fast (constant memory, per device): handles that do not get one use their global tables
std::mutex g_fast_mu;
bool g_fast_used[64][kFastSlots];

}  // namespace

struct GorderXtcDev;
void gorder_xtc_dev_free(GorderXtcDev *d);

// Debugging / A-B switches: read ONCE from the environment when a handle is created, never per batch.
struct Switches {
    int mpt = 0;                 // GORDER_MPT: molecules per lane of the bond kernels (1 | 2 | 4), 0 = size heuristic
    bool no_fast = false;        // GORDER_NO_FAST: generic bond_order_kernel instead of bond_fast_kernel
    bool no_const_tables = false;   // GORDER_NO_CONST_TABLES: K1f tables from global instead of constant memory
    bool no_spec = false, no_spec_leftover = false;   // GORDER_NO_SPEC, GORDER_NO_SPEC_LEFTOVER: centre pre-pass
    bool no_inline_leaflets = false;   // GORDER_NO_INLINE_LEAFLETS: separate leaflet_assign_kernel for Global-every-frame
    bool no_overlap = false;     // GORDER_NO_OVERLAP: one stream instead of the pre / main / post pipeline
    bool ua_exact = false;       // GORDER_UA_EXACT: bit-exact hydrogen construction everywhere
    bool xtc_nt_copy = true;          // GORDER_XTC_NO_NT_COPY: plain memcpy instead of non-temporal stores for the compressed frames -> pinned batch
    bool xtc_host_walk = false;       // GORDER_XTC_HOST_WALK: the host threads walk the XTC control bits (bookmarks) instead of xtc_walk_kernel
    bool no_sorted_normals = false;   // GORDER_NO_SORTED_NORMALS: one lane per lipid in molecule order (dynamic_normal_cell_kernel)
    bool verbose = false;        // GORDER_VERBOSE
    int center_blocks = 0, center_sub = 0;   // GORDER_CENTER_BLOCKS, GORDER_CENTER_SUB
    int cell_min_heads = 2048, lcell_min_atoms = 4096;   // GORDER_CELL_MIN_HEADS, GORDER_LCELL_MIN_ATOMS
    static Switches from_env() {
        Switches w;
        auto flag = [](const char *n) { return getenv(n) != nullptr; };
        auto num = [](const char *n, int dflt) { const char *e = getenv(n); return e ? atoi(e) : dflt; };
        w.mpt = num("GORDER_MPT", 0);
        w.no_fast = flag("GORDER_NO_FAST"); w.no_const_tables = flag("GORDER_NO_CONST_TABLES");
        w.no_spec = flag("GORDER_NO_SPEC"); w.no_spec_leftover = flag("GORDER_NO_SPEC_LEFTOVER");
        w.no_inline_leaflets = flag("GORDER_NO_INLINE_LEAFLETS"); w.no_overlap = flag("GORDER_NO_OVERLAP");
        w.ua_exact = flag("GORDER_UA_EXACT"); w.verbose = flag("GORDER_VERBOSE");
        w.no_sorted_normals = flag("GORDER_NO_SORTED_NORMALS"); w.xtc_host_walk = flag("GORDER_XTC_HOST_WALK"); w.xtc_nt_copy = !flag("GORDER_XTC_NO_NT_COPY");
        w.center_blocks = num("GORDER_CENTER_BLOCKS", 0); w.center_sub = num("GORDER_CENTER_SUB", 0);
        w.cell_min_heads = num("GORDER_CELL_MIN_HEADS", 2048); w.lcell_min_atoms = num("GORDER_LCELL_MIN_ATOMS", 4096);
        return w;
    }
};

struct GorderHandle {
    // ---- configuration (deep copy) ----
    GorderSetup s{};
    Switches sw;
    std::vector<TypeDesc> types;
    std::vector<std::vector<int32_t>> mol_base;       // per type (for error decoding)
    std::vector<std::vector<int32_t>> used_rel;       // per type: used relative atoms (sorted)
    std::vector<std::vector<int32_t>> item_slots;     // per type, per item: the atom slots (rel) involved
    std::vector<int32_t> slot_off, slot_cs;
    std::vector<int> molpad_type_h;
    int n_slots = 0, n_molpad = 0, n_mol_total = 0, n_chunks = 0, mpt = 1;
    long long frame_floats = 0;
    int max_batch = 0;
    bool leaf = false, extra = false, nvec = false, ua = false;

    // ---- device ----
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    DeviceView view{};
    std::vector<void *> owned;   // freed at destroy
    int *d_slot_off = nullptr, *d_slot_cs = nullptr, *d_molpad_type = nullptr;
    int *d_err = nullptr;
    long long *d_err_detail = nullptr;

    // accumulator block: [tot_sum n*3][tot_cnt n*3][map_sum n*3*bins][map_cnt n*3*bins] (int64 words)
    long long *d_block = nullptr;
    long long block_words = 0;
    long long *d_tot_sum = nullptr;
    unsigned long long *d_tot_cnt = nullptr;
    long long *d_map_sum = nullptr;
    unsigned long long *d_map_cnt = nullptr;

    // per-frame accumulators: ring of max_batch rows, or (timewise) all frames
    long long *d_bsum = nullptr;            // timewise: [tw_cap][n_slots][3] rows of every analysed frame
    unsigned long long *d_bcnt = nullptr;
    long long tw_cap = 0;
    // otherwise one ring per staging slot: [max_batch rows of sums][max_batch rows of counts][flagged-frame counter]
    long long *d_ring[2] = {nullptr, nullptr};
    size_t ring_words = 0;
    unsigned *d_nflag_tw = nullptr;         // [2] flagged-frame counters of the timewise mode
    std::vector<long long> frame_index_done;   // frame_index of every analysed frame, in order

    // staging (2-deep)
    float *d_xyz[2] = {nullptr, nullptr};
    float *d_planes[2] = {nullptr, nullptr};
    float *d_box[2] = {nullptr, nullptr};
    FrameAux *h_aux[2] = {nullptr, nullptr};
    FrameAux *d_aux[2] = {nullptr, nullptr};
    int *h_list[2] = {nullptr, nullptr};   // [2*max_batch]: assign list, all list
    int *d_list[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage_free[2] = {nullptr, nullptr};   // compute done with staging slot
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    int cur = 0;

    // leaflets
    unsigned char *d_leaf_rows = nullptr;   // [(1 + max_batch)][n_molpad]
    bool have_leaflets = false;
    long long cur_leaflet_frame = -1;
    unsigned char *d_leaf_collect = nullptr;
    long long leaf_collect_cap = 0, n_leaf_collected = 0;
    std::vector<long long> leaf_frame_index;

    // normals
    float *d_normals = nullptr;          // [max_batch][3][n_molpad]
    int *d_normal_npoints = nullptr;     // [max_batch][n_molpad]
    unsigned char *d_normal_used = nullptr;   // [max_batch][n_molpad] (only geometry + collect)
    float *d_normals_collect = nullptr;  // [cap][3][n_molpad]
    unsigned char *d_used_collect = nullptr;
    long long normals_collect_cap = 0;

    // cell list of the normal heads (K4)
    bool use_cells = false;
    bool normals_sorted = false;   // dynamic_normal_sorted_kernel: every analysed lipid's normal head is in the NormalHeads group
    int *d_head_molpad = nullptr;  // NormalHeads member -> padded molecule whose normal head it is (-1: none)
    int cells_cap = 0;
    int *d_head_cell = nullptr, *d_cell_count = nullptr, *d_cell_start = nullptr, *d_cell_span = nullptr;
    float4 *d_cell_sorted = nullptr;   // head positions (+ index) in cell order

    // 2-D cell list of the membrane atoms (Local leaflets)
    bool use_lcells = false;
    int lcells_cap = 0;
    int *d_matom_cell = nullptr, *d_lcell_count = nullptr, *d_lcell_start = nullptr;
    float4 *d_lcell_sorted = nullptr;

    // centres
    // per staging slot, so that the centre passes of batch k+1 (pre stream) overlap the bond kernel of batch k
    float *d_est2[2] = {nullptr, nullptr}, *d_center2[2] = {nullptr, nullptr};   // [max_batch*3]
    long long *d_partial2[2] = {nullptr, nullptr};    // [max_batch][kCenterBlocks][2] fixed-point partial sums
    unsigned *d_ticket2[2] = {nullptr, nullptr};   // [max_batch]
    float *d_est = nullptr, *d_center = nullptr;   // current slot's buffers
    long long *d_partial = nullptr;
    unsigned *d_ticket = nullptr;
    cudaStream_t stream_pre = nullptr;             // frame setup + centre reduction of the next batch
    cudaEvent_t ev_pre[2] = {nullptr, nullptr};
    // three-stage pipeline of the speculative path: pre (setup of batch k+1) | main (bond kernels back to back) |
    // post (repair + fold of batch k-1)
    cudaStream_t stream_post = nullptr;
    cudaEvent_t ev_bond[2] = {nullptr, nullptr}, ev_post[2] = {nullptr, nullptr};
    cudaEvent_t ev_post_any = nullptr;             // last batch whose tail ran on the post stream
    bool post_used = false;
    struct SegList { Seg *d = nullptr; int n = 0; };
    SegList seg_membrane[3], seg_geom[3];          // per axis
    // spherical-clustering leaflets (gorder_spherical.cuh)
    bool spherical = false;
    float *d_sph_scratch = nullptr;                       // [max_batch][kSphArrays][sph_pad]
    int sph_pad = 0;                                      // n_membrane rounded up to 4 floats
    unsigned char *d_sph_upper = nullptr;                 // [max_batch][n_membrane]
    int *d_sph_index = nullptr;                           // [n_molpad] position of the molecule's head in the ClusterHeads group
    SegList seg_left;                              // leaflet-axis runs of the membrane atoms the bond kernel does not count (SPEC)
    int spec_left_blocks = 0;
    double *d_spec_left_sum = nullptr;      // [max_batch][spec_left_blocks][2]
    float *d_spec_left_mm = nullptr;        // [max_batch][spec_left_blocks][2]

    // optional event timing of the accumulation kernel
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    double prof_ms = 0.0;
    long long prof_n = 0;
    // ... and of the membrane-normal stage (cell list + PCA kernels), when there is one
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events_n;
    size_t prof_used_n = 0;
    double prof_ms_n = 0.0;
    long long prof_n_n = 0;

    bool fast_ok = false;   // K1f applies (bond_fast_kernel)
    bool ua_fast_ok = false;   // K2f applies (ua_fast_kernel)
    size_t ua_fast_smem = 0;
    int ua_fast_tile = 0, ua_fast_items = 0, ua_fast_orders = 0, ua_fast_nbuf = 0, n_sm = 0;
    int fast_slot = -1;     // slot of this handle's tables in constant memory (c_fast), -1: global tables
    // speculative Global leaflets (bond_order_kernel<SPEC> + spec_repair_kernel)
    bool spec_ok = false, spec_disabled = false, spec_ref_valid = false;
    int spec_cur = 0;
    float *d_spec_ref = nullptr;            // [4] ring of provisional centres: batch k reads [k % 4], writes [(k + 1) % 4]
    double *d_spec_sum = nullptr;           // [max_batch][n_chunks][2]
    float *d_spec_mm = nullptr;             // [max_batch][n_chunks][4]
    unsigned *d_spec_ticket = nullptr;      // [max_batch]
    float *d_spec_center = nullptr;         // [max_batch]
    unsigned char *d_spec_flag = nullptr;   // [max_batch]
    unsigned *h_spec_counters = nullptr, *d_spec_counters = nullptr;   // mapped pinned: frames speculated, frames repaired

    struct GorderXtcDev *xtc_dev = nullptr;   // device-side XTC unpacker (gorder_gpu_run_xtc_device)
    // pinned batches of the trajectory feed (gorder_gpu_run_xtc)
    float *xtc_pin[2] = {nullptr, nullptr}, *xtc_pbox[2] = {nullptr, nullptr};
    int xtc_pin_frames = 0;

    long long n_frames = 0;
    long long n_launches = 0;
    long long last_frame_index = -1;

    // error state
    int err_code = 0;
    long long err_detail = -1;
    std::string err_msg;
    std::recursive_mutex mu;   // every entry point that takes a handle holds it (finish -> sync, run_xtc -> submit nest)

    void set_error(int code, const std::string &msg, long long detail = -1) {
        if (!err_code) { err_code = code; err_msg = msg; err_detail = detail; }
    }
};

namespace {

template <typename T>
int dev_alloc(GorderHandle *h, T **out, size_t count, bool zero = false) {
    *out = nullptr;
    if (count == 0) count = 1;
    void *p = nullptr;
    CK(cudaMalloc(&p, count * sizeof(T)));
    if (zero) CK(cudaMemset(p, 0, count * sizeof(T)));
    h->owned.push_back(p);
    *out = static_cast<T *>(p);
    return GORDER_OK;
}

template <typename T>
int dev_upload(GorderHandle *h, T **out, const std::vector<T> &v) {
    int rc = dev_alloc(h, out, v.size());
    if (rc) return rc;
    if (!v.empty()) CK(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return GORDER_OK;
}

int ua_hydrogens(int kind) { return kind == GORDER_UA_CH3 ? 3 : (kind == GORDER_UA_CH2 ? 2 : 1); }

int round_up(int x, int a) { return (x + a - 1) / a * a; }

bool should_assign(const GorderSetup &s, long long frame) {   // leaflets.rs:435-441
    if (s.leaflet_mode == GORDER_LEAFLET_NONE) return false;
    if (s.leaflet_freq_kind == GORDER_FREQ_ONCE) return frame == 0;
    return frame % (s.leaflet_freq > 0 ? s.leaflet_freq : 1) == 0;
}
long long assignment_frame(const GorderSetup &s, long long frame) {   // leaflets.rs:1438-1473
    if (s.leaflet_freq_kind == GORDER_FREQ_ONCE) return 0;
    long long n = s.leaflet_freq > 0 ? s.leaflet_freq : 1;
    return frame / n * n;
}

int upload_group(GorderHandle *h, GroupRef *g, const int32_t *idx, int n) {
    std::vector<int> off(n), cs(n), slot(n);
    for (int i = 0; i < n; i++) { off[i] = h->slot_off[idx[i]]; cs[i] = h->slot_cs[idx[i]]; slot[i] = idx[i]; }
    int *d_off, *d_cs, *d_slot;
    int rc;
    if ((rc = dev_upload(h, &d_off, off))) return rc;
    if ((rc = dev_upload(h, &d_cs, cs))) return rc;
    if ((rc = dev_upload(h, &d_slot, slot))) return rc;
    g->off = d_off; g->cs = d_cs; g->slot = d_slot; g->n = n;
    return GORDER_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------------
template <int MPT, bool PBC, bool NVEC, bool LEAF, bool EXTRA>
void launch_bond(GorderHandle *h, dim3 grid, size_t smem, const float *planes, const FrameAux *aux, AccumOut o) {
    bond_order_kernel<MPT, PBC, NVEC, LEAF, EXTRA><<<grid, kBlock, smem, h->stream>>>(h->view, planes, aux, h->d_leaf_rows, h->d_normals,
                                                                                       h->d_normal_npoints, o);
}
template <int MPT, bool PBC, bool NVEC>
void launch_bond2(GorderHandle *h, dim3 grid, size_t smem, const float *planes, const FrameAux *aux, AccumOut o) {
    if (h->leaf) { if (h->extra) launch_bond<MPT, PBC, NVEC, true, true>(h, grid, smem, planes, aux, o); else launch_bond<MPT, PBC, NVEC, true, false>(h, grid, smem, planes, aux, o); }
    else { if (h->extra) launch_bond<MPT, PBC, NVEC, false, true>(h, grid, smem, planes, aux, o); else launch_bond<MPT, PBC, NVEC, false, false>(h, grid, smem, planes, aux, o); }
}
template <int MPT>
void launch_bond3(GorderHandle *h, dim3 grid, size_t smem, const float *planes, const FrameAux *aux, AccumOut o) {
    const bool pbc = h->s.handle_pbc != 0;
    if (pbc) { if (h->nvec) launch_bond2<MPT, true, true>(h, grid, smem, planes, aux, o); else launch_bond2<MPT, true, false>(h, grid, smem, planes, aux, o); }
    else { if (h->nvec) launch_bond2<MPT, false, true>(h, grid, smem, planes, aux, o); else launch_bond2<MPT, false, false>(h, grid, smem, planes, aux, o); }
}

template <int MPT>
void launch_bond_spec(GorderHandle *h, dim3 grid, size_t smem, const float *planes, const FrameAux *aux, AccumOut o) {
    bond_order_kernel<MPT, true, false, true, false, true><<<grid, kBlock, smem, h->stream>>>(h->view, planes, aux, h->d_leaf_rows, h->d_normals,
                                                                                             h->d_normal_npoints, o);
}
// K1f (gorder_fast.cuh): PBC, static normal, no geometry / maps, 2 or 4 molecules per lane
template <int NP, int BLOCK = kBlock>
void launch_fast(GorderHandle *h, dim3 grid, size_t smem, const float *planes, const FrameAux *aux, AccumOut o, bool spec) {
    if (!h->leaf) bond_fast_kernel<NP, false, false, BLOCK><<<grid, BLOCK, smem, h->stream>>>(h->view, planes, aux, h->d_leaf_rows, o, h->fast_slot);
    else if (spec) bond_fast_kernel<NP, true, true, BLOCK><<<grid, BLOCK, smem, h->stream>>>(h->view, planes, aux, h->d_leaf_rows, o, h->fast_slot);
    else bond_fast_kernel<NP, true, false, BLOCK><<<grid, BLOCK, smem, h->stream>>>(h->view, planes, aux, h->d_leaf_rows, o, h->fast_slot);
}

template <int MPT>
void launch_repair(GorderHandle *h, cudaStream_t st, dim3 grid, const RepairParams &rp, const float *planes, const FrameAux *aux, AccumOut o, int nf) {
    spec_repair_kernel<MPT><<<grid, kBlock, 0, st>>>(h->view, rp, planes, aux, o, nf);
}

template <bool PBC, bool NVEC>
void launch_ua2(GorderHandle *h, dim3 grid, size_t smem, const float *planes, const FrameAux *aux, AccumOut o) {
#define UA_L(L, E) ua_order_kernel<PBC, NVEC, L, E><<<grid, kBlock, smem, h->stream>>>(h->view, planes, aux, h->d_leaf_rows, h->d_normals, h->d_normal_npoints, o)
    if (h->leaf) { if (h->extra) UA_L(true, true); else UA_L(true, false); }
    else { if (h->extra) UA_L(false, true); else UA_L(false, false); }
#undef UA_L
}
void launch_ua(GorderHandle *h, dim3 grid, size_t smem, const float *planes, const FrameAux *aux, AccumOut o) {
    const bool pbc = h->s.handle_pbc != 0;
    if (pbc) { if (h->nvec) launch_ua2<true, true>(h, grid, smem, planes, aux, o); else launch_ua2<true, false>(h, grid, smem, planes, aux, o); }
    else { if (h->nvec) launch_ua2<false, true>(h, grid, smem, planes, aux, o); else launch_ua2<false, false>(h, grid, smem, planes, aux, o); }
}

size_t accum_smem(const GorderHandle *h) {
    int max_items = 0, max_orders = 0;
    for (auto &t : h->types) { max_items = std::max(max_items, t.n_items); max_orders = std::max(max_orders, t.n_orders); }
    const int na = h->leaf ? (h->extra ? 3 : 2) : (h->extra ? 2 : 1);
    const size_t items = h->ua ? (size_t)8 * max_items : (size_t)2 * max_items;
    return (items + (size_t)kWarps * max_orders * na + 2) * sizeof(int);
}

// runs of contiguous native floats of component `axis` of a group, split into pieces of <= 4096
int build_segs(GorderHandle *h, GorderHandle::SegList *out, const int32_t *idx, int n, int axis) {
    std::vector<long long> offs(n);
    for (int i = 0; i < n; i++) offs[i] = (long long)h->slot_off[idx[i]] + (long long)axis * h->slot_cs[idx[i]];
    std::sort(offs.begin(), offs.end());
    std::vector<Seg> segs;
    for (int i = 0; i < n;) {
        int j = i + 1;
        while (j < n && offs[j] == offs[j - 1] + 1 && j - i < 4096) j++;
        segs.push_back(Seg{(int)offs[i], j - i});
        i = j;
    }
    out->n = (int)segs.size();
    return dev_upload(h, &out->d, segs);
}

// centre of a group along the axes in `axis_mask` for the frames in d_list[0..n_list) -> h->d_center[3*i + axis]
int run_group_center(GorderHandle *h, cudaStream_t st, const GorderHandle::SegList *segs, int n_group, int axis_mask, const float *planes,
                     const FrameAux *aux, const int *d_list, int n_list) {
    const bool pbc = h->s.handle_pbc != 0;
    // sub-batches whose axis planes fit comfortably in L2 (126 MB): pass 1 re-reads what pass 0 just streamed
    long long sub = n_list;   // (sub-batching for L2 reuse of the axis planes was measured slower: small kernels, see profiles/README.md)
    if (h->sw.center_sub >= 1) sub = h->sw.center_sub;
    for (int axis = 0; axis < 3; axis++) {
        if (!(axis_mask & (1 << axis))) continue;
        const int want = (h->sw.center_blocks >= 1 && h->sw.center_blocks <= kCenterBlocks) ? h->sw.center_blocks : kCenterBlocks;
        const int nblk = std::max(1, std::min(want, segs[axis].n));
        for (int l0 = 0; l0 < n_list; l0 += (int)sub) {
            const int nl = std::min<int>((int)sub, n_list - l0);
            dim3 grid(nblk, nl);
            for (int pass = 0; pass < (pbc ? 2 : 1); pass++) {
                center_axis_kernel<<<grid, 256, 0, st>>>(h->view, segs[axis].d, segs[axis].n, n_group, axis, planes, aux, d_list + l0, h->d_est + 3 * l0,
                                                         h->d_center + 3 * l0, h->d_partial + (size_t)l0 * kCenterBlocks * 2, h->d_ticket + l0, pass);
                h->n_launches++;
            }
        }
    }
    CK(cudaGetLastError());
    return GORDER_OK;
}

int grow_rows(GorderHandle *h, long long need) {
    if (!h->s.timewise || need <= h->tw_cap) return GORDER_OK;
    long long cap = std::max<long long>(need, h->tw_cap * 2);
    cap = std::max<long long>(cap, 1024);
    const size_t row = (size_t)h->n_slots * 3;
    long long *ns = nullptr;
    unsigned long long *nc = nullptr;
    CK(cudaMalloc((void **)&ns, cap * row * sizeof(long long)));
    if (cudaError_t e = cudaMalloc((void **)&nc, cap * row * sizeof(unsigned long long))) {
        cudaFree(ns);
        h->set_error(e == cudaErrorMemoryAllocation ? GORDER_ERR_OUT_OF_MEMORY : GORDER_ERR_CUDA, std::string("per-frame rows: ") + cudaGetErrorString(e));
        return h->err_code;
    }
    CK(cudaMemsetAsync(ns, 0, cap * row * sizeof(long long), h->stream));
    CK(cudaMemsetAsync(nc, 0, cap * row * sizeof(unsigned long long), h->stream));
    if (h->d_bsum) {
        CK(cudaStreamSynchronize(h->stream_post));
        CK(cudaStreamSynchronize(h->stream_pre));
        CK(cudaMemcpyAsync(ns, h->d_bsum, h->n_frames * row * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(nc, h->d_bcnt, h->n_frames * row * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_bsum); cudaFree(h->d_bcnt);
    }
    h->d_bsum = ns; h->d_bcnt = nc; h->tw_cap = cap;
    return GORDER_OK;
}

template <typename T>
int grow_collect(GorderHandle *h, T **buf, long long *cap, long long used_rows, long long need_rows, size_t row_elems) {
    if (need_rows <= *cap) return GORDER_OK;
    long long ncap = std::max<long long>(need_rows, *cap * 2);
    ncap = std::max<long long>(ncap, 64);
    T *nb = nullptr;
    CK(cudaMalloc((void **)&nb, ncap * row_elems * sizeof(T)));
    if (*buf) {
        CK(cudaMemcpyAsync(nb, *buf, used_rows * row_elems * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(*buf);
    }
    *buf = nb; *cap = ncap;
    return GORDER_OK;
}

// Analyse one batch whose frames are resident on the device in the native layout.
int process_batch(GorderHandle *h, const float *d_planes, const float *d_box, const long long *frame_index, int nf, int slot,
                  bool planes_on_main) {
    const GorderSetup &s = h->s;
    FrameAux *ha = h->h_aux[slot];
    int *list_assign = h->h_list[slot], *list_all = h->h_list[slot] + h->max_batch;
    int n_assign = 0, last_row = 0;
    const long long row0 = s.timewise ? h->n_frames : 0;
    long long last_fi = h->last_frame_index;   // committed with n_frames once the batch is queued
    for (int f = 0; f < nf; f++) {
        FrameAux &a = ha[f];
        memset(&a, 0, sizeof(a));
        const long long fi = frame_index[f];
        if (fi <= last_fi && h->n_frames + f > 0) {
            h->set_error(GORDER_ERR_INVALID_ARGUMENT, "frame_index must be strictly increasing");
            return h->err_code;
        }
        last_fi = fi;
        a.frame_index = fi;
        a.tw_row = (int)(row0 + f);
        a.manual_norm_row = (int)(fi / (s.step > 0 ? s.step : 1));
        a.leaf_row = -1;
        if (h->leaf) {
            if (should_assign(s, fi)) { list_assign[n_assign] = f; last_row = 1 + n_assign; n_assign++; }
            else if (last_row == 0 && (!h->have_leaflets || h->cur_leaflet_frame != assignment_frame(s, fi))) {
                h->set_error(GORDER_ERR_LEAFLET_FRAME_UNAVAILABLE, "leaflet assignment frame not held by this handle", fi);
                return h->err_code;
            }
            a.leaf_row = last_row;
        }
        list_all[f] = f;
    }
    FrameAux *da = h->d_aux[slot];
    int *dl_assign = h->d_list[slot], *dl_all = h->d_list[slot] + h->max_batch;
    h->d_est = h->d_est2[slot]; h->d_center = h->d_center2[slot]; h->d_partial = h->d_partial2[slot]; h->d_ticket = h->d_ticket2[slot];
    // Global leaflets on every analysed frame (AA/CG): the bond kernel classifies inline, no table pass
    const bool inline_leaf = !h->ua && s.leaflet_mode == GORDER_LEAFLET_GLOBAL && s.leaflet_freq_kind == GORDER_FREQ_EVERY &&
                             s.leaflet_freq <= std::max(1, s.step) && n_assign == nf && !h->sw.no_inline_leaflets;
    // Frames already resident on the device: frame setup and the centre passes of this batch go to the
    // pre stream and overlap the bond kernel of the previous batch (both are latency-, not HBM-bound).
    // ... and without any centre pre-pass when the speculative path applies (AccumOut::spec_*)
    if (h->spec_ok && !h->spec_disabled) {   // frames that needed the exact centre so far (mapped counters, no sync)
        const unsigned checked = h->h_spec_counters[0], repaired = h->h_spec_counters[1];
        if (checked >= 16 && repaired * 8ull > checked) h->spec_disabled = true;   // thick membrane / thin water: pre-pass is cheaper
    }
    const bool spec = inline_leaf && h->spec_ok && !h->spec_disabled;
    const bool overlap = inline_leaf && !spec && !planes_on_main && !h->sw.no_overlap;   // (with per-molecule normals too: they are computed on the main stream, in order)
    // speculative path on resident frames: the bond kernels of consecutive batches run back to back on the main stream,
    // the setup of the next batch and the tail (repair + fold) of the previous one run beside them
    // ... and so do runs without leaflets (nothing of a batch depends on the previous one but the totals, which only the
    // post stream touches): for small systems the setup and the fold are a quarter of a batch (S-UA, S-AA-small)
    const bool plain = !h->leaf && !h->nvec;
    const bool pipelined = (spec || plain) && !planes_on_main && !s.collect_leaflets && !h->sw.no_overlap;
    cudaStream_t sp = (overlap || pipelined) ? h->stream_pre : h->stream;
    cudaStream_t spost = pipelined ? h->stream_post : h->stream;
    if (!pipelined && h->post_used) CK(cudaStreamWaitEvent(h->stream, h->ev_post_any, 0));   // totals are touched by one stream at a time
    long long *bsum = s.timewise ? h->d_bsum : h->d_ring[slot];
    unsigned long long *bcnt = s.timewise ? h->d_bcnt : reinterpret_cast<unsigned long long *>(h->d_ring[slot] + h->ring_words);
    unsigned *nflag = s.timewise ? h->d_nflag_tw + slot : reinterpret_cast<unsigned *>(h->d_ring[slot] + 2 * h->ring_words);
    CK(cudaMemcpyAsync(da, ha, sizeof(FrameAux) * nf, cudaMemcpyHostToDevice, sp));
    CK(cudaMemcpyAsync(h->d_list[slot], h->h_list[slot], sizeof(int) * (h->max_batch + nf), cudaMemcpyHostToDevice, sp));

    frame_setup_kernel<<<(nf + 63) / 64, 64, 0, sp>>>(h->view, da, d_box, nf, 0);
    h->n_launches++;
    if (s.geom_kind != GORDER_GEOM_NONE) {
        if (s.geom_ref_kind == GORDER_GEOMREF_SELECTION) {
            int rc = run_group_center(h, sp, h->seg_geom, h->s.n_geom_ref, 7, d_planes, da, dl_all, nf);
            if (rc) return rc;
            store_center_kernel<<<(nf + 63) / 64, 64, 0, sp>>>(da, dl_all, nf, h->d_center);
            h->n_launches++;
        }
        frame_setup_kernel<<<(nf + 63) / 64, 64, 0, sp>>>(h->view, da, d_box, nf, 1);
        h->n_launches++;
    }
    if (spec) {
        if (!h->spec_ref_valid) {   // first batch: the provisional centre is the exact centre of its first frame
            int rc = run_group_center(h, sp, h->seg_membrane, h->s.n_membrane, 1 << s.leaflet_axis, d_planes, da, dl_assign, 1);
            if (rc) return rc;
            CK(cudaMemcpyAsync(h->d_spec_ref + h->spec_cur % 4, h->d_center + s.leaflet_axis, sizeof(float), cudaMemcpyDeviceToDevice, sp));
            h->spec_ref_valid = true;
        }
    } else if (inline_leaf) {
        int rc = run_group_center(h, sp, h->seg_membrane, h->s.n_membrane, 1 << s.leaflet_axis, d_planes, da, dl_assign, n_assign);
        if (rc) return rc;
    } else if (h->leaf && n_assign > 0) {
        if (s.leaflet_mode == GORDER_LEAFLET_GLOBAL) {
            int rc = run_group_center(h, sp, h->seg_membrane, h->s.n_membrane, 1 << s.leaflet_axis, d_planes, da, dl_assign, n_assign);
            if (rc) return rc;
        }
        if (h->use_lcells) {
            const int nm = s.n_membrane;
            CK(cudaMemsetAsync(h->d_lcell_count, 0, (size_t)n_assign * h->lcells_cap * sizeof(int), sp));
            dim3 gm((nm + 255) / 256, n_assign);
            lcell_count_kernel<<<gm, 256, 0, sp>>>(h->view, d_planes, da, dl_assign, h->d_matom_cell, h->d_lcell_count, h->lcells_cap);
            lcell_scan_kernel<<<n_assign, 1024, 0, sp>>>(h->view, da, dl_assign, h->d_lcell_count, h->d_lcell_start, h->lcells_cap);
            lcell_fill_kernel<<<gm, 256, 0, sp>>>(h->view, d_planes, dl_assign, h->d_matom_cell, h->d_lcell_count, h->d_lcell_start, h->d_lcell_sorted, h->lcells_cap);
            h->n_launches += 4;
        }
        dim3 grid((h->n_molpad + 255) / 256, n_assign);
        if (h->spherical) {   // gorder_spherical.cuh
            int rc = run_group_center(h, sp, h->seg_membrane, h->s.n_membrane, 7, d_planes, da, dl_assign, n_assign);
            if (rc) return rc;
            spherical_cluster_kernel<<<n_assign, kSphThreads, 0, sp>>>(h->view, d_planes, da, dl_assign, h->d_center, h->d_sph_scratch, h->sph_pad, h->d_sph_upper);
            spherical_assign_kernel<<<grid, 256, 0, sp>>>(h->view, h->d_molpad_type, h->d_sph_index, h->d_sph_upper, h->d_leaf_rows);
            h->n_launches++;
        } else
            leaflet_assign_kernel<<<grid, 256, 0, sp>>>(h->view, d_planes, da, dl_assign, h->d_center, h->d_molpad_type, h->d_leaf_rows,
                                                         h->use_lcells ? h->d_lcell_start : nullptr, h->d_lcell_sorted, h->lcells_cap);
        h->n_launches++;
    }
    // per-frame accumulator rows
    const size_t row = (size_t)h->n_slots * 3;
    if (s.timewise) {
        int rc = grow_rows(h, h->n_frames + nf);
        if (rc) return rc;
        bsum = h->d_bsum; bcnt = h->d_bcnt;
        if (spec) CK(cudaMemsetAsync(nflag, 0, sizeof(unsigned), sp));
    } else {
        CK(cudaMemsetAsync(bsum, 0, (2 * h->ring_words + 2) * sizeof(long long), sp));   // includes the flagged-frame counter
    }
    if (overlap || pipelined) {
        CK(cudaEventRecord(h->ev_pre[slot], sp));
        CK(cudaStreamWaitEvent(h->stream, h->ev_pre[slot], 0));
    }
    std::pair<cudaEvent_t, cudaEvent_t> *pn = nullptr;
    if (h->nvec && h->profiling) {
        if (h->prof_used_n == h->prof_events_n.size()) {
            std::pair<cudaEvent_t, cudaEvent_t> e;
            CK(cudaEventCreate(&e.first)); CK(cudaEventCreate(&e.second));
            h->prof_events_n.push_back(e);
        }
        pn = &h->prof_events_n[h->prof_used_n++];
        CK(cudaEventRecord(pn->first, h->stream));
    }
    if (h->nvec) {
        if (s.normal_mode == GORDER_NORMAL_DYNAMIC && h->use_cells) {
            const int nh = s.n_normal_heads;
            CK(cudaMemsetAsync(h->d_cell_count, 0, (size_t)nf * h->cells_cap * sizeof(int), h->stream));
            dim3 gh((nh + 255) / 256, nf);
            cell_count_kernel<<<gh, 256, 0, h->stream>>>(h->view, d_planes, da, h->d_head_cell, h->d_cell_count, h->cells_cap);
            cell_span_sum_kernel<<<dim3(kScanBlocks, nf), 1024, 0, h->stream>>>(h->view, da, h->d_cell_count, h->d_cell_span, h->cells_cap);
            cell_scan_kernel<<<dim3(kScanBlocks, nf), 1024, 0, h->stream>>>(h->view, da, h->d_cell_count, h->d_cell_start, h->d_cell_span, h->cells_cap);
            cell_fill_kernel<<<gh, 256, 0, h->stream>>>(h->view, d_planes, da, h->d_head_cell, h->d_cell_count, h->d_cell_start, h->d_cell_sorted, h->cells_cap);
            if (h->normals_sorted) {
                // lanes walk the cell-sorted heads; molecules without a head in the list (padding) keep NaN
                CK(cudaMemsetAsync(h->d_normals, 0xff, (size_t)nf * 3 * h->n_molpad * sizeof(float), h->stream));
                dim3 grid((nh + 127) / 128, nf);
                dynamic_normal_sorted_kernel<<<grid, 128, 0, h->stream>>>(h->view, da, h->d_head_molpad, h->d_cell_start, h->d_cell_sorted,
                                                                         h->cells_cap, h->d_normals, h->d_normal_npoints);
            } else {
                dim3 grid((h->n_molpad + 127) / 128, nf);
                dynamic_normal_cell_kernel<<<grid, 128, 0, h->stream>>>(h->view, d_planes, da, h->d_molpad_type, h->d_cell_start, h->d_cell_sorted,
                                                                       h->cells_cap, h->d_normals, h->d_normal_npoints);
            }
            h->n_launches += 3;
        } else if (s.normal_mode == GORDER_NORMAL_DYNAMIC) {
            dim3 grid((h->n_molpad + 127) / 128, nf);
            dynamic_normal_kernel<<<grid, 128, 0, h->stream>>>(h->view, d_planes, da, h->d_molpad_type, h->d_normals, h->d_normal_npoints);
        } else {
            dim3 grid((h->n_molpad + 255) / 256, nf);
            manual_normal_kernel<<<grid, 256, 0, h->stream>>>(h->view, da, h->d_molpad_type, h->d_normals);
        }
        h->n_launches++;
        if (h->d_normal_used) CK(cudaMemsetAsync(h->d_normal_used, 0, (size_t)nf * h->n_molpad, h->stream));
        if (pn) CK(cudaEventRecord(pn->second, h->stream));
    }
    AccumOut o{};
    o.inline_center = (inline_leaf && !spec) ? h->d_center : nullptr;
    if (spec) {
        o.spec_ref = h->d_spec_ref + h->spec_cur % 4; o.spec_ref_next = h->d_spec_ref + (h->spec_cur + 1) % 4;
        o.spec_sum = h->d_spec_sum; o.spec_mm = h->d_spec_mm; o.spec_ticket = h->d_spec_ticket; o.spec_center = h->d_spec_center + (size_t)slot * h->max_batch;
        o.spec_flag = h->d_spec_flag + (size_t)slot * h->max_batch; o.spec_nflag = nflag; o.n_membrane = s.n_membrane;
        o.spec_left_sum = h->d_spec_left_sum; o.spec_left_mm = h->d_spec_left_mm; o.spec_left_parts = h->spec_left_blocks;
        if (h->spec_left_blocks) {   // on the main stream: the provisional centre comes from the previous batch's bond kernel
            spec_leftover_kernel<<<dim3(h->spec_left_blocks, nf), 256, 0, h->stream>>>(h->view, h->seg_left.d, h->seg_left.n, d_planes, da, o.spec_ref,
                                                                                      h->d_spec_left_sum, h->d_spec_left_mm);
            h->n_launches++;
        }
    }
    o.leaf_out = (inline_leaf && s.collect_leaflets) ? h->d_leaf_rows : nullptr;
    o.bsum = bsum; o.bcnt = bcnt; o.map_sum = h->d_map_sum; o.map_cnt = h->d_map_cnt; o.normal_used = h->d_normal_used;
    dim3 grid(h->n_chunks, nf);
    const size_t smem = accum_smem(h);
    std::pair<cudaEvent_t, cudaEvent_t> *pe = nullptr;
    if (h->profiling) {
        if (h->prof_used == h->prof_events.size()) {
            std::pair<cudaEvent_t, cudaEvent_t> e;
            CK(cudaEventCreate(&e.first)); CK(cudaEventCreate(&e.second));
            h->prof_events.push_back(e);
        }
        pe = &h->prof_events[h->prof_used++];
        CK(cudaEventRecord(pe->first, h->stream));
    }
    if (h->fast_ok && !h->sw.no_fast) {
        if (h->mpt == 4) launch_fast<2>(h, grid, smem, d_planes, da, o, spec);
        else if (h->mpt == 2) launch_fast<1>(h, grid, smem, d_planes, da, o, spec);
        else launch_fast<2, 64>(h, grid, smem, d_planes, da, o, spec);   // one tile of 256 molecules: 64 lanes x 4 molecules
    } else if (spec) {
        if (h->mpt == 4) launch_bond_spec<4>(h, grid, smem, d_planes, da, o);
        else if (h->mpt == 2) launch_bond_spec<2>(h, grid, smem, d_planes, da, o);
        else launch_bond_spec<1>(h, grid, smem, d_planes, da, o);
    } else if (h->ua_fast_ok && (reinterpret_cast<uintptr_t>(d_planes) & 15) == 0) {   // bulk copies need 16-byte aligned tiles
        const int ctas = (int)std::min<long long>((long long)h->n_chunks * nf, h->n_sm);   // persistent: one CTA per SM
        if (h->leaf) ua_fast_kernel<true><<<ctas, kUaThreads, h->ua_fast_smem, h->stream>>>(h->view, d_planes, da, h->d_leaf_rows, o, nf, h->ua_fast_tile, h->ua_fast_items, h->ua_fast_orders, h->ua_fast_nbuf);
        else ua_fast_kernel<false><<<ctas, kUaThreads, h->ua_fast_smem, h->stream>>>(h->view, d_planes, da, h->d_leaf_rows, o, nf, h->ua_fast_tile, h->ua_fast_items, h->ua_fast_orders, h->ua_fast_nbuf);
    } else if (h->ua) launch_ua(h, grid, smem, d_planes, da, o);
    else if (h->mpt == 4) launch_bond3<4>(h, grid, smem, d_planes, da, o);
    else if (h->mpt == 2) launch_bond3<2>(h, grid, smem, d_planes, da, o);
    else launch_bond3<1>(h, grid, smem, d_planes, da, o);
    h->n_launches++;
    if (pe) CK(cudaEventRecord(pe->second, h->stream));
    if (pipelined) {
        CK(cudaEventRecord(h->ev_bond[slot], h->stream));
        CK(cudaStreamWaitEvent(spost, h->ev_bond[slot], 0));
    }
    if (spec) {   // frames whose leaflets are not provably those of the exact centre (normally none: the kernel returns at once)
        RepairParams rp;
        rp.segs = h->seg_membrane[s.leaflet_axis].d; rp.n_segs = h->seg_membrane[s.leaflet_axis].n;
        rp.n_blocks = std::max(1, std::min(kCenterBlocks, rp.n_segs)); rp.n_group = s.n_membrane;
        rp.ref = o.spec_ref; rp.flag = o.spec_flag; rp.nflag = nflag; rp.center = h->d_spec_center + (size_t)slot * h->max_batch;
        rp.host_counters = h->d_spec_counters; rp.leaf_out = o.leaf_out;
        dim3 rg(h->n_chunks, std::min(nf, 4));
        if (h->mpt == 4) launch_repair<4>(h, spost, rg, rp, d_planes, da, o, nf);
        else if (h->mpt == 2) launch_repair<2>(h, spost, rg, rp, d_planes, da, o, nf);
        else launch_repair<1>(h, spost, rg, rp, d_planes, da, o, nf);
        h->n_launches++;
        h->spec_cur = (h->spec_cur + 1) % 4;
    }
    if (h->n_slots > 0)
        fold_kernel<<<h->n_slots, 128, 0, spost>>>(h->n_slots, nf, h->leaf ? 1 : 0, bsum + row0 * row, bcnt + row0 * row, h->d_tot_sum, h->d_tot_cnt);
    h->n_launches++;
    CK(cudaGetLastError());
    if (pipelined) {
        CK(cudaEventRecord(h->ev_post[slot], spost));
        CK(cudaEventRecord(h->ev_post_any, spost));
        h->post_used = true;
    }
    if (h->leaf && n_assign > 0 && s.collect_leaflets) {
        int rc = grow_collect(h, &h->d_leaf_collect, &h->leaf_collect_cap, h->n_leaf_collected, h->n_leaf_collected + n_assign, (size_t)h->n_molpad);
        if (rc) return rc;
        if (inline_leaf) {   // rows 1 + f were written by the accumulation kernel for every frame f
            CK(cudaMemcpyAsync(h->d_leaf_collect + (size_t)h->n_leaf_collected * h->n_molpad, h->d_leaf_rows + h->n_molpad,
                               (size_t)nf * h->n_molpad, cudaMemcpyDeviceToDevice, h->stream));
        } else {
            CK(cudaMemcpyAsync(h->d_leaf_collect + (size_t)h->n_leaf_collected * h->n_molpad, h->d_leaf_rows + h->n_molpad,
                               (size_t)n_assign * h->n_molpad, cudaMemcpyDeviceToDevice, h->stream));
        }
        for (int a = 0; a < n_assign; a++) h->leaf_frame_index.push_back(frame_index[list_assign[a]]);
        h->n_leaf_collected += n_assign;
    }
    if (h->leaf && n_assign > 0 && !inline_leaf) {   // keep the newest table for the frames of the next batch
        CK(cudaMemcpyAsync(h->d_leaf_rows, h->d_leaf_rows + (size_t)n_assign * h->n_molpad, h->n_molpad, cudaMemcpyDeviceToDevice, h->stream));
        h->have_leaflets = true;
        h->cur_leaflet_frame = frame_index[list_assign[n_assign - 1]];
    }
    if (s.collect_normals && s.normal_mode == GORDER_NORMAL_DYNAMIC) {
        // normals that were never requested are exported as NaN (normal.rs:211-227)
        if (h->d_normal_used) {
            dim3 g2((h->n_molpad + 255) / 256, nf);
            mask_normals_kernel<<<g2, 256, 0, h->stream>>>(h->n_molpad, h->d_normal_used, h->d_normals);
            h->n_launches++;
        }
        int rc = grow_collect(h, &h->d_normals_collect, &h->normals_collect_cap, h->n_frames, h->n_frames + nf, (size_t)3 * h->n_molpad);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h->d_normals_collect + (size_t)h->n_frames * 3 * h->n_molpad, h->d_normals, (size_t)nf * 3 * h->n_molpad * sizeof(float),
                           cudaMemcpyDeviceToDevice, h->stream));
    }
    for (int f = 0; f < nf; f++) h->frame_index_done.push_back(frame_index[f]);
    h->n_frames += nf;
    h->last_frame_index = last_fi;
    return GORDER_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char *gorder_gpu_version(void) { return "gorder-b200 0.1.0 (sm_100a)"; }

void gorder_gpu_destroy(GorderHandle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    if (h->stream_pre) cudaStreamSynchronize(h->stream_pre);
    if (h->stream_post) cudaStreamSynchronize(h->stream_post);
    for (void *p : h->owned) cudaFree(p);
    cudaFree(h->d_bsum); cudaFree(h->d_bcnt);
    cudaFree(h->d_leaf_collect); cudaFree(h->d_normals_collect); cudaFree(h->d_used_collect);
    for (int i = 0; i < 2; i++) {
        cudaFree(h->d_xyz[i]);
        if (h->h_aux[i]) cudaFreeHost(h->h_aux[i]);
        if (h->h_list[i]) cudaFreeHost(h->h_list[i]);
        if (h->ev_stage_free[i]) cudaEventDestroy(h->ev_stage_free[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    }
    if (h->h_spec_counters) cudaFreeHost(h->h_spec_counters);
    for (int i = 0; i < 2; i++) { if (h->xtc_pin[i]) cudaFreeHost(h->xtc_pin[i]); if (h->xtc_pbox[i]) cudaFreeHost(h->xtc_pbox[i]); }
    if (h->xtc_dev) gorder_xtc_dev_free(h->xtc_dev);
    if (h->fast_slot >= 0) {
        std::lock_guard<std::mutex> lock(g_fast_mu);
        g_fast_used[h->device][h->fast_slot] = false;
    }
    for (auto &e : h->prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : h->prof_events_n) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream_pre) cudaStreamDestroy(h->stream_pre);
    if (h->stream_post) cudaStreamDestroy(h->stream_post);
    for (int i = 0; i < 2; i++) {
        if (h->ev_pre[i]) cudaEventDestroy(h->ev_pre[i]);
        if (h->ev_bond[i]) cudaEventDestroy(h->ev_bond[i]);
        if (h->ev_post[i]) cudaEventDestroy(h->ev_post[i]);
    }
    if (h->ev_post_any) cudaEventDestroy(h->ev_post_any);
    delete h;
}

static int create_impl(const GorderSetup *s, GorderHandle *h) {
    h->s = *s;
    h->sw = Switches::from_env();
    const bool ua = s->kind == GORDER_KIND_UA;
    h->ua = ua;
    h->leaf = s->leaflet_mode != GORDER_LEAFLET_NONE;
    h->extra = s->geom_kind != GORDER_GEOM_NONE || s->map_enabled;
    h->nvec = s->normal_mode != GORDER_NORMAL_STATIC;
    if (s->n_atoms <= 0 || s->n_moltypes < 0) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "bad n_atoms / n_moltypes"); return h->err_code; }
    {   // enumerators and counts: a value outside its range would index past a 3-vector or skip every branch on the device
        auto in = [](int v, int lo, int hi) { return v >= lo && v <= hi; };
        const char *what = nullptr;
        if (!in(s->kind, GORDER_KIND_AA, GORDER_KIND_UA)) what = "kind";
        else if (!in(s->normal_mode, GORDER_NORMAL_STATIC, GORDER_NORMAL_MANUAL)) what = "normal_mode";
        else if (!in(s->normal_axis, GORDER_AXIS_X, GORDER_AXIS_Z)) what = "normal_axis";
        else if (!in(s->leaflet_axis, GORDER_AXIS_X, GORDER_AXIS_Z)) what = "leaflet_axis";
        else if (!in(s->leaflet_freq_kind, GORDER_FREQ_EVERY, GORDER_FREQ_ONCE)) what = "leaflet_freq_kind";
        else if (!in(s->geom_kind, GORDER_GEOM_NONE, GORDER_GEOM_SPHERE)) what = "geom_kind";
        else if (!in(s->geom_ref_kind, GORDER_GEOMREF_POINT, GORDER_GEOMREF_BOX_CENTER)) what = "geom_ref_kind";
        else if (!in(s->geom_axis, GORDER_AXIS_X, GORDER_AXIS_Z)) what = "geom_axis";
        else if (s->map_enabled && !in(s->map_plane, GORDER_PLANE_XY, GORDER_PLANE_YZ)) what = "map_plane";
        else if (s->n_membrane < 0 || s->n_geom_ref < 0 || s->n_normal_heads < 0 || (s->n_moltypes > 0 && !s->moltypes)) what = "group sizes / moltypes";
        else if (s->normal_mode == GORDER_NORMAL_DYNAMIC && !(s->dynamic_radius > 0.0f)) what = "dynamic_radius";
        else if (s->leaflet_mode == GORDER_LEAFLET_LOCAL && !(s->leaflet_radius > 0.0f)) what = "leaflet_radius";
        if (what) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, (std::string("value out of range: ") + what).c_str()); return h->err_code; }
    }
    h->spherical = s->leaflet_mode == GORDER_LEAFLET_SPHERICAL;
    if (s->leaflet_mode < GORDER_LEAFLET_NONE || s->leaflet_mode > GORDER_LEAFLET_SPHERICAL) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "unknown leaflet mode"); return h->err_code; }
    if (h->spherical && (s->n_membrane < 2 || !s->membrane)) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "spherical clustering needs the ClusterHeads group in `membrane`"); return h->err_code; }

    for (int t = 0; t < s->n_moltypes; t++) {   // the classifiers read the head (and the methyls) of every molecule: a missing one would index before the type's planes
        const GorderMolType &m = s->moltypes[t];
        const int lm = s->leaflet_mode;
        const bool needs_head = lm == GORDER_LEAFLET_GLOBAL || lm == GORDER_LEAFLET_LOCAL || lm == GORDER_LEAFLET_INDIVIDUAL || lm == GORDER_LEAFLET_SPHERICAL;
        if (needs_head && m.head_rel < 0) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "leaflet classification needs head_rel >= 0 for every molecule type", t); return h->err_code; }
        if (m.n_methyls < 0 || (m.n_methyls > 0 && !m.methyl_rel)) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "bad methyl table", t); return h->err_code; }
        for (int k = 0; k < m.n_methyls; k++) if (m.methyl_rel[k] < 0) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "negative methyl relative index", t); return h->err_code; }
        if (lm == GORDER_LEAFLET_INDIVIDUAL && m.n_methyls == 0) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "individual leaflet classification needs methyls", t); return h->err_code; }
        if (lm == GORDER_LEAFLET_MANUAL && (!m.manual_leaflets || m.n_manual_leaflet_frames <= 0)) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "manual leaflet classification needs a table", t); return h->err_code; }
    }

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || s->device < 0 || s->device >= n_dev) {
        cudaGetLastError();
        h->set_error(GORDER_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
        return h->err_code;
    }
    h->device = s->device;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));

    // molecules per thread of the bond kernel (tunable: GORDER_MPT)
    int max_mol = 0;
    for (int t = 0; t < s->n_moltypes; t++) max_mol = std::max(max_mol, s->moltypes[t].n_molecules);
    // a batch supplies the parallelism (grid = tiles x frames), so the vector width only has to leave most molecules in
    // full tiles (kBlock * mpt molecules: the fast kernel's unit; a partial last tile runs the generic body)
    h->mpt = (max_mol >= 8 * kBlock * 4 || (max_mol >= kBlock * 4 && max_mol % (kBlock * 4) == 0)) ? 4
           : (max_mol >= 8 * kBlock * 2 || (max_mol >= kBlock * 2 && max_mol % (kBlock * 2) == 0)) ? 2 : 1;
    // geometry / maps: two molecules per lane (S-AA-large: 5.1e10 -> 6.0e10 samples/s in the step; the four-fold unrolled
    // filter + map code does not fit 64 registers and its CTAs are too coarse for the single wave they run in)
    if (h->extra && h->mpt == 4) h->mpt = 2;
    if (h->sw.mpt == 1 || h->sw.mpt == 2 || h->sw.mpt == 4) h->mpt = h->sw.mpt;
    if (ua) h->mpt = 1;

    // ---- native layout -------------------------------------------------------------------------
    h->slot_off.assign(s->n_atoms, -1);
    h->slot_cs.assign(s->n_atoms, 0);
    std::vector<BondItem> bonds;
    std::vector<UAItem> uas;
    std::vector<int> methyl_offs;
    std::vector<Chunk> chunks;
    std::vector<unsigned char> manual_leaf;
    std::vector<float> manual_norm;
    long long off = 0;
    int slot = 0, molpad = 0, mol = 0;
    auto bad_rel = [&](int r) { return r < 0; };
    // speculative Global leaflets sum every membrane atom's displacement exactly once: whole planes (relative atom u
    // of ALL molecules of a type) that some bond loads are counted in the bond kernel's registers -- the first bond
    // item that loads such a plane carries the "count it" bit -- and the rest goes through spec_leftover_kernel.
    std::vector<char> mem_done(s->n_atoms, 0);
    std::vector<char> in_mem(s->n_atoms, 0);
    bool mem_cover = !ua && s->leaflet_mode == GORDER_LEAFLET_GLOBAL && s->n_membrane > 0 && s->membrane;
    for (int i = 0; mem_cover && i < s->n_membrane; i++) {
        const int a = s->membrane[i];
        if (a < 0 || a >= s->n_atoms || in_mem[a]) mem_cover = false; else in_mem[a] = 1;
    }
    long long mem_counted = 0;
    for (int t = 0; t < s->n_moltypes; t++) {
        const GorderMolType &m = s->moltypes[t];
        if (m.n_molecules <= 0 || !m.mol_base) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "molecule type without molecules"); return h->err_code; }
        TypeDesc td{};
        td.n_mol = m.n_molecules;
        td.tile = kBlock * (ua ? 1 : h->mpt);   // one tile = the molecules of one CTA
        td.mpad = round_up(m.n_molecules, td.tile);
        std::vector<int32_t> used;
        if (ua) for (int i = 0; i < m.n_ua_atoms; i++) for (int k = 0; k < 4; k++) { int r = m.ua_rel[4 * i + k]; if (r >= 0) used.push_back(r); else if (k < 3) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "negative UA relative index"); return h->err_code; } }
        else for (int i = 0; i < 2 * m.n_bond_types; i++) { if (bad_rel(m.bond_rel[i])) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "negative bond relative index"); return h->err_code; } used.push_back(m.bond_rel[i]); }
        if (m.head_rel >= 0) used.push_back(m.head_rel);
        if (m.normal_head_rel >= 0) used.push_back(m.normal_head_rel);
        for (int k = 0; k < m.n_methyls; k++) used.push_back(m.methyl_rel[k]);
        std::sort(used.begin(), used.end());
        used.erase(std::unique(used.begin(), used.end()), used.end());
        auto u_of = [&](int rel) { return (int)(std::lower_bound(used.begin(), used.end(), rel) - used.begin()); };
        td.plane_base = (int)off;
        td.cstride = (int)used.size() * td.tile;
        const long long tile_stride = 3LL * td.cstride;
        for (size_t u = 0; u < used.size(); u++)
            for (int mm = 0; mm < m.n_molecules; mm++) {
                long long sl = (long long)m.mol_base[mm] + used[u];
                if (sl < 0 || sl >= s->n_atoms) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "atom slot out of range", sl); return h->err_code; }
                if (h->slot_off[sl] >= 0) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "atom belongs to two molecules", sl); return h->err_code; }
                h->slot_off[sl] = (int)(off + (long long)(mm / td.tile) * tile_stride + (long long)u * td.tile + (mm % td.tile));
                h->slot_cs[sl] = td.cstride;
            }
        td.tile_stride = (int)tile_stride;
        off += (long long)(td.mpad / td.tile) * tile_stride;
        td.slot0 = slot; td.molpad0 = molpad; td.mol0 = mol;
        td.head_off = m.head_rel >= 0 ? u_of(m.head_rel) * td.tile : -1;
        td.nhead_off = m.normal_head_rel >= 0 ? u_of(m.normal_head_rel) * td.tile : -1;
        td.n_methyls = m.n_methyls; td.methyl_off = (int)methyl_offs.size();
        for (int k = 0; k < m.n_methyls; k++) methyl_offs.push_back(u_of(m.methyl_rel[k]) * td.tile);
        std::vector<int32_t> islots;
        if (ua) {
            td.item_off = (int)uas.size(); td.n_items = m.n_ua_atoms;
            int k = 0;
            for (int i = 0; i < m.n_ua_atoms; i++) {
                UAItem it{};
                it.kind = m.ua_kind[i];
                const int32_t *r = m.ua_rel + 4 * i;
                it.t_off = u_of(r[0]) * td.tile; it.h1_off = u_of(r[1]) * td.tile; it.h2_off = u_of(r[2]) * td.tile;
                it.h3_off = r[3] >= 0 ? u_of(r[3]) * td.tile : 0;
                if (it.kind == GORDER_UA_CH1_SAT && r[3] < 0) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "CH1_SAT needs three helpers"); return h->err_code; }
                it.slot_rel = k; k += ua_hydrogens(it.kind);
                uas.push_back(it);
                islots.push_back(r[0]);
            }
            td.n_orders = k;
        } else {
            td.item_off = (int)bonds.size(); td.n_items = m.n_bond_types; td.n_orders = m.n_bond_types;
            std::vector<char> plane_mem(used.size(), 0), plane_done(used.size(), 0);
            for (size_t u = 0; mem_cover && u < used.size(); u++) {
                int cnt = 0;
                for (int mm = 0; mm < m.n_molecules; mm++) cnt += in_mem[m.mol_base[mm] + used[u]];
                if (cnt == m.n_molecules) plane_mem[u] = 1;
            }
            for (int i = 0; i < m.n_bond_types; i++) {
                // plane offsets are multiples of 32: the two low bits of a_off carry the register-reuse hint
                int reuse = 0;
                if (i > 0 && m.bond_rel[2 * i] == m.bond_rel[2 * (i - 1)]) reuse = 1;            // same first atom as the previous bond
                else if (i > 0 && m.bond_rel[2 * i] == m.bond_rel[2 * (i - 1) + 1]) reuse = 2;   // previous bond's second atom
                const int ua_ = u_of(m.bond_rel[2 * i]), ub_ = u_of(m.bond_rel[2 * i + 1]);
                if (plane_mem[ua_] && !plane_done[ua_]) { reuse |= 4; plane_done[ua_] = 1; mem_counted += m.n_molecules; }
                if (plane_mem[ub_] && !plane_done[ub_]) { reuse |= 8; plane_done[ub_] = 1; mem_counted += m.n_molecules; }
                bonds.push_back(BondItem{ua_ * td.tile + reuse, ub_ * td.tile});
                islots.push_back(m.bond_rel[2 * i]);
            }
            for (size_t u = 0; u < used.size(); u++)
                if (plane_done[u]) for (int mm = 0; mm < m.n_molecules; mm++) mem_done[m.mol_base[mm] + used[u]] = 1;
        }
        td.manual_leaf_off = -1; td.n_manual_leaf = 0; td.manual_norm_off = -1; td.n_manual_norm = 0;
        if (m.manual_leaflets && m.n_manual_leaflet_frames > 0) {
            td.manual_leaf_off = (int)manual_leaf.size(); td.n_manual_leaf = m.n_manual_leaflet_frames;
            manual_leaf.insert(manual_leaf.end(), m.manual_leaflets, m.manual_leaflets + (size_t)m.n_manual_leaflet_frames * m.n_molecules);
        }
        if (m.manual_normals && m.n_manual_normal_frames > 0) {
            td.manual_norm_off = (int)manual_norm.size(); td.n_manual_norm = m.n_manual_normal_frames;
            manual_norm.insert(manual_norm.end(), m.manual_normals, m.manual_normals + (size_t)3 * m.n_manual_normal_frames * m.n_molecules);
        }
        const int per_cta = kBlock * (ua ? 1 : h->mpt);
        for (int first = 0; first < td.n_mol; first += per_cta) chunks.push_back(Chunk{t, first});
        for (int i = 0; i < td.mpad; i++) h->molpad_type_h.push_back(t);
        slot += td.n_orders; molpad += td.mpad; mol += td.n_mol;
        h->types.push_back(td);
        h->mol_base.emplace_back(m.mol_base, m.mol_base + m.n_molecules);
        h->used_rel.push_back(used);
        h->item_slots.push_back(islots);
    }
    h->n_slots = slot; h->n_molpad = molpad; h->n_mol_total = mol; h->n_chunks = (int)chunks.size();
    // extra atoms: members of the groups that are not part of an analysed molecule
    std::vector<int32_t> extra;
    auto scan_group = [&](const int32_t *g, int n) -> int {
        for (int i = 0; i < n; i++) {
            if (g[i] < 0 || g[i] >= s->n_atoms) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "group atom out of range", g[i]); return h->err_code; }
            if (h->slot_off[g[i]] < 0) { h->slot_off[g[i]] = -2; extra.push_back(g[i]); }
        }
        return GORDER_OK;
    };
    if (int rc = scan_group(s->membrane, s->n_membrane)) return rc;
    if (int rc = scan_group(s->geom_ref, s->n_geom_ref)) return rc;
    if (int rc = scan_group(s->normal_heads, s->n_normal_heads)) return rc;
    const int n_extra_pad = round_up((int)extra.size(), kMolAlign);
    for (size_t e = 0; e < extra.size(); e++) { h->slot_off[extra[e]] = (int)(off + (long long)e); h->slot_cs[extra[e]] = n_extra_pad; }
    off += 3LL * n_extra_pad;
    if (off >= (1LL << 31)) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "frame too large for 32-bit plane offsets"); return h->err_code; }
    h->frame_floats = std::max<long long>(off, kMolAlign);

    // ---- batch size ------------------------------------------------------------------------------
    const size_t frame_bytes = (size_t)h->frame_floats * sizeof(float);
    long long mb = s->max_batch_frames > 0 ? s->max_batch_frames : (long long)((512ull << 20) / std::max<size_t>(frame_bytes, 1));
    mb = std::max<long long>(1, std::min<long long>(mb, s->normal_mode == GORDER_NORMAL_DYNAMIC ? 32 : 4096));
    if (s->leaflet_mode == GORDER_LEAFLET_LOCAL && s->handle_pbc) mb = std::min<long long>(mb, 32);   // per-frame cell list of the membrane atoms
    if (h->spherical) mb = std::min<long long>(mb, 256);   // per-frame scratch of the mixture fit
    h->max_batch = (int)mb;

    // ---- device tables ---------------------------------------------------------------------------
    int rc;
    TypeDesc *d_types; Chunk *d_chunks; BondItem *d_bonds; UAItem *d_ua; int *d_methyl; unsigned char *d_mleaf; float *d_mnorm;
    if ((rc = dev_upload(h, &d_types, h->types))) return rc;
    if ((rc = dev_upload(h, &d_chunks, chunks))) return rc;
    if ((rc = dev_upload(h, &d_bonds, bonds))) return rc;
    if ((rc = dev_upload(h, &d_ua, uas))) return rc;
    if ((rc = dev_upload(h, &d_methyl, methyl_offs))) return rc;
    if ((rc = dev_upload(h, &d_mleaf, manual_leaf))) return rc;
    if ((rc = dev_upload(h, &d_mnorm, manual_norm))) return rc;
    if ((rc = dev_upload(h, &h->d_slot_off, h->slot_off))) return rc;
    if ((rc = dev_upload(h, &h->d_slot_cs, h->slot_cs))) return rc;
    if ((rc = dev_upload(h, &h->d_molpad_type, h->molpad_type_h))) return rc;
    if ((rc = dev_alloc(h, &h->d_err, 2, true))) return rc;
    if ((rc = dev_alloc(h, &h->d_err_detail, 1, true))) return rc;

    DeviceView &v = h->view;
    v.types = d_types; v.chunks = d_chunks; v.bonds = d_bonds; v.ua = d_ua; v.methyl_offs = d_methyl;
    v.n_types = s->n_moltypes; v.n_chunks = h->n_chunks; v.n_slots = h->n_slots; v.n_molpad = h->n_molpad; v.n_mol_total = h->n_mol_total;
    v.frame_floats = h->frame_floats;
    v.kind = s->kind; v.handle_pbc = s->handle_pbc; v.step = s->step;
    v.normal_mode = s->normal_mode; v.normal_axis = s->normal_axis; v.dynamic_radius = s->dynamic_radius; v.collect_normals = s->collect_normals;
    v.leaflet_mode = s->leaflet_mode; v.leaflet_axis = s->leaflet_axis; v.leaflet_flip = s->leaflet_flip;
    v.leaflet_freq_kind = s->leaflet_freq_kind; v.leaflet_freq = s->leaflet_freq; v.leaflet_radius = s->leaflet_radius;
    v.shape.kind = s->geom_kind; v.shape.invert = s->geom_invert; v.shape.axis = s->geom_axis; v.shape.ref_kind = s->geom_ref_kind;
    for (int k = 0; k < 3; k++) { v.shape.ref_point[k] = s->geom_ref_point[k]; v.shape.structure_box[k] = s->structure_box[k]; }
    for (int k = 0; k < 6; k++) v.shape.dims[k] = s->geom_dims[k];
    v.manual_leaflets = d_mleaf; v.manual_normals = d_mnorm;
    v.err = h->d_err; v.err_detail = h->d_err_detail;
    v.ua_exact = h->sw.ua_exact ? 1 : 0;
    // rotation constants with the host libm (the reference's sin/cos of the same f32 angles)
    v.tet_s = sinf(1.910633f); v.tet_c = cosf(1.910633f);
    v.tet_half_s = sinf(0.9553165f); v.tet_half_c = cosf(0.9553165f);
    v.ch3_s = sinf(2.0943952f); v.ch3_c = cosf(2.0943952f);
    if ((rc = upload_group(h, &v.membrane, s->membrane, s->n_membrane))) return rc;
    if ((rc = upload_group(h, &v.geom_ref, s->geom_ref, s->n_geom_ref))) return rc;
    if ((rc = upload_group(h, &v.normal_heads, s->normal_heads, s->n_normal_heads))) return rc;
    for (int axis = 0; axis < 3; axis++) {
        if ((s->leaflet_mode == GORDER_LEAFLET_GLOBAL && axis == s->leaflet_axis) || h->spherical)
            if ((rc = build_segs(h, &h->seg_membrane[axis], s->membrane, s->n_membrane, axis))) return rc;
        if (s->geom_kind != GORDER_GEOM_NONE && s->geom_ref_kind == GORDER_GEOMREF_SELECTION)
            if ((rc = build_segs(h, &h->seg_geom[axis], s->geom_ref, s->n_geom_ref, axis))) return rc;
    }
    long long mem_left = 0;
    if (mem_cover) {
        std::vector<int32_t> left;
        for (int i = 0; i < s->n_membrane; i++) if (!mem_done[s->membrane[i]]) left.push_back(s->membrane[i]);
        mem_left = (long long)left.size();
        if (h->sw.no_spec_leftover && mem_left) mem_cover = false;
        else if (mem_left && (rc = build_segs(h, &h->seg_left, left.data(), (int)left.size(), s->leaflet_axis))) return rc;
    }

    // order maps: Map::new (ordermap.rs:40-96); node count = round(span / bin) + 1
    v.map.enabled = 0; v.map.n_bins = 0;
    if (s->map_enabled) {
        const float sx = s->map_span_x[1] - s->map_span_x[0], sy = s->map_span_y[1] - s->map_span_y[0];
        if (!(s->map_bin[0] > 0) || !(s->map_bin[1] > 0) || s->map_bin[0] > sx || s->map_bin[1] > sy) {
            h->set_error(GORDER_ERR_ORDERMAP_BIN_TOO_LARGE, "ordermap bin larger than span");
            return h->err_code;
        }
        v.map.enabled = 1; v.map.plane = s->map_plane;
        v.map.nx = (int)roundf(sx / s->map_bin[0]) + 1; v.map.ny = (int)roundf(sy / s->map_bin[1]) + 1;
        v.map.x0 = s->map_span_x[0]; v.map.y0 = s->map_span_y[0]; v.map.binx = s->map_bin[0]; v.map.biny = s->map_bin[1];
        v.map.inv_binx = 1.0f / v.map.binx; v.map.inv_biny = 1.0f / v.map.biny;
        v.map.n_bins = (long long)v.map.nx * v.map.ny;
        if (v.map.n_bins >= (1LL << 31)) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "ordermap with more than 2^31 bins"); return h->err_code; }
    }

    // ---- accumulators ----------------------------------------------------------------------------
    const long long na = (long long)h->n_slots * 3;
    h->block_words = 2 * na + 2 * na * v.map.n_bins;
    if ((rc = dev_alloc(h, &h->d_block, (size_t)h->block_words, true))) return rc;
    h->d_tot_sum = h->d_block;
    h->d_tot_cnt = reinterpret_cast<unsigned long long *>(h->d_block + na);
    h->d_map_sum = h->d_block + 2 * na;
    h->d_map_cnt = reinterpret_cast<unsigned long long *>(h->d_block + 2 * na + na * v.map.n_bins);
    if (!s->timewise) {
        // one allocation: [max_batch rows of sums][max_batch rows of counts] -> a single memset per batch
        h->ring_words = std::max<size_t>(1, (size_t)h->max_batch * na);
        for (int i = 0; i < 2; i++)
            if ((rc = dev_alloc(h, &h->d_ring[i], 2 * h->ring_words + 2, true))) return rc;
    } else {
        if ((rc = dev_alloc(h, &h->d_nflag_tw, 2, true))) return rc;
    }

    // ---- per-batch buffers -------------------------------------------------------------------------
    const size_t B = (size_t)h->max_batch;
    for (int i = 0; i < 2; i++) {
        if ((rc = dev_alloc(h, &h->d_planes[i], B * (size_t)h->frame_floats))) return rc;
        if ((rc = dev_alloc(h, &h->d_box[i], B * 3))) return rc;
        if ((rc = dev_alloc(h, &h->d_aux[i], B))) return rc;
        if ((rc = dev_alloc(h, &h->d_list[i], 2 * B))) return rc;
        CK(cudaMallocHost((void **)&h->h_aux[i], B * sizeof(FrameAux)));
        CK(cudaMallocHost((void **)&h->h_list[i], 2 * B * sizeof(int)));
        CK(cudaEventCreateWithFlags(&h->ev_stage_free[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
    }
    // planes may contain padding that is never written: keep it finite
    for (int i = 0; i < 2; i++) CK(cudaMemset(h->d_planes[i], 0, B * (size_t)h->frame_floats * sizeof(float)));
    if (h->leaf) { if ((rc = dev_alloc(h, &h->d_leaf_rows, (1 + B) * (size_t)h->n_molpad, true))) return rc; }
    if (s->normal_mode == GORDER_NORMAL_DYNAMIC && s->handle_pbc) {
        // cell list for the neighbour search unless the head group is small (brute force keeps the oracle's summation order)
        h->use_cells = s->n_normal_heads >= h->sw.cell_min_heads;
        if (h->use_cells) {
            h->cells_cap = kCellBudget;
            if ((rc = dev_alloc(h, &h->d_head_cell, B * (size_t)s->n_normal_heads))) return rc;
            if ((rc = dev_alloc(h, &h->d_cell_sorted, B * (size_t)s->n_normal_heads))) return rc;
            if ((rc = dev_alloc(h, &h->d_cell_count, B * (size_t)h->cells_cap))) return rc;
            if ((rc = dev_alloc(h, &h->d_cell_start, B * ((size_t)h->cells_cap + 1)))) return rc;
            if ((rc = dev_alloc(h, &h->d_cell_span, B * (size_t)kScanBlocks))) return rc;
            // NormalHeads member -> the padded molecule whose normal head it is
            std::unordered_map<int, int> head_of;
            size_t need = 0;
            for (int t = 0; t < s->n_moltypes; t++) {
                const int rel = s->moltypes[t].normal_head_rel;
                if (rel < 0) continue;
                for (int m = 0; m < h->types[t].n_mol; m++, need++) head_of[h->mol_base[t][m] + rel] = h->types[t].molpad0 + m;
            }
            std::vector<int> hm(s->n_normal_heads, -1);
            std::vector<char> seen(h->n_molpad, 0);
            size_t found = 0;
            for (int i = 0; i < s->n_normal_heads; i++) {
                auto it = head_of.find(s->normal_heads[i]);
                if (it == head_of.end() || seen[it->second]) continue;   // a repeated member counts in the clouds, not as a second lipid
                seen[it->second] = 1; hm[i] = it->second; found++;
            }
            h->normals_sorted = found == need && !h->sw.no_sorted_normals;
            if ((rc = dev_upload(h, &h->d_head_molpad, hm))) return rc;
        }
    }
    if (s->leaflet_mode == GORDER_LEAFLET_LOCAL && s->handle_pbc) {
        h->use_lcells = s->n_membrane >= h->sw.lcell_min_atoms;   // below: brute force over the membrane group
        if (h->use_lcells) {
            h->lcells_cap = kLCellMaxDim * kLCellMaxDim;
            if ((rc = dev_alloc(h, &h->d_matom_cell, B * (size_t)s->n_membrane))) return rc;
            if ((rc = dev_alloc(h, &h->d_lcell_sorted, B * (size_t)s->n_membrane))) return rc;
            if ((rc = dev_alloc(h, &h->d_lcell_count, B * (size_t)h->lcells_cap))) return rc;
            if ((rc = dev_alloc(h, &h->d_lcell_start, B * ((size_t)h->lcells_cap + 1)))) return rc;
        }
    }
    if (h->nvec) {
        if ((rc = dev_alloc(h, &h->d_normals, B * 3 * (size_t)h->n_molpad))) return rc;
        if ((rc = dev_alloc(h, &h->d_normal_npoints, B * (size_t)h->n_molpad, true))) return rc;
        if (s->collect_normals && s->geom_kind != GORDER_GEOM_NONE && !ua) { if ((rc = dev_alloc(h, &h->d_normal_used, B * (size_t)h->n_molpad, true))) return rc; }
    }
    for (int i = 0; i < 2; i++) {
        if ((rc = dev_alloc(h, &h->d_est2[i], B * 3))) return rc;
        if ((rc = dev_alloc(h, &h->d_center2[i], B * 3))) return rc;
        if ((rc = dev_alloc(h, &h->d_partial2[i], B * kCenterBlocks * 2))) return rc;
        if ((rc = dev_alloc(h, &h->d_ticket2[i], B, true))) return rc;
        CK(cudaEventCreateWithFlags(&h->ev_pre[i], cudaEventDisableTiming));
    }
    h->d_est = h->d_est2[0]; h->d_center = h->d_center2[0]; h->d_partial = h->d_partial2[0]; h->d_ticket = h->d_ticket2[0];
    {   // higher priority: the small centre kernels of the next batch must not queue behind the bond kernel's CTAs
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&h->stream_pre, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&h->stream_post, cudaStreamNonBlocking, hi));
        for (int i = 0; i < 2; i++) {
            CK(cudaEventCreateWithFlags(&h->ev_bond[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_post[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&h->ev_post_any, cudaEventDisableTiming));
    }

    h->fast_ok = !ua && !h->nvec && !h->extra && s->handle_pbc;
    if (h->fast_ok && s->n_moltypes <= kFastTypes && (int)bonds.size() <= kFastBonds && h->device < 64 && !h->sw.no_const_tables) {
        {
            std::lock_guard<std::mutex> lock(g_fast_mu);
            for (int i = 0; i < kFastSlots && h->fast_slot < 0; i++)
                if (!g_fast_used[h->device][i]) { g_fast_used[h->device][i] = true; h->fast_slot = i; }
        }
        if (h->fast_slot >= 0) {
            FastTables ft{};
            ft.n_types = s->n_moltypes;
            int c = 0;
            for (int t = 0; t < s->n_moltypes; t++) { ft.chunk0[t] = c; c += h->types[t].mpad / h->types[t].tile; ft.types[t] = h->types[t]; }
            ft.chunk0[s->n_moltypes] = c;
            for (size_t i = 0; i < bonds.size(); i++) ft.bonds[i] = bonds[i];
            CK(cudaMemcpyToSymbol(c_fast, &ft, sizeof(ft), (size_t)h->fast_slot * sizeof(FastTables)));
        }
    }
    // K2f (gorder_ua_fast.cuh): UA, PBC, static normal, no geometry / maps, streaming hydrogen construction; the tile of a CTA
    // (3 x used atoms x 256 molecules floats) must fit in shared memory next to the tables
    if (ua && !h->nvec && !h->extra && s->handle_pbc && !h->sw.ua_exact && !h->sw.no_fast && h->frame_floats % 4 == 0) {
        bool ok = true;
        int max_tile = 0, max_orders = 0;
        for (auto &t : h->types) {
            ok = ok && t.tile == kUaTile && (t.plane_base % 4) == 0;
            max_tile = std::max(max_tile, t.tile_stride); max_orders = std::max(max_orders, t.n_orders);
        }
        int max_optin = 0;
        CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
        CK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device));
        const size_t tables = uas.size() * sizeof(UAItem) + (size_t)2 * kUaPairWarps * max_orders * 2 * sizeof(int) + 256;
        for (int nbuf = 2; ok && nbuf >= 1 && !h->ua_fast_ok; nbuf--) {   // two tile buffers when they fit (the copy of the next tile overlaps the arithmetic), else one
            const size_t need = (size_t)nbuf * max_tile * sizeof(float) + tables;
            if (need + 1024 > (size_t)max_optin) continue;
            h->ua_fast_ok = true; h->ua_fast_smem = need; h->ua_fast_nbuf = nbuf;
            h->ua_fast_tile = max_tile; h->ua_fast_items = (int)uas.size(); h->ua_fast_orders = max_orders;
            CK(cudaFuncSetAttribute(ua_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            CK(cudaFuncSetAttribute(ua_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        }
    }
    // speculative Global leaflets: AA/CG, static normal along the leaflet axis, PBC, assignment on every analysed frame,
    // no geometry / maps, membrane covered by the bond kernel's loads (above)
    h->spec_ok = mem_cover && mem_counted + mem_left == s->n_membrane && !ua && !h->nvec && !h->extra && s->handle_pbc &&
                 s->leaflet_mode == GORDER_LEAFLET_GLOBAL && s->leaflet_freq_kind == GORDER_FREQ_EVERY && s->leaflet_freq <= std::max(1, s->step) &&
                 s->leaflet_axis == s->normal_axis && h->n_chunks > 0 && !h->sw.no_spec;
    if (h->spec_ok) {
        if ((rc = dev_alloc(h, &h->d_spec_ref, 4, true))) return rc;
        if ((rc = dev_alloc(h, &h->d_spec_sum, 2 * B * (size_t)h->n_chunks))) return rc;
        if ((rc = dev_alloc(h, &h->d_spec_mm, B * (size_t)h->n_chunks * 4))) return rc;
        if ((rc = dev_alloc(h, &h->d_spec_ticket, B, true))) return rc;
        h->spec_left_blocks = std::min(kSpecLeftBlocks, h->seg_left.n);
        if (h->spec_left_blocks) {
            if ((rc = dev_alloc(h, &h->d_spec_left_sum, 2 * B * (size_t)h->spec_left_blocks))) return rc;
            if ((rc = dev_alloc(h, &h->d_spec_left_mm, 2 * B * (size_t)h->spec_left_blocks))) return rc;
        }
        if ((rc = dev_alloc(h, &h->d_spec_center, 2 * B, true))) return rc;
        if ((rc = dev_alloc(h, &h->d_spec_flag, 2 * B, true))) return rc;
        CK(cudaHostAlloc((void **)&h->h_spec_counters, 2 * sizeof(unsigned), cudaHostAllocMapped));
        h->h_spec_counters[0] = h->h_spec_counters[1] = 0;
        CK(cudaHostGetDevicePointer((void **)&h->d_spec_counters, h->h_spec_counters, 0));
    }

    if (h->spherical) {   // scratch of spherical_cluster_kernel and the head -> ClusterHeads position table
        std::vector<int> pos_of(s->n_atoms, -1), index((size_t)h->n_molpad, 0);
        for (int i = 0; i < s->n_membrane; i++) {
            if (s->membrane[i] < 0 || s->membrane[i] >= s->n_atoms) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "group atom out of range", s->membrane[i]); return h->err_code; }
            pos_of[s->membrane[i]] = i;
        }
        for (int t = 0; t < s->n_moltypes; t++) {
            const GorderMolType &m = s->moltypes[t];
            for (int mm = 0; mm < m.n_molecules; mm++) {
                const int head = m.head_rel >= 0 ? m.mol_base[mm] + m.head_rel : -1;
                if (head < 0 || pos_of[head] < 0) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "molecule head is not in the ClusterHeads group", head); return h->err_code; }
                index[(size_t)h->types[t].molpad0 + mm] = pos_of[head];
            }
        }
        if ((rc = dev_upload(h, &h->d_sph_index, index))) return rc;
        h->sph_pad = round_up(s->n_membrane, 4);
        if ((rc = dev_alloc(h, &h->d_sph_scratch, B * (size_t)kSphArrays * (size_t)h->sph_pad))) return rc;
        if ((rc = dev_alloc(h, &h->d_sph_upper, B * (size_t)s->n_membrane))) return rc;
    }

    // dynamic shared memory of the accumulation kernels (small; no opt-in needed below 48 KB)
    if (accum_smem(h) > 48 * 1024) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "too many order slots per molecule type for shared memory"); return h->err_code; }
    CK(cudaDeviceSynchronize());
    return GORDER_OK;
}

int gorder_gpu_create(const GorderSetup *setup, GorderHandle **out) {
    if (!setup || !out || setup->abi_version != GORDER_ABI_VERSION) return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    GorderHandle *h = new GorderHandle();
    int rc = create_impl(setup, h);
    if (rc) {
        if (h->sw.verbose || getenv("GORDER_VERBOSE")) fprintf(stderr, "gorder_gpu_create failed: %s\n", h->err_msg.c_str());
        // pointers inside the copied setup are not owned
        gorder_gpu_destroy(h);
        return rc;
    }
    // the copied setup must not keep caller pointers
    h->s.moltypes = nullptr; h->s.normal_heads = nullptr; h->s.membrane = nullptr; h->s.geom_ref = nullptr;
    *out = h;
    return GORDER_OK;
}

// everything queued by this handle, on all of its streams (the tail of a batch may run on the post stream)
static int sync_all(GorderHandle *h) {
    CK(cudaStreamSynchronize(h->copy_stream));
    CK(cudaStreamSynchronize(h->stream_pre));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaStreamSynchronize(h->stream_post));
    return GORDER_OK;
}

// first deferred device error -> host error state
static int poll_device_error(GorderHandle *h) {
    int code[2] = {0, 0};
    long long detail = 0;
    CK(cudaMemcpyAsync(code, h->d_err, sizeof(code), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&detail, h->d_err_detail, sizeof(detail), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (code[0]) {
        long long d = detail;
        if (code[0] == GORDER_ERR_UNDEFINED_POSITION && (detail >> 32) != 0) {   // (type, item, molecule) -> atom slot
            int t = (int)(detail >> 48), item = (int)((detail >> 32) & 0xffff), m = (int)(detail & 0xffffffff);
            if (t < (int)h->mol_base.size() && m < (int)h->mol_base[t].size() && item < (int)h->item_slots[t].size())
                d = (long long)h->mol_base[t][m] + h->item_slots[t][item];
        }
        h->set_error(code[0], code[0] == GORDER_ERR_INVALID_ARGUMENT ? "corrupt XTC frame (detail: its number in the file), found by the walk on the device" : "deferred device error", d);
    }
    return h->err_code;
}

static int begin_slot(GorderHandle *h, int *slot) {
    *slot = h->cur;
    h->cur ^= 1;
    CK(cudaEventSynchronize(h->ev_stage_free[*slot]));
    CK(cudaEventSynchronize(h->ev_post[*slot]));   // tail of the batch that used this slot (post stream)
    return GORDER_OK;
}

static int check_args(GorderHandle *h, const void *frames, const void *box, const int64_t *fi, int32_t n) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    if (h->err_code) return h->err_code;
    if (n < 0 || (n > 0 && (!frames || !fi || (h->s.handle_pbc && !box)))) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "null argument"); return h->err_code; }
    cudaSetDevice(h->device);
    return GORDER_OK;
}

static int launch_relayout(GorderHandle *h, const float *d_xyz, float *d_planes, int nf) {
    dim3 grid((h->s.n_atoms + 255) / 256, nf);
    relayout_kernel<<<grid, 256, 0, h->stream>>>(h->view, d_xyz, d_planes, h->d_slot_off, h->d_slot_cs, h->s.n_atoms);
    h->n_launches++;
    CK(cudaGetLastError());
    return GORDER_OK;
}

int gorder_gpu_submit(GorderHandle *h, const float *xyz, const float *box, const int64_t *frame_index, int32_t n_frames) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (int rc = check_args(h, xyz, box, frame_index, n_frames)) return rc;
    const size_t fstride = (size_t)h->s.n_atoms * 3;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        const int nf = std::min(h->max_batch, n_frames - f0);
        int slot;
        if (int rc = begin_slot(h, &slot)) return rc;
        if (!h->d_xyz[slot]) CK(cudaMalloc((void **)&h->d_xyz[slot], (size_t)h->max_batch * fstride * sizeof(float)));
        CK(cudaMemcpyAsync(h->d_xyz[slot], xyz + (size_t)f0 * fstride, (size_t)nf * fstride * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
        if (h->s.handle_pbc) CK(cudaMemcpyAsync(h->d_box[slot], box + 3 * (size_t)f0, (size_t)nf * 3 * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->ev_h2d[slot], h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, h->ev_h2d[slot], 0));
        if (int rc = launch_relayout(h, h->d_xyz[slot], h->d_planes[slot], nf)) return rc;
        if (int rc = process_batch(h, h->d_planes[slot], h->d_box[slot], (const long long *)frame_index + f0, nf, slot, true)) return rc;
        CK(cudaEventRecord(h->ev_stage_free[slot], h->stream));
        CK(cudaEventSynchronize(h->ev_h2d[slot]));   // the caller's buffers may be reused from here on
    }
    return GORDER_OK;
}

int gorder_gpu_submit_device(GorderHandle *h, const float *d_xyz, const float *d_box, const int64_t *frame_index, int32_t n_frames) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (int rc = check_args(h, d_xyz, d_box, frame_index, n_frames)) return rc;
    const size_t fstride = (size_t)h->s.n_atoms * 3;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        const int nf = std::min(h->max_batch, n_frames - f0);
        int slot;
        if (int rc = begin_slot(h, &slot)) return rc;
        if (int rc = launch_relayout(h, d_xyz + (size_t)f0 * fstride, h->d_planes[slot], nf)) return rc;
        if (int rc = process_batch(h, h->d_planes[slot], d_box ? d_box + 3 * (size_t)f0 : nullptr, (const long long *)frame_index + f0, nf, slot, true)) return rc;
        CK(cudaEventRecord(h->ev_stage_free[slot], h->stream));
    }
    return GORDER_OK;
}

int gorder_gpu_native_layout(GorderHandle *h, int64_t *frame_floats, int32_t *plane_offset, int32_t *plane_cstride) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    if (frame_floats) *frame_floats = h->frame_floats;
    if (plane_offset) for (int i = 0; i < h->s.n_atoms; i++) plane_offset[i] = h->slot_off[i] >= 0 ? h->slot_off[i] : -1;
    if (plane_cstride) for (int i = 0; i < h->s.n_atoms; i++) plane_cstride[i] = h->slot_cs[i];
    return GORDER_OK;
}

int gorder_gpu_submit_native(GorderHandle *h, const float *planes_host, const float *box, const int64_t *frame_index, int32_t n_frames) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (int rc = check_args(h, planes_host, box, frame_index, n_frames)) return rc;
    const size_t fstride = (size_t)h->frame_floats;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        const int nf = std::min(h->max_batch, n_frames - f0);
        int slot;
        if (int rc = begin_slot(h, &slot)) return rc;
        CK(cudaMemcpyAsync(h->d_planes[slot], planes_host + (size_t)f0 * fstride, (size_t)nf * fstride * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
        if (h->s.handle_pbc) CK(cudaMemcpyAsync(h->d_box[slot], box + 3 * (size_t)f0, (size_t)nf * 3 * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->ev_h2d[slot], h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, h->ev_h2d[slot], 0));
        if (int rc = process_batch(h, h->d_planes[slot], h->d_box[slot], (const long long *)frame_index + f0, nf, slot, true)) return rc;
        CK(cudaEventRecord(h->ev_stage_free[slot], h->stream));
        CK(cudaEventSynchronize(h->ev_h2d[slot]));
    }
    return GORDER_OK;
}

int gorder_gpu_submit_native_device(GorderHandle *h, const float *d_planes, const float *d_box, const int64_t *frame_index, int32_t n_frames) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (int rc = check_args(h, d_planes, d_box, frame_index, n_frames)) return rc;
    const size_t fstride = (size_t)h->frame_floats;
    for (int f0 = 0; f0 < n_frames; f0 += h->max_batch) {
        const int nf = std::min(h->max_batch, n_frames - f0);
        int slot;
        if (int rc = begin_slot(h, &slot)) return rc;
        if (int rc = process_batch(h, d_planes + (size_t)f0 * fstride, d_box ? d_box + 3 * (size_t)f0 : nullptr, (const long long *)frame_index + f0, nf, slot, false)) return rc;
        CK(cudaEventRecord(h->ev_stage_free[slot], h->stream));
    }
    return GORDER_OK;
}

int gorder_gpu_reserve_frames(GorderHandle *h, int64_t n_frames) {
    if (!h || n_frames < 0) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (h->err_code) return h->err_code;
    cudaSetDevice(h->device);
    return grow_rows(h, n_frames);
}

int gorder_gpu_set_leaflets(GorderHandle *h, const uint8_t *table, int64_t frame_index) {
    if (!h || !table) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (h->err_code) return h->err_code;
    if (!h->leaf) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "leaflets are not enabled"); return h->err_code; }
    cudaSetDevice(h->device);
    std::vector<unsigned char> row(h->n_molpad, GORDER_UPPER);
    for (auto &td : h->types) for (int m = 0; m < td.n_mol; m++) row[td.molpad0 + m] = table[td.mol0 + m];
    if (int rc = sync_all(h)) return rc;   // batches in flight (any stream) may still read row 0
    CK(cudaMemcpy(h->d_leaf_rows, row.data(), row.size(), cudaMemcpyHostToDevice));
    h->have_leaflets = true; h->cur_leaflet_frame = frame_index;
    return GORDER_OK;
}

int gorder_gpu_sync(GorderHandle *h) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (h->err_code) return h->err_code;
    cudaSetDevice(h->device);
    if (int rc = sync_all(h)) return rc;
    return poll_device_error(h);
}

int gorder_gpu_result_sizes(GorderHandle *h, GorderResults *r) {
    if (!h || !r) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    r->n_slots = h->n_slots; r->n_frames = h->n_frames; r->n_map_bins = h->view.map.n_bins;
    r->map_nx = h->view.map.enabled ? h->view.map.nx : 0; r->map_ny = h->view.map.enabled ? h->view.map.ny : 0;
    r->n_leaflet_frames = h->n_leaf_collected; r->n_molecules_total = h->n_mol_total;
    return GORDER_OK;
}

int gorder_gpu_finish(GorderHandle *h, GorderResults *r) {
    if (!h || !r) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (int rc = gorder_gpu_sync(h)) return rc;
    gorder_gpu_result_sizes(h, r);
    const size_t na = (size_t)h->n_slots * 3;
    if (r->sum) CK(cudaMemcpy(r->sum, h->d_tot_sum, na * sizeof(long long), cudaMemcpyDeviceToHost));
    if (r->count) CK(cudaMemcpy(r->count, h->d_tot_cnt, na * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (h->s.timewise && h->n_frames > 0) {
        if (r->tw_sum) CK(cudaMemcpy(r->tw_sum, h->d_bsum, na * h->n_frames * sizeof(long long), cudaMemcpyDeviceToHost));
        if (r->tw_count) CK(cudaMemcpy(r->tw_count, h->d_bcnt, na * h->n_frames * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    if (r->tw_frame_index) for (long long i = 0; i < h->n_frames; i++) r->tw_frame_index[i] = h->frame_index_done[i];
    if (h->view.map.enabled) {
        const size_t nb = (size_t)h->view.map.n_bins;
        if (r->map_sum) {
            CK(cudaMemcpy(r->map_sum, h->d_map_sum, na * nb * sizeof(long long), cudaMemcpyDeviceToHost));
            if (h->leaf)   // total map = upper + lower (bond.rs:184-215)
                for (int sl = 0; sl < h->n_slots; sl++)
                    for (size_t b = 0; b < nb; b++)
                        r->map_sum[((size_t)sl * 3 + GORDER_TOTAL) * nb + b] = r->map_sum[((size_t)sl * 3 + GORDER_ACC_UPPER) * nb + b] + r->map_sum[((size_t)sl * 3 + GORDER_ACC_LOWER) * nb + b];
        }
        if (r->map_count) {
            CK(cudaMemcpy(r->map_count, h->d_map_cnt, na * nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            if (h->leaf)
                for (int sl = 0; sl < h->n_slots; sl++)
                    for (size_t b = 0; b < nb; b++)
                        r->map_count[((size_t)sl * 3 + GORDER_TOTAL) * nb + b] = r->map_count[((size_t)sl * 3 + GORDER_ACC_UPPER) * nb + b] + r->map_count[((size_t)sl * 3 + GORDER_ACC_LOWER) * nb + b];
        }
    }
    if (r->leaflets && h->n_leaf_collected > 0) {
        std::vector<unsigned char> rows((size_t)h->n_leaf_collected * h->n_molpad);
        CK(cudaMemcpy(rows.data(), h->d_leaf_collect, rows.size(), cudaMemcpyDeviceToHost));
        for (long long a = 0; a < h->n_leaf_collected; a++)
            for (auto &td : h->types)
                memcpy(r->leaflets + (size_t)a * h->n_mol_total + td.mol0, rows.data() + (size_t)a * h->n_molpad + td.molpad0, td.n_mol);
    }
    if (r->leaflet_frame_index) for (long long a = 0; a < h->n_leaf_collected; a++) r->leaflet_frame_index[a] = h->leaf_frame_index[a];
    if (r->normals && h->d_normals_collect && h->n_frames > 0) {
        std::vector<float> buf((size_t)h->n_frames * 3 * h->n_molpad);
        CK(cudaMemcpy(buf.data(), h->d_normals_collect, buf.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (long long f = 0; f < h->n_frames; f++)
            for (auto &td : h->types)
                for (int m = 0; m < td.n_mol; m++)
                    for (int c = 0; c < 3; c++)
                        r->normals[((size_t)f * h->n_mol_total + td.mol0 + m) * 3 + c] = buf[((size_t)f * 3 + c) * h->n_molpad + td.molpad0 + m];
    }
    return GORDER_OK;
}

int gorder_gpu_accumulator_block(GorderHandle *h, void **d_ptr, int64_t *n_words) {
    if (!h || !d_ptr || !n_words) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int rc = sync_all(h)) return rc;
    *d_ptr = h->d_block; *n_words = h->block_words;
    return GORDER_OK;
}

int gorder_gpu_read_block(GorderHandle *h, void *d_dst) {
    if (!h || !d_dst) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int rc = sync_all(h)) return rc;
    CK(cudaMemcpyAsync(d_dst, h->d_block, (size_t)h->block_words * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GORDER_OK;
}

int gorder_gpu_write_block(GorderHandle *h, const void *d_src) {
    if (!h || !d_src) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int rc = sync_all(h)) return rc;
    CK(cudaMemcpyAsync(h->d_block, d_src, (size_t)h->block_words * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GORDER_OK;
}

static int drain_profile(GorderHandle *h) {
    if (h->prof_used == 0 && h->prof_used_n == 0) return GORDER_OK;
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < h->prof_used; i++) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second));
        h->prof_ms += ms; h->prof_n++;
    }
    h->prof_used = 0;
    for (size_t i = 0; i < h->prof_used_n; i++) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->prof_events_n[i].first, h->prof_events_n[i].second));
        h->prof_ms_n += ms; h->prof_n_n++;
    }
    h->prof_used_n = 0;
    return GORDER_OK;
}

int gorder_gpu_profile(GorderHandle *h, int enable) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int rc = drain_profile(h)) return rc;
    h->profiling = enable != 0;
    return GORDER_OK;
}

int gorder_gpu_profile_read(GorderHandle *h, double *hot_kernel_ms, int64_t *hot_kernel_launches) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int rc = drain_profile(h)) return rc;
    if (hot_kernel_ms) *hot_kernel_ms = h->prof_ms;
    if (hot_kernel_launches) *hot_kernel_launches = h->prof_n;
    h->prof_ms = 0.0; h->prof_n = 0;
    return GORDER_OK;
}

int gorder_gpu_profile_read_normals(GorderHandle *h, double *stage_ms, int64_t *batches) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int rc = drain_profile(h)) return rc;
    if (stage_ms) *stage_ms = h->prof_ms_n;
    if (batches) *batches = h->prof_n_n;
    h->prof_ms_n = 0.0; h->prof_n_n = 0;
    return GORDER_OK;
}

int gorder_gpu_stats(GorderHandle *h, int64_t *kernel_launches, int64_t *frames) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (kernel_launches) *kernel_launches = h->n_launches;
    if (frames) *frames = h->n_frames;
    return GORDER_OK;
}

int gorder_gpu_speculation_stats(GorderHandle *h, int32_t *enabled, int64_t *frames_speculated, int64_t *frames_repaired) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (int rc = sync_all(h)) return rc;
    if (enabled) *enabled = (h->spec_ok && !h->spec_disabled) ? 1 : 0;
    if (frames_speculated) *frames_speculated = h->h_spec_counters ? h->h_spec_counters[0] : 0;
    if (frames_repaired) *frames_repaired = h->h_spec_counters ? h->h_spec_counters[1] : 0;
    return GORDER_OK;
}

// Frames per batch for which the grid of the accumulation kernel (tiles x frames) is a whole number of waves: the kernels
// are compiled for a fixed number of resident CTAs per SM, and a last, partly filled wave costs as much as a full one
// (S-AA-large: 128 -> 148 frames per batch = +15 % in the kernel).  0: no preference (persistent kernel).
int gorder_gpu_wave_frames(GorderHandle *h, int32_t *frames) {
    if (!h || !frames) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    *frames = 0;
    if (h->ua && h->ua_fast_ok) return GORDER_OK;
    int per_sm = 4;                                                   // bond_fast_kernel<.., 256>, bond_order_kernel, ua_order_kernel
    const bool fast = h->fast_ok && !h->sw.no_fast;
    if (fast && h->mpt == 1) per_sm = 16;                             // bond_fast_kernel<.., 64>
    else if (!fast && !h->ua && h->extra) per_sm = GORDER_EXTRA_MINB;
    else if (!fast && !h->ua && h->nvec) per_sm = GORDER_NVEC_MINB;
    int n_sm = h->n_sm;
    if (n_sm <= 0) { cudaSetDevice(h->device); CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device)); }
    const long long slots = (long long)per_sm * n_sm;
    long long a = slots, b = std::max(1, h->n_chunks);
    while (b) { const long long t = a % b; a = b; b = t; }            // gcd
    *frames = (int32_t)std::min<long long>(slots / a, 1 << 20);
    return GORDER_OK;
}

int gorder_gpu_fence(GorderHandle *h) {
    if (!h) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    cudaSetDevice(h->device);
    if (h->post_used) CK(cudaStreamWaitEvent(h->stream, h->ev_post_any, 0));
    return GORDER_OK;
}

void *gorder_gpu_stream(GorderHandle *h) { return h ? (void *)h->stream : nullptr; }

int gorder_gpu_last_error(GorderHandle *h, char *buf, size_t len) {
    if (!h || !buf || !len) return GORDER_ERR_INVALID_ARGUMENT;
    snprintf(buf, len, "%s", h->err_msg.c_str());
    return h->err_code;
}

int64_t gorder_gpu_error_detail(GorderHandle *h) { return h ? h->err_detail : -1; }

}  // extern "C"

#include "gorder_xtc.inl"
#include "gorder_results.inl"
#include "gorder_multi.inl"
#include "gorder_topology.inl"

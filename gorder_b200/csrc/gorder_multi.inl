// gorder_multi.inl — multi-GPU merge behind the C ABI (included by gorder_capi.cu).
//
// Replaces ParallelTrajData::reduce over the per-thread clones (reference: src/analysis/topology/mod.rs:256-272; the Add
// chain bond.rs:449-465, order.rs:160-176, timewise.rs:34-51, ordermap.rs:116-138, normal.rs:234-256, 478-498, and the
// interleave of per-frame vectors, common.rs:380-404) when the analysed frames are sharded over GPUs in contiguous ranges
// (SURVEY.md §8e).  All accumulators are integers, so the merge is exact and order-free:
//   * the contiguous accumulator block (sums, counts, map bins) is SUMMED on the root;
//   * per-frame rows (error estimation), collected leaflet tables and normals are GATHERED in frame order (every shard
//     owns disjoint frames; shards are ordered by the frame_index of their first frame);
//   * the first error of any shard (in shard order) becomes the root's error, as the first Err aborts the reference's
//     map-reduce (groan traj_iter_map_reduce).
// Two transports:
//   gorder_gpu_reduce        one process drives all devices: ONE kernel on the root device reads the peers' blocks through
//                            NVLink peer access and adds them (no staging copy, no library collective); the row gathers are
//                            peer-to-peer copies.
//   gorder_gpu_reduce_comm   one process per GPU (torchrun / MPI): NCCL (ncclReduce int64 + grouped send / recv of the rows)
//                            on a communicator the library creates from a broadcast unique id.  NCCL is loaded at run time
//                            (dlopen of libnccl.so.2), so the library has no link-time dependency on it.
#include <dlfcn.h>

#include <chrono>

namespace {

// ---- block sum over peer memory -------------------------------------------------------------------------------------
constexpr int kMaxPeers = 15;
struct PeerBlocks { const long long *p[kMaxPeers]; int n; };

__global__ void __launch_bounds__(256) peer_block_sum_kernel(long long *__restrict__ dst, PeerBlocks peers, long long n_words) {
    const long long stride = (long long)gridDim.x * blockDim.x * 2;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n_words; i += stride) {
        if (i + 1 < n_words) {   // 128-bit loads over NVLink (the block is 16-byte aligned: cudaMalloc)
            longlong2 a = *reinterpret_cast<const longlong2 *>(dst + i);
            for (int k = 0; k < peers.n; k++) {
                const longlong2 b = *reinterpret_cast<const longlong2 *>(peers.p[k] + i);
                a.x += b.x; a.y += b.y;
            }
            *reinterpret_cast<longlong2 *>(dst + i) = a;
        } else {
            long long a = dst[i];
            for (int k = 0; k < peers.n; k++) a += peers.p[k][i];
            dst[i] = a;
        }
    }
}

// ---- what a shard contributes besides its block (host view) ------------------------------------------------------------
struct ShardMeta {
    long long n_frames = 0, n_leaf = 0;
    int err_code = 0;
    long long err_detail = -1;
    std::vector<long long> frame_index, leaf_frame_index;
};

int check_same_shape(GorderHandle *a, GorderHandle *b) {
    if (a->n_slots != b->n_slots || a->block_words != b->block_words || a->n_molpad != b->n_molpad || a->n_mol_total != b->n_mol_total ||
        a->s.timewise != b->s.timewise || a->s.collect_leaflets != b->s.collect_leaflets || a->s.collect_normals != b->s.collect_normals) {
        a->set_error(GORDER_ERR_INVALID_ARGUMENT, "shards were created from different setups");
        return a->err_code;
    }
    return GORDER_OK;
}

// order of the shards in the merged per-frame arrays: by the frame_index of their first frame (empty shards last)
std::vector<int> shard_order(const std::vector<ShardMeta> &m, bool by_leaf) {
    std::vector<int> idx(m.size());
    for (size_t i = 0; i < m.size(); i++) idx[i] = (int)i;
    auto key = [&](int i) {
        const auto &v = by_leaf ? m[i].leaf_frame_index : m[i].frame_index;
        return v.empty() ? std::numeric_limits<long long>::max() : v[0];
    };
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return key(a) < key(b); });
    return idx;
}

// The merged per-frame arrays of the root: allocated here, filled by `fetch(shard, dst pointer, src offset 0, bytes, what)`.
struct MergePlan {
    std::vector<int> order, leaf_order;
    std::vector<long long> row_off, leaf_off;   // per shard: first row in the merged arrays
    long long total_frames = 0, total_leaf = 0;
};

MergePlan make_plan(const std::vector<ShardMeta> &m) {
    MergePlan p;
    p.order = shard_order(m, false); p.leaf_order = shard_order(m, true);
    p.row_off.assign(m.size(), 0); p.leaf_off.assign(m.size(), 0);
    for (int i : p.order) { p.row_off[i] = p.total_frames; p.total_frames += m[i].n_frames; }
    for (int i : p.leaf_order) { p.leaf_off[i] = p.total_leaf; p.total_leaf += m[i].n_leaf; }
    return p;
}

// New per-frame arrays of the root (device memory on the root's device), sized for the whole trajectory.
struct MergedArrays {
    long long *bsum = nullptr;
    unsigned long long *bcnt = nullptr;
    unsigned char *leaf = nullptr;
    float *normals = nullptr;
    bool rows_in_place = false;   // the root's own rows already sit at the front of arrays large enough for all shards
};

// `root`: index of the root's shard.  When the root owns the first frames and gorder_gpu_reserve_frames announced the whole
// trajectory to it, the peers' rows land directly behind its own: no allocation, no copy of the root's rows.
int alloc_merged(GorderHandle *h, const MergePlan &p, int root, MergedArrays *a) {
    const size_t row = (size_t)h->n_slots * 3;
    if (h->s.timewise && p.total_frames > 0) {
        if (p.row_off[(size_t)root] == 0 && h->d_bsum && h->tw_cap >= p.total_frames) {
            a->bsum = h->d_bsum; a->bcnt = h->d_bcnt; a->rows_in_place = true;
        } else {
            CK(cudaMalloc((void **)&a->bsum, (size_t)p.total_frames * row * sizeof(long long)));
            CK(cudaMalloc((void **)&a->bcnt, (size_t)p.total_frames * row * sizeof(unsigned long long)));
        }
    }
    if (h->s.collect_leaflets && p.total_leaf > 0) CK(cudaMalloc((void **)&a->leaf, (size_t)p.total_leaf * h->n_molpad));
    if (h->s.collect_normals && h->s.normal_mode == GORDER_NORMAL_DYNAMIC && p.total_frames > 0)
        CK(cudaMalloc((void **)&a->normals, (size_t)p.total_frames * 3 * h->n_molpad * sizeof(float)));
    return GORDER_OK;
}

// the root handle takes the merged arrays and the merged frame lists over
void adopt_merged(GorderHandle *h, const std::vector<ShardMeta> &m, const MergePlan &p, const MergedArrays &a) {
    if (a.bsum && !a.rows_in_place) { cudaFree(h->d_bsum); cudaFree(h->d_bcnt); h->d_bsum = a.bsum; h->d_bcnt = a.bcnt; h->tw_cap = p.total_frames; }
    if (a.leaf) { cudaFree(h->d_leaf_collect); h->d_leaf_collect = a.leaf; h->leaf_collect_cap = p.total_leaf; }
    if (a.normals) { cudaFree(h->d_normals_collect); h->d_normals_collect = a.normals; h->normals_collect_cap = p.total_frames; }
    h->frame_index_done.clear(); h->leaf_frame_index.clear();
    for (int i : p.order) h->frame_index_done.insert(h->frame_index_done.end(), m[i].frame_index.begin(), m[i].frame_index.end());
    for (int i : p.leaf_order) h->leaf_frame_index.insert(h->leaf_frame_index.end(), m[i].leaf_frame_index.begin(), m[i].leaf_frame_index.end());
    h->n_frames = p.total_frames; h->n_leaf_collected = p.total_leaf;
    if (!h->frame_index_done.empty()) h->last_frame_index = *std::max_element(h->frame_index_done.begin(), h->frame_index_done.end());
}

ShardMeta meta_of(GorderHandle *h) {
    ShardMeta m;
    m.n_frames = h->n_frames; m.n_leaf = h->n_leaf_collected; m.err_code = h->err_code; m.err_detail = h->err_detail;
    m.frame_index = h->frame_index_done; m.leaf_frame_index = h->leaf_frame_index;
    return m;
}

// ---- NCCL, loaded at run time ------------------------------------------------------------------------------------------
struct NcclId { char b[128]; };   // ncclUniqueId (passed by value to ncclCommInitRank)
struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
};
constexpr int kNcclInt8 = 0, kNcclInt64 = 4, kNcclSum = 0;

Nccl *nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("GORDER_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) { n.why = "libnccl.so.2 not found (set GORDER_NCCL_LIB)"; return; }
        auto sym = [&](const char *s) { void *p = dlsym(n.lib, s); if (!p && n.why.empty()) n.why = std::string("missing NCCL symbol ") + s; return p; };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.Reduce = (decltype(n.Reduce))sym("ncclReduce");
        n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
        n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
        n.Send = (decltype(n.Send))sym("ncclSend");
        n.Recv = (decltype(n.Recv))sym("ncclRecv");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    });
    return (n.lib && n.why.empty()) ? &n : nullptr;
}

#define NK(call)                                                                                                   \
    do {                                                                                                           \
        int r__ = (call);                                                                                          \
        if (r__ != 0) {                                                                                            \
            h->set_error(GORDER_ERR_NCCL, std::string(#call) + ": " + (N->GetErrorString ? N->GetErrorString(r__) : "NCCL error")); \
            return h->err_code;                                                                                    \
        }                                                                                                          \
    } while (0)

}  // namespace

struct GorderComm {
    void *comm = nullptr;
    int n_ranks = 0, rank = 0, device = 0;
    long long *d_scratch = nullptr;   // headers and frame lists of a merge (grown on demand)
    size_t scratch_words = 0;
};

namespace {
int comm_scratch(GorderHandle *h, GorderComm *c, size_t words, long long **out) {
    if (words > c->scratch_words) {
        if (c->d_scratch) cudaFree(c->d_scratch);
        c->d_scratch = nullptr; c->scratch_words = 0;
        const size_t cap = std::max<size_t>(words, 1 << 20);
        CK(cudaMalloc((void **)&c->d_scratch, cap * sizeof(long long)));
        c->scratch_words = cap;
    }
    *out = c->d_scratch;
    return GORDER_OK;
}
}  // namespace

extern "C" {

int gorder_gpu_reduce(GorderHandle **hs, int32_t n, int32_t root) {
    if (!hs || n < 1 || n > kMaxPeers + 1 || root < 0 || root >= n) return GORDER_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n; i++) if (!hs[i]) return GORDER_ERR_INVALID_ARGUMENT;
    GorderHandle *h = hs[root];
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) if (hs[i] == hs[j]) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "a handle appears twice"); return h->err_code; }
    // every shard finishes its queued work; the first error (in shard order) is the job's error
    std::vector<ShardMeta> meta((size_t)n);
    for (int i = 0; i < n; i++) {
        if (int rc = check_same_shape(h, hs[i])) return rc;
        const int rc = gorder_gpu_sync(hs[i]);
        meta[(size_t)i] = meta_of(hs[i]);
        meta[(size_t)i].err_code = rc ? rc : hs[i]->err_code;
    }
    for (int i = 0; i < n; i++)
        if (meta[(size_t)i].err_code) {
            if (i != root) { h->err_code = 0; h->set_error(meta[(size_t)i].err_code, "error in shard " + std::to_string(i) + ": " + hs[i]->err_msg, meta[(size_t)i].err_detail); }
            return h->err_code;
        }
    CK(cudaSetDevice(h->device));
    // peer access root -> every other device (a handle on the root's own device needs none)
    PeerBlocks peers{};
    for (int i = 0; i < n; i++) {
        if (i == root) continue;
        if (hs[i]->device != h->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, h->device, hs[i]->device));
            if (!can) { h->set_error(GORDER_ERR_CUDA, "no peer access between the shards' devices"); return h->err_code; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(hs[i]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->set_error(GORDER_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return h->err_code; }
            cudaGetLastError();
        }
        peers.p[peers.n++] = hs[i]->d_block;
    }
    const MergePlan plan = make_plan(meta);
    MergedArrays merged;
    if (int rc = alloc_merged(h, plan, root, &merged)) return rc;
    const size_t row = (size_t)h->n_slots * 3;
    for (int i = 0; i < n; i++) {   // gathers: peer-to-peer copies into the merged arrays (the root's own rows included)
        GorderHandle *s = hs[i];
        const ShardMeta &m = meta[(size_t)i];
        if (merged.bsum && m.n_frames > 0 && !(i == root && merged.rows_in_place)) {
            CK(cudaMemcpyPeerAsync(merged.bsum + (size_t)plan.row_off[(size_t)i] * row, h->device, s->d_bsum, s->device, (size_t)m.n_frames * row * sizeof(long long), h->stream));
            CK(cudaMemcpyPeerAsync(merged.bcnt + (size_t)plan.row_off[(size_t)i] * row, h->device, s->d_bcnt, s->device, (size_t)m.n_frames * row * sizeof(long long), h->stream));
        }
        if (merged.leaf && m.n_leaf > 0)
            CK(cudaMemcpyPeerAsync(merged.leaf + (size_t)plan.leaf_off[(size_t)i] * h->n_molpad, h->device, s->d_leaf_collect, s->device, (size_t)m.n_leaf * h->n_molpad, h->stream));
        if (merged.normals && m.n_frames > 0)
            CK(cudaMemcpyPeerAsync(merged.normals + (size_t)plan.row_off[(size_t)i] * 3 * h->n_molpad, h->device, s->d_normals_collect, s->device,
                                   (size_t)m.n_frames * 3 * h->n_molpad * sizeof(float), h->stream));
    }
    if (peers.n > 0 && h->block_words > 0) {   // the sum: one kernel reads the peers' blocks over NVLink
        const int blocks = (int)std::min<long long>((h->block_words / 2 + 255) / 256 + 1, 148 * 8);
        peer_block_sum_kernel<<<blocks, 256, 0, h->stream>>>(h->d_block, peers, h->block_words);
        h->n_launches++;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));
    adopt_merged(h, meta, plan, merged);
    return GORDER_OK;
}

int gorder_comm_unique_id(uint8_t *id) {
    if (!id) return GORDER_ERR_INVALID_ARGUMENT;
    Nccl *N = nccl();
    if (!N) return GORDER_ERR_NCCL;
    return N->GetUniqueId(id) == 0 ? GORDER_OK : GORDER_ERR_NCCL;
}

int gorder_comm_create(const uint8_t *id, int32_t n_ranks, int32_t rank, int32_t device, GorderComm **out) {
    if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    Nccl *N = nccl();
    if (!N) return GORDER_ERR_NCCL;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) { cudaGetLastError(); return GORDER_ERR_NO_DEVICE; }
    if (cudaSetDevice(device) != cudaSuccess) return GORDER_ERR_CUDA;
    NcclId uid;
    memcpy(uid.b, id, 128);
    GorderComm *c = new GorderComm();
    c->n_ranks = n_ranks; c->rank = rank; c->device = device;
    if (N->CommInitRank(&c->comm, n_ranks, uid, rank) != 0) { delete c; return GORDER_ERR_NCCL; }
    {   // NCCL connects its channels lazily, at the first use of every collective / peer pair (hundreds of milliseconds): do
        // it here, once, with the operations gorder_gpu_reduce_comm uses (all-gather, reduce, broadcast, send / recv between all pairs)
        cudaStream_t st = nullptr;
        long long *d = nullptr;
        bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess && cudaMalloc((void **)&d, (size_t)(2 * n_ranks + 2) * sizeof(long long)) == cudaSuccess;
        if (ok) {
            cudaMemsetAsync(d, 0, (size_t)(2 * n_ranks + 2) * sizeof(long long), st);
            ok = N->AllGather(d + n_ranks, d, 1, kNcclInt64, c->comm, st) == 0 && N->Reduce(d, d, 1, kNcclInt64, kNcclSum, 0, c->comm, st) == 0 &&
                 N->Broadcast(d, d, 8, kNcclInt8, 0, c->comm, st) == 0;
            if (ok && n_ranks > 1) {
                ok = N->GroupStart() == 0;
                for (int r = 0; ok && r < n_ranks; r++)
                    if (r != rank) ok = N->Send(d + n_ranks, 1, kNcclInt64, r, c->comm, st) == 0 && N->Recv(d + n_ranks + 1 + r, 1, kNcclInt64, r, c->comm, st) == 0;
                ok = N->GroupEnd() == 0 && ok;
            }
            ok = cudaStreamSynchronize(st) == cudaSuccess && ok;
        }
        if (d) cudaFree(d);
        if (st) cudaStreamDestroy(st);
        if (!ok) { N->CommDestroy(c->comm); delete c; cudaGetLastError(); return GORDER_ERR_NCCL; }
        // headers and frame lists of a merge: allocated now -- with peer access enabled (NCCL), cudaMalloc / cudaFree have to update
        // the peer mappings and take tens to hundreds of milliseconds, which must not happen inside the merge
        if (cudaMalloc((void **)&c->d_scratch, ((size_t)1 << 20) * sizeof(long long)) == cudaSuccess) c->scratch_words = (size_t)1 << 20;
        else cudaGetLastError();
    }
    *out = c;
    return GORDER_OK;
}

void gorder_comm_destroy(GorderComm *c) {
    if (!c) return;
    Nccl *N = nccl();
    if (N && c->comm) { cudaSetDevice(c->device); N->CommDestroy(c->comm); }
    if (c->d_scratch) cudaFree(c->d_scratch);
    delete c;
}

// Frequency::Once (leaflets.rs:435-441): the table computed from analysed frame 0 by the shard that owns it reaches the other
// shards before they accumulate per-leaflet sums.  Collective: rank `root` must have analysed frame 0.
int gorder_comm_broadcast_leaflets(GorderHandle *h, GorderComm *c, int32_t root) {
    if (!h || !c || root < 0 || root >= c->n_ranks) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    if (!h->leaf) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "leaflets are not enabled"); return h->err_code; }
    Nccl *N = nccl();
    if (!N) { h->set_error(GORDER_ERR_NCCL, "NCCL is not available"); return h->err_code; }
    CK(cudaSetDevice(h->device));
    if (int rc = sync_all(h)) return rc;
    // row 0 of the leaflet rows is the table carried from batch to batch (padded per molecule type)
    NK(N->Broadcast(h->d_leaf_rows, h->d_leaf_rows, (size_t)h->n_molpad, kNcclInt8, root, c->comm, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (c->rank != root) { h->have_leaflets = true; h->cur_leaflet_frame = 0; }
    return GORDER_OK;
}

int gorder_gpu_reduce_comm(GorderHandle *h, GorderComm *c, int32_t root) {
    if (!h || !c || root < 0 || root >= c->n_ranks) return GORDER_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(h->mu);
    Nccl *N = nccl();
    if (!N) { h->set_error(GORDER_ERR_NCCL, "NCCL is not available"); return h->err_code; }
    cudaSetDevice(h->device);
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (h->sw.verbose) fprintf(stderr, "[gorder_gpu_reduce_comm rank %d] %-28s %8.3f ms\n", c->rank, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    const int own_rc = gorder_gpu_sync(h);   // do not return yet: the collective below must be entered by every rank
    lap("own work finished");
    const int R = c->n_ranks, me = c->rank;
    const bool is_root = me == root;
    // ---- 1. header of every shard on every rank: frames, collected tables, error, shape ----
    constexpr int kHdr = 8;
    std::vector<long long> hdr_all((size_t)R * kHdr, 0);
    long long *d_hdr = nullptr;
    if (int rc = comm_scratch(h, c, (size_t)(R + 1) * kHdr, &d_hdr)) return rc;
    const long long mine[kHdr] = {h->n_frames, h->n_leaf_collected, own_rc ? own_rc : h->err_code, h->err_detail, h->block_words, h->n_slots, h->n_molpad,
                                  (long long)h->s.timewise | ((long long)h->s.collect_leaflets << 1) | ((long long)h->s.collect_normals << 2)};
    cudaMemcpyAsync(d_hdr + (size_t)R * kHdr, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream);
    int nrc = N->AllGather(d_hdr + (size_t)R * kHdr, d_hdr, kHdr, kNcclInt64, c->comm, h->stream);
    cudaMemcpyAsync(hdr_all.data(), d_hdr, (size_t)R * kHdr * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    if (nrc != 0) { h->err_code = 0; h->set_error(GORDER_ERR_NCCL, std::string("ncclAllGather: ") + N->GetErrorString(nrc)); return h->err_code; }
    lap("headers gathered");
    std::vector<ShardMeta> meta((size_t)R);
    for (int r = 0; r < R; r++) {
        const long long *q = hdr_all.data() + (size_t)r * kHdr;
        meta[(size_t)r].n_frames = q[0]; meta[(size_t)r].n_leaf = q[1]; meta[(size_t)r].err_code = (int)q[2]; meta[(size_t)r].err_detail = q[3];
        if (q[4] != mine[4] || q[5] != mine[5] || q[6] != mine[6] || q[7] != mine[7]) { h->set_error(GORDER_ERR_INVALID_ARGUMENT, "shards were created from different setups"); return h->err_code; }
    }
    for (int r = 0; r < R; r++)   // the first error in rank order is everybody's result (no rank is left inside a collective)
        if (meta[(size_t)r].err_code) {
            if (r != me) { h->err_code = 0; h->set_error(meta[(size_t)r].err_code, "error in shard " + std::to_string(r), meta[(size_t)r].err_detail); }
            return h->err_code;
        }
    // ---- 2. frame lists -> root (int64 lists over NCCL send / recv) ----
    long long total_idx = 0;
    for (int r = 0; r < R; r++) total_idx += meta[(size_t)r].n_frames + meta[(size_t)r].n_leaf;
    long long *d_idx = nullptr;
    const long long my_idx = h->n_frames + h->n_leaf_collected;
    if (int rc = comm_scratch(h, c, (size_t)std::max<long long>(1, is_root ? total_idx : my_idx), &d_idx)) return rc;
    std::vector<long long> own(h->frame_index_done);
    own.insert(own.end(), h->leaf_frame_index.begin(), h->leaf_frame_index.end());
    std::vector<long long> off_idx((size_t)R + 1, 0);
    for (int r = 0; r < R; r++) off_idx[(size_t)r + 1] = off_idx[(size_t)r] + meta[(size_t)r].n_frames + meta[(size_t)r].n_leaf;
    if (my_idx) CK(cudaMemcpyAsync(d_idx + (is_root ? off_idx[(size_t)me] : 0), own.data(), (size_t)my_idx * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    NK(N->GroupStart());
    if (is_root) {
        for (int r = 0; r < R; r++) { const long long cnt = off_idx[(size_t)r + 1] - off_idx[(size_t)r]; if (r != me && cnt) NK(N->Recv(d_idx + off_idx[(size_t)r], (size_t)cnt, kNcclInt64, r, c->comm, h->stream)); }
    } else if (my_idx) NK(N->Send(d_idx, (size_t)my_idx, kNcclInt64, root, c->comm, h->stream));
    NK(N->GroupEnd());
    MergePlan plan;
    MergedArrays merged;
    if (is_root) {
        std::vector<long long> all((size_t)std::max<long long>(1, total_idx));
        CK(cudaMemcpyAsync(all.data(), d_idx, (size_t)total_idx * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < R; r++) {
            const long long *q = all.data() + off_idx[(size_t)r];
            meta[(size_t)r].frame_index.assign(q, q + meta[(size_t)r].n_frames);
            meta[(size_t)r].leaf_frame_index.assign(q + meta[(size_t)r].n_frames, q + meta[(size_t)r].n_frames + meta[(size_t)r].n_leaf);
        }
        plan = make_plan(meta);
        if (int rc = alloc_merged(h, plan, me, &merged)) return rc;
    }
    lap("frame lists on the root");
    // ---- 3. the single sum-reduce of the accumulator block + the gathers of per-frame data, one NCCL group ----
    const size_t row = (size_t)h->n_slots * 3;
    const bool tw = h->s.timewise != 0, lf = h->s.collect_leaflets != 0, nm = h->s.collect_normals && h->s.normal_mode == GORDER_NORMAL_DYNAMIC;
    if (h->block_words > 0) NK(N->Reduce(h->d_block, h->d_block, (size_t)h->block_words, kNcclInt64, kNcclSum, root, c->comm, h->stream));
    NK(N->GroupStart());
    if (is_root) {
        for (int r = 0; r < R; r++) {
            const ShardMeta &m = meta[(size_t)r];
            const size_t ro = (size_t)plan.row_off[(size_t)r], lo = (size_t)plan.leaf_off[(size_t)r];
            if (r == me) {
                if (tw && m.n_frames && !merged.rows_in_place) {
                    CK(cudaMemcpyAsync(merged.bsum + ro * row, h->d_bsum, (size_t)m.n_frames * row * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
                    CK(cudaMemcpyAsync(merged.bcnt + ro * row, h->d_bcnt, (size_t)m.n_frames * row * sizeof(long long), cudaMemcpyDeviceToDevice, h->stream));
                }
                if (lf && m.n_leaf) CK(cudaMemcpyAsync(merged.leaf + lo * h->n_molpad, h->d_leaf_collect, (size_t)m.n_leaf * h->n_molpad, cudaMemcpyDeviceToDevice, h->stream));
                if (nm && m.n_frames) CK(cudaMemcpyAsync(merged.normals + ro * 3 * h->n_molpad, h->d_normals_collect, (size_t)m.n_frames * 3 * h->n_molpad * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
                continue;
            }
            if (tw && m.n_frames) {
                NK(N->Recv(merged.bsum + ro * row, (size_t)m.n_frames * row, kNcclInt64, r, c->comm, h->stream));
                NK(N->Recv(merged.bcnt + ro * row, (size_t)m.n_frames * row, kNcclInt64, r, c->comm, h->stream));
            }
            if (lf && m.n_leaf) NK(N->Recv(merged.leaf + lo * h->n_molpad, (size_t)m.n_leaf * h->n_molpad, kNcclInt8, r, c->comm, h->stream));
            if (nm && m.n_frames) NK(N->Recv(merged.normals + ro * 3 * h->n_molpad, (size_t)m.n_frames * 3 * h->n_molpad * sizeof(float), kNcclInt8, r, c->comm, h->stream));
        }
    } else {
        if (tw && h->n_frames) {
            NK(N->Send(h->d_bsum, (size_t)h->n_frames * row, kNcclInt64, root, c->comm, h->stream));
            NK(N->Send(h->d_bcnt, (size_t)h->n_frames * row, kNcclInt64, root, c->comm, h->stream));
        }
        if (lf && h->n_leaf_collected) NK(N->Send(h->d_leaf_collect, (size_t)h->n_leaf_collected * h->n_molpad, kNcclInt8, root, c->comm, h->stream));
        if (nm && h->n_frames) NK(N->Send(h->d_normals_collect, (size_t)h->n_frames * 3 * h->n_molpad * sizeof(float), kNcclInt8, root, c->comm, h->stream));
    }
    NK(N->GroupEnd());
    CK(cudaStreamSynchronize(h->stream));
    lap("block reduced, rows gathered");
    if (is_root) adopt_merged(h, meta, plan, merged);
    lap("done");
    return GORDER_OK;
}

}  // extern "C"

// cg_order: a C++ host on top of the C ABI (include/gorder_b200.h) -- what the reference's CGOrder driver
// (src/analysis/cgorder.rs:60-140: classify, read the trajectory, analyse every frame, convert) looks like when the
// per-frame engine is this library.  Topology: N Martini-POPC-like lipids of 12 beads stored one after another
// (NC3 PO4 GL1 GL2 C1A D2A C3A C4A C1B C2B C3B C4B, the 11 bonds of the reference's CG validation system), Global
// leaflets from the PO4 beads, membrane = all beads.
//
//   g++ -O2 -std=c++17 -Iinclude examples/cg_order.cpp -Lgorder_b200 -lgorder_b200 -Wl,-rpath,$PWD/gorder_b200 -o cg_order
//   ./cg_order trajectory.xtc [n_blocks]
//
// There is no CPU fallback: without a CUDA device gorder_gpu_create fails with GORDER_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gorder_b200.h"

static const char *kBeads[12] = {"NC3", "PO4", "GL1", "GL2", "C1A", "D2A", "C3A", "C4A", "C1B", "C2B", "C3B", "C4B"};
static const int32_t kBonds[11][2] = {{0, 1}, {1, 2}, {2, 3}, {2, 4}, {3, 8}, {4, 5}, {5, 6}, {6, 7}, {8, 9}, {9, 10}, {10, 11}};

static int fail(GorderHandle *h, const char *what, int rc) {
    char msg[256] = "";
    if (h) gorder_gpu_last_error(h, msg, sizeof msg);
    std::fprintf(stderr, "cg_order: %s failed with code %d %s\n", what, rc, msg);
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s trajectory.xtc [n_blocks]\n", argv[0]); return 2; }
    const int n_blocks = argc > 2 ? std::atoi(argv[2]) : 5;

    GorderXtc *xtc = nullptr;
    int rc = gorder_xtc_open(argv[1], &xtc);
    if (rc) return fail(nullptr, "gorder_xtc_open", rc);
    int32_t n_atoms = 0;
    int64_t n_frames = 0;
    float precision = 0;
    gorder_xtc_info(xtc, &n_atoms, &n_frames, &precision);
    if (n_atoms < 12 || n_atoms % 12) { std::fprintf(stderr, "cg_order: %d atoms is not a whole number of 12-bead lipids\n", n_atoms); return 2; }
    const int n_lipids = n_atoms / 12;

    // the classified topology (reference: SystemTopology after classify.rs): one molecule type, 11 bond types
    std::vector<int32_t> mol_base(n_lipids), membrane(n_atoms);
    for (int m = 0; m < n_lipids; m++) mol_base[m] = 12 * m;
    for (int a = 0; a < n_atoms; a++) membrane[a] = a;
    GorderMolType popc{};
    popc.n_molecules = n_lipids; popc.mol_base = mol_base.data();
    popc.n_bond_types = 11; popc.bond_rel = &kBonds[0][0];
    popc.head_rel = 1; popc.normal_head_rel = -1;
    GorderSetup s{};
    s.abi_version = GORDER_ABI_VERSION; s.kind = GORDER_KIND_CG; s.n_atoms = n_atoms; s.handle_pbc = 1; s.step = 1;
    s.n_moltypes = 1; s.moltypes = &popc;
    s.normal_mode = GORDER_NORMAL_STATIC; s.normal_axis = GORDER_AXIS_Z;
    s.leaflet_mode = GORDER_LEAFLET_GLOBAL; s.leaflet_axis = GORDER_AXIS_Z; s.leaflet_freq_kind = GORDER_FREQ_EVERY; s.leaflet_freq = 1;
    s.n_membrane = n_atoms; s.membrane = membrane.data();
    s.timewise = n_blocks > 0;   // per-frame sums for the block-averaging error

    GorderHandle *h = nullptr;
    if ((rc = gorder_gpu_create(&s, &h))) return fail(h, "gorder_gpu_create", rc);

    // the frame loop (reference: traj_iter_map_reduce + analyze_frame, common.rs:201-339), decode on the device
    int64_t bytes = 0;
    if ((rc = gorder_gpu_run_xtc_device(h, xtc, nullptr, 0, n_frames, 1, 0, 8, 32, &bytes))) return fail(h, "gorder_gpu_run_xtc_device", rc);

    // reduce + fetch (reference: ParallelTrajData::reduce), then the conversion (converter.rs)
    GorderResults r{};
    if ((rc = gorder_gpu_result_sizes(h, &r))) return fail(h, "gorder_gpu_result_sizes", rc);
    std::vector<int64_t> sum(3 * r.n_slots), tw_sum(3 * r.n_slots * r.n_frames), tw_index(r.n_frames);
    std::vector<uint64_t> count(3 * r.n_slots), tw_count(3 * r.n_slots * r.n_frames);
    r.sum = sum.data(); r.count = count.data(); r.tw_frame_index = tw_index.data();
    if (s.timewise) { r.tw_sum = tw_sum.data(); r.tw_count = tw_count.data(); }
    if ((rc = gorder_gpu_finish(h, &r))) return fail(h, "gorder_gpu_finish", rc);

    GorderRaw raw{};
    raw.n_slots = (int32_t)r.n_slots; raw.n_frames = s.timewise ? r.n_frames : 0;
    raw.sum = sum.data(); raw.count = count.data();
    if (s.timewise) { raw.tw_sum = tw_sum.data(); raw.tw_count = tw_count.data(); }
    std::printf("# %d lipids, %lld frames, %.1f MB of compressed frames sent to the GPU\n", n_lipids, (long long)r.n_frames, bytes / 1e6);
    std::printf("# %-12s %20s %20s %20s\n", "bond", "total", "upper", "lower");
    std::vector<int32_t> all;
    for (int32_t b = 0; b <= 11; b++) {
        float v[3], e[3];
        if (b < 11) { all.push_back(b); rc = gorder_results_order(&raw, &b, 1, n_blocks, 1, 1.0f, v, e); }
        else rc = gorder_results_order(&raw, all.data(), 11, n_blocks, 1, 1.0f, v, e);   // molecule average (OrderSummer)
        if (rc) return fail(h, "gorder_results_order", rc);
        char name[32];
        if (b < 11) std::snprintf(name, sizeof name, "%s - %s", kBeads[kBonds[b][0]], kBeads[kBonds[b][1]]);
        else std::snprintf(name, sizeof name, "average");
        std::printf("  %-12s", name);
        for (int k = 0; k < 3; k++) std::printf("   %8.4f +- %6.4f", v[k], e[k]);
        std::printf("\n");
    }
    gorder_gpu_destroy(h);
    gorder_xtc_close(xtc);
    return 0;
}

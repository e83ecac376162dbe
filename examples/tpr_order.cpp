// tpr_order: coarse-grained order parameters from a GROMACS run file and a trajectory, without the Rust front end --
// the reference's `gorder` run with
//     structure: system.tpr   trajectory: traj.xtc   analysis_type: !CGOrder "@membrane"
//     leaflets: !Global { membrane: "@membrane", heads: "name PO4" }   estimate_error: { n_blocks: 5 }
// (src/analysis/cgorder.rs:60-140) written against the C ABI only: gorder_system_from_tpr (structure.rs:27-88),
// gorder_classify_bonds (topology/classify.rs:45-315), gorder_gpu_create / gorder_gpu_run_xtc_device / gorder_gpu_finish (the
// per-frame engine), gorder_results_order (presentation/converter.rs).  The selection language is not part of the library:
// "@membrane" is a list of residue names here and the heads are picked by atom name.
//
//   g++ -O2 -std=c++17 -Iinclude examples/tpr_order.cpp -Lgorder_b200 -lgorder_b200 -Wl,-rpath,$PWD/gorder_b200 -o tpr_order
//   ./tpr_order system.tpr traj.xtc [POPC,POPE,POPG,...] [PO4] [n_blocks]
//
// There is no CPU fallback: without a CUDA device gorder_gpu_create fails with GORDER_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gorder_b200.h"

static int fail(GorderHandle *h, const char *what, int rc) {
    char msg[256] = "";
    if (h) gorder_gpu_last_error(h, msg, sizeof msg);
    std::fprintf(stderr, "tpr_order: %s failed with code %d %s%s\n", what, rc, msg, h ? "" : gorder_topology_last_error());
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s system.tpr traj.xtc [lipid residues, comma separated] [head atom] [n_blocks]\n", argv[0]); return 2; }
    const std::string lipids = std::string(",") + (argc > 3 ? argv[3] : "POPC,POPE,POPG,POPS,DOPC,DPPC,DOPE,DOPS") + ",";
    const char *head_name = argc > 4 ? argv[4] : "PO4";
    const int n_blocks = argc > 5 ? std::atoi(argv[5]) : 5;

    // structure + topology
    GorderSystem *sys = nullptr;
    int rc = gorder_system_from_tpr(argv[1], &sys);
    if (rc) return fail(nullptr, "gorder_system_from_tpr", rc);
    const int n_atoms = gorder_system_n_atoms(sys);
    std::vector<char> names(8 * (size_t)n_atoms), resn(8 * (size_t)n_atoms);
    gorder_system_atoms(sys, names.data(), resn.data(), nullptr, nullptr, nullptr, nullptr);
    std::vector<int32_t> beads, heads;
    for (int i = 0; i < n_atoms; i++) {
        if (lipids.find(std::string(",") + &resn[8 * (size_t)i] + ",") == std::string::npos) continue;
        beads.push_back(i);
        if (!std::strcmp(&names[8 * (size_t)i], head_name)) heads.push_back(i);
    }

    // molecule types
    GorderClassification *cls = nullptr;
    rc = gorder_classify_bonds(sys, beads.data(), (int32_t)beads.size(), beads.data(), (int32_t)beads.size(), heads.data(), (int32_t)heads.size(),
                               nullptr, 0, nullptr, 0, &cls);
    if (rc) return fail(nullptr, "gorder_classify_bonds", rc);
    const int n_types = gorder_classification_n_types(cls);
    if (n_types == 0) { std::fprintf(stderr, "tpr_order: %s\n", gorder_classification_warning(cls)); return 0; }
    const GorderMolType *mts = gorder_classification_moltypes(cls);

    // the engine: one slot per atom of the run file (the device decoder drops the atoms no group needs)
    GorderSetup s{};
    s.abi_version = GORDER_ABI_VERSION; s.kind = GORDER_KIND_CG; s.n_atoms = n_atoms; s.handle_pbc = 1; s.step = 1;
    s.n_moltypes = n_types; s.moltypes = mts;
    s.normal_mode = GORDER_NORMAL_STATIC; s.normal_axis = GORDER_AXIS_Z;
    s.leaflet_mode = GORDER_LEAFLET_GLOBAL; s.leaflet_axis = GORDER_AXIS_Z; s.leaflet_freq_kind = GORDER_FREQ_EVERY; s.leaflet_freq = 1;
    s.n_membrane = (int32_t)beads.size(); s.membrane = beads.data();
    s.timewise = n_blocks > 0;
    GorderHandle *h = nullptr;
    if ((rc = gorder_gpu_create(&s, &h))) return fail(h, "gorder_gpu_create", rc);

    // the frame loop
    GorderXtc *xtc = nullptr;
    if ((rc = gorder_xtc_open(argv[2], &xtc))) return fail(h, "gorder_xtc_open", rc);
    int32_t xtc_atoms = 0;
    int64_t n_frames = 0;
    float precision = 0;
    gorder_xtc_info(xtc, &xtc_atoms, &n_frames, &precision);
    if (xtc_atoms != n_atoms) { std::fprintf(stderr, "tpr_order: the trajectory has %d atoms, the run file %d\n", xtc_atoms, n_atoms); return 2; }
    int64_t bytes = 0;
    if ((rc = gorder_gpu_run_xtc_device(h, xtc, nullptr, 0, n_frames, 1, 0, 8, 32, &bytes))) return fail(h, "gorder_gpu_run_xtc_device", rc);

    // reduce + fetch, then the conversion
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

    std::printf("# %s: %d atoms, %lld bonds, tpx %d; %lld frames\n", argv[1], n_atoms, (long long)gorder_system_n_bonds(sys), gorder_system_tpx_version(sys),
                (long long)r.n_frames);
    auto row = [&](const char *label, const int32_t *slots, int n) -> int {
        float v[3], e[3];
        if (int rc2 = gorder_results_order(&raw, slots, n, n_blocks, 1, 1.0f, v, e)) return rc2;
        std::printf("  %-40s", label);
        for (int k = 0; k < 3; k++) std::printf("   %8.4f +- %6.4f", v[k], e[k]);
        std::printf("\n");
        return 0;
    };
    std::vector<int32_t> all;
    int32_t slot = 0;
    for (int t = 0; t < n_types; t++) {
        std::printf("# molecule type %s: %d molecules; %40s %20s %20s\n", gorder_classification_type_name(cls, t), mts[t].n_molecules, "total", "upper", "lower");
        std::vector<int32_t> mine;
        for (int b = 0; b < mts[t].n_bond_types; b++, slot++) {
            mine.push_back(slot); all.push_back(slot);
            if ((rc = row(gorder_classification_item_name(cls, t, b), &slot, 1))) return fail(h, "gorder_results_order", rc);
        }
        if ((rc = row("average", mine.data(), (int)mine.size()))) return fail(h, "gorder_results_order", rc);
    }
    std::printf("# system\n");
    if ((rc = row("average", all.data(), (int)all.size()))) return fail(h, "gorder_results_order", rc);
    gorder_gpu_destroy(h);
    gorder_xtc_close(xtc);
    gorder_classification_free(cls);
    gorder_system_free(sys);
    return 0;
}

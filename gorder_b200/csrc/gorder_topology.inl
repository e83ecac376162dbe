// The step BEFORE the hot path, host only: structure + topology of the system and the classification of its
// lipids into molecule types, i.e. everything gorder_gpu_create needs besides the options.
//
//   gorder_system_from_tpr        groan_rs System::from_file (TPR) -> minitpr 0.2.3 (Cargo.lock:939; not vendored).  The
//                                 reader below follows the published layout of GROMACS' tpxio.cpp (do_tpxheader, do_tpx_body,
//                                 do_mtop, do_ffparams / do_iparams, do_moltype, do_atoms, do_ilists, do_molblock, do_groups):
//                                 header in XDR, body in XDR (tpx < 119) or in the big-endian in-memory serialisation
//                                 (tpx >= 119: 1-byte bool / uchar, 2-byte ushort, strings as a 64-bit length + bytes).
//                                 Bonds = the interaction lists BONDS .. RESTRBONDS, CONSTR, CONSTRNC and the two O-H pairs
//                                 of SETTLE.  Pinned on the reference's 14 test TPRs (tpx 103, 122, 127): atom / residue names
//                                 against the .gro / .pdb files of the same systems, bonds against cg.bnd / pcpepg.bnd / the
//                                 CONECT records of ua_nobox.pdb, coordinates and boxes against the values SURVEY.md §8c
//                                 located by other means (tests/test_topology_cpu.py).
//   gorder_system_read_bonds      structure.rs:91-165 (read_bonds / parse_bonds_file) with BondsError's cases.
//   gorder_classify_bonds / _ua   topology/classify.rs:45-315 (MoleculesClassifier: molecule = connected component, molecule type
//                                 = topology of (relative index, residue name, atom name) + bonds, types in order of first
//                                 appearance, solve_name_conflicts :262-294, sanity_check_molecules :297-315), :355-420 (order
//                                 bonds), bond.rs:77-81 (order of the bond types), uaorder.rs:580-665 (get_atom_type),
//                                 common.rs:345-375 (get_reference_head), leaflets.rs:743-775 (methyls).
// Groups arrive as index lists: the selection language (GSL) stays with the host (out of scope).

namespace gtopo {

struct ParseError { std::string what; };

// ---- big-endian reader with bounds checks (a corrupt file must fail, not crash) -----------------------------------------
struct Reader {
    const unsigned char *b = nullptr;
    size_t n = 0, p = 0;
    bool mem = false;    // in-memory serialisation (tpx >= 119) instead of XDR
    bool dbl = false;    // double-precision file: a `real` is 8 bytes
    void need(size_t k) const { if (k > n - p) throw ParseError{"file ends inside a record"}; }
    uint32_t u32() { need(4); uint32_t v = ((uint32_t)b[p] << 24) | ((uint32_t)b[p + 1] << 16) | ((uint32_t)b[p + 2] << 8) | b[p + 3]; p += 4; return v; }
    int32_t i32() { return (int32_t)u32(); }
    uint64_t u64() { uint64_t hi = u32(); uint64_t lo = u32(); return (hi << 32) | lo; }
    float f32() { uint32_t v = u32(); float f; memcpy(&f, &v, 4); return f; }
    double f64() { uint64_t v = u64(); double d; memcpy(&d, &v, 8); return d; }
    double real() { return dbl ? f64() : (double)f32(); }
    void skip(size_t k) { need(k); p += k; }
    void skip_reals(size_t k) { if (k > (n - p) / (dbl ? 8 : 4)) throw ParseError{"file ends inside a record"}; p += k * (dbl ? 8 : 4); }
    void skip_ints(size_t k) { if (k > (n - p) / 4) throw ParseError{"file ends inside a record"}; p += 4 * k; }
    size_t count() { int32_t v = i32(); if (v < 0) throw ParseError{"negative count"}; return (size_t)v; }
    int uchar_() { if (mem) { need(1); return b[p++]; } return i32(); }
    int ushort_() { if (mem) { need(2); int v = (b[p] << 8) | b[p + 1]; p += 2; return v; } return i32(); }
    bool boolean() { if (mem) { need(1); return b[p++] != 0; } return i32() != 0; }
    std::string string_() {
        if (mem) { uint64_t len = u64(); if (len > n - p) throw ParseError{"string longer than the file"}; std::string s((const char *)b + p, (size_t)len); p += (size_t)len; return s; }
        (void)i32();   // gmx_fio_do_string: strlen + 1, then the XDR string (length, bytes, padding to 4)
        size_t len = count();
        size_t pad = (len + 3) / 4 * 4;
        if (pad > n - p) throw ParseError{"string longer than the file"};
        std::string s((const char *)b + p, len); p += pad;
        return s;
    }
};

// Interaction functions in the order of GROMACS' enumeration (ifunc); `since` = tpx version that introduced the entry (files
// written before it number their types without it and hold no interaction list for it: tpxio.cpp ftupd).
struct FuncDef { const char *name; int since; int n_real; int n_int_before; int n_int_after; };
enum : int { FT_BONDS = 0, FT_RESTRBONDS = 9, FT_CONSTR = 62, FT_CONSTRNC = 63, FT_SETTLE = 64 };
static const FuncDef kFuncs[] = {
    {"BONDS", 0, 4, 0, 0}, {"G96BONDS", 0, 4, 0, 0}, {"MORSE", 0, 6, 0, 0}, {"CUBICBONDS", 0, 3, 0, 0}, {"CONNBONDS", 0, 0, 0, 0},
    {"HARMONIC", 0, 4, 0, 0}, {"FENEBONDS", 0, 2, 0, 0}, {"TABBONDS", 0, -1, 0, 0}, {"TABBONDSNC", 0, -1, 0, 0}, {"RESTRBONDS", 0, 8, 0, 0},
    {"ANGLES", 0, 4, 0, 0}, {"G96ANGLES", 0, 4, 0, 0}, {"RESTRANGLES", 98, 2, 0, 0}, {"LINEAR_ANGLES", 0, 4, 0, 0}, {"CROSS_BOND_BONDS", 0, 3, 0, 0},
    {"CROSS_BOND_ANGLES", 0, 4, 0, 0}, {"UREY_BRADLEY", 0, 8, 0, 0}, {"QUARTIC_ANGLES", 0, 6, 0, 0}, {"TABANGLES", 0, -1, 0, 0},
    {"PDIHS", 0, 4, 0, 1}, {"RBDIHS", 0, 12, 0, 0}, {"RESTRDIHS", 98, 2, 0, 0}, {"CBTDIHS", 98, 6, 0, 0}, {"FOURDIHS", 0, 12, 0, 0},
    {"IDIHS", 0, 4, 0, 0}, {"PIDIHS", 0, 4, 0, 1}, {"TABDIHS", 0, -1, 0, 0}, {"CMAP", 0, 0, 2, 0},
    {"GB12", 0, 5, 0, 0}, {"GB13", 0, 5, 0, 0}, {"GB14", 0, 5, 0, 0}, {"GBPOL", 0, 0, 0, 0}, {"NPSOLVATION", 0, 0, 0, 0},
    {"LJ14", 0, 4, 0, 0}, {"COUL14", 0, 0, 0, 0}, {"LJC14_Q", 0, 5, 0, 0}, {"LJC_PAIRS_NB", 0, 4, 0, 0}, {"LJ", 0, 2, 0, 0}, {"BHAM", 0, 3, 0, 0},
    {"LJ_LR", 0, 0, 0, 0}, {"BHAM_LR", 0, 0, 0, 0}, {"DISPCORR", 0, 0, 0, 0}, {"COUL_SR", 0, 0, 0, 0}, {"COUL_LR", 0, 0, 0, 0}, {"RF_EXCL", 0, 0, 0, 0},
    {"COUL_RECIP", 0, 0, 0, 0}, {"LJ_RECIP", 0, 0, 0, 0}, {"DPD", 0, 0, 0, 0},
    {"POLARIZATION", 0, 1, 0, 0}, {"WATER_POL", 0, 6, 0, 0}, {"THOLE_POL", 0, 4, 0, 0}, {"ANHARM_POL", 0, 3, 0, 0},
    {"POSRES", 0, 12, 0, 0}, {"FBPOSRES", 0, 5, 1, 0}, {"DISRES", 0, 4, 2, 0}, {"DISRESVIOL", 0, 0, 0, 0}, {"ORIRES", 0, 3, 3, 0}, {"ORIRESDEV", 0, 0, 0, 0},
    {"ANGRES", 0, 4, 0, 1}, {"ANGRESZ", 0, 4, 0, 1}, {"DIHRES", 0, 6, 0, 0}, {"DIHRESVIOL", 0, 0, 0, 0},
    {"CONSTR", 0, 2, 0, 0}, {"CONSTRNC", 0, 2, 0, 0}, {"SETTLE", 0, 2, 0, 0},
    {"VSITE1", 121, 0, 0, 0}, {"VSITE2", 0, 1, 0, 0}, {"VSITE2FD", 118, 1, 0, 0}, {"VSITE3", 0, 2, 0, 0}, {"VSITE3FD", 0, 2, 0, 0}, {"VSITE3FAD", 0, 2, 0, 0},
    {"VSITE3OUT", 0, 3, 0, 0}, {"VSITE4FD", 0, 3, 0, 0}, {"VSITE4FDN", 0, 3, 0, 0}, {"VSITEN", 0, 1, 1, 0},
    {"COM_PULL", 0, 0, 0, 0}, {"DENSITYFITTING", 117, 0, 0, 0}, {"EQM", 0, 0, 0, 0}, {"EPOT", 0, 0, 0, 0}, {"EKIN", 0, 0, 0, 0}, {"ETOT", 0, 0, 0, 0},
    {"ECONSERVED", 0, 0, 0, 0}, {"TEMP", 0, 0, 0, 0}, {"VTEMP", 0, 0, 0, 0}, {"PDISPCORR", 0, 0, 0, 0}, {"PRES", 0, 0, 0, 0}, {"DVDL_CONSTR", 0, 0, 0, 0},
    {"DVDL", 0, 0, 0, 0}, {"DKDL", 0, 0, 0, 0}, {"DVDL_COUL", 0, 0, 0, 0}, {"DVDL_VDW", 0, 0, 0, 0}, {"DVDL_BONDED", 0, 0, 0, 0},
    {"DVDL_RESTRAINT", 0, 0, 0, 0}, {"DVDL_TEMPERATURE", 0, 0, 0, 0},
};
constexpr int kNFuncs = (int)(sizeof(kFuncs) / sizeof(kFuncs[0]));
constexpr int kTpxMin = 103, kTpxMax = 127;   // GROMACS 5.1 .. 2022: what the reference's test tree holds and what is verified

// parameters of one interaction type (do_iparams): only their size matters here
static void skip_iparams(Reader &r, int ft) {
    const FuncDef &d = kFuncs[ft];
    if (d.n_real < 0) { r.skip_reals(1); r.skip_ints(1); r.skip_reals(1); return; }   // tabulated: kA, table, kB
    r.skip_ints(d.n_int_before);
    r.skip_reals(d.n_real);
    r.skip_ints(d.n_int_after);
}

}  // namespace gtopo

struct GorderSystem {
    int n_atoms = 0;
    std::vector<std::string> name, resname;
    std::vector<int32_t> resid, atomic_number;
    std::vector<float> mass, charge;
    std::vector<float> xyz;                 // [n][3], empty if the file held no coordinates
    float box9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool has_box = false;
    int tpx_version = 0;
    std::vector<std::vector<int32_t>> bonded;   // per atom, sorted, unique (groan's AtomContainer order)
    size_t n_bonds() const { size_t k = 0; for (const auto &v : bonded) k += v.size(); return k / 2; }
    void add_bond(int a, int b) { bonded[a].push_back(b); bonded[b].push_back(a); }
    void finish_bonds() { for (auto &v : bonded) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); } }
};

namespace gtopo {

struct MolTypeTpr {
    std::vector<std::string> atom_name, res_name;
    std::vector<int32_t> res_nr, res_index, atomic_number;
    int n_res = 0;
    std::vector<float> mass, charge;
    std::vector<std::pair<int, int>> bonds;
};

static void parse_tpr(Reader &r, GorderSystem &sys) {
    // ---- header (always XDR): do_tpxheader ----
    const std::string version = r.string_();
    if (version.compare(0, 7, "VERSION") != 0) throw ParseError{"not a TPR file (no VERSION string)"};
    const int prec = r.i32();
    if (prec != 4 && prec != 8) throw ParseError{"unknown precision"};
    r.dbl = prec == 8;
    const int ver = r.i32();
    sys.tpx_version = ver;
    if (ver < kTpxMin || ver > kTpxMax)
        throw ParseError{"TPR version " + std::to_string(ver) + " is outside the supported range " + std::to_string(kTpxMin) + ".." + std::to_string(kTpxMax) +
                         " (GROMACS 5.1 - 2022)"};
    const int gen = r.i32();
    (void)r.string_();   // file tag
    const int natoms = r.i32(), ngtc = r.i32();
    if (natoms < 0 || ngtc < 0) throw ParseError{"negative atom count"};
    (void)r.i32();       // fep_state
    (void)r.real();      // lambda
    const bool b_ir = r.i32() != 0, b_top = r.i32() != 0, b_x = r.i32() != 0, b_v = r.i32() != 0, b_f = r.i32() != 0, b_box = r.i32() != 0;
    (void)b_ir; (void)b_v; (void)b_f;
    if (ver >= 119 && gen >= 27) { (void)r.u64(); r.mem = true; }   // size of the body; the body is serialised in memory
    // ---- do_tpx_state_first ----
    if (b_box) {
        for (int i = 0; i < 9; i++) sys.box9[i] = (float)r.real();
        sys.has_box = true;
        r.skip_reals(18);   // relative box, box velocity
    }
    r.skip_reals((size_t)ngtc);   // the former Berendsen lambdas
    if (!b_top) throw ParseError{"the TPR file holds no topology"};
    // ---- do_mtop ----
    std::vector<int> file_ft;   // function types as the writing version numbered them
    for (int i = 0; i < kNFuncs; i++) if (kFuncs[i].since <= ver) file_ft.push_back(i);
    const size_t nsym = r.count();
    std::vector<std::string> sym;
    for (size_t i = 0; i < nsym; i++) sym.push_back(r.string_());
    auto symstr = [&]() -> const std::string & { size_t k = r.count(); if (k >= sym.size()) throw ParseError{"symbol index out of range"}; return sym[k]; };
    (void)symstr();   // system name
    (void)r.i32();    // atnr
    const size_t ntypes = r.count();
    if (ntypes > (r.n - r.p) / 4) throw ParseError{"file ends inside a record"};
    std::vector<int> functype(ntypes);
    for (auto &f : functype) { f = r.i32(); if (f < 0 || f >= (int)file_ft.size()) throw ParseError{"unknown interaction function"}; }
    (void)r.f64();    // reppow
    (void)r.real();   // fudgeQQ
    for (size_t i = 0; i < ntypes; i++) skip_iparams(r, file_ft[functype[i]]);
    const size_t nmoltype = r.count();
    std::vector<MolTypeTpr> mts;
    for (size_t m = 0; m < nmoltype; m++) {
        MolTypeTpr mt;
        (void)symstr();
        const size_t nr = r.count(), nres = r.count();
        if (nr > (r.n - r.p) / 16) throw ParseError{"file ends inside a record"};
        std::vector<int> resind(nr);
        for (size_t a = 0; a < nr; a++) {   // do_atom
            mt.mass.push_back((float)r.real()); mt.charge.push_back((float)r.real());
            r.skip_reals(2);                 // mB, qB
            (void)r.ushort_(); (void)r.ushort_();   // type, typeB
            (void)r.i32();                   // ptype
            resind[a] = r.i32();
            mt.atomic_number.push_back(r.i32());
        }
        for (size_t a = 0; a < nr; a++) mt.atom_name.push_back(symstr());
        r.skip_ints(2 * nr);                 // type names A, B
        std::vector<std::string> rname; std::vector<int32_t> rnr;
        for (size_t j = 0; j < nres; j++) { rname.push_back(symstr()); rnr.push_back(r.i32()); (void)r.uchar_(); }
        for (size_t a = 0; a < nr; a++) {
            if (resind[a] < 0 || (size_t)resind[a] >= nres) throw ParseError{"residue index out of range"};
            mt.res_name.push_back(rname[resind[a]]); mt.res_nr.push_back(rnr[resind[a]]); mt.res_index.push_back(resind[a]);
        }
        mt.n_res = (int)nres;
        for (int ft : file_ft) {             // do_ilists
            const size_t len = r.count();
            if (len > (r.n - r.p) / 4) throw ParseError{"file ends inside a record"};
            const bool two = ft <= FT_RESTRBONDS || ft == FT_CONSTR || ft == FT_CONSTRNC;
            const size_t stride = two ? 3 : (ft == FT_SETTLE ? 4 : 0);
            if (stride && len % stride == 0) {
                for (size_t k = 0; k < len; k += stride) {
                    (void)r.i32();
                    int a[3] = {0, 0, 0};
                    for (size_t j = 1; j < stride; j++) { a[j - 1] = r.i32(); if (a[j - 1] < 0 || (size_t)a[j - 1] >= nr) throw ParseError{"bonded atom out of range"}; }
                    if (two) mt.bonds.emplace_back(a[0], a[1]);
                    else { mt.bonds.emplace_back(a[0], a[1]); mt.bonds.emplace_back(a[0], a[2]); }   // SETTLE: O-H1, O-H2
                }
            } else {
                if (stride) throw ParseError{"interaction list of unexpected length"};
                r.skip_ints(len);
            }
        }
        { const size_t ncg = r.count(); r.skip_ints(ncg + 1); }                       // the obsolete charge-group block
        { const size_t ne = r.count(), nra = r.count(); r.skip_ints(ne + 1); r.skip_ints(nra); }   // exclusions
        mts.push_back(std::move(mt));
    }
    const size_t nmolblock = r.count();
    sys.n_atoms = 0;
    // Residue numbers as GROMACS shows them (mtop_util: molecule types of ONE residue are renumbered consecutively over the
    // molecules of the system, starting behind the largest number a multi-residue type stores; the others keep theirs)
    long long next_res = 1;
    for (const auto &mt : mts) if (mt.n_res > 1) for (int v : mt.res_nr) next_res = std::max<long long>(next_res, (long long)v + 1);
    for (size_t bidx = 0; bidx < nmolblock; bidx++) {   // do_molblock
        const size_t type = r.count(), nmol = r.count();
        (void)r.i32();   // atoms per molecule
        { const size_t k = r.count(); r.skip_reals(3 * k); }
        { const size_t k = r.count(); r.skip_reals(3 * k); }
        if (type >= mts.size()) throw ParseError{"molecule block of an unknown type"};
        const MolTypeTpr &mt = mts[type];
        const size_t na = mt.atom_name.size();
        if (na && nmol > ((size_t)natoms - sys.n_atoms) / na) throw ParseError{"more atoms in the molecule blocks than in the header"};
        for (size_t mol = 0; mol < nmol; mol++) {
            const int base = sys.n_atoms;
            for (size_t a = 0; a < na; a++) {
                sys.name.push_back(mt.atom_name[a]); sys.resname.push_back(mt.res_name[a]);
                sys.resid.push_back(mt.n_res <= 1 ? (int32_t)(next_res + (long long)mol * mt.n_res + mt.res_index[a]) : mt.res_nr[a]);
                sys.atomic_number.push_back(mt.atomic_number[a]); sys.mass.push_back(mt.mass[a]); sys.charge.push_back(mt.charge[a]);
            }
            sys.n_atoms += (int)na;
            sys.bonded.resize(sys.n_atoms);
            for (const auto &bd : mt.bonds) if (bd.first != bd.second) sys.add_bond(base + bd.first, base + bd.second);
        }
        if (mt.n_res <= 1) next_res += (long long)nmol * mt.n_res;
    }
    if (r.i32() != natoms || sys.n_atoms != natoms) throw ParseError{"atom counts of header, topology and molecule blocks disagree"};
    sys.finish_bonds();
    if (ver >= 103 && r.boolean()) {      // intermolecular interactions: one more set of lists over global indices
        for (int ft : file_ft) {
            const size_t len = r.count();
            if (len > (r.n - r.p) / 4) throw ParseError{"file ends inside a record"};
            const bool two = ft <= FT_RESTRBONDS || ft == FT_CONSTR || ft == FT_CONSTRNC;
            if (two && len % 3 == 0) {
                for (size_t k = 0; k < len; k += 3) {
                    (void)r.i32();
                    const int a = r.i32(), b2 = r.i32();
                    if (a < 0 || a >= natoms || b2 < 0 || b2 >= natoms) throw ParseError{"bonded atom out of range"};
                    if (a != b2) sys.add_bond(a, b2);
                }
            } else r.skip_ints(len);
        }
        sys.finish_bonds();
    }
    if (!b_x) return;
    // ---- the rest of the topology, only to reach the coordinates ----
    {   // do_atomtypes
        const size_t nat = r.count();
        if (ver < 115) r.skip_reals(3 * nat);
        r.skip_ints(nat);
        if (ver < 115) r.skip_reals(2 * nat);
    }
    {   // do_cmap
        const size_t ngrid = r.count(), spacing = r.count();
        if (spacing > 4096) throw ParseError{"CMAP grid spacing out of range"};
        r.skip_reals(ngrid * spacing * spacing * 4);
    }
    {   // do_groups
        constexpr int kGroupTypes = 10;
        for (int g = 0; g < kGroupTypes; g++) { const size_t k = r.count(); r.skip_ints(k); }
        { const size_t k = r.count(); r.skip_ints(k); }
        for (int g = 0; g < kGroupTypes; g++) { const size_t k = r.count(); r.skip(r.mem ? k : 4 * k); }
    }
    if (ver >= 120) { const uint64_t k = r.u64(); if (k > (r.n - r.p) / 4) throw ParseError{"file ends inside a record"}; r.skip_ints((size_t)k); }
    // ---- do_tpx_state_second ----
    if ((size_t)natoms > (r.n - r.p) / (r.dbl ? 24 : 12)) throw ParseError{"file ends inside the coordinates"};
    sys.xyz.resize(3 * (size_t)natoms);
    for (auto &c : sys.xyz) c = (float)r.real();
}

}  // namespace gtopo

static thread_local std::string g_topology_error;

extern "C" {

const char *gorder_topology_last_error(void) { return g_topology_error.c_str(); }

int gorder_system_from_tpr(const char *path, GorderSystem **out) {
    if (!path || !out) return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) { g_topology_error = std::string("could not open '") + path + "'"; return GORDER_ERR_IO; }
    std::vector<unsigned char> buf;
    unsigned char chunk[1 << 16];
    size_t k;
    while ((k = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + k);
    fclose(f);
    auto *sys = new GorderSystem();
    gtopo::Reader r;
    r.b = buf.data(); r.n = buf.size();
    try {
        gtopo::parse_tpr(r, *sys);
    } catch (const gtopo::ParseError &e) {
        g_topology_error = std::string(path) + ": " + e.what + " (byte " + std::to_string(r.p) + ")";
        delete sys;
        return GORDER_ERR_TPR_FORMAT;
    } catch (const std::exception &e) {
        g_topology_error = std::string(path) + ": " + e.what();
        delete sys;
        return GORDER_ERR_TPR_FORMAT;
    }
    *out = sys;
    return GORDER_OK;
}

}  // extern "C"

// ---- GRO and PDB (groan_rs System::from_file for the two text formats gorder's tests use) ---------------------------------
namespace gtopo {

static bool read_text(const char *path, std::vector<std::string> &lines) {
    FILE *f = fopen(path, "r");
    if (!f) return false;
    std::string cur;
    char buf[4096];
    while (fgets(buf, sizeof(buf), f)) {
        cur += buf;
        if (!cur.empty() && cur.back() == '\n') { cur.pop_back(); if (!cur.empty() && cur.back() == '\r') cur.pop_back(); lines.push_back(cur); cur.clear(); }
    }
    if (!cur.empty()) lines.push_back(cur);
    fclose(f);
    return true;
}
static std::string field(const std::string &ln, size_t a, size_t b) {   // columns [a, b), trimmed
    if (a >= ln.size()) return "";
    std::string t = ln.substr(a, std::min(b, ln.size()) - a);
    const size_t i = t.find_first_not_of(' '), j = t.find_last_not_of(' ');
    return i == std::string::npos ? "" : t.substr(i, j - i + 1);
}
static bool to_float(const std::string &t, float &v) { if (t.empty()) return false; char *e = nullptr; v = strtof(t.c_str(), &e); return e && *e == 0; }
static bool to_int(const std::string &t, int &v) { if (t.empty()) return false; char *e = nullptr; long x = strtol(t.c_str(), &e, 10); v = (int)x; return e && *e == 0; }

// GRO: title, atom count, `%5d%-5s%5s%5d` + three positions of width n + 5 (n decimals: GROMACS writes any precision; the width
// is the distance between the decimal points), box line (3 or 9 numbers: xx yy zz xy xz yx yz zx zy)
static void parse_gro(const std::vector<std::string> &ln, GorderSystem &sys) {
    if (ln.size() < 3) throw ParseError{"GRO file with fewer than three lines"};
    int n;
    if (!to_int(field(ln[1], 0, ln[1].size()), n) || n < 0 || (size_t)n + 3 > ln.size()) throw ParseError{"GRO atom count does not match the file"};
    sys.n_atoms = n;
    sys.xyz.resize(3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const std::string &l = ln[2 + i];
        if (l.size() < 20 + 3 * 4) throw ParseError{"GRO atom line too short (line " + std::to_string(3 + i) + ")"};
        int resid;
        if (!to_int(field(l, 0, 5), resid)) throw ParseError{"GRO residue number (line " + std::to_string(3 + i) + ")"};
        sys.resid.push_back(resid); sys.resname.push_back(field(l, 5, 10)); sys.name.push_back(field(l, 10, 15));
        sys.atomic_number.push_back(0); sys.mass.push_back(0.0f); sys.charge.push_back(0.0f);
        const size_t d1 = l.find('.', 20), d2 = d1 == std::string::npos ? d1 : l.find('.', d1 + 1);
        if (d2 == std::string::npos) throw ParseError{"GRO positions (line " + std::to_string(3 + i) + ")"};
        const size_t w = d2 - d1;
        for (int k = 0; k < 3; k++)
            if (!to_float(field(l, 20 + k * w, 20 + (k + 1) * w), sys.xyz[3 * (size_t)i + k])) throw ParseError{"GRO positions (line " + std::to_string(3 + i) + ")"};
    }
    float b[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int nb = 0;
    {
        const std::string &l = ln[2 + (size_t)n];
        size_t i = 0;
        while (i < l.size() && nb < 9) {
            while (i < l.size() && isspace((unsigned char)l[i])) i++;
            size_t j = i;
            while (j < l.size() && !isspace((unsigned char)l[j])) j++;
            if (j > i && !to_float(l.substr(i, j - i), b[nb++])) throw ParseError{"GRO box line"};
            i = j;
        }
    }
    if (nb == 3 || nb == 9) {   // row-major box matrix: rows = box vectors
        sys.box9[0] = b[0]; sys.box9[4] = b[1]; sys.box9[8] = b[2];
        if (nb == 9) { sys.box9[1] = b[3]; sys.box9[2] = b[4]; sys.box9[3] = b[5]; sys.box9[5] = b[6]; sys.box9[6] = b[7]; sys.box9[7] = b[8]; }
        sys.has_box = true;
    }
    sys.bonded.resize(n);
}

// PDB: ATOM / HETATM of the first model (Angstrom -> nm), CRYST1, CONECT.  Returns false when the serial numbers repeat
// (ParsePdbConnectivityError::DuplicateAtomNumbers: the CONECT records are ambiguous then).
static bool parse_pdb(const std::vector<std::string> &ln, GorderSystem &sys, bool &has_conect) {
    std::vector<std::pair<int, int>> serial_index;
    bool model_done = false, dup = false;
    std::vector<std::pair<int, int>> conect;
    has_conect = false;
    for (const std::string &l : ln) {
        const std::string rec = l.substr(0, std::min<size_t>(6, l.size()));
        if ((rec == "ATOM  " || rec == "HETATM") && !model_done) {
            if (l.size() < 54) throw ParseError{"PDB atom record too short"};
            int serial = 0, resid = 0;
            to_int(field(l, 6, 11), serial);
            if (!to_int(field(l, 22, 26), resid)) resid = 0;
            float x, y, z;
            if (!to_float(field(l, 30, 38), x) || !to_float(field(l, 38, 46), y) || !to_float(field(l, 46, 54), z)) throw ParseError{"PDB coordinates"};
            serial_index.emplace_back(serial, sys.n_atoms);
            sys.name.push_back(field(l, 12, 16)); sys.resname.push_back(field(l, 17, 21)); sys.resid.push_back(resid);
            sys.atomic_number.push_back(0); sys.mass.push_back(0.0f); sys.charge.push_back(0.0f);
            sys.xyz.push_back(x / 10.0f); sys.xyz.push_back(y / 10.0f); sys.xyz.push_back(z / 10.0f);
            sys.n_atoms++;
        } else if (rec == "CRYST1") {
            float a, b, c, al = 90, be = 90, ga = 90;
            if (to_float(field(l, 6, 15), a) && to_float(field(l, 15, 24), b) && to_float(field(l, 24, 33), c)) {
                to_float(field(l, 33, 40), al); to_float(field(l, 40, 47), be); to_float(field(l, 47, 54), ga);
                const double d2r = 3.14159265358979323846 / 180.0, ca = cos(al * d2r), cb = cos(be * d2r), cg = cos(ga * d2r), sg = sin(ga * d2r);
                a /= 10.0f; b /= 10.0f; c /= 10.0f;
                auto clean = [](double v) { return fabs(v) < 1e-6 ? 0.0f : (float)v; };
                sys.box9[0] = a; sys.box9[3] = clean(b * cg); sys.box9[4] = clean(b * sg);
                sys.box9[6] = clean(c * cb); sys.box9[7] = clean(c * (ca - cb * cg) / sg);
                const double z2 = (double)c * c - (double)sys.box9[6] * sys.box9[6] - (double)sys.box9[7] * sys.box9[7];
                sys.box9[8] = (float)sqrt(z2 > 0 ? z2 : 0.0);
                sys.has_box = true;
            }
        } else if (rec == "CONECT") {
            has_conect = true;
            int first = 0;
            if (!to_int(field(l, 6, 11), first)) continue;
            for (size_t k = 11; k + 1 <= l.size() && k < 31; k += 5) {
                int other;
                if (to_int(field(l, k, k + 5), other)) conect.emplace_back(first, other);
            }
        } else if (rec.compare(0, 6, "ENDMDL") == 0) model_done = true;
    }
    sys.bonded.resize(sys.n_atoms);
    std::sort(serial_index.begin(), serial_index.end());
    for (size_t i = 1; i < serial_index.size(); i++) if (serial_index[i].first == serial_index[i - 1].first) dup = true;
    if (dup && has_conect) return false;
    auto find = [&](int serial) -> int {
        auto it = std::lower_bound(serial_index.begin(), serial_index.end(), std::make_pair(serial, -1));
        return it != serial_index.end() && it->first == serial ? it->second : -1;
    };
    for (const auto &c : conect) {
        const int a = find(c.first), b = find(c.second);
        if (a < 0 || b < 0) throw ParseError{"CONECT record names an atom that does not exist"};
        if (a != b) sys.add_bond(a, b);
    }
    sys.finish_bonds();
    return true;
}

static bool ends_with(const std::string &s, const char *suffix) {
    const size_t n = strlen(suffix);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; i++) if (tolower((unsigned char)s[s.size() - n + i]) != suffix[i]) return false;
    return true;
}

}  // namespace gtopo

extern "C" int gorder_system_read_bonds(GorderSystem *s, const char *bonds_file);

// read_structure_and_topology (structure.rs:27-88): a bonds file, when given, replaces whatever topology the structure holds;
// without one a TPR brings its own bonds, a PDB its CONECT records (none: NoTopology; repeated atom numbers:
// InvalidPdbTopology) and a GRO file has no topology (NoTopology).  Box checks (check_box) are the engine's.
extern "C" int gorder_system_from_file(const char *structure, const char *bonds_file, GorderSystem **out) {
    if (!structure || !out) return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    const std::string path(structure);
    GorderSystem *sys = nullptr;
    int rc = GORDER_OK;
    if (gtopo::ends_with(path, ".tpr")) {
        if ((rc = gorder_system_from_tpr(structure, &sys))) return rc;
    } else if (gtopo::ends_with(path, ".gro") || gtopo::ends_with(path, ".pdb")) {
        const bool pdb = gtopo::ends_with(path, ".pdb");
        std::vector<std::string> lines;
        if (!gtopo::read_text(structure, lines)) { g_topology_error = "could not open '" + path + "'"; return GORDER_ERR_IO; }
        sys = new GorderSystem();
        bool has_conect = false, ok = true;
        try {
            if (pdb) ok = gtopo::parse_pdb(lines, *sys, has_conect); else gtopo::parse_gro(lines, *sys);
        } catch (const gtopo::ParseError &e) {
            g_topology_error = path + ": " + e.what;
            delete sys;
            return GORDER_ERR_STRUCTURE_FORMAT;
        }
        if (!bonds_file) {
            if (pdb && !ok) { g_topology_error = "cannot parse topology from the provided PDB file '" + path + "' - non-unique atom numbers make the CONECT information ambiguous"; delete sys; return GORDER_ERR_PDB_TOPOLOGY; }
            if (!pdb || !has_conect || sys->n_bonds() == 0) {
                g_topology_error = "the input structure file '" + path + "' does not contain topology information (hint: provide a `bonds` file)";
                delete sys;
                return GORDER_ERR_NO_TOPOLOGY;
            }
        }
    } else {
        g_topology_error = "the provided structure file '" + path + "' has an unknown, invalid, or unsupported format";
        return GORDER_ERR_STRUCTURE_FORMAT;
    }
    if (bonds_file && (rc = gorder_system_read_bonds(sys, bonds_file))) { delete sys; return rc; }
    *out = sys;
    return GORDER_OK;
}

extern "C" {

int gorder_system_from_arrays(int32_t n_atoms, const char *const *atom_names, const char *const *res_names, const int32_t *res_ids,
                              const float *xyz, const float *box9, GorderSystem **out) {
    if (!out || n_atoms < 0 || (n_atoms > 0 && (!atom_names || !res_names))) return GORDER_ERR_INVALID_ARGUMENT;
    auto *sys = new GorderSystem();
    sys->n_atoms = n_atoms;
    for (int i = 0; i < n_atoms; i++) {
        sys->name.emplace_back(atom_names[i] ? atom_names[i] : ""); sys->resname.emplace_back(res_names[i] ? res_names[i] : "");
        sys->resid.push_back(res_ids ? res_ids[i] : 0); sys->atomic_number.push_back(0); sys->mass.push_back(0.0f); sys->charge.push_back(0.0f);
    }
    if (xyz) sys->xyz.assign(xyz, xyz + 3 * (size_t)n_atoms);
    if (box9) { memcpy(sys->box9, box9, sizeof(sys->box9)); sys->has_box = true; }
    sys->bonded.resize(n_atoms);
    *out = sys;
    return GORDER_OK;
}

void gorder_system_free(GorderSystem *s) { delete s; }

int32_t gorder_system_n_atoms(const GorderSystem *s) { return s ? s->n_atoms : -1; }
int64_t gorder_system_n_bonds(const GorderSystem *s) { return s ? (int64_t)s->n_bonds() : -1; }
int32_t gorder_system_tpx_version(const GorderSystem *s) { return s ? s->tpx_version : -1; }

// names: [n][8] bytes each, NUL-padded (GROMACS names are at most 5 characters; longer names are cut at 7)
int gorder_system_atoms(const GorderSystem *s, char *atom_names8, char *res_names8, int32_t *res_ids, int32_t *atomic_numbers, float *masses, float *charges) {
    if (!s) return GORDER_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < s->n_atoms; i++) {
        if (atom_names8) { memset(atom_names8 + 8 * (size_t)i, 0, 8); strncpy(atom_names8 + 8 * (size_t)i, s->name[i].c_str(), 7); }
        if (res_names8) { memset(res_names8 + 8 * (size_t)i, 0, 8); strncpy(res_names8 + 8 * (size_t)i, s->resname[i].c_str(), 7); }
        if (res_ids) res_ids[i] = s->resid[i];
        if (atomic_numbers) atomic_numbers[i] = s->atomic_number[i];
        if (masses) masses[i] = s->mass[i];
        if (charges) charges[i] = s->charge[i];
    }
    return GORDER_OK;
}
// pairs: [n_bonds][2], i < j, sorted
int gorder_system_bonds(const GorderSystem *s, int32_t *pairs) {
    if (!s || !pairs) return GORDER_ERR_INVALID_ARGUMENT;
    size_t k = 0;
    for (int i = 0; i < s->n_atoms; i++) for (int j : s->bonded[i]) if (j > i) { pairs[k++] = i; pairs[k++] = j; }
    return GORDER_OK;
}
// returns 1 / 0 in *has: a structure may come without coordinates or box
int gorder_system_positions(const GorderSystem *s, float *xyz, int32_t *has) {
    if (!s) return GORDER_ERR_INVALID_ARGUMENT;
    const bool h = !s->xyz.empty() || s->n_atoms == 0;
    if (has) *has = h ? 1 : 0;
    if (h && xyz) memcpy(xyz, s->xyz.data(), s->xyz.size() * sizeof(float));
    return GORDER_OK;
}
int gorder_system_box(const GorderSystem *s, float *box9, int32_t *has) {
    if (!s) return GORDER_ERR_INVALID_ARGUMENT;
    if (has) *has = s->has_box ? 1 : 0;
    if (box9) memcpy(box9, s->box9, sizeof(s->box9));
    return GORDER_OK;
}
int gorder_system_set_bonds(GorderSystem *s, const int32_t *pairs, int64_t n_pairs) {
    if (!s || n_pairs < 0 || (n_pairs > 0 && !pairs)) return GORDER_ERR_INVALID_ARGUMENT;
    for (int64_t k = 0; k < n_pairs; k++) {
        const int a = pairs[2 * k], b = pairs[2 * k + 1];
        if (a < 0 || b < 0 || a >= s->n_atoms || b >= s->n_atoms) { g_topology_error = "bond with an atom outside the system"; return GORDER_ERR_BONDS_ATOM_NOT_FOUND; }
        if (a == b) { g_topology_error = "atom bonded to itself"; return GORDER_ERR_BONDS_SELF; }
    }
    for (auto &v : s->bonded) v.clear();
    for (int64_t k = 0; k < n_pairs; k++) s->add_bond(pairs[2 * k], pairs[2 * k + 1]);
    s->finish_bonds();
    return GORDER_OK;
}

// structure.rs:91-165: `i j k ...` (serial numbers from 1) = atom i is bonded to j, k, ...; '#' starts a comment; lines with
// fewer than two fields are skipped; duplicates are ignored; all bonds set before are dropped.
int gorder_system_read_bonds(GorderSystem *s, const char *bonds_file) {
    if (!s || !bonds_file) return GORDER_ERR_INVALID_ARGUMENT;
    FILE *f = fopen(bonds_file, "r");
    if (!f) { g_topology_error = std::string("could not open the bonds file '") + bonds_file + "'"; return GORDER_ERR_IO; }
    std::vector<std::vector<int32_t>> bonded(s->n_atoms);
    std::string line;
    int rc = GORDER_OK;
    char buf[4096];
    auto parse = [&](const std::string &tok, long long &v) -> bool {   // Rust's usize::from_str: digits only (an optional '+')
        size_t i = 0;
        if (i < tok.size() && tok[i] == '+') i++;
        if (i == tok.size()) return false;
        v = 0;
        for (; i < tok.size(); i++) { if (tok[i] < '0' || tok[i] > '9') return false; v = v * 10 + (tok[i] - '0'); if (v > (1LL << 40)) return false; }
        return true;
    };
    bool more = true;
    while (more && rc == GORDER_OK) {
        line.clear();
        for (;;) {   // one line of any length
            if (!fgets(buf, sizeof(buf), f)) { more = false; break; }
            line += buf;
            if (!line.empty() && line.back() == '\n') break;
        }
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.resize(hash);
        std::vector<std::string> tok;
        size_t i = 0;
        while (i < line.size()) {
            while (i < line.size() && isspace((unsigned char)line[i])) i++;
            size_t j = i;
            while (j < line.size() && !isspace((unsigned char)line[j])) j++;
            if (j > i) tok.push_back(line.substr(i, j - i));
            i = j;
        }
        if (tok.size() < 2) continue;
        long long target;
        if (!parse(tok[0], target)) { g_topology_error = "could not read '" + tok[0] + "' as an atom serial number"; rc = GORDER_ERR_BONDS_PARSE; break; }
        if (target > s->n_atoms) { g_topology_error = "atom with serial number '" + tok[0] + "' does not exist"; rc = GORDER_ERR_BONDS_ATOM_NOT_FOUND; break; }
        for (size_t k = 1; k < tok.size(); k++) {
            long long a;
            if (!parse(tok[k], a)) { g_topology_error = "could not read '" + tok[k] + "' as an atom serial number"; rc = GORDER_ERR_BONDS_PARSE; break; }
            if (a == target) { g_topology_error = "atom with serial number '" + tok[k] + "' claims to be bonded to itself"; rc = GORDER_ERR_BONDS_SELF; break; }
            if (a > s->n_atoms) { g_topology_error = "atom with serial number '" + tok[k] + "' does not exist"; rc = GORDER_ERR_BONDS_ATOM_NOT_FOUND; break; }
            if (a < 1 || target < 1) { g_topology_error = "atom serial numbers start at 1"; rc = GORDER_ERR_BONDS_ATOM_NOT_FOUND; break; }   // the reference panics on 0
            bonded[target - 1].push_back((int32_t)(a - 1));
            bonded[a - 1].push_back((int32_t)(target - 1));
        }
    }
    fclose(f);
    if (rc != GORDER_OK) return rc;
    s->bonded.swap(bonded);
    s->finish_bonds();
    return GORDER_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------------
// GROMACS index files (groan_rs Groups::from_ndx) and the leaflet tables read from them (leaflets.rs:1030-1215)
// ---------------------------------------------------------------------------------------------------------------------------
struct GorderNdx {
    std::vector<std::string> names;                 // in file order; a repeated name replaces the earlier group (and is reported)
    std::vector<std::vector<int32_t>> atoms;        // 0-based, as listed
    std::vector<std::string> invalid, duplicate;    // names refused ('"&|!@()<>= are not allowed), names that occurred twice
    int find(const std::string &n) const { for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return (int)i; return -1; }
};

namespace gtopo {

static int parse_ndx(const char *path, int n_atoms, GorderNdx &nd) {
    std::vector<std::string> lines;
    if (!read_text(path, lines)) { g_topology_error = std::string("could not open the ndx file '") + path + "'"; return GORDER_ERR_IO; }
    int cur = -1;       // group being filled; -1: none yet, -2: a group with an invalid name (its atoms are skipped)
    for (size_t li = 0; li < lines.size(); li++) {
        const std::string &l = lines[li];
        const size_t a = l.find_first_not_of(" \t");
        if (a == std::string::npos) continue;
        if (l[a] == '[') {
            const size_t b = l.find(']', a);
            if (b == std::string::npos) { g_topology_error = std::string(path) + ": line " + std::to_string(li + 1) + ": group header without ']'"; return GORDER_ERR_NDX_PARSE; }
            const std::string name = field(l, a + 1, b);
            if (name.empty() || name.find_first_of("'\"&|!@()<>=") != std::string::npos) { nd.invalid.push_back(name); cur = -2; continue; }
            const int old = nd.find(name);
            if (old >= 0) { nd.duplicate.push_back(name); nd.atoms[old].clear(); cur = old; }
            else { nd.names.push_back(name); nd.atoms.emplace_back(); cur = (int)nd.names.size() - 1; }
            continue;
        }
        size_t i = a;
        while (i < l.size()) {
            while (i < l.size() && isspace((unsigned char)l[i])) i++;
            size_t j = i;
            while (j < l.size() && !isspace((unsigned char)l[j])) j++;
            if (j > i) {
                int v;
                if (!to_int(l.substr(i, j - i), v) || v < 1) { g_topology_error = std::string(path) + ": line " + std::to_string(li + 1) + ": '" + l.substr(i, j - i) + "' is not an atom number"; return GORDER_ERR_NDX_PARSE; }
                if (n_atoms >= 0 && v > n_atoms) { g_topology_error = std::string(path) + ": atom number " + std::to_string(v) + " does not exist (the system has " + std::to_string(n_atoms) + " atoms)"; return GORDER_ERR_NDX_PARSE; }
                if (cur == -1) { g_topology_error = std::string(path) + ": atom numbers before the first group header"; return GORDER_ERR_NDX_PARSE; }
                if (cur >= 0) nd.atoms[cur].push_back(v - 1);
            }
            i = j;
        }
    }
    return GORDER_OK;
}

}  // namespace gtopo

extern "C" {

// n_atoms < 0: atom numbers are not range-checked
int gorder_ndx_open(const char *path, int32_t n_atoms, GorderNdx **out) {
    if (!path || !out) return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    auto nd = std::make_unique<GorderNdx>();
    if (int rc = gtopo::parse_ndx(path, n_atoms, *nd)) return rc;
    *out = nd.release();
    return GORDER_OK;
}
void gorder_ndx_close(GorderNdx *n) { delete n; }
int32_t gorder_ndx_n_groups(const GorderNdx *n) { return n ? (int32_t)n->names.size() : -1; }
const char *gorder_ndx_group_name(const GorderNdx *n, int32_t g) { return n && g >= 0 && g < (int32_t)n->names.size() ? n->names[g].c_str() : nullptr; }
int64_t gorder_ndx_group_size(const GorderNdx *n, int32_t g) { return n && g >= 0 && g < (int32_t)n->names.size() ? (int64_t)n->atoms[g].size() : -1; }
const int32_t *gorder_ndx_group_atoms(const GorderNdx *n, int32_t g) { return n && g >= 0 && g < (int32_t)n->names.size() ? n->atoms[g].data() : nullptr; }
int32_t gorder_ndx_find(const GorderNdx *n, const char *name) { return n && name ? n->find(name) : -1; }

// LeafletClassification::FromNdx (leaflets.rs:1030-1215, NdxClassification): one ndx file per assignment frame; a molecule is
// Upper if its head is in the group `upper`, else Lower if it is in `lower`, else AssignmentNotFound.  table[f][m] for the
// n_molecules heads (absolute atom indices, in molecule order) = GORDER_UPPER / GORDER_LOWER: GorderMolType.manual_leaflets of
// GORDER_LEAFLET_MANUAL (a run longer than n_files assignment frames then fails with GORDER_ERR_MANUAL_LEAFLET_FRAME =
// NdxLeafletClassificationError::FrameNotFound).  Invalid / repeated group names matter only when they are one of the two.
int gorder_leaflets_from_ndx(const char *const *ndx_files, int32_t n_files, int32_t n_atoms, const char *upper, const char *lower,
                             const int32_t *heads, int32_t n_molecules, uint8_t *table) {
    if (n_files < 0 || n_molecules < 0 || !upper || !lower || (n_files > 0 && !ndx_files) || (n_molecules > 0 && (!heads || !table))) return GORDER_ERR_INVALID_ARGUMENT;
    std::vector<char> in_up, in_lo;
    for (int f = 0; f < n_files; f++) {
        GorderNdx nd;
        if (int rc = gtopo::parse_ndx(ndx_files[f], n_atoms, nd)) return rc;
        for (const auto &nm : nd.invalid)
            if (nm == upper || nm == lower) { g_topology_error = "group name '" + nm + "' specified in an ndx file '" + ndx_files[f] + "' is invalid and cannot be used"; return GORDER_ERR_NDX_INVALID_NAME; }
        for (const auto &nm : nd.duplicate)
            if (nm == upper || nm == lower) { g_topology_error = "multiple groups named '" + nm + "' are specified in an ndx file '" + ndx_files[f] + "'"; return GORDER_ERR_NDX_DUPLICATE_NAME; }
        const int gu = nd.find(upper), gl = nd.find(lower);
        if (gu < 0 || gl < 0) {
            g_topology_error = std::string("group '") + (gu < 0 ? upper : lower) + "' expected to specify " + (gu < 0 ? "upper" : "lower") + "-leaflet lipids not found in the ndx file '" + ndx_files[f] + "'";
            return GORDER_ERR_NDX_GROUP_NOT_FOUND;
        }
        int hi = 0;
        for (int m = 0; m < n_molecules; m++) hi = std::max(hi, heads[m]);
        for (int a : nd.atoms[gu]) hi = std::max(hi, a);
        for (int a : nd.atoms[gl]) hi = std::max(hi, a);
        in_up.assign((size_t)hi + 1, 0); in_lo.assign((size_t)hi + 1, 0);
        for (int a : nd.atoms[gu]) in_up[a] = 1;
        for (int a : nd.atoms[gl]) in_lo[a] = 1;
        for (int m = 0; m < n_molecules; m++) {
            if (heads[m] < 0) return GORDER_ERR_INVALID_ARGUMENT;
            if (in_up[heads[m]]) table[(size_t)f * n_molecules + m] = GORDER_UPPER;
            else if (in_lo[heads[m]]) table[(size_t)f * n_molecules + m] = GORDER_LOWER;
            else {
                g_topology_error = "could not assign molecule " + std::to_string(m) + " (head atom index " + std::to_string(heads[m]) + ") to a leaflet: not in '" + upper + "' nor '" + lower + "' of '" + ndx_files[f] + "'";
                return GORDER_ERR_NDX_ASSIGNMENT_NOT_FOUND;
            }
        }
    }
    return GORDER_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------------
// classification
// ---------------------------------------------------------------------------------------------------------------------------
struct GorderClassification {
    struct Type {
        std::string name;
        std::vector<int32_t> mol_base, atoms_rel;
        std::vector<int32_t> bond_rel;             // [n][2]
        std::vector<int32_t> ua_kind, ua_rel;      // [n], [n][4]
        std::vector<int32_t> methyl_rel;
        int32_t head_rel = -1, normal_head_rel = -1;
        std::vector<std::string> item_names;
    };
    std::vector<Type> types;
    std::vector<GorderMolType> c_types;
    std::string warning;
};

namespace gtopo {

struct Component {
    std::vector<int32_t> atoms;   // sorted
};

static Component component_of(const GorderSystem &s, int start, std::vector<char> &visited) {
    Component c;
    std::vector<int32_t> stack{start};
    visited[start] = 1;
    while (!stack.empty()) {
        const int a = stack.back(); stack.pop_back();
        c.atoms.push_back(a);
        for (int b : s.bonded[a]) if (!visited[b]) { visited[b] = 1; stack.push_back(b); }
    }
    std::sort(c.atoms.begin(), c.atoms.end());
    return c;
}

struct RawType {
    std::vector<int32_t> atoms_rel;                      // relative indices of the molecule's atoms
    std::vector<std::pair<int32_t, int32_t>> bonds_rel;  // sorted, i < j
    std::vector<int32_t> mol_base;
    std::string name;
};

// classify.rs:140-242: walk the order group in index order; a molecule is the connected component of its atom; two molecules
// are of one type iff their (relative index, residue name, atom name) atoms and their relative bonds agree.
static std::vector<RawType> classify_molecules(const GorderSystem &s, std::vector<int32_t> order_atoms) {
    std::sort(order_atoms.begin(), order_atoms.end());
    order_atoms.erase(std::unique(order_atoms.begin(), order_atoms.end()), order_atoms.end());
    std::vector<char> visited(s.n_atoms, 0);
    std::vector<RawType> types;
    for (int a : order_atoms) {
        if (visited[a]) continue;
        Component c = component_of(s, a, visited);
        if (c.atoms.size() < 2) continue;   // an atom without bonds: no molecule bonds, nothing to analyse
        const int mn = c.atoms[0];
        std::vector<std::pair<int32_t, int32_t>> rel;
        for (int i : c.atoms) for (int j : s.bonded[i]) if (j > i) rel.emplace_back(i - mn, j - mn);
        std::sort(rel.begin(), rel.end());
        RawType *hit = nullptr;
        for (auto &t : types) {
            if (t.atoms_rel.size() != c.atoms.size() || t.bonds_rel != rel) continue;
            bool same = true;
            const int b0 = t.mol_base[0];
            for (size_t k = 0; same && k < c.atoms.size(); k++) {
                const int x = c.atoms[k], y = b0 + t.atoms_rel[k];
                same = (x - mn) == t.atoms_rel[k] && s.resname[x] == s.resname[y] && s.name[x] == s.name[y];
            }
            if (same) { hit = &t; break; }
        }
        if (hit) { hit->mol_base.push_back(mn); continue; }
        RawType t;
        for (int i : c.atoms) t.atoms_rel.push_back(i - mn);
        t.bonds_rel = rel;
        t.mol_base.push_back(mn);
        std::vector<std::string> residues;   // molecule name: residue names in order of appearance, joined by '-'
        for (int i : c.atoms) if (std::find(residues.begin(), residues.end(), s.resname[i]) == residues.end()) residues.push_back(s.resname[i]);
        for (size_t k = 0; k < residues.size(); k++) t.name += (k ? "-" : "") + residues[k];
        types.push_back(std::move(t));
    }
    return types;
}

// classify.rs:262-294 (solve_name_conflicts): names that occur once are left alone; namesakes get their running count
// appended, walking the types from the last to the first (A, B, C named POPC -> POPC1, POPC2, POPC3)
static void solve_name_conflicts(std::vector<GorderClassification::Type> &types) {
    std::vector<std::pair<std::string, int>> counts;
    for (auto &t : types) {
        bool hit = false;
        for (auto &c : counts) if (c.first == t.name) { c.second++; hit = true; break; }
        if (!hit) counts.emplace_back(t.name, 1);
    }
    counts.erase(std::remove_if(counts.begin(), counts.end(), [](const std::pair<std::string, int> &c) { return c.second <= 1; }), counts.end());
    for (auto it = types.rbegin(); it != types.rend(); ++it)
        for (auto &c : counts)
            if (c.first == it->name) { it->name += std::to_string(c.second); c.second--; break; }
}

struct Groups {
    std::vector<char> heads, methyls, normal_heads;
    bool has_heads = false, has_methyls = false, has_normal_heads = false;
};

static int fill_group(const GorderSystem &s, const int32_t *idx, int n, std::vector<char> &mask, bool &has) {
    has = idx != nullptr;
    mask.assign(s.n_atoms, 0);
    if (!idx) return GORDER_OK;
    for (int i = 0; i < n; i++) { if (idx[i] < 0 || idx[i] >= s.n_atoms) return GORDER_ERR_INVALID_ARGUMENT; mask[idx[i]] = 1; }
    return GORDER_OK;
}

// common.rs:345-375 for every molecule of the type (MoleculeLeafletClassification::insert / MoleculeMembraneNormal::insert);
// the engine addresses the head by its relative index, which therefore has to be the same in all molecules of the type
static int single_rel(const RawType &t, const std::vector<char> &mask, int32_t &rel) {
    rel = -1;
    for (size_t m = 0; m < t.mol_base.size(); m++) {
        int found = -1, n = 0;
        for (int r : t.atoms_rel) if (mask[t.mol_base[m] + r]) { if (!n) found = r; n++; }
        if (n == 0) { g_topology_error = "molecule starting with atom index '" + std::to_string(t.mol_base[m]) + "' contains no head group atom"; return GORDER_ERR_TOPOLOGY_NO_HEAD; }
        if (n > 1) { g_topology_error = "molecule starting with atom index '" + std::to_string(t.mol_base[m]) + "' contains multiple head group atoms"; return GORDER_ERR_TOPOLOGY_MULTIPLE_HEADS; }
        if (m == 0) rel = found;
        else if (found != rel) { g_topology_error = "head group atoms at different positions in molecules of one type"; return GORDER_ERR_INVALID_ARGUMENT; }
    }
    return GORDER_OK;
}
// leaflets.rs:743-775: at least one methyl per molecule, the same number in all molecules of a type
static int methyl_rels(const RawType &t, const std::vector<char> &mask, std::vector<int32_t> &out) {
    out.clear();
    for (size_t m = 0; m < t.mol_base.size(); m++) {
        std::vector<int32_t> mine;
        for (int r : t.atoms_rel) if (mask[t.mol_base[m] + r]) mine.push_back(r);
        if (mine.empty()) { g_topology_error = "molecule starting with atom index '" + std::to_string(t.mol_base[m]) + "' contains no methyl group atom"; return GORDER_ERR_TOPOLOGY_NO_METHYL; }
        if (m == 0) out = mine;
        else if (mine.size() != out.size()) {
            g_topology_error = "molecule starting with atom index '" + std::to_string(t.mol_base[m]) + "' contains a number of methyl group atoms ('" +
                               std::to_string(mine.size()) + "') not consistent with other molecules ('" + std::to_string(out.size()) + "')";
            return GORDER_ERR_TOPOLOGY_INCONSISTENT_METHYLS;
        } else if (mine != out) { g_topology_error = "methyl group atoms at different positions in molecules of one type"; return GORDER_ERR_INVALID_ARGUMENT; }
    }
    return GORDER_OK;
}

static int common_groups(const RawType &rt, const Groups &g, GorderClassification::Type &t) {
    int rc;
    if (g.has_heads && (rc = single_rel(rt, g.heads, t.head_rel))) return rc;
    if (g.has_methyls && (rc = methyl_rels(rt, g.methyls, t.methyl_rel))) return rc;
    if (g.has_normal_heads && (rc = single_rel(rt, g.normal_heads, t.normal_head_rel))) return rc;
    return GORDER_OK;
}

static std::string atom_label(const GorderSystem &s, int base, int rel) {
    return s.resname[base + rel] + " " + s.name[base + rel] + " (" + std::to_string(rel) + ")";
}

static void finish(GorderClassification &c) {
    solve_name_conflicts(c.types);
    for (auto &t : c.types) {
        GorderMolType m{};
        m.n_molecules = (int32_t)t.mol_base.size(); m.mol_base = t.mol_base.data();
        m.n_bond_types = (int32_t)(t.bond_rel.size() / 2); m.bond_rel = t.bond_rel.empty() ? nullptr : t.bond_rel.data();
        m.n_ua_atoms = (int32_t)t.ua_kind.size(); m.ua_kind = t.ua_kind.empty() ? nullptr : t.ua_kind.data(); m.ua_rel = t.ua_rel.empty() ? nullptr : t.ua_rel.data();
        m.head_rel = t.head_rel; m.n_methyls = (int32_t)t.methyl_rel.size(); m.methyl_rel = t.methyl_rel.empty() ? nullptr : t.methyl_rel.data();
        m.normal_head_rel = t.normal_head_rel;
        c.c_types.push_back(m);
    }
}

}  // namespace gtopo

extern "C" {

// AA (group1 = heavy atoms, group2 = hydrogens) and CG (group1 = group2 = beads): classify.rs:53-76, 355-420
int gorder_classify_bonds(const GorderSystem *s, const int32_t *group1, int32_t n1, const int32_t *group2, int32_t n2,
                          const int32_t *heads, int32_t n_heads, const int32_t *methyls, int32_t n_methyls,
                          const int32_t *normal_heads, int32_t n_normal_heads, GorderClassification **out) {
    if (!s || !out || n1 < 0 || n2 < 0 || (n1 > 0 && !group1) || (n2 > 0 && !group2)) return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    std::vector<char> g1(s->n_atoms, 0), g2(s->n_atoms, 0);
    for (int i = 0; i < n1; i++) { if (group1[i] < 0 || group1[i] >= s->n_atoms) return GORDER_ERR_INVALID_ARGUMENT; g1[group1[i]] = 1; }
    for (int i = 0; i < n2; i++) { if (group2[i] < 0 || group2[i] >= s->n_atoms) return GORDER_ERR_INVALID_ARGUMENT; g2[group2[i]] = 1; }
    gtopo::Groups g;
    int rc;
    if ((rc = gtopo::fill_group(*s, heads, n_heads, g.heads, g.has_heads))) return rc;
    if ((rc = gtopo::fill_group(*s, methyls, n_methyls, g.methyls, g.has_methyls))) return rc;
    if ((rc = gtopo::fill_group(*s, normal_heads, n_normal_heads, g.normal_heads, g.has_normal_heads))) return rc;
    auto raw = gtopo::classify_molecules(*s, std::vector<int32_t>(group1, group1 + n1));
    auto c = std::make_unique<GorderClassification>();
    bool empty_type = false;
    for (const auto &rt : raw) {
        const int base = rt.mol_base[0];
        GorderClassification::Type t;
        t.name = rt.name; t.mol_base = rt.mol_base; t.atoms_rel = rt.atoms_rel;
        for (const auto &b : rt.bonds_rel) {   // already in bond.rs:77-81 order: sorted by (lower, higher) relative index
            const int i = base + b.first, j = base + b.second;
            if ((g1[i] && g2[j]) || (g2[i] && g1[j])) {
                t.bond_rel.push_back(b.first); t.bond_rel.push_back(b.second);
                t.item_names.push_back(gtopo::atom_label(*s, base, b.first) + " - " + gtopo::atom_label(*s, base, b.second));
            }
        }
        if ((rc = gtopo::common_groups(rt, g, t))) return rc;
        if (t.bond_rel.empty()) empty_type = true;
        c->types.push_back(std::move(t));
    }
    // classify.rs:297-315: no molecules, or a molecule type without order bonds -> nothing is analysed (a warning, not an error)
    if (c->types.empty()) { c->warning = "No molecules suitable for analysis detected."; }
    else if (empty_type) { c->warning = "No bonds/atoms suitable for analysis detected."; c->types.clear(); }
    gtopo::finish(*c);
    *out = c.release();
    return GORDER_OK;
}

// UA: order group = saturated + unsaturated carbons; classify.rs:77-90, uaorder.rs:580-665
int gorder_classify_ua(const GorderSystem *s, const int32_t *saturated, int32_t n_sat, const int32_t *unsaturated, int32_t n_unsat,
                       const int32_t *ignore, int32_t n_ignore, const int32_t *heads, int32_t n_heads, const int32_t *methyls, int32_t n_methyls,
                       const int32_t *normal_heads, int32_t n_normal_heads, GorderClassification **out) {
    if (!s || !out || n_sat < 0 || n_unsat < 0 || n_ignore < 0 || (n_sat > 0 && !saturated) || (n_unsat > 0 && !unsaturated) || (n_ignore > 0 && !ignore))
        return GORDER_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (n_sat + n_unsat == 0) { g_topology_error = "no carbons for the calculation of united-atom order parameters were specified"; return GORDER_ERR_TOPOLOGY_NO_UA_CARBONS; }
    std::vector<char> sat(s->n_atoms, 0), unsat(s->n_atoms, 0), ign(s->n_atoms, 0);
    std::vector<int32_t> order;
    for (int i = 0; i < n_sat; i++) { if (saturated[i] < 0 || saturated[i] >= s->n_atoms) return GORDER_ERR_INVALID_ARGUMENT; sat[saturated[i]] = 1; order.push_back(saturated[i]); }
    for (int i = 0; i < n_unsat; i++) { if (unsaturated[i] < 0 || unsaturated[i] >= s->n_atoms) return GORDER_ERR_INVALID_ARGUMENT; unsat[unsaturated[i]] = 1; order.push_back(unsaturated[i]); }
    for (int i = 0; i < n_ignore; i++) { if (ignore[i] < 0 || ignore[i] >= s->n_atoms) return GORDER_ERR_INVALID_ARGUMENT; ign[ignore[i]] = 1; }
    for (int i = 0; i < s->n_atoms; i++) if (sat[i] && unsat[i]) { g_topology_error = "atom " + std::to_string(i) + " is part of both 'Saturated' and 'Unsaturated'"; return GORDER_ERR_INVALID_ARGUMENT; }
    gtopo::Groups g;
    int rc;
    if ((rc = gtopo::fill_group(*s, heads, n_heads, g.heads, g.has_heads))) return rc;
    if ((rc = gtopo::fill_group(*s, methyls, n_methyls, g.methyls, g.has_methyls))) return rc;
    if ((rc = gtopo::fill_group(*s, normal_heads, n_normal_heads, g.normal_heads, g.has_normal_heads))) return rc;
    auto raw = gtopo::classify_molecules(*s, order);
    auto c = std::make_unique<GorderClassification>();
    bool empty_type = false;
    for (const auto &rt : raw) {
        const int base = rt.mol_base[0];
        GorderClassification::Type t;
        t.name = rt.name; t.mol_base = rt.mol_base; t.atoms_rel = rt.atoms_rel;
        for (int r : rt.atoms_rel) {   // carbon types sorted by relative index (topology/uatom.rs:39-41)
            const int a = base + r;
            if (!sat[a] && !unsat[a]) continue;
            std::vector<int32_t> bonded;
            for (int b : s->bonded[a]) if (!ign[b]) bonded.push_back(b);
            const int missing = bonded.size() >= 4 ? 0 : 4 - (int)bonded.size();
            int kind = -1;
            int32_t rel4[4] = {r, -1, -1, -1};
            if (missing == 0 || (!sat[a] && missing == 1)) continue;
            if (sat[a] && missing == 1) { kind = GORDER_UA_CH1_SAT; rel4[1] = bonded[0] - base; rel4[2] = bonded[1] - base; rel4[3] = bonded[2] - base; }
            else if (sat[a] && missing == 2) { kind = GORDER_UA_CH2; rel4[1] = bonded[0] - base; rel4[2] = bonded[1] - base; }
            else if (sat[a] && missing == 3) {
                const int h1 = bonded[0];
                int h2 = -1;
                for (int x : s->bonded[h1]) if (x != a) { h2 = x; break; }
                if (h2 < 0) continue;   // an isolated chain of two carbons: ignored with a warning in the reference
                kind = GORDER_UA_CH3; rel4[1] = h1 - base; rel4[2] = h2 - base;
            } else if (!sat[a] && missing == 2) { kind = GORDER_UA_CH1_UNSAT; rel4[1] = bonded[0] - base; rel4[2] = bonded[1] - base; }
            else continue;               // four missing hydrogens / unsupported: ignored with a warning in the reference
            t.ua_kind.push_back(kind);
            for (int k = 0; k < 4; k++) t.ua_rel.push_back(rel4[k]);
            t.item_names.push_back(gtopo::atom_label(*s, base, r));
        }
        if ((rc = gtopo::common_groups(rt, g, t))) return rc;
        if (t.ua_kind.empty()) empty_type = true;
        c->types.push_back(std::move(t));
    }
    if (c->types.empty()) { c->warning = "No molecules suitable for analysis detected."; }
    else if (empty_type) { c->warning = "No bonds/atoms suitable for analysis detected."; c->types.clear(); }
    gtopo::finish(*c);
    *out = c.release();
    return GORDER_OK;
}

void gorder_classification_free(GorderClassification *c) { delete c; }
int32_t gorder_classification_n_types(const GorderClassification *c) { return c ? (int32_t)c->types.size() : -1; }
// ready for GorderSetup.moltypes; the arrays live as long as the classification
const GorderMolType *gorder_classification_moltypes(const GorderClassification *c) { return c && !c->c_types.empty() ? c->c_types.data() : nullptr; }
const char *gorder_classification_type_name(const GorderClassification *c, int32_t t) { return c && t >= 0 && t < (int32_t)c->types.size() ? c->types[t].name.c_str() : nullptr; }
// "POPC C22 (20) - POPC H2R (21)" for a bond type, "POPC C22 (20)" for a united-atom carbon
const char *gorder_classification_item_name(const GorderClassification *c, int32_t t, int32_t i) {
    if (!c || t < 0 || t >= (int32_t)c->types.size() || i < 0 || i >= (int32_t)c->types[t].item_names.size()) return nullptr;
    return c->types[t].item_names[i].c_str();
}
const char *gorder_classification_warning(const GorderClassification *c) { return c ? c->warning.c_str() : nullptr; }
int32_t gorder_classification_n_atoms_rel(const GorderClassification *c, int32_t t) { return c && t >= 0 && t < (int32_t)c->types.size() ? (int32_t)c->types[t].atoms_rel.size() : -1; }
const int32_t *gorder_classification_atoms_rel(const GorderClassification *c, int32_t t) { return c && t >= 0 && t < (int32_t)c->types.size() ? c->types[t].atoms_rel.data() : nullptr; }

}  // extern "C"

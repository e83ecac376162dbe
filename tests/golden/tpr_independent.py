import struct, sys
import numpy as np

# function types, GROMACS 2022 order (with VSITE1)
F = """BONDS G96BONDS MORSE CUBICBONDS CONNBONDS HARMONIC FENEBONDS TABBONDS TABBONDSNC RESTRBONDS ANGLES G96ANGLES RESTRANGLES
LINEAR_ANGLES CROSS_BOND_BONDS CROSS_BOND_ANGLES UREY_BRADLEY QUARTIC_ANGLES TABANGLES PDIHS RBDIHS RESTRDIHS CBTDIHS FOURDIHS IDIHS PIDIHS
TABDIHS CMAP GB12 GB13 GB14 GBPOL NPSOLVATION LJ14 COUL14 LJC14_Q LJC_PAIRS_NB LJ BHAM LJ_LR BHAM_LR DISPCORR COUL_SR COUL_LR RF_EXCL COUL_RECIP
LJ_RECIP DPD POLARIZATION WATER_POL THOLE_POL ANHARM_POL POSRES FBPOSRES DISRES DISRESVIOL ORIRES ORIRESDEV ANGRES ANGRESZ DIHRES DIHRESVIOL
CONSTR CONSTRNC SETTLE VSITE1 VSITE2 VSITE2FD VSITE3 VSITE3FD VSITE3FAD VSITE3OUT VSITE4FD VSITE4FDN VSITEN COM_PULL DENSITYFITTING EQM EPOT EKIN ETOT ECONSERVED
TEMP VTEMP PDISPCORR PRES DVDL_CONSTR DVDL DKDL DVDL_COUL DVDL_VDW DVDL_BONDED DVDL_RESTRAINT DVDL_TEMPERATURE""".split()
ADDED = {'VSITE1': 121, 'VSITE2FD': 118, 'DENSITYFITTING': 117, 'RESTRANGLES': 98, 'RESTRDIHS': 98, 'CBTDIHS': 98}

class R:
    def __init__(s, b): s.b = b; s.p = 0; s.mem = False
    def i32(s): v = struct.unpack_from('>i', s.b, s.p)[0]; s.p += 4; return v
    def i64(s): v = struct.unpack_from('>q', s.b, s.p)[0]; s.p += 8; return v
    def f32(s): v = struct.unpack_from('>f', s.b, s.p)[0]; s.p += 4; return v
    def f64(s): v = struct.unpack_from('>d', s.b, s.p)[0]; s.p += 8; return v
    def real(s): return s.f32()
    def reals(s, n): v = np.frombuffer(s.b, '>f4', n, s.p); s.p += 4 * n; return v
    def ints(s, n): v = np.frombuffer(s.b, '>i4', n, s.p); s.p += 4 * n; return v
    def uchar(s):
        if s.mem: v = s.b[s.p]; s.p += 1; return v
        return s.i32()
    def ushort(s):
        if s.mem: v = struct.unpack_from('>H', s.b, s.p)[0]; s.p += 2; return v
        return s.i32()
    def boolean(s):
        if s.mem: v = s.b[s.p]; s.p += 1; return v
        return s.i32()
    def string(s):
        if s.mem:
            n = s.i64(); v = s.b[s.p:s.p + n]; s.p += n; return v.decode()
        n = s.i32(); m = s.i32(); v = s.b[s.p:s.p + m]; s.p += (m + 3) // 4 * 4; return v.decode()

def iparams(r, ft, ver):
    n = F[ft]
    def rl(k): r.p += 4 * k
    if n in ('ANGLES', 'G96ANGLES', 'BONDS', 'G96BONDS', 'HARMONIC', 'IDIHS'): rl(4)
    elif n == 'RESTRANGLES': rl(2)
    elif n == 'LINEAR_ANGLES': rl(4)
    elif n == 'FENEBONDS': rl(2)
    elif n == 'RESTRBONDS': rl(8)
    elif n in ('TABBONDS', 'TABBONDSNC', 'TABANGLES', 'TABDIHS'): rl(3)
    elif n == 'CROSS_BOND_BONDS': rl(3)
    elif n == 'CROSS_BOND_ANGLES': rl(4)
    elif n == 'UREY_BRADLEY': rl(8)
    elif n == 'QUARTIC_ANGLES': rl(6)
    elif n == 'BHAM': rl(3)
    elif n == 'MORSE': rl(6)
    elif n == 'CUBICBONDS': rl(3)
    elif n == 'CONNBONDS': pass
    elif n == 'POLARIZATION': rl(1)
    elif n == 'ANHARM_POL': rl(3)
    elif n == 'WATER_POL': rl(6)
    elif n == 'THOLE_POL': rl(4)
    elif n == 'LJ': rl(2)
    elif n == 'LJ14': rl(4)
    elif n == 'LJC14_Q': rl(5)
    elif n == 'LJC_PAIRS_NB': rl(4)
    elif n in ('PDIHS', 'PIDIHS', 'ANGRES', 'ANGRESZ'): rl(5)
    elif n == 'RESTRDIHS': rl(2)
    elif n == 'DISRES': rl(6)
    elif n == 'ORIRES': rl(6)
    elif n == 'DIHRES': rl(6)
    elif n == 'POSRES': rl(12)
    elif n == 'FBPOSRES': rl(6)
    elif n == 'CBTDIHS': rl(6)
    elif n in ('RBDIHS', 'FOURDIHS'): rl(12)
    elif n in ('CONSTR', 'CONSTRNC', 'SETTLE'): rl(2)
    elif n == 'VSITE1': pass
    elif n in ('VSITE2', 'VSITE2FD'): rl(1)
    elif n in ('VSITE3', 'VSITE3FD', 'VSITE3FAD'): rl(2)
    elif n in ('VSITE3OUT', 'VSITE4FD', 'VSITE4FDN'): rl(3)
    elif n == 'VSITEN': rl(2)
    elif n in ('GB12', 'GB13', 'GB14'): rl(5)
    elif n == 'CMAP': rl(2)
    else: raise ValueError('iparams ' + n)

def parse(path, verbose=True):
    b = open(path, 'rb').read()
    r = R(b)
    vs = r.string(); prec = r.i32(); ver = r.i32(); gen = r.i32(); tag = r.string()
    natoms = r.i32(); ngtc = r.i32(); fep = r.i32(); lam = r.real()
    bIr, bTop, bX, bV, bF, bBox = [r.i32() for _ in range(6)]
    if ver >= 119: size = r.i64(); r.mem = True
    if verbose: print(path, vs, prec, ver, gen, tag, natoms, ngtc, 'hdr end', r.p)
    assert prec == 4
    box = r.reals(9) if bBox else None
    if bBox: r.reals(9); r.reals(9)
    if ngtc > 0: r.reals(ngtc)
    # file functype enumeration
    ftypes = [i for i, n in enumerate(F) if ADDED.get(n, 0) <= ver]
    # mtop
    nsym = r.i32(); sym = [r.string() for _ in range(nsym)]
    name = sym[r.i32()]
    atnr = r.i32(); ntypes = r.i32(); functype = r.ints(ntypes).copy()
    reppow = r.f64(); fudge = r.real()
    for i in range(ntypes): iparams(r, ftypes[functype[i]], ver)
    nmt = r.i32()
    mts = []
    for m in range(nmt):
        mname = sym[r.i32()]
        nr = r.i32(); nres = r.i32()
        atoms = []
        for a in range(nr):
            mass = r.real(); q = r.real(); r.real(); r.real(); r.ushort(); r.ushort(); ptype = r.i32(); resind = r.i32(); atomnumber = r.i32()
            atoms.append((mass, q, resind, atomnumber))
        an = [sym[x] for x in r.ints(nr)]; r.ints(nr); r.ints(nr)
        res = []
        for j in range(nres):
            rn = sym[r.i32()]; rnr = r.i32(); ic = r.uchar(); res.append((rn, rnr))
        il = {}
        for ft in ftypes:
            n = r.i32(); ia = r.ints(n)
            if n: il[F[ft]] = ia.copy()
        nb = r.i32(); r.ints(nb + 1)          # cgs
        ne = r.i32(); nra = r.i32(); r.ints(ne + 1); r.ints(nra)
        mts.append((mname, atoms, an, res, il))
        if verbose: print('  moltype', mname, nr, nres, {k: len(v) for k, v in il.items()})
    nmb = r.i32(); mbs = []
    for i in range(nmb):
        t = r.i32(); nmol = r.i32(); nat = r.i32()
        na = r.i32(); r.reals(3 * na); nb_ = r.i32(); r.reals(3 * nb_)
        mbs.append((t, nmol, nat))
    nat_tot = r.i32()
    if verbose: print('  molblocks', mbs, nat_tot)
    assert nat_tot == natoms
    if ver >= 103:
        inter = r.boolean()
        if inter:
            for ft in ftypes:
                n = r.i32(); r.ints(n)
    # atomtypes
    nat_t = r.i32()
    if ver < 115: r.reals(3 * nat_t)
    r.ints(nat_t)
    if ver < 115: r.reals(2 * nat_t)
    # cmap
    ngrid = r.i32(); gs = r.i32(); r.reals(ngrid * gs * gs * 4)
    # groups
    ngt = 10
    for g in range(ngt):
        n = r.i32(); r.ints(n)
    ngn = r.i32(); r.ints(ngn)
    for g in range(ngt):
        n = r.i32()
        if n:
            if r.mem: r.p += n
            else: r.p += 4 * n
    if ver >= 120:
        n = r.i64(); r.ints(n)
    x = r.reals(3 * natoms).reshape(-1, 3) if bX else None
    if verbose: print('  x at', r.p - 12 * natoms, x[:2])
    return dict(box=box, sym=sym, mts=mts, mbs=mbs, x=x, natoms=natoms)

if __name__ == '__main__':
    for p in sys.argv[1:]:
        parse(p)

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, refuses to run without a device (no CPU fallback), and the ctypes mirror matches the header."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gorder_b200 import abi, synthetic
from gorder_b200._lib import SYMBOLS, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gorder_b200.h")


def test_library_exports_every_declared_symbol():
    decl = set(re.findall(r"\b(gorder_(?:gpu|xtc|results|comm|system|classify|classification|topology|ndx|leaflets)_[a-z_]+)\s*\(", open(HEADER).read()))
    assert decl == set(SYMBOLS), decl ^ set(SYMBOLS)
    L = lib()
    for s in SYMBOLS:
        assert hasattr(L, s), s
    assert b"sm_100a" in L.gorder_gpu_version()


def test_enum_values_match_header():
    src = open(HEADER).read()
    for name, val in re.findall(r"GORDER_([A-Z0-9_]+)\s*=\s*(\d+)", src):
        py = {"ACC_UPPER": "ACC_UPPER", "ACC_LOWER": "ACC_LOWER"}.get(name, name)
        if hasattr(abi, py):
            assert getattr(abi, py) == int(val), name
    assert abi.ABI_VERSION == int(re.search(r"#define GORDER_ABI_VERSION (\d+)", src).group(1))


def test_struct_layout_matches_c(tmp_path):
    """Compile a tiny C program against the header and compare sizeof / offsetof with ctypes."""
    import subprocess
    fields = [("GorderSetup", "n_moltypes"), ("GorderSetup", "geom_ref_point"), ("GorderSetup", "map_bin"), ("GorderSetup", "max_batch_frames"),
              ("GorderMolType", "manual_normals"), ("GorderResults", "normals")]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "gorder_b200.h"\nint main(){\n'
    prog += 'printf("%zu %zu %zu\\n", sizeof(GorderSetup), sizeof(GorderMolType), sizeof(GorderResults));\n'
    for st, f in fields:
        prog += f'printf("%zu\\n", offsetof({st}, {f}));\n'
    prog += "return 0;}\n"
    src = tmp_path / "layout.c"
    src.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out[:3]]
    assert sizes == [C.sizeof(abi.CGorderSetup), C.sizeof(abi.CGorderMolType), C.sizeof(abi.CGorderResults)]
    cls = {"GorderSetup": abi.CGorderSetup, "GorderMolType": abi.CGorderMolType, "GorderResults": abi.CGorderResults}
    for (st, f), off in zip(fields, out[3:]):
        assert getattr(cls[st], f).offset == int(off), (st, f)


def test_no_cpu_fallback():
    """Without a CUDA device gorder_gpu_create must fail with GORDER_ERR_NO_DEVICE, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gorder_b200 import SystemTopology
    s = synthetic.s_cg(64)
    with pytest.raises(abi.GorderError) as e:
        SystemTopology(s.setup)
    assert e.value.code == abi.ERR_NO_DEVICE


def test_invalid_arguments_are_rejected():
    L = lib()
    h = C.c_void_p()
    assert L.gorder_gpu_create(None, C.byref(h)) == abi.ERR_INVALID_ARGUMENT
    s = synthetic.s_cg(8).setup.to_c()
    s.abi_version = 99
    assert L.gorder_gpu_create(C.byref(s), C.byref(h)) == abi.ERR_INVALID_ARGUMENT
    assert L.gorder_gpu_sync(None) == abi.ERR_INVALID_ARGUMENT


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gorder_b200/ or include/ may reference it."""
    for base in ("gorder_b200", "include"):
        for dp, _dn, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "gorder_oracle_" not in txt, os.path.join(dp, fn)


def test_synthetic_frames_are_reproducible():
    s = synthetic.s_cg(100)
    a, ba = s.frame(7)
    b, bb = s.frame(7)
    assert np.array_equal(a, b) and np.array_equal(ba, bb)
    c, _ = s.frame(8)
    assert not np.array_equal(a, c)
    assert s.setup.samples_per_frame() == 1100


def test_create_rejects_values_outside_their_range():
    """Checked before any device is touched (so also on this box): an axis or a mode outside its enumeration must not reach
    a kernel, where it would index past a 3-vector."""
    import numpy as np
    import pytest
    from gorder_b200 import SystemTopology, abi
    mt = abi.MolType(name="LIP", mol_base=np.array([0, 2]), bond_rel=[(0, 1)], head_rel=0)
    bad = [dict(kind=7), dict(normal_axis=3), dict(normal_axis=-1), dict(leaflet_axis=5), dict(normal_mode=9), dict(geom_kind=4), dict(geom_ref_kind=-2),
           dict(geom_axis=3), dict(leaflet_freq_kind=2), dict(leaflet_mode=17), dict(map_enabled=True, map_plane=3),
           dict(normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=0.0), dict(leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=-1.0)]
    for kw in bad:
        setup = abi.EngineSetup(**{**dict(kind=abi.KIND_CG, n_atoms=4, moltypes=[mt]), **kw})
        with pytest.raises(abi.GorderError) as e:
            SystemTopology(setup)
        assert e.value.code == abi.ERR_INVALID_ARGUMENT, kw


def test_create_rejects_classifiers_without_heads_or_methyls():
    """A molecule type without a head (or, for the individual method, without methyls) would make the classifiers read one
    float before the type's planes; the reference cannot reach that state (its classification guarantees a head), the C ABI
    can.  Checked before any device is touched."""
    import numpy as np
    import pytest
    from gorder_b200 import SystemTopology, abi
    base = dict(kind=abi.KIND_CG, n_atoms=4, membrane=np.array([0, 2]))
    no_head = abi.MolType(name="LIP", mol_base=np.array([0, 2]), bond_rel=[(0, 1)], head_rel=-1)
    head = abi.MolType(name="LIP", mol_base=np.array([0, 2]), bond_rel=[(0, 1)], head_rel=0)
    neg_methyl = abi.MolType(name="LIP", mol_base=np.array([0, 2]), bond_rel=[(0, 1)], head_rel=0, methyl_rel=[-1])
    cases = [(no_head, dict(leaflet_mode=abi.LEAFLET_GLOBAL)), (no_head, dict(leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=1.0)),
             (no_head, dict(leaflet_mode=abi.LEAFLET_INDIVIDUAL)), (no_head, dict(leaflet_mode=abi.LEAFLET_SPHERICAL)),
             (head, dict(leaflet_mode=abi.LEAFLET_INDIVIDUAL)), (neg_methyl, dict(leaflet_mode=abi.LEAFLET_INDIVIDUAL)),
             (head, dict(leaflet_mode=abi.LEAFLET_MANUAL))]
    for mt, kw in cases:
        with pytest.raises(abi.GorderError) as e:
            SystemTopology(abi.EngineSetup(moltypes=[mt], **base, **kw))
        assert e.value.code == abi.ERR_INVALID_ARGUMENT, kw


def test_multi_gpu_entry_points_fail_cleanly_without_devices():
    """gorder_gpu_reduce / gorder_comm_* validate their arguments before touching a device or NCCL."""
    import ctypes as C
    from gorder_b200 import abi
    L = lib()
    assert L.gorder_gpu_reduce(None, 2, 0) == abi.ERR_INVALID_ARGUMENT
    arr = (C.c_void_p * 2)(None, None)
    assert L.gorder_gpu_reduce(arr, 2, 0) == abi.ERR_INVALID_ARGUMENT
    assert L.gorder_gpu_reduce(arr, 2, 5) == abi.ERR_INVALID_ARGUMENT
    assert L.gorder_comm_unique_id(None) == abi.ERR_INVALID_ARGUMENT
    out = C.c_void_p()
    assert L.gorder_comm_create(None, 2, 0, 0, C.byref(out)) == abi.ERR_INVALID_ARGUMENT
    buf = C.create_string_buffer(128)
    assert L.gorder_comm_create(buf, 2, 3, 0, C.byref(out)) == abi.ERR_INVALID_ARGUMENT
    assert L.gorder_gpu_reduce_comm(None, None, 0) == abi.ERR_INVALID_ARGUMENT
    L.gorder_comm_destroy(None)


def _header_struct(name):
    """[(field, c type, is pointer, array length)] of a struct of include/gorder_b200.h, in declaration order."""
    hdr = open(HEADER).read()
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        const, typ, rest = re.match(r"(const )?(\w+) (.*)$", decl).groups()
        for var in rest.split(","):
            var = var.strip()
            ptr = var.startswith("*")
            var = var.lstrip("*").strip()
            arr = None
            m = re.match(r"(\w+)\[(\d+)\]", var)
            if m:
                var, arr = m.group(1), int(m.group(2))
            out.append((var, typ, bool(const), ptr, arr))
    return out


def test_rust_shim_structs_and_functions_match_the_header():
    """integration/gpu.rs (the module a maintainer adds to the gorder crate) cannot be compiled here (no Rust toolchain): its
    #[repr(C)] structs must at least list the header's fields in order with the matching types, and its extern block must
    declare functions the library exports, with the header's argument counts."""
    rs = open(os.path.join(ROOT, "integration", "gpu.rs")).read()
    rust_type = {"int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8", "float": "f32", "GorderMolType": "GorderMolType"}
    for name in ("GorderMolType", "GorderSetup", "GorderResults"):
        body = re.search(r"#\[repr\(C\)\]\s*pub struct " + name + r" \{(.*?)\n\}", rs, re.S).group(1)
        got = [(m.group(1), m.group(2).strip()) for m in re.finditer(r"pub (\w+): ([^,]+),", body)]
        want = []
        for var, typ, const, ptr, arr in _header_struct(name):
            t = rust_type[typ]
            if ptr:
                t = ("*const " if const else "*mut ") + t
            if arr:
                t = f"[{t}; {arr}]"
            want.append((var, t))
        assert got == want, (name, [x for x in zip(got, want) if x[0] != x[1]][:3])
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(gorder_\w+)\s*\(([^)]*)\)\s*;", hdr)}
    ext = re.search(r'extern "C" \{(.*?)\n\}', rs, re.S).group(1)
    L = lib()
    seen = 0
    for m in re.finditer(r"pub fn (gorder_\w+)\(([^)]*)\)", ext):
        fn, args = m.group(1), m.group(2)
        assert fn in decl and hasattr(L, fn), fn
        n_c = 0 if decl[fn].strip() in ("", "void") else decl[fn].count(",") + 1
        n_rs = 0 if not args.strip() else args.count(",") + 1
        assert n_c == n_rs, (fn, decl[fn], args)
        seen += 1
    assert seen >= 15
    for name, val in re.findall(r"pub const (GORDER_[A-Z0-9_]+): i32 = (\d+);", rs):
        m = re.search(r"\b" + name + r"\s*=\s*(\d+)", hdr) or re.search(r"#define " + name + r" (\d+)", hdr)
        assert m and int(m.group(1)) == int(val), name


def test_every_struct_field_offset_matches_c(tmp_path):
    """sizeof / offsetof of EVERY field of GorderSetup / GorderMolType / GorderResults / GorderRaw as gcc lays them out ==
    the ctypes mirror the tests and the bench drive the library with."""
    import subprocess
    structs = {"GorderSetup": abi.CGorderSetup, "GorderMolType": abi.CGorderMolType, "GorderResults": abi.CGorderResults, "GorderRaw": abi.CGorderRaw}
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "gorder_b200.h"\nint main(){\n'
    order = []
    for st in structs:
        prog += f'printf("%zu\\n", sizeof({st}));\n'
        order.append((st, None))
        for var, *_ in _header_struct(st):
            prog += f'printf("%zu\\n", offsetof({st}, {var}));\n'
            order.append((st, var))
    prog += "return 0;}\n"
    src = tmp_path / "layout_all.c"
    src.write_text(prog)
    exe = tmp_path / "layout_all"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert len(out) == len(order)
    for (st, var), val in zip(order, out):
        if var is None:
            assert C.sizeof(structs[st]) == val, st
        else:
            assert getattr(structs[st], var).offset == val, (st, var)

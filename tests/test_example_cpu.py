"""examples/cg_order.cpp: a C++ host on top of the C ABI builds against include/gorder_b200.h and, on a box without a GPU,
fails loudly (GORDER_ERR_NO_DEVICE) instead of falling back to anything."""
import os
import subprocess

import numpy as np
import pytest

from gorder_b200 import SystemTopology, abi, results, synthetic
from gorder_b200.xtc import write_xtc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_builds_and_needs_a_gpu(tmp_path):
    exe = str(tmp_path / "cg_order")
    lib_dir = os.path.join(ROOT, "gorder_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cg_order.cpp"),
                    "-L", lib_dir, "-lgorder_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True, capture_output=True, text=True)
    s = synthetic.s_cg(64, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
    xyz, box, idx = s.frames(0, 10)
    path = str(tmp_path / "t.xtc")
    write_xtc(path, xyz, box)
    r = subprocess.run([exe, path, "5"], capture_output=True, text=True, timeout=300)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if not have_gpu:
        assert r.returncode == 1 and f"code {abi.ERR_NO_DEVICE}" in r.stderr, (r.returncode, r.stderr)
        return
    # with a GPU: the table equals the Python mirror's on the decoded frames
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.splitlines() if not ln.startswith("#")]
    from gorder_b200.xtc import XtcFile
    with XtcFile(path) as x:
        dec, box9, _, _ = x.read()
    eng = SystemTopology(s.setup)
    eng.analyze_frames(dec, box9[:, [0, 4, 8]], idx)
    res = results.convert(eng.finish(), s.setup, n_blocks=5, native=True)
    eng.close()
    mol = next(iter(res.molecules.values()))
    got = np.array([[float(t) for t in row[-9:] if t != "+-"] for row in rows[:11]], np.float64).reshape(11, 3, 2)
    for b, item in enumerate(mol.items):
        for k, key in enumerate(("total", "upper", "lower")):
            o = getattr(item.order, key)
            assert abs(got[b, k, 0] - o.value) < 5.1e-5 and abs(got[b, k, 1] - o.error) < 5.1e-5


def test_tpr_host_builds_and_needs_a_gpu(tmp_path):
    """examples/tpr_order.cpp: run file + trajectory -> order parameters with no Rust and no Python in the loop.  On a GPU its
    table is the reference's cg_order_asymmetric_errors.yaml (tests_cg.rs:2246-2309)."""
    import golden_cases as gc
    exe = str(tmp_path / "tpr_order")
    lib_dir = os.path.join(ROOT, "gorder_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "tpr_order.cpp"),
                    "-L", lib_dir, "-lgorder_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True, capture_output=True, text=True)
    # the trajectory of the fixture holds the membrane beads only: put them back among the 9170 atoms of the run file
    setup, xyz, box, cases = gc.cg_asym()
    z = np.load(os.path.join(gc.GOLDEN, "cg_asym.npz"))
    full = np.zeros((xyz.shape[0], 9170, 3), np.float32)
    full[:, z["keep"], :] = xyz
    path = str(tmp_path / "t.xtc")
    write_xtc(path, full, box, precision=100.0)
    tpr = os.path.join(gc.GOLDEN, "tpr", "cg_asym.tpr")
    r = subprocess.run([exe, tpr, path, "POPE,POPG", "PO4", "5"], capture_output=True, text=True, timeout=300)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if not have_gpu:
        assert r.returncode == 1 and f"code {abi.ERR_NO_DEVICE}" in r.stderr, (r.returncode, r.stderr)
        return
    assert r.returncode == 0, r.stderr
    rows = [ln for ln in r.stdout.splitlines() if not ln.startswith("#")]
    assert len(rows) == 2 * 12 + 1 and "POPE NH3 (0) - POPE PO4 (1)" in rows[0]
    val = lambda ln: [float(t) for t in ln.split()[-9:] if t != "+-"]   # total, error, upper, error, lower, error
    got = val(rows[-1])                       # flatten_yaml order: system average, then per molecule its average and its bonds
    for m in range(2):
        got += val(rows[12 * m + 11])
        for b in range(11):
            got += val(rows[12 * m + b])
    exp = np.array(cases["errors"]["expected"], np.float64)
    np.testing.assert_allclose(np.array(got), exp, atol=2e-4, rtol=0, equal_nan=True)

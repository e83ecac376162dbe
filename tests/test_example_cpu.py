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

"""Small runs of every kernel family, for compute-sanitizer (memcheck / racecheck): `bash profiles/sanitize.sh <tool>`.
Each case goes through the C ABI like the tests do and is checked against the oracle, so a race that corrupts a sum shows up
twice.  Sizes are small: the sanitizer slows kernels by one to two orders of magnitude."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gorder_b200 import SystemTopology, abi, synthetic   # noqa: E402
from gorder_b200.xtc import XtcFile, write_xtc            # noqa: E402
from parity import assert_raw_parity, run_both            # noqa: E402


def case(name, s, n_frames, batches=2, **kw):
    xyz, box, idx = s.frames(0, n_frames)
    g, r = run_both(s.setup, xyz, box, idx, batches=batches, **kw)
    assert_raw_parity(g, r, s.setup, what=name)
    print("ok", name, int(g.count[:, 0].sum()), "samples", flush=True)


def main():
    G = abi.LEAFLET_GLOBAL
    # K1f + speculative Global leaflets + repair + fold (full tiles and a partial one), timewise rows, collected tables
    case("K1f spec, 64-thread tiles", synthetic.s_cg(700, leaflet_mode=G, timewise=True, collect_leaflets=True), 6, batches=3)
    os.environ["GORDER_MPT"] = "4"
    case("K1f spec, 1024-molecule tiles", synthetic.s_cg(2500, leaflet_mode=G, timewise=True), 4)
    os.environ["GORDER_MPT"] = "2"
    case("K1f table leaflets every 3", synthetic.s_cg(1300, leaflet_mode=G, leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=3), 7)
    del os.environ["GORDER_MPT"]
    os.environ["GORDER_NO_SPEC"] = "1"
    case("centre pre-pass (center_axis_kernel) + inline leaflets", synthetic.s_cg(900, leaflet_mode=G), 4)
    del os.environ["GORDER_NO_SPEC"]
    # generic bond kernel: geometry around a group centre, order maps, no PBC
    s = synthetic.s_aa(40, n_water=50, leaflet_mode=G, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=(0.5, 0.5), map_span_x=(0.0, 4.0),
                       map_span_y=(0.0, 4.0), geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_SELECTION, geom_ref=np.arange(0, 134 * 5),
                       geom_dims=(1.5, float("-inf"), float("inf")), geom_axis=abi.AXIS_Z)
    case("K1 maps + cylinder around a selection", s, 3)
    case("K1 no PBC, Individual leaflets", synthetic.s_cg(300, leaflet_mode=abi.LEAFLET_INDIVIDUAL, handle_pbc=False), 3)
    # local leaflets: brute force and 2-D cell list
    case("Local leaflets, brute force", synthetic.s_cg(200, leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=2.0), 2)
    os.environ["GORDER_LCELL_MIN_ATOMS"] = "64"
    case("Local leaflets, cell list", synthetic.s_cg(400, leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=2.0), 2)
    del os.environ["GORDER_LCELL_MIN_ATOMS"]
    # dynamic normals: brute force and cell list; collected normals
    case("dynamic normals, brute force", synthetic.s_cg(300, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0, collect_normals=True), 2)
    os.environ["GORDER_CELL_MIN_HEADS"] = "64"
    case("dynamic normals, cell list + vesicle + spherical clustering", synthetic.s_ves(1200, timewise=True, collect_leaflets=True), 3)
    del os.environ["GORDER_CELL_MIN_HEADS"]
    # united atom: K2f (persistent, bulk copies + mbarriers) and the exact kernel with a geometry selection
    case("K2f", synthetic.s_ua(300, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, with_ch1_sat=True), 5, batches=2)
    case("K2 exact + cuboid", synthetic.s_ua(60, geom_kind=abi.GEOM_CUBOID, geom_ref_kind=abi.GEOMREF_BOX_CENTER,
                                             geom_dims=(-1.5, 1.5, -1.5, 1.5, float("-inf"), float("inf"))), 2)
    # trajectory feed: host decode and device decode
    s = synthetic.s_cg(600, leaflet_mode=G, timewise=True)
    xyz, box, idx = s.frames(0, 5)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "t.xtc")
        write_xtc(path, xyz, box)
        with XtcFile(path) as x:
            a = SystemTopology(s.setup)
            a.run_xtc(x, batch_frames=2, n_threads=2)
            ra = a.finish()
            a.close()
            b = SystemTopology(s.setup)
            b.run_xtc_device(x, batch_frames=2, n_threads=2)
            rb = b.finish()
            b.close()
    np.testing.assert_array_equal(ra.tw_sum, rb.tw_sum)
    print("ok xtc host decode == device decode", flush=True)


if __name__ == "__main__":
    main()

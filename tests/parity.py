"""Shared helpers of the parity tests: run the CUDA engine and the oracle on the same inputs and
compare raw accumulators.

Bars (north_star / SURVEY.md §8c): sample counts, leaflet tables, molecule counts bit-exact; order
parameters (sum / count), order-map values and block errors within 1e-5 absolute.
"""
from __future__ import annotations

import numpy as np

from gorder_b200 import abi

TOL = 1e-5   # absolute tolerance on order parameters stated by north_star


def mean_order(sum_, count):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(count > 0, sum_.astype(np.float64) / np.maximum(count, 1).astype(np.float64) / 1e6, np.nan)


def assert_raw_parity(gpu: abi.RawResults, ref: abi.RawResults, setup: abi.EngineSetup, tol: float = TOL, what: str = ""):
    assert gpu.n_slots == ref.n_slots and gpu.n_frames == ref.n_frames, what
    # integer work: bit-exact
    np.testing.assert_array_equal(gpu.count, ref.count, err_msg=f"{what}: sample counts differ")
    # fixed-point sums: every sample is within 1 unit of 1e-6 -> means within tol
    mg, mr = mean_order(gpu.sum, gpu.count), mean_order(ref.sum, ref.count)
    np.testing.assert_allclose(mg, mr, atol=tol, rtol=0, equal_nan=True, err_msg=f"{what}: order parameters differ")
    # and far tighter in aggregate: |sum difference| <= 1 unit of 1e-6 per sample (static / manual normals: the
    # bond vectors are bit-identical); PCA normals carry their own f32 noise (different summation order): 3 units
    per_sample = 3 if setup.normal_mode == abi.NORMAL_DYNAMIC else 1
    assert np.all(np.abs(gpu.sum - ref.sum) <= per_sample * ref.count.astype(np.int64)), f"{what}: a sample moved by more than {per_sample}e-6"
    if setup.timewise:
        np.testing.assert_array_equal(gpu.tw_frame_index, ref.tw_frame_index)
        np.testing.assert_array_equal(gpu.tw_count, ref.tw_count, err_msg=f"{what}: per-frame counts differ")
        np.testing.assert_allclose(mean_order(gpu.tw_sum, gpu.tw_count), mean_order(ref.tw_sum, ref.tw_count), atol=tol, rtol=0,
                                   equal_nan=True, err_msg=f"{what}: per-frame order differs")
    if setup.map_enabled:
        assert gpu.map_shape == ref.map_shape
        # a midpoint may sit within an ulp of a bin edge: allow a handful of samples to move bins
        moved = np.abs(gpu.map_count.astype(np.int64) - ref.map_count.astype(np.int64)).sum()
        assert moved <= max(4, 2e-5 * ref.map_count.sum()), f"{what}: {moved} map samples moved bins"
        np.testing.assert_array_equal(gpu.map_count.sum(axis=(2, 3)), ref.map_count.sum(axis=(2, 3)))
        same = gpu.map_count == ref.map_count
        mg, mr = mean_order(gpu.map_sum, gpu.map_count), mean_order(ref.map_sum, ref.map_count)
        np.testing.assert_allclose(mg[same], mr[same], atol=tol, rtol=0, equal_nan=True, err_msg=f"{what}: map values differ")
    if setup.collect_leaflets:
        np.testing.assert_array_equal(gpu.leaflet_frame_index, ref.leaflet_frame_index)
        np.testing.assert_array_equal(gpu.leaflets, ref.leaflets, err_msg=f"{what}: leaflet assignment differs")
    if setup.collect_normals and setup.normal_mode == abi.NORMAL_DYNAMIC:
        g, r = gpu.normals, ref.normals
        assert np.array_equal(np.isnan(g), np.isnan(r)), f"{what}: NaN pattern of normals differs"
        ok = ~np.isnan(r[..., 0])
        # SIGNED: the reference exports the last row of V^T of nalgebra's SVD as it comes (normal.rs:443-457) and compares the
        # components signed (tests/common/mod.rs:84-87); oracle and device both reproduce that sign
        # (a sign decision of the algorithm -- e.g. the sign of the cloud's x-y covariance -- that sits within rounding of zero
        # comes out either way, in the reference too: at most one normal in a thousand may differ in sign)
        dots = np.sum(g[ok] * r[ok], axis=-1)
        assert np.all(np.abs(dots) > 1 - 1e-4), f"{what}: normals differ (min |cos| {np.abs(dots).min()})"
        assert (dots < 0).sum() <= max(1, 1e-3 * dots.size), f"{what}: {int((dots < 0).sum())} of {dots.size} normals with the opposite sign"


def run_both(setup: abi.EngineSetup, xyz, box, frame_index=None, batches=1, oracle_threads=2, native=False):
    """Feed the same frames to the CUDA engine (through the C ABI) and to the oracle."""
    from gorder_b200 import SystemTopology
    from oracle import oracle as orc

    xyz = np.asarray(xyz, np.float32)
    n = xyz.shape[0]
    if frame_index is None:
        frame_index = np.arange(n, dtype=np.int64) * setup.step
    eng = SystemTopology(setup)
    ref = orc.Oracle(setup, n_threads=oracle_threads)
    try:
        edges = np.linspace(0, n, batches + 1).astype(int)
        for a, b in zip(edges[:-1], edges[1:]):
            if b == a:
                continue
            if native:
                eng.analyze_frames_native(eng.to_native(xyz[a:b]), box[a:b], frame_index[a:b])
            else:
                eng.analyze_frames(xyz[a:b], box[a:b], frame_index[a:b])
            ref.analyze_frames(xyz[a:b], box[a:b], frame_index[a:b])
        return eng.finish(), ref.finish()
    finally:
        eng.close()
        ref.close()

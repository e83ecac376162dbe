"""Results conversion through the C ABI (gorder_results_*, csrc/gorder_results.inl): the step after the hot path
(converter.rs:52-559, order.rs:97-107, timewise.rs:191-274).  Host only -- runs without a GPU.

Checked against the reference's known answers, against the oracle's restatement and against the numpy converter
(gorder_b200/results.py) on random accumulators; the reference's YAML / XVG fixtures go through it in
tests/test_oracle_pins.py and tests/test_gpu_golden.py (golden_cases.assert_matches_yaml).
"""
import ctypes as C

import numpy as np
import pytest

from gorder_b200 import _lib, abi, results, synthetic
from oracle import oracle


def _raw(n_slots, n_frames, seed=0, empty_frames=(), scale=800_000):
    rng = np.random.default_rng(seed)
    tw_count = rng.integers(0, 40, size=(n_frames, n_slots, 3)).astype(np.uint64)
    tw_count[:, :, 0] = tw_count[:, :, 1] + tw_count[:, :, 2]
    for f in empty_frames:
        tw_count[f] = 0
    per = rng.integers(-scale // 2, scale, size=(n_frames, n_slots, 3))
    tw_sum = (per * tw_count.astype(np.int64)).astype(np.int64)
    return abi.RawResults(n_slots=n_slots, n_frames=n_frames, sum=tw_sum.sum(axis=0), count=tw_count.sum(axis=0, dtype=np.uint64),
                          tw_sum=tw_sum, tw_count=tw_count)


def _order(raw, slots, n_blocks=0, min_samples=1, sign=1.0, want_error=True):
    L = _lib.lib()
    r = abi.CGorderRaw()
    keep = [np.ascontiguousarray(raw.sum), np.ascontiguousarray(raw.count)]
    r.n_slots, r.n_frames, r.sum, r.count = raw.n_slots, raw.n_frames if raw.tw_sum is not None else 0, keep[0].ctypes.data, keep[1].ctypes.data
    if raw.tw_sum is not None:
        keep += [np.ascontiguousarray(raw.tw_sum), np.ascontiguousarray(raw.tw_count)]
        r.tw_sum, r.tw_count = keep[2].ctypes.data, keep[3].ctypes.data
    sl = np.ascontiguousarray(slots, np.int32)
    val, err = np.zeros(3, np.float32), np.zeros(3, np.float32)
    rc = L.gorder_results_order(C.byref(r), sl.ctypes.data, len(sl), n_blocks, min_samples, sign, val.ctypes.data, err.ctypes.data if want_error else None)
    return rc, val, err


def test_known_answers_of_the_reference():
    # converter.rs:787-794: AnalysisOrder::new(45.32, 56) -> 0.8092857
    raw = abi.RawResults(1, 0, np.array([[45320000, -7, 10]], np.int64), np.array([[56, 2, 3]], np.uint64))
    rc, val, _ = _order(raw, [0], min_samples=1, want_error=False)
    assert rc == abi.OK
    assert val[0] == pytest.approx(0.8092857, abs=1e-6)
    assert val[1] == np.float32(-3e-6)          # integer division truncates toward zero (order.rs:34-41)
    assert val[2] == np.float32(3e-6)
    rc, val, _ = _order(raw, [0], min_samples=5, sign=-1.0, want_error=False)
    assert val[0] == pytest.approx(-0.8092857, abs=1e-6) and np.isnan(val[1]) and np.isnan(val[2])
    # timewise.rs:594-616: estimate_error(5) == 0.0514468
    order = [10.0, 15.0, 18.0, 12.0, 14.0, 15.0, 16.0, 20.0, 21.0, 18.0, 9.0, 11.0, 13.0, 14.0, 19.0, 16.0, 17.0]
    samples = np.array([10, 12, 15, 11, 13, 11, 11, 17, 18, 15, 8, 10, 12, 13, 17, 14, 15], np.uint64)
    sums = np.array([oracle.order_value(x) for x in order], np.int64)
    tw_sum, tw_cnt = np.zeros((17, 1, 3), np.int64), np.zeros((17, 1, 3), np.uint64)
    tw_sum[:, 0, 0], tw_cnt[:, 0, 0] = sums, samples
    raw = abi.RawResults(1, 17, tw_sum.sum(axis=0), tw_cnt.sum(axis=0, dtype=np.uint64), tw_sum, tw_cnt)
    rc, val, err = _order(raw, [0], n_blocks=5)
    assert rc == abi.OK and err[0] == pytest.approx(0.0514468, abs=1.2e-7)
    assert err[0] == np.float32(oracle.estimate_error(sums, samples, 5))
    assert np.isnan(err[1]) and np.isnan(val[1])      # no samples in the leaflets
    # timewise.rs:624-647: prefix averages
    order = [10.0, 12.0, 15.0, 10.0, 9.0, 12.0, 98432.0]
    samples = np.array([13, 15, 20, 12, 11, 14, 98432], np.uint64)
    sums = np.array([oracle.order_value(x) for x in order], np.int64)
    tw_sum, tw_cnt = np.zeros((7, 1, 3), np.int64), np.zeros((7, 1, 3), np.uint64)
    tw_sum[:, 0, 0], tw_cnt[:, 0, 0] = sums, samples
    raw = abi.RawResults(1, 7, tw_sum.sum(axis=0), tw_cnt.sum(axis=0, dtype=np.uint64), tw_sum, tw_cnt)
    be = results._NativeBackend(raw, 1.0, False, None, 1, 1)
    got = be.convergence([0])["total"]
    np.testing.assert_allclose(got, [0.769230769, 0.785714286, 0.770833333, 0.783333333, 0.788732394, 0.8, 0.999827441], atol=1e-5)
    np.testing.assert_array_equal(got, oracle.prefix_average(sums, samples).astype(np.float32))


@pytest.mark.parametrize("n_blocks", [2, 5, 8, 13, 64])
def test_block_errors_equal_numpy_and_oracle(n_blocks):
    raw = _raw(7, 203, seed=n_blocks)
    for slots in ([3], [0, 1, 2], list(range(7))):
        rc, val, err = _order(raw, slots, n_blocks=n_blocks, sign=-1.0)
        assert rc == abi.OK
        ts, tc = raw.tw_sum[:, slots, :].sum(axis=1), raw.tw_count[:, slots, :].sum(axis=1, dtype=np.uint64)
        for k in range(3):
            assert err[k] == np.float32(results.estimate_error(ts[:, k], tc[:, k], n_blocks))
            assert err[k] == np.float32(oracle.estimate_error(ts[:, k], tc[:, k], n_blocks))
            assert val[k] == np.float32(-results.calc_order(ts[:, k].sum(), tc[:, k].sum()))


def test_empty_blocks_few_frames_and_min_samples():
    raw = _raw(3, 40, seed=3, empty_frames=range(8, 16))       # block 1 of 5 has no samples
    rc, val, err = _order(raw, [0, 1, 2], n_blocks=5)
    assert rc == abi.OK and np.all(np.isnan(err)) and not np.any(np.isnan(val))
    rc, val, err = _order(raw, [0, 1, 2], n_blocks=2)
    assert not np.any(np.isnan(err))
    rc, val, err = _order(raw, [1], n_blocks=2, min_samples=10**9)
    assert np.all(np.isnan(val)) and np.all(np.isnan(err))
    # fewer frames than blocks (block size 0 divides by zero in the reference, timewise.rs:201-207) and fewer than two
    # blocks (timewise.rs:196-199 panics) are refused, by the C converter and by the numpy one
    for nb in (5, 1):
        assert _order(_raw(2, 3, seed=5), [0], n_blocks=nb)[0] == abi.ERR_INVALID_ARGUMENT
        with pytest.raises(abi.GorderError):
            results.estimate_error(np.ones(3, np.int64), np.ones(3, np.uint64), nb)
    no_tw = abi.RawResults(2, 0, raw.sum[:2], raw.count[:2])
    rc, val, err = _order(no_tw, [0, 1], n_blocks=5)
    assert rc == abi.OK and np.all(np.isnan(err)) and not np.any(np.isnan(val))


def test_invalid_arguments():
    raw = _raw(2, 4)
    assert _order(raw, [2])[0] == abi.ERR_INVALID_ARGUMENT
    assert _order(raw, [-1])[0] == abi.ERR_INVALID_ARGUMENT
    L = _lib.lib()
    assert L.gorder_results_order(None, None, 0, 0, 1, 1.0, None, None) == abi.ERR_INVALID_ARGUMENT
    assert L.gorder_results_map(None, None, 4, 1, 1.0, None) == abi.ERR_INVALID_ARGUMENT
    assert L.gorder_results_map(None, None, 0, 1, 1.0, None) == abi.OK


def test_maps():
    rng = np.random.default_rng(11)
    cnt = rng.integers(0, 9, size=(5, 3, 17, 13)).astype(np.uint64)
    sm = rng.integers(-900_000, 900_000, size=cnt.shape) * cnt.astype(np.int64)
    raw = abi.RawResults(5, 0, np.zeros((5, 3), np.int64), np.zeros((5, 3), np.uint64), map_sum=sm, map_count=cnt, map_shape=(17, 13))
    for min_samples in (1, 4):
        py = results._PythonBackend(raw, -1.0, True, None, 1, min_samples)
        nat = results._NativeBackend(raw, -1.0, True, None, 1, min_samples)
        for s in range(5):
            a, b = py.slots_map([s]), nat.slots_map([s])
            np.testing.assert_array_equal(a, b)
            assert np.array_equal(np.isnan(b), cnt[s] < min_samples)
        np.testing.assert_array_equal(py.slots_map([1, 2, 4]), nat.slots_map([1, 2, 4]))   # the bonds of one atom, merged


@pytest.mark.parametrize("kind", ["cg", "aa", "ua"])
def test_whole_tree_equals_numpy_converter(kind):
    s = {"cg": lambda: synthetic.s_cg(40, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True),
         "aa": lambda: synthetic.s_aa(24, n_water=10, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True),
         "ua": lambda: synthetic.s_ua(24, timewise=True)}[kind]()
    xyz, box, idx = s.frames(0, 23)
    o = oracle.Oracle(s.setup, n_threads=4)
    o.analyze_frames(xyz, box, idx)
    raw = o.finish()
    o.close()
    for nb, ms in ((5, 1), (None, 1), (3, 30)):
        a = results.convert(raw, s.setup, n_blocks=nb, min_samples=ms)
        b = results.convert(raw, s.setup, n_blocks=nb, min_samples=ms, native=True)

        def same(x, y):
            for k in ("total", "upper", "lower"):
                p, q = getattr(x, k), getattr(y, k)
                assert (p is None) == (q is None)
                if p is None:
                    continue
                assert np.float32(p.value).tobytes() == np.float32(q.value).tobytes() or (np.isnan(p.value) and np.isnan(q.value))
                assert (p.error is None) == (q.error is None)
                if p.error is not None:
                    assert np.float32(p.error) == np.float32(q.error) or (np.isnan(p.error) and np.isnan(q.error))

        same(a.average, b.average)
        assert list(a.molecules) == list(b.molecules)
        for name, m in a.molecules.items():
            n = b.molecules[name]
            same(m.average, n.average)
            assert len(m.items) == len(n.items)
            for i, j in zip(m.items, n.items):
                assert i.label == j.label
                same(i.order, j.order)
                for p, q in zip(i.bonds, j.bonds):
                    same(p, q)
            for k in m.convergence:
                np.testing.assert_array_equal(m.convergence[k], n.convergence[k])

#!/usr/bin/env python
"""bench.py — throughput of the per-frame order-parameter engine on BASELINE.json's workload.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the oracle port on host cores)

Workload (BASELINE.json configs[1], SURVEY.md §8d "S-CG"): CGOrder on a synthetic Martini bilayer of
83 334 lipids = 1 000 008 beads, 11 bond types (916 674 S evaluations per frame), Global leaflet
assignment every frame, static z normal, PBC.  A "step" is one pass of the hot path over one batch of
``--frames`` frames.  Weak scaling: every rank analyses its own frames (contiguous frame ranges, no
data-path collective); the integer accumulators are combined with ONE NCCL sum-reduce at the end.

value = whole-job S evaluations / s with the frames already resident in HBM in the engine's plane
        layout (CUDA events on the engine's stream, max over ranks);
e2e   = the same metric through the C ABI entry point a host calls (gorder_gpu_submit) with pinned
        HOST buffers in the decoder's [atom][xyz] layout: H2D copy, device re-layout, analysis and a
        D2H read of the accumulators inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bond-frame samples/sec (S_CH evaluations/s)"
UNIT = "samples/s"
N_LIPIDS = 83334
BYTES_PER_ATOM = 12.0   # algorithmic: every coordinate of the frame is read once


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""

    REASONS = {0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x8: "hw_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                t = time.perf_counter()
                clk = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((t, clk, mask))
            except Exception:
                pass
            time.sleep(0.001)

    def summary(self, t0=None, t1=None):
        """Samples inside the timed window [t0, t1] (the sampler runs from before the warm-up on)."""
        win = [x for x in self.samples if t0 is None or (t0 <= x[0] <= t1)]
        if not win and self.samples:   # window shorter than one NVML query: take the sample closest to it
            win = [min(self.samples, key=lambda x: abs(x[0] - (t0 + t1) / 2))]
        reasons = set()
        for _t, _c, mask in win:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([x[1] for x in win])) if win else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(win), "window_ms": None if t0 is None else 1e3 * (t1 - t0)}


WORKLOADS = {
    # name: (description, default lipids, default frames per step)
    "cg": ("S-CG: CGOrder Martini bilayer, Global leaflets every frame (BASELINE configs[1])", N_LIPIDS, 256),
    "aa": ("S-AA-small: AAOrder POPC-like, 64 C-H bond types, static z (BASELINE configs[0] shape)", 256, 2048),
    "ua": ("S-UA: UAOrder Berger-like, 64 virtual C-H, error blocks (BASELINE configs[2])", 256, 2048),
    "aa_maps": ("S-AA-large: AAOrder 4096 lipids, Global leaflets, XY order maps 0.1 nm, cylinder r=8 nm (BASELINE configs[3])", 4096, 128),
    "cg_dyn": ("S-DYN: CGOrder flat bilayer, dynamic PCA normals r=2 nm, Global leaflets (BASELINE configs[4] kernel mix)", 100000, 8),
}


def make_system(args):
    from gorder_b200 import abi, synthetic
    w, n, F = args.workload, args.lipids, args.frames
    if w == "cg":
        return synthetic.s_cg(n, leaflet_mode=abi.LEAFLET_GLOBAL, max_batch_frames=F)
    if w == "aa":
        return synthetic.s_aa(n, n_water=30000 * n // 256, max_batch_frames=F)
    if w == "ua":
        return synthetic.s_ua(n, timewise=True, max_batch_frames=F)
    if w == "aa_maps":
        s = synthetic.s_aa(n, n_water=0, leaflet_mode=abi.LEAFLET_GLOBAL, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=(0.1, 0.1),
                           geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_BOX_CENTER, geom_dims=(8.0, float("-inf"), float("inf")),
                           geom_axis=abi.AXIS_Z, max_batch_frames=F)
        s.setup.map_span_x = (0.0, float(s.box[0]))
        s.setup.map_span_y = (0.0, float(s.box[1]))
        return s
    if w == "cg_dyn":
        return synthetic.s_cg(n, leaflet_mode=abi.LEAFLET_GLOBAL, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0, max_batch_frames=F)
    raise SystemExit(f"unknown workload {w}")


def config(args, s, extra=None):
    st = s.setup
    c = {"workload": WORKLOADS[args.workload][0], "lipids": args.lipids, "atoms_per_frame": s.n_atoms,
         "order_slots": st.n_slots, "samples_per_frame": st.samples_per_frame(), "frames_per_step": args.frames,
         "leaflets": {0: "none", 1: "Global, every frame"}.get(st.leaflet_mode, str(st.leaflet_mode)),
         "normal": {0: "static z", 1: "dynamic PCA", 2: "manual"}[st.normal_mode], "pbc": bool(st.handle_pbc),
         "l2_policy": "inputs larger than L2 (one step streams frames_per_step frames; 126 MB L2)", "parallelism": f"frame-sharded x{args.gpus}"}
    if extra:
        c.update(extra)
    return c


def hot_kernel_name(st, lipids):
    """Name of the accumulation kernel gorder_gpu_profile brackets for this configuration (gorder_capi.cu dispatch)."""
    if st.kind == 2:
        return "ua_order_kernel"
    plain = st.handle_pbc and st.normal_mode == 0 and not st.map_enabled and st.geom_kind == 0
    return "bond_fast_kernel" if plain and lipids >= 1024 and not os.environ.get("GORDER_NO_FAST") else "bond_order_kernel"


def cpu_oracle_rate(s, xyz, box, threads, min_seconds):
    """Time the oracle port (all host threads) on the given frames, repeated until min_seconds."""
    from oracle import oracle as orc
    n = xyz.shape[0]
    done, t_total, res = 0, 0.0, None
    while t_total < min_seconds:
        o = orc.Oracle(s.setup, n_threads=threads)
        t0 = time.perf_counter()
        o.analyze_frames(xyz, box, np.arange(n, dtype=np.int64))
        t_total += time.perf_counter() - t0
        res = o.finish()
        o.close()
        done += n
        if t_total > 30:
            break
    return done * s.setup.samples_per_frame() / t_total, done, t_total, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s = make_system(args)
    threads = os.cpu_count() or 1
    nf = min(args.frames, args.ref_frames)
    xyz, box, _ = s.frames(0, nf)
    from oracle import oracle as orc
    for _ in range(min(args.warmup, 1)):
        o = orc.Oracle(s.setup, n_threads=threads)
        o.analyze_frames(xyz[:1], box[:1], np.arange(1, dtype=np.int64))
        o.close()
    t0 = time.perf_counter()
    for k in range(args.steps):
        o = orc.Oracle(s.setup, n_threads=threads)
        o.analyze_frames(xyz, box, np.arange(nf, dtype=np.int64))
        o.finish()
        o.close()
    dt = time.perf_counter() - t0
    value = args.steps * nf * s.setup.samples_per_frame() / dt
    sample = f"{nf} frames of the {args.workload} workload per step, oracle port (oracle/gorder_oracle.c) with {threads} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config(args, s, {"frames_per_step": nf}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    # libraries (NCCL, torchrun) may print to stdout: keep fd 1 for the ONE JSON line, send everything else to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    from gorder_b200 import SystemTopology

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gorder_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = make_system(args)
    s.setup.device = local
    F, K, W = args.frames, args.steps, args.warmup
    spf = s.setup.samples_per_frame()

    # ---- inputs: every rank generates its own frames (weak scaling) ---------------------------------
    first = rank * F
    xyz, box, _ = s.frames(first, F)
    eng = SystemTopology(s.setup)
    planes = eng.to_native(xyz)
    d_planes = torch.from_numpy(planes).cuda()
    d_box = torch.from_numpy(box).cuda()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    frame_bytes = eng.frame_floats * 4
    needed_atoms = int((eng.native_layout()[1] >= 0).sum())

    def step(k):
        base = (k * world + rank) * F   # disjoint, strictly increasing frame ranges per rank
        eng.analyze_frames_device(d_planes.data_ptr(), d_box.data_ptr(), F, frame_index=base + np.arange(F, dtype=np.int64), native=True)

    eng.reserve_frames((W + K + 6) * F)
    sampler = ClockSampler(local)
    sampler.start()
    for k in range(W):
        step(k)
    eng.sync()
    block_ptr, n_words = eng.accumulator_block()

    class _Block:   # zero-copy view of the engine's contiguous int64 accumulator block (sums, counts, maps)
        __cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (block_ptr, False), "version": 2}

    block = torch.as_tensor(_Block(), device=torch.device("cuda", local))
    assert block.data_ptr() == block_ptr and block.dtype == torch.int64
    if world > 1:
        # warm the communicator with the same collective the job ends with (channel setup is not part of a step)
        warm = torch.zeros(n_words, dtype=torch.int64, device="cuda")
        dist.reduce(warm, dst=0, op=dist.ReduceOp.SUM)
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.stats()["kernel_launches"]
    eng.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_win0 = time.perf_counter()
    with torch.cuda.stream(stream):   # the engine's main stream is torch's current stream: NCCL orders itself after it
        if world > 1:   # align the DEVICE timelines of the ranks (the host barrier above leaves ~1 ms of launch skew)
            dist.all_reduce(torch.zeros(1, device="cuda"))
        ev0.record()
        for k in range(W, W + K):
            step(k)
        eng.fence()   # the tail (repair + fold) of the last batch runs on a helper stream: the main stream waits for it
        if world > 1:   # the single collective of the job: sum the integer accumulators in place on rank 0 (no host sync)
            dist.reduce(block, dst=0, op=dist.ReduceOp.SUM)
        ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_win1 = time.perf_counter()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    hot_ms, hot_n = eng.profile_read()
    eng.profile(False)
    launches = eng.stats()["kernel_launches"] - l0
    spec_stats = eng.speculation_stats()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = K * F * spf * world / (ms * 1e-3)
    res = eng.finish()
    total_samples = int(res.count[:, 0].sum())
    # the same kernel timed WITHOUT the overlapped centre kernels of the next batch (explains the in-step number)
    iso_ms = iso_n = None
    if rank == 0 and s.setup.leaflet_mode == 1:
        os.environ["GORDER_NO_OVERLAP"] = "1"
        eng.profile(True)
        for k in range(W + K, W + K + 5):
            step(k)
        eng.sync()
        iso_ms, iso_n = eng.profile_read()
        eng.profile(False)
        del os.environ["GORDER_NO_OVERLAP"]

    # ---- end to end: pinned host AoS frames through gorder_gpu_submit, D2H of the sums every step -------
    eng2 = SystemTopology(s.setup)
    eng2.reserve_frames((K + 4) * F)
    pin = torch.from_numpy(xyz).pin_memory()
    pin_box = torch.from_numpy(box).pin_memory()
    hx, hb = pin.numpy(), pin_box.numpy()

    def e2e_step(k):
        base = (k * world + rank) * F
        eng2.analyze_frames(hx, hb, base + np.arange(F, dtype=np.int64))
        return eng2.finish()

    for k in range(min(W, 3)):
        e2e_step(k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(3, 3 + K):
        r2 = e2e_step(k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = K * F * spf * world / float(t.item())
    h2d = int(xyz.nbytes + box.nbytes)
    d2h = int(r2.sum.nbytes + r2.count.nbytes)
    eng2.close()

    # ---- end to end from a trajectory FILE: host XTC decode (all host threads) -> pinned plane batches -> engine -----
    e2e_xtc = None
    if rank == 0 and args.xtc_frames > 0:
        try:   # an optional leg: its failure must not take the headline line with it
            import tempfile
            from gorder_b200.xtc import XtcFile, write_xtc
            nx = min(args.xtc_frames, F)
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "bench.xtc")
                write_xtc(path, xyz[:nx], box[:nx])
                fbytes = os.path.getsize(path)
                with XtcFile(path) as xf:
                    threads_x = os.cpu_count() or 1
                    eng4 = SystemTopology(s.setup)
                    eng4.reserve_frames(2 * nx + 8)
                    eng4.run_xtc(xf, n_threads=threads_x, batch_frames=args.xtc_batch)   # warm-up: page mappings of the file, pinned buffers
                    eng4.sync()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    dec_s = eng4.run_xtc(xf, n_threads=threads_x, batch_frames=args.xtc_batch, frame_index0=nx)
                    rx = eng4.finish()
                    dt_x = time.perf_counter() - t0
                    eng4.close()
                    # the same file with the decode on the device: host threads only copy compressed bytes
                    eng5 = SystemTopology(s.setup)
                    eng5.reserve_frames(4 * nx + 8)
                    eng5.run_xtc_device(xf, n_threads=threads_x, batch_frames=args.xtc_dev_batch)   # warm-up: buffers, page mappings of the file
                    eng5.sync()
                    torch.cuda.synchronize()
                    reps = 3
                    t0 = time.perf_counter()
                    for r_ in range(reps):
                        moved = eng5.run_xtc_device(xf, n_threads=threads_x, batch_frames=args.xtc_dev_batch, frame_index0=(r_ + 1) * nx)
                    rd = eng5.finish()
                    dt_d = (time.perf_counter() - t0) / reps
                    eng5.close()
            e2e_xtc = {"value": nx * spf / dt_x, "unit": UNIT, "frames": nx, "file_bytes": fbytes, "bytes_per_atom": fbytes / nx / s.n_atoms,
                       "decode_threads": threads_x, "decode_thread_seconds": dec_s, "wall_seconds": dt_x,
                       "decode_atoms_per_s_per_thread": nx * s.n_atoms / max(dec_s, 1e-9),
                       "entry": "gorder_gpu_run_xtc (host XTC decode + H2D + analysis + D2H of the sums; rank 0)",
                       "samples_accumulated_incl_warmup": int(rx.count[:, 0].sum()),
                       "device_decode": {"value": nx * spf / dt_d, "unit": UNIT, "wall_seconds": dt_d, "batch_frames": args.xtc_dev_batch, "h2d_bytes": moved, "h2d_bytes_per_atom": moved / nx / s.n_atoms,
                                         "entry": "gorder_gpu_run_xtc_device (host copies + bookmarks the compressed frames; xtc_decode_kernel unpacks them on the GPU)"}}
        except Exception as exc:   # noqa: BLE001
            e2e_xtc = {"error": f"{type(exc).__name__}: {exc}"}

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes: every coordinate the path needs, once (12 B x used atoms; S-CG: all 1 000 008 beads)
        launch_bytes = K * F * needed_atoms * BYTES_PER_ATOM / max(hot_n, 1)   # a step may be several launches
        achieved = launch_bytes / (hot_ms / max(hot_n, 1) * 1e-3) / 1e9 if hot_n else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tr = json.load(f)
            if args.workload == "cg" and args.lipids == N_LIPIDS:   # the capture is of this workload only
                traffic = tr["dram_bytes_per_frame"] * K * F / max(hot_n, 1)
        except Exception:
            pass
        # ---- CPU baseline (oracle port) on a bounded sample + parity of the GPU sums against it ------
        threads = os.cpu_count() or 1
        nb = min(F, args.ref_frames)
        cpu_rate, cpu_frames, cpu_t, ref = cpu_oracle_rate(s, xyz[:nb], box[:nb], threads, args.cpu_seconds)
        eng3 = SystemTopology(s.setup)
        eng3.analyze_frames(xyz[:nb], box[:nb], np.arange(nb, dtype=np.int64))
        g = eng3.finish()
        eng3.close()
        parity = {"counts_equal": bool(np.array_equal(g.count, ref.count)),
                  "max_abs_dS": float(np.abs(g.sum / np.maximum(g.count, 1).astype(np.float64) - ref.sum / np.maximum(ref.count, 1).astype(np.float64)).max() / 1e6),
                  "frames": nb}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config(args, s, {"resident_layout": "engine planes (gorder_gpu_native_layout)", "native_frame_bytes": frame_bytes}),
            "clocks": sampler.summary(t_win0, t_win1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "entry": "gorder_gpu_submit (pinned host [atom][xyz] frames) + gorder_gpu_finish", "timer": "host wall clock between device syncs"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "kernel": hot_kernel_name(s.setup, args.lipids), "launches_timed": hot_n, "avg_launch_ms": hot_ms / max(hot_n, 1),
                         "algorithmic_bytes_per_launch": launch_bytes, "peak_source": peak_src,
                         "step_share": (hot_ms / ms) if ms else None,
                         "note": ("in-step duration; speculative Global leaflets: no centre pre-pass, the kernel runs alone" if spec_stats["enabled"] else
                                  "in-step duration; with Global leaflets the centre kernels of the NEXT batch run concurrently on a second stream"),
                         "isolated": ({"avg_launch_ms": iso_ms / iso_n, "achieved": launch_bytes / (iso_ms / iso_n * 1e-3) / 1e9,
                                       "frac": launch_bytes / (iso_ms / iso_n * 1e-3) / 1e9 / peak} if iso_n else None)},
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{cpu_frames} frames of the same workload ({cpu_t:.1f} s), oracle port with {threads} OpenMP threads"},
            "e2e_xtc": e2e_xtc, "parity": parity, "total_samples_accumulated": total_samples, "speculative_leaflets": spec_stats,
        }
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        os.write(json_fd, (json.dumps(out) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cg", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step (batch); 0 = workload default")
    ap.add_argument("--lipids", type=int, default=0, help="0 = workload default")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per step of the CPU arms; 0 = one per host thread")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--xtc-frames", type=int, default=256, help="frames of the XTC end-to-end leg (0 = skip)")
    ap.add_argument("--xtc-batch", type=int, default=16, help="frames per decoded batch of the XTC leg")
    ap.add_argument("--xtc-dev-batch", type=int, default=64, help="frames per batch of the device-decode XTC leg (>= 4 per staging thread lets a thread walk 4 frames at once)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.lipids = args.lipids or WORKLOADS[args.workload][1]
    args.frames = args.frames or WORKLOADS[args.workload][2]
    args.ref_frames = args.ref_frames or max(8, os.cpu_count() or 8)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

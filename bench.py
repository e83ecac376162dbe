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
    # name: (description, default lipids, frames of the resident window = frames per launch, windows per step)
    "cg": ("S-CG: CGOrder Martini bilayer, Global leaflets every frame (BASELINE configs[1]); 20 steps of 5120 frames = the 100k-frame trajectory", N_LIPIDS, 256, 20),
    "aa": ("S-AA-small: AAOrder POPC-like, 64 C-H bond types, static z (BASELINE configs[0] shape)", 256, 2048, 4),
    "ua": ("S-UA: UAOrder Berger-like, 64 virtual C-H, error blocks (BASELINE configs[2])", 256, 2048, 4),
    "aa_maps": ("S-AA-large: AAOrder 4096 lipids, Global leaflets, XY order maps 0.1 nm, cylinder r=8 nm (BASELINE configs[3])", 4096, 148, 4),   # 148 frames x 8 tiles = two whole waves of the map kernel
    "cg_dyn": ("S-DYN: CGOrder flat bilayer, dynamic PCA normals r=2 nm, Global leaflets (BASELINE configs[4] kernel mix)", 100000, 32, 2),   # 32 = the engine's batch limit with dynamic normals
    "ves": ("S-VES: CGOrder Martini vesicle (100 000 lipids, outer radius 51 nm, box 114 nm), dynamic PCA normals r=2 nm, "
            "spherical-clustering leaflets assigned once (BASELINE configs[4])", 100000, 32, 4),
}


def make_system(args):
    from gorder_b200 import abi, synthetic
    w, n, F = args.workload, args.lipids, args.frames
    tw = args.scaling == "strong"   # per-frame rows (error estimation): the merge of a sharded trajectory then gathers 48 B x slots per frame
    if w == "cg":
        return synthetic.s_cg(n, leaflet_mode=abi.LEAFLET_GLOBAL, max_batch_frames=F, timewise=tw)
    if w == "aa":
        return synthetic.s_aa(n, n_water=30000 * n // 256, max_batch_frames=F, timewise=tw)
    if w == "ua":
        return synthetic.s_ua(n, timewise=True, max_batch_frames=F)
    if w == "aa_maps":
        s = synthetic.s_aa(n, n_water=0, leaflet_mode=abi.LEAFLET_GLOBAL, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=(0.1, 0.1),
                           geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_BOX_CENTER, geom_dims=(8.0, float("-inf"), float("inf")),
                           geom_axis=abi.AXIS_Z, max_batch_frames=F)
        s.setup.map_span_x = (0.0, float(s.box[0]))
        s.setup.map_span_y = (0.0, float(s.box[1]))
        return s
    if w == "ves":
        return synthetic.s_ves(n, max_batch_frames=F, timewise=tw)
    if w == "cg_dyn":
        return synthetic.s_cg(n, leaflet_mode=abi.LEAFLET_GLOBAL, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0, max_batch_frames=F)
    raise SystemExit(f"unknown workload {w}")


def config(args, s, extra=None):
    st = s.setup
    c = {"workload": WORKLOADS[args.workload][0], "lipids": args.lipids, "atoms_per_frame": s.n_atoms,
         "order_slots": st.n_slots, "samples_per_frame": st.samples_per_frame(), "frames_per_step": args.frames * args.windows,
         "leaflets": {0: "none", 1: "Global, every frame", 5: "spherical clustering, once"}.get(st.leaflet_mode, str(st.leaflet_mode)),
         "normal": {0: "static z", 1: "dynamic PCA", 2: "manual"}[st.normal_mode], "pbc": bool(st.handle_pbc),
         "l2_policy": "inputs larger than L2 (a launch streams the resident window, 3.1 GB for S-CG; 126 MB L2)", "parallelism": f"frame-sharded x{args.gpus}"}
    if extra:
        c.update(extra)
    return c


def hot_kernel_name(st):
    """Name of the accumulation kernel gorder_gpu_profile brackets for this configuration (gorder_capi.cu dispatch)."""
    plain = st.handle_pbc and st.normal_mode == 0 and not st.map_enabled and st.geom_kind == 0 and not os.environ.get("GORDER_NO_FAST")
    if st.kind == 2:
        return "ua_fast_kernel" if plain and not os.environ.get("GORDER_UA_EXACT") else "ua_order_kernel"
    return "bond_fast_kernel" if plain else "bond_order_kernel"


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
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config(args, s, {"frames_per_step": nf}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def my_share(args, world, rank):
    """Frames of one step this rank analyses, as batches of <= F frames of the resident window.
    strong scaling: a step is `windows x F` frames of ONE trajectory, split over the ranks in contiguous ranges
    (gorder_b200.sharding.frame_ranges, the reference's thread model with contiguous instead of interleaved ownership);
    weak scaling: every rank analyses `windows x F` frames of its own."""
    from gorder_b200 import sharding
    F, total = args.frames, args.frames * args.windows
    n = total
    if args.scaling == "strong":
        lo, hi = sharding.frame_ranges(total, world, 1)[rank]
        n = hi - lo
    return [F] * (n // F) + ([n % F] if n % F else []), n, total


def run_ours(args):
    # libraries (NCCL, torchrun) may print to stdout: keep fd 1 for the ONE JSON line, send everything else to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    from gorder_b200 import SystemTopology
    from gorder_b200.topology import Comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gorder_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the library's own communicator (gorder_comm_*): rank 0 makes the id, torch.distributed only carries its 128 bytes
        uid = torch.zeros(Comm.ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        comm = Comm(bytes(uid.cpu().numpy().tobytes()), world, rank, local)
    s = make_system(args)
    s.setup.device = local
    F, K, W = args.frames, args.steps, args.warmup
    spf = s.setup.samples_per_frame()
    batches, n_mine, step_frames = my_share(args, world, rank)
    job_frames_per_step = step_frames if args.scaling == "strong" else step_frames * world

    # ---- inputs: a window of F frames resident in HBM, re-used with fresh frame indices ------------
    xyz, box, _ = s.frames(rank * F if args.scaling == "weak" else 0, F)
    eng = SystemTopology(s.setup)
    planes = eng.to_native(xyz)
    d_planes = torch.from_numpy(planes).cuda()
    d_box = torch.from_numpy(box).cuda()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    frame_bytes = eng.frame_floats * 4
    needed_atoms = int((eng.native_layout()[1] >= 0).sum())
    n_steps_total = W + K + 6
    cursor = [rank * n_steps_total * max(n_mine, 1)]   # contiguous, strictly increasing frame range per rank

    def step():
        for b in batches:
            eng.analyze_frames_device(d_planes.data_ptr(), d_box.data_ptr(), b, frame_index=cursor[0] + np.arange(b, dtype=np.int64), native=True)
            cursor[0] += b

    # rank 0 owns the first frames: with the whole trajectory announced to it, the merge lands the other shards' per-frame rows
    # directly behind its own (gorder_gpu_reserve_frames; no allocation inside the timed merge)
    eng.reserve_frames(n_steps_total * (step_frames if (rank == 0 and args.scaling == "strong") else n_mine) + 8)
    sampler = ClockSampler(local)
    sampler.start()
    once = s.setup.leaflet_mode != 0 and s.setup.leaflet_freq_kind == 1

    def first_step(engine, fn):
        """Frequency::Once: the table of analysed frame 0 (rank 0's first frame) reaches every shard before it accumulates."""
        if not once:
            return fn()
        r = fn() if rank == 0 else None
        if world > 1:
            engine.broadcast_leaflets(comm, 0)
        return r if rank == 0 else fn()

    eng.profile(True)   # the warm-up launches are the burst measurement: the first milliseconds of load after an idle period
    first_step(eng, step)
    eng.sync()
    eng.profile_read()   # ... without the very first step (clock ramp-up from idle, first-touch of the accumulators)
    for _ in range(W - 1):
        step()
    eng.sync()
    burst_ms, burst_n = eng.profile_read()
    eng.profile_read_normals()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.stats()["kernel_launches"]
    eng.profile(True)
    ev0, ev1, ev_r = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_win0 = time.perf_counter()
    with torch.cuda.stream(stream):   # the engine's main stream is torch's current stream
        if world > 1:   # align the DEVICE timelines of the ranks (the host barrier above leaves ~1 ms of launch skew)
            dist.all_reduce(torch.zeros(1, device="cuda"))
        ev0.record()
        for _ in range(K):
            step()
        eng.fence()   # the tail (repair + fold) of the last batch runs on a helper stream: the main stream waits for it
        ev_r.record()
        if world > 1:   # the single merge of the job, behind the C ABI: ncclReduce of the block + gather of the per-frame rows
            eng.reduce_comm(comm, 0)
        ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_win1 = time.perf_counter()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    reduce_ms = ev_r.elapsed_time(ev1)
    hot_ms, hot_n = eng.profile_read()
    nrm_ms, nrm_n = eng.profile_read_normals()
    eng.profile(False)
    launches = eng.stats()["kernel_launches"] - l0
    spec_stats = eng.speculation_stats()
    t = torch.tensor([ms, reduce_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, reduce_ms = float(t[0].item()), float(t[1].item())
    value = K * job_frames_per_step * spf / (ms * 1e-3)
    total_samples = rows_merged = None
    if rank == 0:
        res = eng.finish()
        total_samples = int(res.count[:, 0].sum())
        rows_merged = int(res.n_frames)

    # ---- end to end: pinned host AoS frames through gorder_gpu_submit, D2H of the sums every step -------
    eng2 = SystemTopology(s.setup)
    Ke = max(1, min(K, args.e2e_steps))
    eng2.reserve_frames((Ke + 2) * n_mine + 8)
    pin = torch.from_numpy(xyz).pin_memory()
    pin_box = torch.from_numpy(box).pin_memory()
    hx, hb = pin.numpy(), pin_box.numpy()
    cur2 = [rank * (Ke + 2) * max(n_mine, 1)]

    def e2e_step():
        for b in batches:
            eng2.analyze_frames(hx[:b], hb[:b], cur2[0] + np.arange(b, dtype=np.int64))
            cur2[0] += b
        return eng2.finish(totals_only=True)   # the step's result: running sums and counts (maps / per-frame rows stay on the device)

    r2 = first_step(eng2, e2e_step)   # warm-up: staging buffers, page locking
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(Ke):
        r2 = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = Ke * job_frames_per_step * spf / float(t.item())
    h2d = int(sum(batches) * (xyz[0].nbytes + box[0].nbytes))
    d2h = int(r2.sum.nbytes + r2.count.nbytes)
    eng2.close()

    # ---- end to end from a trajectory FILE on EVERY rank: host XTC decode / device XTC decode -> engine ------------
    e2e_xtc = None
    if args.xtc_frames > 0:
        xt = np.zeros(4, dtype=np.float64)   # [host-decode wall s, device-decode wall s, decode thread-seconds, ok]
        info = {}
        try:   # an optional leg: its failure must not take the headline line with it
            import tempfile
            from gorder_b200.xtc import XtcFile, write_xtc
            nx = min(args.xtc_frames, F)
            threads_x = max(1, (os.cpu_count() or 1) // world)
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, f"bench{rank}.xtc")
                write_xtc(path, xyz[:nx], box[:nx])
                fbytes = os.path.getsize(path)
                # the device-decode leg reads a longer file (the same frames appended `xtc_repeat` times): its pipeline
                # (host copy | H2D | walk + unpack + analysis) needs more than a handful of batches to show its steady state
                repeat = max(1, args.xtc_repeat // world)   # every rank writes its own file: keep the job's temporary files below ~5 GB
                for r_ in range(1, repeat):
                    write_xtc(path, xyz[:nx], box[:nx], append=True, first_step=r_ * nx)
                nx_dev = nx * repeat
                with XtcFile(path) as xf:
                    eng4 = SystemTopology(s.setup)
                    eng4.reserve_frames(2 * nx + 8)
                    eng4.run_xtc(xf, last=nx, n_threads=threads_x, batch_frames=args.xtc_batch)   # warm-up: page mappings of the file, pinned buffers
                    eng4.sync()
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    dec_s = eng4.run_xtc(xf, last=nx, n_threads=threads_x, batch_frames=args.xtc_batch, frame_index0=nx)
                    eng4.finish(totals_only=True)
                    xt[0], xt[2] = time.perf_counter() - t0, dec_s
                    eng4.close()
                    # the same file with the decode on the device: host threads only copy compressed bytes
                    eng5 = SystemTopology(s.setup)
                    reps = 2
                    eng5.reserve_frames((reps + 1) * nx_dev + 8)
                    eng5.run_xtc_device(xf, n_threads=threads_x, batch_frames=args.xtc_dev_batch)   # warm-up
                    eng5.sync()
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for r_ in range(reps):
                        moved = eng5.run_xtc_device(xf, n_threads=threads_x, batch_frames=args.xtc_dev_batch, frame_index0=(r_ + 1) * nx_dev)
                    eng5.finish(totals_only=True)
                    xt[1] = (time.perf_counter() - t0) / reps
                    eng5.close()
            xt[3] = 1.0
            info = {"frames_per_rank": nx, "file_bytes": fbytes, "bytes_per_atom": fbytes / nx / s.n_atoms, "decode_threads_per_rank": threads_x,
                    "h2d_bytes": moved, "h2d_bytes_per_atom": moved / nx_dev / s.n_atoms, "device_decode_frames_per_rank": nx_dev}
        except Exception as exc:   # noqa: BLE001
            info = {"error": f"{type(exc).__name__}: {exc}"}
        tx = torch.from_numpy(xt).cuda()
        ok = torch.tensor([xt[3]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tx, op=dist.ReduceOp.MAX)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 1.0 and rank == 0:
            nx = info["frames_per_rank"]
            e2e_xtc = {"value": world * nx * spf / float(tx[0]), "unit": UNIT, "ranks": world, **info, "wall_seconds": float(tx[0]),
                       "decode_atoms_per_s_per_thread": nx * s.n_atoms / max(float(tx[2]), 1e-9),
                       "entry": "gorder_gpu_run_xtc on every rank (host XTC decode + H2D + analysis + D2H of the sums), whole-job rate, max time over ranks",
                       "device_decode": {"value": world * info["device_decode_frames_per_rank"] * spf / float(tx[1]), "unit": UNIT, "wall_seconds": float(tx[1]), "batch_frames": args.xtc_dev_batch,
                                         "frames_per_rank": info["device_decode_frames_per_rank"],
                                         "entry": "gorder_gpu_run_xtc_device on every rank (host threads copy the compressed frames into pinned batches; xtc_walk_kernel + xtc_decode_kernel unpack them on the GPU)"}}
        elif rank == 0:
            e2e_xtc = info if "error" in info else {"error": "the XTC leg failed on another rank"}

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes: every coordinate the path needs, once (12 B x used atoms; S-CG: all 1 000 008 beads)
        launch_bytes = K * n_mine * needed_atoms * BYTES_PER_ATOM / max(hot_n, 1)   # a step is several launches
        achieved = launch_bytes / (hot_ms / max(hot_n, 1) * 1e-3) / 1e9 if hot_n else None
        traffic = traffic_src = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tr = json.load(f)
            if args.workload == "cg" and args.lipids == N_LIPIDS:   # the capture is of this workload only
                traffic = tr["dram_bytes_per_frame"] * K * n_mine / max(hot_n, 1)
                traffic_src = "replayed, not measured in this run: dram__bytes (read + write) per frame of ONE ncu --set full capture (profiles/roofline_traffic.json) x frames per launch"
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
        tw_bytes = rows_merged * s.setup.n_slots * 3 * 16 if s.setup.timewise else 0
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config(args, s, {"frames_per_step": job_frames_per_step, "frames_per_step_per_gpu": n_mine, "resident_window_frames": F,
                                       "trajectory_frames_timed": K * job_frames_per_step,
                                       "per_frame_rows": bool(s.setup.timewise),
                                       "resident_layout": "engine planes (gorder_gpu_native_layout)", "native_frame_bytes": frame_bytes}),
            "clocks": sampler.summary(t_win0, t_win1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "entry": "gorder_gpu_submit (pinned host [atom][xyz] frames) + gorder_gpu_finish (sums and counts)", "timer": "host wall clock between device syncs, max over ranks"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": hot_kernel_name(s.setup), "launches_timed": hot_n, "avg_launch_ms": hot_ms / max(hot_n, 1),
                         "algorithmic_bytes_per_launch": launch_bytes, "peak_source": peak_src,
                         "step_share": (hot_ms / (ms - reduce_ms)) if ms else None,
                         # the same kernel in the warm-up steps, i.e. before the board reaches its power cap
                         "burst": ({"avg_launch_ms": burst_ms / burst_n, "launches": burst_n, "achieved": launch_bytes / (burst_ms / burst_n * 1e-3) / 1e9,
                                    "frac": launch_bytes / (burst_ms / burst_n * 1e-3) / 1e9 / peak, "what": "warm-up steps after the first one: the first milliseconds of load after an idle period"}
                                   if burst_n and burst_ms > 0 else None),
                         "note": ("in-step duration, CUDA events around every launch of the kernel on the engine's stream (gorder_gpu_profile); "
                                  + ("speculative Global leaflets: no centre pre-pass" if spec_stats["enabled"] else "rank 0"))},
            # dynamic normals: the neighbour search + PCA of every lipid (normal.rs:160-199) dominates the step.  It is a gather through
            # L2 (per lipid ~27 cells x a few heads x 16 B), not an HBM stream: its floor is the one read of the head coordinates
            "normals_stage": ({"kernels": "cell_count / cell_scan / cell_fill / dynamic_normal_cell_kernel", "ms_per_frame": nrm_ms / max(K * n_mine, 1),
                               "step_share": nrm_ms / (ms - reduce_ms), "lipids_per_s": K * n_mine * s.setup.n_molecules_total / (nrm_ms * 1e-3) if nrm_ms else None,
                               "algorithmic_bytes_per_frame": 12 * len(s.setup.normal_heads),
                               "hbm_frac_of_floor": (12 * len(s.setup.normal_heads) * K * n_mine / (nrm_ms * 1e-3) / 1e9 / peak) if nrm_ms else None}
                              if nrm_n else None),
            "reduce": ({"ms": reduce_ms, "entry": "gorder_gpu_reduce_comm: one ncclReduce (int64 sum) of the accumulator block + grouped send / recv of the per-frame rows to rank 0",
                        "per_frame_row_bytes_gathered": tw_bytes, "rows_on_root": rows_merged} if world > 1 else None),
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{cpu_frames} frames of the same workload ({cpu_t:.1f} s), oracle port (oracle/gorder_oracle.c, -O3 -march=native) with {threads} OpenMP threads; "
                                       "the reference's published single-thread runs correspond to ~60-80 ns per sample on an i7-11700 (BASELINE.md), this port needs more (libm fmodf / acosf / cosf)"},
            "e2e_xtc": e2e_xtc, "parity": parity, "total_samples_accumulated": total_samples, "speculative_leaflets": spec_stats,
        }
    eng.close()
    if comm is not None:
        comm.close()
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
    ap.add_argument("--frames", type=int, default=0, help="frames of the resident window = frames per launch; 0 = workload default")
    ap.add_argument("--windows", type=int, default=0, help="passes over the window per step (a step = windows x frames trajectory frames); 0 = workload default")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: a step's frames belong to ONE trajectory and are split over the GPUs; weak: every GPU analyses a step's worth of its own")
    ap.add_argument("--e2e-steps", type=int, default=2, help="timed steps of the end-to-end leg (each moves frames_per_step x 12 B x atoms over PCIe)")
    ap.add_argument("--lipids", type=int, default=0, help="0 = workload default")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per step of the CPU arms; 0 = one per host thread")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--xtc-frames", type=int, default=256, help="frames of the XTC end-to-end leg (0 = skip)")
    ap.add_argument("--xtc-repeat", type=int, default=4, help="the device-decode leg reads the XTC frames appended this many times (one call)")
    ap.add_argument("--xtc-batch", type=int, default=16, help="frames per decoded batch of the XTC leg")
    ap.add_argument("--xtc-dev-batch", type=int, default=256, help="frames per batch of the device-decode XTC leg (the walk of a batch takes ~6 ms whatever its size: large batches hide it behind the copies)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.lipids = args.lipids or WORKLOADS[args.workload][1]
    args.frames = args.frames or WORKLOADS[args.workload][2]
    args.windows = args.windows or WORKLOADS[args.workload][3]
    args.ref_frames = args.ref_frames or max(8, os.cpu_count() or 8)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

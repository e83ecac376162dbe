"""AA 4096 lipids, Global leaflets, frames resident in HBM: time per frame with / without the SPEC side pass."""
import os, sys, time
import numpy as np, torch
from gorder_b200 import SystemTopology, abi, synthetic

F = 64
s = synthetic.s_aa(4096, n_water=0, leaflet_mode=abi.LEAFLET_GLOBAL, max_batch_frames=F)
xyz, box, idx = s.frames(0, F)
for mode in ("side-pass", "pre-pass"):
    if mode == "pre-pass": os.environ["GORDER_NO_SPEC_LEFTOVER"] = "1"
    eng = SystemTopology(s.setup)
    planes = torch.from_numpy(eng.to_native(xyz)).cuda(); dbox = torch.from_numpy(box).cuda()
    k = 0
    def step():
        global k
        eng.analyze_frames_device(planes.data_ptr(), dbox.data_ptr(), F, frame_index=idx + k * F, native=True); k += 1
    for _ in range(3): step()
    eng.sync(); t0 = time.perf_counter()
    for _ in range(20): step()
    eng.sync(); dt = (time.perf_counter() - t0) / 20
    st = eng.speculation_stats(); raw = eng.finish(); eng.close()
    print(f"{mode}: {1e6 * dt / F:.1f} us/frame  {int(raw.count.sum()) / (k * F) / (dt / F) :.3e} samples/s  spec={st}", flush=True)

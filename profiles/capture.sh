#!/bin/bash
# Re-captures the evidence under profiles/ on a B200 box.  Run from the repo root THROUGH gpurun, e.g.
#   gpurun --timeout 900 -- 'bash profiles/capture.sh r02 all'
# Writes into gpurun_out/ (scratch); copy what should be judged into profiles/<tag>_*.
# Parts: tests | bench | workloads | launches | ncu | xtc | all
set -u
TAG=${1:-rXX}; PART=${2:-all}; OUT=gpurun_out; mkdir -p $OUT
want() { [ "$PART" = all ] || [ "$PART" = "$1" ]; }

if want tests; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.txt
  python __graft_entry__.py --smoke 2>&1 | tail -2
fi
if want bench; then       # the headline line (S-CG), then the CPU arm
  python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err; tail -c 600 $OUT/${TAG}_bench_1gpu.json
  python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>/dev/null
fi
if want workloads; then   # the other BASELINE configs (parity + roofline of their hot kernel)
  for w in aa ua aa_maps cg_dyn; do
    timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 2 --xtc-frames 0 > $OUT/${TAG}_bench_workload_$w.json 2> $OUT/${TAG}_bench_$w.err
  done
  PYTHONPATH=. python profiles/aa_leaflets_time.py 2>&1 | tail -3 | tee $OUT/${TAG}_aa_leaflets.txt
fi
if want launches; then    # per-launch durations of one bench step (cold cache, serialised: shares, not absolutes)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --frames 128 --cpu-seconds 0.2 --xtc-frames 0 > $OUT/${TAG}_ncu_launches.log 2>&1
fi
if want ncu; then         # one full capture of the dominant kernel (source-level, lineinfo)
  ncu --set full --clock-control none --import-source on -k regex:bond_fast_kernel -s 3 -c 1 -o $OUT/${TAG}_bond_fast_kernel \
      python bench.py --steps 2 --warmup 1 --frames 128 --cpu-seconds 0.2 --xtc-frames 0 > $OUT/${TAG}_ncu_full.log 2>&1
fi
if want xtc; then         # launch list of the device-decode leg
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_xtc_device.csv \
      python bench.py --steps 1 --warmup 1 --frames 32 --cpu-seconds 0.2 --xtc-frames 64 > $OUT/${TAG}_ncu_xtc.log 2>&1
fi

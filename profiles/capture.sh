#!/bin/bash
# Re-captures the evidence under profiles/ on a B200 box.  Run from the repo root THROUGH gpurun, e.g.
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r02 bench'
# Writes into gpurun_out/ (scratch); copy what should be judged into profiles/<tag>_*.
# Parts: tests | bench | workloads | launches | ncu | ncu_ua | ncu_aa | ncu_dyn | ncu_maps | xtc | multi     (ONE ncu part per gpurun call)
# multi: under `gpurun --gpus N`: the 2-GPU engine tests, the strong-scaling bench and the platform's H2D ceiling
set -u
TAG=${1:-rXX}; PART=${2:-bench}; OUT=gpurun_out; mkdir -p $OUT
want() { [ "$PART" = "$1" ]; }

if want tests; then
  python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_gpu.txt
  python __graft_entry__.py --smoke 2>&1 | tail -2
fi
if want bench; then       # the headline line (S-CG, 20 steps x 5120 frames), then the CPU arm
  python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err; tail -c 400 $OUT/${TAG}_bench_1gpu.json
  python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>/dev/null
  python bench.py --windows 1 --cpu-seconds 1 --xtc-frames 0 --e2e-steps 1 > $OUT/${TAG}_bench_1gpu_burst.json 2>/dev/null   # 11 ms timed region: no power capping
fi
if want workloads; then   # the other BASELINE configs (parity + roofline of their hot kernel)
  for w in aa ua aa_maps cg_dyn ves; do
    timeout 500 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 2 --xtc-frames 0 --e2e-steps 1 > $OUT/${TAG}_bench_workload_$w.json 2> $OUT/${TAG}_bench_$w.err
  done
fi
if want launches; then    # per-launch durations of one bench step (cold cache, serialised: shares, not absolutes)
  python bench.py --steps 2 --warmup 1 --windows 2 --cpu-seconds 0.2 --xtc-frames 0 --e2e-steps 1 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --windows 2 --cpu-seconds 0.2 --xtc-frames 0 --e2e-steps 1 > $OUT/${TAG}_ncu_launches.log 2>&1
fi
ncu_full() {   # $1 kernel regex, $2 name, rest: bench arguments
  local k=$1 n=$2; shift 2
  python bench.py "$@" > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $OUT/${TAG}_$n python bench.py "$@" > $OUT/${TAG}_ncu_$n.log 2>&1
}
if want ncu; then ncu_full bond_fast_kernel bond_fast_kernel --steps 2 --warmup 1 --windows 1 --frames 128 --cpu-seconds 0.2 --xtc-frames 0 --e2e-steps 1; fi
if want ncu_ua; then ncu_full ua_fast_kernel ua_fast_kernel --workload ua --steps 2 --warmup 1 --windows 1 --cpu-seconds 0.2 --xtc-frames 0 --e2e-steps 1; fi
if want ncu_aa; then ncu_full bond_fast_kernel bond_fast_kernel_aa_small --workload aa --steps 2 --warmup 1 --windows 1 --cpu-seconds 0.2 --xtc-frames 0 --e2e-steps 1; fi
if want ncu_dyn; then ncu_full dynamic_normal_sorted_kernel dynamic_normal_sorted_kernel --workload cg_dyn --steps 2 --warmup 1 --windows 1 --cpu-seconds 0.2 --xtc-frames 0 --e2e-steps 1; fi
if want xtc; then         # launch list of the device-decode leg
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_xtc_device.csv \
      python bench.py --steps 1 --warmup 1 --windows 1 --frames 32 --cpu-seconds 0.2 --xtc-frames 64 > $OUT/${TAG}_ncu_xtc.log 2>&1
fi
if want ncu_maps; then ncu_full bond_order_kernel aa_maps_kernel --workload aa_maps --steps 2 --warmup 1 --windows 1 --cpu-seconds 0.2 --xtc-frames 0 --e2e-steps 1; fi
if want multi; then
  N=${3:-2}
  python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_multi_${N}gpu.txt
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 profiles/h2d_scaling.py > $OUT/${TAG}_h2d_scaling_${N}gpu.json 2>/dev/null
fi

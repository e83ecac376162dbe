#!/bin/bash
# compute-sanitizer over one small run of every kernel family (profiles/sanitizer_cases.py).  ONE tool per gpurun call
# (B200_PROFILING.md): gpurun --timeout 1500 -- 'bash profiles/sanitize.sh memcheck'   (then racecheck, in another call)
set -u
TOOL=${1:-memcheck}; OUT=gpurun_out; mkdir -p $OUT
python profiles/sanitizer_cases.py > $OUT/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/sanitizer_plain.log; exit 1; }
timeout 1300 compute-sanitizer --tool $TOOL --print-limit 20 --log-file $OUT/sanitizer_$TOOL.log python profiles/sanitizer_cases.py > $OUT/sanitizer_${TOOL}_stdout.log 2>&1
echo "exit $?"; tail -5 $OUT/sanitizer_${TOOL}_stdout.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" $OUT/sanitizer_$TOOL.log | sort | uniq -c | head -20

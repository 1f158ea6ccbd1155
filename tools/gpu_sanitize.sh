#!/bin/bash
# compute-sanitizer over every entry point (tools/sanitize_small.py): memcheck, racecheck (shared-memory hazards), synccheck
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/${tool}_r02.log 2>&1; echo "$tool rc=$?"; grep -c "ERROR\|Error\|hazard" gpurun_out/${tool}_r02.log; tail -3 gpurun_out/${tool}_r02.log
done

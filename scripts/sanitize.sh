#!/bin/bash
# compute-sanitizer passes over the kernel tests: racecheck (shared-memory hazards between the
# producer / MMA / epilogue roles), synccheck (barrier misuse), memcheck on a subset.  Logs -> gpurun_out/sanitizer_*.log
SEL='gemm or attention or layernorm or conv3x3 or stem or batchnorm or head or adam or preprocess or u8'
for tool in synccheck racecheck memcheck; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" -p no:cacheprovider \
    > gpurun_out/sanitizer_${tool}_r02.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_${tool}_r02.log | tail -5
done

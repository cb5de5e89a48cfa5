#!/usr/bin/env bash
# compute-sanitizer pass over the kernel-level GPU tests (race detection / sanitizers).
# Run on a B200 box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'   (memcheck ~10x slower than a plain run)
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 --log-file gpurun_out/sanitize_$tool.log \
      python -m pytest tests/test_gpu_features.py tests/test_gpu_tc.py -m gpu -x -q -k "not 1024 and not 4096 and not 38188" \
      > gpurun_out/sanitize_${tool}_pytest.log 2>&1
  echo "$tool: exit $? ; $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$tool.log) summaries"
  grep "ERROR SUMMARY" gpurun_out/sanitize_$tool.log | sort | uniq -c | head
done

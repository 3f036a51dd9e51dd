#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled mbarrier pipelines / aliasing staging buffers (conv_gemm.cu), the
# warp-collective ProbEn kernels (fuse.cu) and one whole detector forward.  Run on the GPU box:
#     tools/sanitize.sh [seconds per run, default 420]
# Logs: gpurun_out/san_<tool>_<suite>.log, one-line verdicts: gpurun_out/sanitizer_summary.txt (copied to profiles/).
LIM=${1:-420}
mkdir -p gpurun_out
: > gpurun_out/sanitizer_summary.txt
run() {  # tool suite-name pytest-args...
  local tool=$1 name=$2; shift 2
  local log=gpurun_out/san_${tool}_${name}.log
  timeout $LIM compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 --log-file $log \
      python -m pytest "$@" -x -q -p no:cacheprovider > gpurun_out/san_${tool}_${name}.pytest 2>&1
  local rc=$?
  local errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
  echo "$tool $name rc=$rc | $(tail -1 gpurun_out/san_${tool}_${name}.pytest) | ${errs}" >> gpurun_out/sanitizer_summary.txt
}
run memcheck conv tests/test_conv_gpu.py
run memcheck fusion tests/test_fusion_gpu.py
run memcheck detector tests/test_pipeline_gpu.py -k "pipeline_fusion_equals or fused_frame_resize"
run racecheck conv tests/test_conv_gpu.py
run racecheck fusion tests/test_fusion_gpu.py
run racecheck detector tests/test_pipeline_gpu.py -k "pipeline_fusion_equals"
run synccheck conv tests/test_conv_gpu.py
cat gpurun_out/sanitizer_summary.txt

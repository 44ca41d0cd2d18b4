#!/bin/bash
# compute-sanitizer passes over the native kernels.  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash bench/sanitize_gpu.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 1500 -- 'bash bench/sanitize_gpu.sh'   (adds the multi-GPU collectives under memcheck)
# Each tool runs a reduced test selection (the sanitizers slow kernels down 10-100x); reports land in gpurun_out/sanitize_*.log and
# the script exits non-zero if any tool reports an error.
set -x
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
rc=0
sel_ops='tests/test_gpu_ops.py -k "not trainer"'
sel_gemm='tests/test_gpu_gemm.py -k "reference and (128-64-64 or 256-256-64) or nt_wgrad and 256-64-64 or two_cta and 256-128-64 or conv_implicit and 8-64-16"'
for tool in memcheck racecheck synccheck initcheck; do
  eval timeout 900 $CS --tool $tool --error-exitcode 9 --target-processes all python -m pytest $sel_ops -m gpu -x -q > gpurun_out/sanitize_ops_$tool.log 2>&1 || rc=1
  eval timeout 900 $CS --tool $tool --error-exitcode 9 --target-processes all python -m pytest $sel_gemm -m gpu -x -q > gpurun_out/sanitize_gemm_$tool.log 2>&1 || rc=1
  tail -3 gpurun_out/sanitize_ops_$tool.log gpurun_out/sanitize_gemm_$tool.log
done
export SHIPYARD_TEST_QUICK=1
timeout 1200 $CS --tool memcheck --error-exitcode 9 --target-processes all python -m pytest tests/test_gpu_coll.py -m gpu -x -q -k "single or multi_gpu_collectives" > gpurun_out/sanitize_coll_memcheck.log 2>&1 || rc=1
tail -3 gpurun_out/sanitize_coll_memcheck.log
grep -l "ERROR SUMMARY: [1-9]" gpurun_out/sanitize_*.log && rc=1
exit $rc

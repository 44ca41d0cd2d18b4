#!/bin/bash
# time-boxed compute-sanitizer memcheck over one launch of each kernel family
mkdir -p gpurun_out
timeout 170 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python bench/sanitize_quick.py > gpurun_out/sanitize_quick_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_quick_memcheck.log
tail -5 gpurun_out/sanitize_quick_memcheck.log

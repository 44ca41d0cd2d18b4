#!/bin/bash
# Time-boxed compute-sanitizer passes (memcheck, racecheck, synccheck, initcheck) over one launch of every native kernel family
# (bench/sanitize_quick.py) and, under racecheck + memcheck, the single-GPU collective battery.  ~4 minutes on one B200.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
CS=/usr/local/cuda/bin/compute-sanitizer
rc=0
for tool in memcheck racecheck synccheck initcheck; do
  timeout 240 $CS --tool $tool --error-exitcode 9 python bench/sanitize_quick.py > gpurun_out/sanitize_quick_$tool.log 2>&1 || rc=1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|kernels ran' gpurun_out/sanitize_quick_$tool.log | tr '\n' ' ')"
done
for tool in memcheck racecheck; do
  SHIPYARD_COLL_TIMEOUT_MS=600000 timeout 400 $CS --tool $tool --error-exitcode 9 python tests/_coll_worker.py --rank 0 --world 1 --session san$$ --device 0 --quick > gpurun_out/sanitize_coll_$tool.log 2>&1 || rc=1
  echo "== coll $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY| OK' gpurun_out/sanitize_coll_$tool.log | tr '\n' ' ' | cut -c1-300)"
done
exit $rc

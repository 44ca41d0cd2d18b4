#!/bin/bash
# Round 23 (last GPU shot of the round, ~3 min budget): hardware check of the halo-load 3x3 convolution.
#  1. numerics with base_offset = (addr >> 7) & 7 and with base_offset = 0, each in its own process (a trap is sticky)
#  2. timing of the mode that passed against the im2col kernels and cuDNN
#  3. the new GPU tests, 4. a short bench with the halo candidates in the autotuner
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for m in 1 0; do
  timeout 40 python bench/halo_check.py numerics $m > gpurun_out/halo_numerics_$m.jsonl 2> gpurun_out/halo_numerics_$m.err
  echo "numerics base_mode=$m rc=$?"; tail -3 gpurun_out/halo_numerics_$m.err
  cat gpurun_out/halo_numerics_$m.jsonl
done
good=$(python - <<'PY'
import json
best = ""
for m in ("1", "0"):
    try:
        rows = [json.loads(l) for l in open(f"gpurun_out/halo_numerics_{m}.jsonl") if l.strip()]
    except Exception:
        rows = []
    one = [r for r in rows if not r["pair"]]
    if one and all("error" not in r and r.get("fprop_rel", 1) < 0.02 and r.get("dgrad_rel", 0) < 0.02 for r in one) and len(one) >= 6:
        best = m
        break
print(best)
PY
)
echo "passing base_mode: '$good'"
if [ -n "$good" ]; then
  timeout 40 python bench/halo_check.py timing $good > gpurun_out/halo_timing.jsonl 2> gpurun_out/halo_timing.err
  echo "timing rc=$?"; cat gpurun_out/halo_timing.jsonl; tail -3 gpurun_out/halo_timing.err
  SHIPYARD_HALO_BASE_MODE=$good timeout 60 python -m pytest tests/test_gpu_conv_halo.py -x -q > gpurun_out/halo_pytest.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/halo_pytest.log
  SHIPYARD_HALO_BASE_MODE=$good SHIPYARD_CONV_HALO=1 SHIPYARD_CONV_PLAN_DUMP=1 timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/halo_bench.json 2> gpurun_out/halo_bench.err
  echo "bench rc=$?"; cat gpurun_out/halo_bench.json; tail -3 gpurun_out/halo_bench.err
fi

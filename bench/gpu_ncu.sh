#!/bin/bash
# ncu captures for the epilogue question (1 GPU, ~5 min): the 64 -> 256 1x1 layer as a GEMM (M = 802 816, N = 256, K = 64), CTA pairs
# with fused statistics, once with the TMA-store epilogue and once with SHIPYARD_GEMM_DIRECT_STORE=1; plus the 64 -> 64 layer (BN = 64).
# Read the reports here with `ncu -i gpurun_out/r2_ncu_*.ncu-rep --page raw --csv` (B200_PROFILING.md): compare
# smsp__average_warps_issue_stalled_* of the epilogue warps, l1tex / lts store throughput and sm__inst_executed_pipe_uniform.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
NCU="ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -c 1 -s 2"
$NCU -o gpurun_out/r2_ncu_n256_tma -f python bench/gemm_epilogue_probe.py --one 1 2cta stats > gpurun_out/r2_ncu_n256_tma.log 2>&1
SHIPYARD_GEMM_DIRECT_STORE=1 $NCU -o gpurun_out/r2_ncu_n256_direct -f python bench/gemm_epilogue_probe.py --one 1 2cta stats > gpurun_out/r2_ncu_n256_direct.log 2>&1
$NCU -o gpurun_out/r2_ncu_n64_tma -f python bench/gemm_epilogue_probe.py --one 0 > gpurun_out/r2_ncu_n64_tma.log 2>&1
SHIPYARD_GEMM_DIRECT_STORE=1 SHIPYARD_GEMM_EPI_ALT=1 $NCU -o gpurun_out/r2_ncu_n64_direct_alt -f python bench/gemm_epilogue_probe.py --one 0 > gpurun_out/r2_ncu_n64_direct_alt.log 2>&1
ls -la gpurun_out/r2_ncu_*.ncu-rep
for f in gpurun_out/r2_ncu_*.ncu-rep; do
  ncu -i "$f" --page raw --csv > "$f.csv" 2>/dev/null
  python - "$f" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1] + ".csv")))
if len(rows) >= 3:
    hdr, vals = rows[0], rows[2]
    want = ("gpu__time_duration.sum", "dram__bytes_write.sum", "dram__bytes_read.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio")
    print(sys.argv[1], {h: v for h, v in zip(hdr, vals) if h in want})
PY
done

#!/usr/bin/env python3
"""Turn the scratch measurements under gpurun_out/ into the tracked summaries in profiles/.
Run after a GPU round: `python profiles/make_reports.py`.  Missing inputs are skipped."""
import collections
import csv
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(PEAKS.get("hbm_gbs", 6583.8)); TF = float(PEAKS.get("bf16_tflops", 1710.4))


def latest(pattern):
    fs = sorted(glob.glob(os.path.join(G, pattern)), key=os.path.getmtime)
    return fs[-1] if fs else None


def jsonl(path):
    out = []
    for l in open(path):
        l = l.strip()
        if l.startswith("{"):
            try:
                out.append(json.loads(l))
            except Exception:
                pass
    return out


def ncu_table(csv_path, out_md, title, note=""):
    rows = list(csv.reader(open(csv_path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("dram__bytes.sum.per_second", "DRAM bandwidth"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__waves_per_multiprocessor", "waves/SM"),
            ("launch__occupancy_limit_registers", "blocks/SM (register limit)"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2 -> SM read bandwidth"), ("launch__cluster_size", "cluster size"),
            ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
            ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall: long scoreboard %")]
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\n{note}\n\nSource: `ncu --set full --clock-control none --import-source on` (raw page exported with `ncu -i ... --page raw --csv`).\n"
                f"Roofline denominators: measured HBM copy {HBM} GB/s, measured cuBLAS bf16 {TF} TFLOP/s (MEASURED_PEAKS.json).\n\n")
        for d in data:
            name = re.sub(r"\(.*", "", d[idx["Kernel Name"]])
            f.write(f"## {name}\n\n| metric | value |\n|---|---|\n")
            for key, label in want:
                if key in idx and d[idx[key]] not in ("", "n/a"):
                    f.write(f"| {label} (`{key}`) | {d[idx[key]]} {units[idx[key]]} |\n")
            try:
                bw = d[idx["dram__bytes.sum.per_second"]]; u = units[idx["dram__bytes.sum.per_second"]]
                gbs = float(bw.replace(",", "")) * {"Tbyte/s": 1e3, "Gbyte/s": 1.0}.get(u, 1.0)
                f.write(f"| **fraction of measured HBM copy bandwidth** | **{gbs / HBM:.3f}** |\n")
            except Exception:
                pass
            f.write("\n")


def ncu_hot_lines(rep, kernels, out_md):
    """Top SASS instructions by warp-stall samples (ncu source page) for each kernel of a report."""
    import subprocess
    with open(out_md, "a") as f:
        f.write("\n# Hottest SASS instructions (warp stall samples, `ncu --page source --csv`)\n\n")
        for k in kernels:
            try:
                out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{k}"], capture_output=True, text=True, timeout=120).stdout
            except Exception:
                continue
            rows = list(csv.reader(out.splitlines()))
            hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
            if hi is None:
                continue
            hdr = rows[hi]; body = []
            for r in rows[hi + 1:]:
                if r and r[0] in ("Address", "Kernel Name"):
                    break                                   # next launch of the same kernel
                if len(r) == len(hdr):
                    body.append(r)
            si = hdr.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in hdr else 2
            tot = sum(int(r[si] or 0) for r in body) or 1
            f.write(f"## {k}\n\n| samples | share | SASS |\n|---|---|---|\n")
            for r in sorted(body, key=lambda r: -int(r[si] or 0))[:8]:
                f.write(f"| {r[si]} | {100 * int(r[si] or 0) / tot:.1f}% | `{r[1].strip()}` |\n")
            f.write("\n")


def step_breakdown(csv_path, out_md, steps_hint=None):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        k = re.sub(r"^void ", "", k).replace("<unnamed>::", "")
        if not k.startswith(("gemm_bf16", "k_")):
            k = re.sub(r"<.*", "", k)                  # library kernels: drop the template noise
        k = k[:80]
        agg[k][0] += 1; agg[k][1] += v; tot += v
    with open(out_md, "w") as f:
        f.write("# ResNet-50 training step: kernel time by kernel name (all convolution passes forced onto the tcgen05 kernels)\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` over a short "
                f"`bench.py` run (`{os.path.basename(csv_path)}`; warm-up + timed steps, eager mode so every kernel is visible). Numbers under a profiler are for "
                "attribution only, never bench values.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
            f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
        f.write(f"\nTotal kernel time in capture: {tot / 1e3:.2f} ms\n")


def coll_tables(out_md):
    with open(out_md, "w") as f:
        f.write("# Collective sweeps (shipyard kernels vs NCCL 2.28.9 on the same box)\n\nDevice-timed (CUDA events), max over ranks, fp32 payloads unless noted. "
                "`busbw` uses the usual factors (all-reduce 2(N-1)/N, all-gather / reduce-scatter / all-to-all (N-1)/N). NVLink 5 nominal 900 GB/s per direction per GPU.\n\n")
        for path in sorted(glob.glob(os.path.join(G, "coll_sweep*_n*.jsonl"))):
            rows = jsonl(path)
            if not rows:
                continue
            f.write(f"## {os.path.basename(path)} (world {rows[0].get('world')})\n\n| op | bytes | NCCL us | shipyard auto us | best algo us | speedup vs NCCL | shipyard busbw GB/s | NCCL busbw GB/s | busbw / 900 |\n|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                bb = r.get("sy_busbw_GBs") or 0
                f.write(f"| {r['op']} | {r['bytes']} | {r.get('nccl_us', 0):.1f} | {r.get('sy_auto_us', 0):.1f} | {r.get('sy_best_us', 0):.1f} | {r.get('speedup_vs_nccl')} | {bb} | {r.get('nccl_busbw_GBs')} | {bb / 900:.2f} |\n")
            f.write("\n")


def simple_table(rows, out_md, title, note):
    if not rows:
        return
    keys = list(rows[0].keys())
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\n{note}\n\n| " + " | ".join(keys) + " |\n|" + "---|" * len(keys) + "\n")
        for r in rows:
            f.write("| " + " | ".join(str(r.get(k, "")) for k in keys) + " |\n")


def bench_lines(out_md):
    rows = []
    for path in sorted(glob.glob(os.path.join(G, "bench*_n*.log")) + glob.glob(os.path.join(G, "base*_n*.log")), key=os.path.getmtime):
        for r in jsonl(path):
            if "value" in r:
                rows.append({"file": os.path.basename(path), "impl": r.get("impl"), "n_gpus": r.get("n_gpus"), "images_per_sec": r.get("value"),
                             "ms_per_step": r.get("ms_per_step"), "e2e_images_per_sec": (r.get("e2e") or {}).get("value"),
                             "own_kernels_per_step": r.get("own_kernels_per_step"), "transport": (r.get("config") or {}).get("collective_transport"),
                             "sm_mhz": (r.get("clocks") or {}).get("sm_mhz"), "throttle": ",".join((r.get("clocks") or {}).get("reasons") or []) or "none"})
    simple_table(rows, out_md, "bench.py results collected during the round (ResNet-50, batch 256/GPU, bf16, synthetic)",
                 "Chronological; later rows include later optimisations. `impl=nccl-baseline` is the plain PyTorch DDP/NCCL arm of the same recipe.")


def main():
    p = latest("ncu_bn*_raw.csv")
    if p:
        ncu_table(p, os.path.join(P, "ncu_bn_kernels.md"), "Fused BatchNorm kernels on the largest ResNet-50 layer (256x56x56x256 bf16 = 411 MB)",
                  "These four kernels are ~45% of the training step (see step_breakdown.md), all HBM-bound.")
        rep = p.replace("_raw.csv", ".ncu-rep")
        if os.path.exists(rep):
            ncu_hot_lines(rep, ["k_bn_stats", "k_bn_apply_fwd", "k_bn_bwd_reduce", "k_bn_bwd_apply"], os.path.join(P, "ncu_bn_kernels.md"))
    p = latest("ncu_gemm*_raw.csv")
    if p:
        ncu_table(p, os.path.join(P, "ncu_gemm_kernels.md"), "tcgen05 kernels: 8192^3 GEMM (CTA pair and 1-CTA) and 256-channel 3x3 convolution at 28x28, batch 256 (fprop+stats, dgrad on CTA pairs; split-K wgrad)",
                  "Captured back to back from one script (`bench/gpu_round15.sh`).")
        rep = p.replace("_raw.csv", ".ncu-rep")
        if os.path.exists(rep):
            ncu_hot_lines(rep, ["gemm_bf16_tn_2cta_kernel", "gemm_bf16_tn_kernel", "gemm_bf16_nt_splitk_kernel"], os.path.join(P, "ncu_gemm_kernels.md"))
    p = os.path.join(G, "launches9_conv1_notc0.csv")          # SHIPYARD_CONV_IMPL=tc: every conv pass on our kernels
    if not os.path.exists(p):
        p = latest("launches_shipyard*.csv")
    if p:
        step_breakdown(p, os.path.join(P, "step_breakdown.md"))
    coll_tables(os.path.join(P, "coll_sweeps.md"))
    for name, title in (("gemm_bench.jsonl", "tcgen05 TN GEMM (1-CTA and CTA-pair) vs cuBLAS"), ("wgrad_bench.jsonl", "split-K MN-major wgrad / MN-major-B dgrad vs cuBLAS"),
                        ("conv_bench.jsonl", "TMA-im2col implicit-GEMM convolution (fprop / dgrad / wgrad) vs cuDNN, ResNet-50 shapes at batch 256")):
        fp = os.path.join(G, name)
        if os.path.exists(fp):
            simple_table(jsonl(fp), os.path.join(P, name.replace(".jsonl", ".md")), title,
                         f"CUDA-event timing, 256 MB L2 flush between iterations, median. Roofline = max(flops / {TF} TFLOP/s measured cuBLAS, bytes / {HBM} GB/s measured copy).")
    p = latest("bn_bench*.log")
    if p:
        simple_table(jsonl(p), os.path.join(P, "bn_bench.md"), "Fused BN(+ReLU) forward / backward bandwidth per ResNet-50 layer shape (batch 256)",
                     "Cold operands (buffer ring > 2x L2). Includes autograd/Python launch overhead, which dominates the small layers (in the model the step is a CUDA graph).")
    bench_lines(os.path.join(P, "bench_history.md"))
    cp = os.path.join(G, "conv_plan.json")
    if os.path.exists(cp):
        tab = json.load(open(cp))
        rows = [dict({"shape n x cin x h x w x cout x k x stride": k}, fprop=v["fprop"], fused_bn_stats=v["stats"], dgrad=v["dgrad"], wgrad=v["wgrad"],
                     **{kk: vv for kk, vv in (v.get("timings_us") or {}).items()}) for k, v in tab.items()]
        keys = []
        for r in rows:
            for k in r:
                if k not in keys:
                    keys.append(k)
        rows = [{k: r.get(k, "") for k in keys} for r in rows]
        simple_table(rows, os.path.join(P, "conv_dispatch_plan.md"), "Convolution dispatcher: measured per-pass choice per layer shape (us)",
                     "`ops/conv.py` times the tcgen05 implicit-GEMM kernels against cuDNN the first time a shape is seen and keeps the faster per pass; "
                     "fprop timings include the separate BN-statistics pass when the epilogue does not produce them.")
    # K10 / HPCG one-line results: keep every N, latest run per N
    with open(os.path.join(P, "k10_hpcg_results.md"), "w") as f:
        f.write("# Fused GEMM + collective (K10), HPCG and TensorFlow-Distributed recipe results\n\nOne line per run, straight from the benchmark programs (`tests/_k10_worker.py --bench`, "
                "`recipes/HPCG-Infiniband-IntelMPI/run_hpcg.py`).  K10: m = n = 8192, K split across ranks, times are device-timed max over ranks.\n\n")
        for pat in ("k10*bench_n*.log", "hpcg*_n*.log", "tfdist*_n*.log"):
            for path in sorted(glob.glob(os.path.join(G, pat)), key=os.path.getmtime):
                txt = [l.strip() for l in open(path) if ("bench m=n" in l or l.startswith("{"))]
                if txt:
                    f.write(f"* `{os.path.basename(path)}`: {txt[-1]}\n")
    for old in ("k10_bench_n8.txt", "hpcg6_n1.txt", "hpcg7_n1.txt"):
        try:
            os.remove(os.path.join(P, old))
        except OSError:
            pass


if __name__ == "__main__":
    main()

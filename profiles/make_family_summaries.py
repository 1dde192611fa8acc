"""profiles/<tag>_families_*: launch shares and ncu --set full metrics of the DiffSTG / DiffusionTS kernels.

    python profiles/make_family_summaries.py r01
Inputs (gpurun_out/, produced on a B200):
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/r01_families_launches.csv python profiles/tools/gpu_prof_families.py
    ncu --set full --clock-control none --import-source on -k 'regex:stg_tcn_ln|dts_fourier_topk|stg_gated_aggregate|nsx_step' -c 40 \
        -o gpurun_out/prof_families python profiles/tools/gpu_prof_families.py
The profiled command: 2 DiffSTG denoise steps on 16 384 replica rows (BASELINE config 5 architecture, BA-100 graph) and one
DiffusionTS loop iteration t=99->98 (x0 prediction, DDIM mean, 3 Langevin iterations, infill) on 1000 rows (config 4).
"""
import collections
import csv
import json
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OWN = ("stg_tcn_ln_kernel", "stg_gated_aggregate_kernel", "stg_posterior_kernel", "gauss_fill_kernel", "nsx_step_kernel",
       "sigma_estimation_kernel", "fx_add_ln_split_small_kernel", "fx_embed_split_kernel",
       "dts_fourier_topk_fwd_kernel", "dts_fourier_topk_bwd_kernel", "dts_ddim_step_kernel", "dts_adagrad_kernel",
       "dts_infill_kernel", "dts_attn_fwd_kernel", "dts_attn_bwd_kernel", "fx_split_kernel",
       "dts_attn_tc_fwd_kernel", "dts_attn_tc_bwd_kernel", "dts_ln_fwd_a3_kernel", "dts_ln_fwd_kernel", "dts_ln_bwd_kernel",
       "gemm3_pair_kernel", "gemm3_kernel", "stg_tcn_ln_cat_kernel", "stg_tcn_mma_kernel", "stg_conv1d_kernel",
       "stg_gated_aggregate_smem_kernel")
lines = [l for l in open("gpurun_out/%s_families_launches.csv" % tag) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
# split the launch list at the first DiffusionTS-only kernel: everything before belongs to DiffSTG
first_dts = next(i for i, r in enumerate(rows) if "dts_" in r["Kernel Name"] or "fourier" in r["Kernel Name"])
# the DiffusionTS part starts with its own gauss_fill-free torch.randn; be conservative: cut at the last stg kernel
last_stg = max(i for i, r in enumerate(rows[:first_dts]) if "stg_" in r["Kernel Name"])
last_dts = max(i for i, r in enumerate(rows) if re.search("dts_ddim|dts_adagrad|dts_infill|dts_fourier", r["Kernel Name"]))
parts = {"DiffSTG (2 denoise steps, 16384 rows)": rows[: last_stg + 1],
         "DiffusionTS (1 loop iteration, K=3, 1000 rows)": rows[last_stg + 1:last_dts + 1],
         "NsDiff_spatial (f(x) + g(x) on 200 rows, 2-step chain on 16000 rows)": rows[last_dts + 1:]}
with open("profiles/%s_families_launches_summary.txt" % tag, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none   python profiles/tools/gpu_prof_families.py\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
    for name, rs in parts.items():
        agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
        for r in rs:
            v = float(r["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[r["Metric Unit"]]
            m = re.search("(" + "|".join(OWN) + ")", r["Kernel Name"])
            short = m.group(1) if m else re.sub(r"<.*", "", r["Kernel Name"].replace("<unnamed>::", "")).replace("void ", "")[:72]
            agg[short][0] += 1
            agg[short][1] += v
            tot += v
        f.write("\n## %s: total GPU time %.2f ms over %d launches\n" % (name, tot / 1e6, len(rs)))
        f.write("%12s %8s %7s  %s\n" % ("time_ms", "share", "n", "kernel"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
            f.write("%12.3f %7.2f%% %7d  %s%s\n" % (t / 1e6, 100 * t / tot, n, k, "   <- own" if k in OWN else ""))
        mine = sum(t for k, (n, t) in agg.items() if k in OWN)
        f.write("# own kernels: %.1f%% of GPU time; the rest are library GEMMs / elementwise ops\n" % (100 * mine / tot))

# optional: argv[2] = ncu report, argv[3] = name of the metrics file (default: the families capture)
rep = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prof_families.ncu-rep"
metrics_name = sys.argv[3] if len(sys.argv) > 3 else "families"
import os
if not os.path.exists(rep):
    print(open("profiles/%s_families_launches_summary.txt" % tag).read())
    sys.exit(0)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
out = []
for vals in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in keep:
            d[h] = v if not u else "%s %s" % (v, u)
    out.append(d)
json.dump(out, open("profiles/%s_%s_ncu_full_metrics.json" % (tag, metrics_name), "w"), indent=1)
print(open("profiles/%s_families_launches_summary.txt" % tag).read())
for d in out:
    print({k: d.get(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size")})

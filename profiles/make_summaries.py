"""Turn the raw ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries in profiles/.

    python profiles/make_summaries.py r01                # reads gpurun_out/{launches.csv,sampler_raw.csv,sampler_src.csv}
    python profiles/make_summaries.py r02 sampler_ws     # round 2: the warp-specialised kernel (file names follow the kernel)

Inputs are produced on the GPU box by (see scratch job scripts / B200_PROFILING.md):
    UPD_BENCH_SKIP_CPU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
        --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1
    UPD_BENCH_SKIP_CPU=1 ncu --set full --clock-control none --import-source on -k regex:sampler_tc -s 1 -c 1 \
        -o gpurun_out/prof_sampler python bench.py --steps 1 --warmup 1
and here by  ncu -i gpurun_out/prof_sampler.ncu-rep --page raw|source --csv.
"""
import collections
import csv
import json
import re
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
kern = sys.argv[2] if len(sys.argv) > 2 else "sampler_tc"
OWN = ("sampler_ws_kernel", "gemm3_pair_kernel", "gemm3_kernel", "gram_centered_kernel", "prediction_error_kernel",
       "sampler_tc_kernel", "sampler_simt_kernel", "welford_over_samples", "window_means", "sigma_estimation_kernel",
       "fx_attention_kernel", "fx_add_ln_split_kernel", "fx_split_kernel", "fx_embed_split_kernel")

# ---- launch list -> shares ----
lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
    m = re.search("(" + "|".join(OWN) + ")", row["Kernel Name"])
    short = m.group(1) if m else re.sub(r"<.*", "", row["Kernel Name"].replace("<unnamed>::", "")).replace("void ", "")[:70]
    agg[short][0] += 1
    agg[short][1] += v
    tot += v
with open("profiles/%s_bench_launches_summary.txt" % tag, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 9000   python bench.py --steps 1 --warmup 1\n")
    f.write("# (UPD_BENCH_SKIP_CPU=1 UPD_BENCH_SKIP_CONFIGS=1: the CPU baseline / parity and other-configs legs are skipped under the profiler).  Per-launch times are\n")
    f.write("# cold-cache and serialised: compare SHARES.  The command runs 2 resident sweeps + 2 end-to-end sweeps\n# + the one sweep of the wall-time measurement (5 sampler launches).\n")
    f.write("# total GPU time %.1f ms over %d launches\n" % (tot / 1e6, sum(a[0] for a in agg.values())))
    f.write("%12s %8s %7s  %s\n" % ("time_ms", "share", "n", "kernel"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        f.write("%12.2f %7.2f%% %7d  %s\n" % (t / 1e6, 100 * t / tot, n, k))
    mine = sum(t for k, (n, t) in agg.items() if k in OWN)
    f.write("# own kernels: %.2f%% of GPU time; the rest are torch elementwise / copy kernels of the f(x) embedding and any library GEMM left\n" % (100 * mine / tot))

# ---- sampler: ncu --set full ----
rows = list(csv.reader(open("gpurun_out/sampler_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
d = {}
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        d[h] = {"unit": u, "value": v}
json.dump(d, open("profiles/%s_%s_ncu_full_metrics.json" % (tag, kern), "w"), indent=1)
rd = float(d["dram__bytes_read.sum"]["value"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[d["dram__bytes_read.sum"]["unit"]]
wr = float(d["dram__bytes_write.sum"]["value"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[d["dram__bytes_write.sum"]["unit"]]
json.dump({"kernel": "%s_kernel<NsDiff,F=1>" % kern,
           "source": "profiles/%s_%s_ncu_full_metrics.json (ncu --set full on the bench workload)" % (tag, kern),
           "dram_bytes_per_launch_bench_workload": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
           "algorithmic_bytes_per_launch": 181 * 100 * 100 * 100 * 4 + 2 * 181 * 100 * 100 * 4},
          open("profiles/sampler_tc_ncu_full.json", "w"), indent=1)

# ---- sampler: SASS opcode mix and hottest stall sites ----
rows = list(csv.reader(open("gpurun_out/sampler_src.csv")))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op = collections.defaultdict(lambda: [0, 0])
top, ti, ts = [], 0, 0
stall_tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    s = r[ix["Source"]]
    inst, samp = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)
    toks = s.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDTM", "STTM", "UTC", "SYNCS", "BAR")) else op.split(".")[0]
    by_op[op][0] += inst
    by_op[op][1] += samp
    ti += inst
    ts += samp
    for c in stall_cols:
        stall_tot[c] += int(r[ix[c]] or 0)
    top.append((samp, s, sorted([(c, int(r[ix[c]] or 0)) for c in stall_cols if int(r[ix[c]] or 0) > 0], key=lambda x: -x[1])[:2]))
with open("profiles/%s_%s_sass_mix.txt" % (tag, kern), "w") as f:
    f.write("# ncu --page source of the same capture: SASS opcode mix (warp instructions) and sampled stall sites\n")
    f.write("# total warp instructions %d, samples %d\n" % (ti, ts))
    for op, (i, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:24]:
        f.write("%-16s inst %6.2f%%   samples %6.2f%%\n" % (op, 100 * i / ti, 100 * s / ts))
    f.write("# stall reasons over all samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / ts) for k, v in stall_tot.most_common(8)) + "\n")
    f.write("# hottest sites\n")
    for samp, s, st in sorted(top, key=lambda t: -t[0])[:14]:
        f.write("%5.2f%%  %-64s %s\n" % (100 * samp / ts, s[:64], st))
print(open("profiles/%s_bench_launches_summary.txt" % tag).read())
print(open("profiles/%s_%s_sass_mix.txt" % (tag, kern)).read())
for k in ("gpu__time_duration.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    print(k, d[k])

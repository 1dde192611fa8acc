"""Key `ncu --set full` metrics of one captured kernel -> profiles/<tag>_<name>_ncu_full_metrics.json

    ncu -i gpurun_out/prof_gemm3_pair.ncu-rep --page raw --csv > gpurun_out/gemm3_pair_raw.csv
    python profiles/make_kernel_metrics.py r02 gemm3_pair gpurun_out/gemm3_pair_raw.csv
"""
import csv
import json
import sys

tag, name, path = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active")
out = []
for vals in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            d[h] = v if not u else {"unit": u, "value": v}
    out.append(d)
json.dump(out if len(out) > 1 else out[0], open("profiles/%s_%s_ncu_full_metrics.json" % (tag, name), "w"), indent=1)
for d in out:
    for k, v in d.items():
        print(k, v)

"""Summarises an .ncu-rep (ncu --set full) as JSON: the metrics profiles/*_ncu_full_summary.json keep, per launch.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_full_summary.json"""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum"]
res = []
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = {"report": rep.split("/")[-1], "Kernel Name": r[hdr.index("Kernel Name")].replace("rsr::", "")}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        res.append(d)
print(json.dumps(res, indent=1))

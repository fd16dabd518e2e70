"""profiles/ncu_summary.json from one `ncu --set full` report that holds the launches of ONE frame (tools/profile_step.py under
`ncu --set full --clock-control none -k regex:conv_|pack_ -s <launches of the warm-up frames> -c <launches per frame>`):
per launch name (rsr_kernel_name) the DRAM bytes, duration, tensor-pipe and DRAM utilisation under the profiler.
usage: python tools/ncu_to_summary.py report.ncu-rep > profiles/ncu_summary.json"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = lambda r, k: float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] not in ("", "n/a") else None
unit = lambda k: rows[1][hdr.index(k)] if k in hdr else ""
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[hdr.index("Kernel Name")]
    if "pack_input" in name:   # (reports made before conv0 read the image itself)
        key = "pack_input"
    elif "conv_fused" in name:
        key = "conv5|6|7+conv8|9 (fused)"
    elif "conv_pair" in name:
        key = "conv1|2|3"
    elif "<5, 1," in name or "(int)5, (int)1," in name:
        key = "conv0"
    elif "<5, 5," in name or "(int)5, (int)5," in name:
        key = "conv1|2|3"
    else:   # 3x3 strip kernels: SHUF (template argument 11) > 0 marks the last launch
        args = name.split("ConvCfg<")[1].split(">")[0].replace("(int)", "").replace("(bool)", "").split(",")
        ng, shuf = int(args[3]), int(args[10])
        key = "conv10+expand+skip" if shuf > 0 else ("conv5|6|7" if ng >= 3 and "conv5|6|7" not in res else "conv8|9")
    rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
    res[key] = {
        "kernel": name.replace("rsr::", "")[:160],
        "dram_bytes_per_launch": (rd or 0) * scale.get(unit("dram__bytes_read.sum"), 1.0) + (wr or 0) * scale.get(unit("dram__bytes_write.sum"), 1.0),
        "duration_us_under_ncu": col(r, "gpu__time_duration.sum"),
        "tensor_pipe_active_pct": col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "mem_tensor_active_pct": col(r, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_throughput_pct": col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_clock_ghz_under_ncu": col(r, "sm__cycles_elapsed.avg.per_second"),
        "registers_per_thread": col(r, "launch__registers_per_thread"),
        "dynamic_smem_kb": col(r, "launch__shared_mem_per_block_dynamic"),
        "grid": col(r, "launch__grid_size"), "block": col(r, "launch__block_size"),
    }
print(json.dumps(res, indent=1))

#!/usr/bin/env python
"""tools/make_profiles.py ROUND — turn the gpurun_out/ captures of a round (launches_rNN.csv,
prof_rNN.ncu-rep, bench_full.log, opbench_rNN.log) into the committed summaries under profiles/.
See profiles/r01_summary.md for the ncu command lines that produce the inputs."""
import csv
import json
import os
import subprocess
import sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)
rows = [r for r in csv.reader(open(f"gpurun_out/launches_{rnd}.csv")) if len(r) > 10]
hdr = rows[0]
ik, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
launches = [(int(r[iid]), r[ik], float(r[iv].replace(",", ""))) for r in rows[1:]]
with open(f"profiles/{rnd}_launches.csv", "w") as f:
    f.write("id,kernel,gpu__time_duration_ns\n")
    for i, k, v in launches:
        f.write(f'{i},"{k}",{v:.0f}\n')
subprocess.run(f"ncu -i gpurun_out/prof_{rnd}.ncu-rep --page raw --csv > gpurun_out/prof_{rnd}_raw.csv 2>/dev/null", shell=True)
raw = list(csv.reader(open(f"gpurun_out/prof_{rnd}_raw.csv")))
h, units = raw[0], raw[1]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
traffic = {}
for r in raw[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    traffic[name] = float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
json.dump(traffic, open(f"profiles/{rnd}_traffic.json", "w"), indent=1)
# keep the columns the summary reads (the raw page has ~1500 metric columns)
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
keep = [k for k in keep if k in h]
with open(f"profiles/{rnd}_fullset.csv", "w", newline="") as f:
    wr_ = csv.writer(f)
    for r in raw:
        wr_.writerow([r[h.index(k)] for k in keep])
print("wrote profiles/", rnd)

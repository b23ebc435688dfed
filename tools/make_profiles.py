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
if os.path.exists(f"gpurun_out/prof_{rnd}.ncu-rep"):
    subprocess.run(f"ncu -i gpurun_out/prof_{rnd}.ncu-rep --page raw --csv > gpurun_out/prof_{rnd}_raw.csv 2>/dev/null", shell=True)
raw = list(csv.reader(open(f"gpurun_out/prof_{rnd}_raw.csv")))
h, units = raw[0], raw[1]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
traffic = {}
for r in raw[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    traffic[name] = float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
sha = open(f"gpurun_out/so_sha16_{rnd}.txt").read().strip() if os.path.exists(f"gpurun_out/so_sha16_{rnd}.txt") else None
# bench.py reads this file: per-launch DRAM bytes (read + write) of the step kernels at 1e9 rows, with the hash of the
# kernel sources the capture was taken on (bench.so_sha16(): sha256 over csrc/ + the header; roofline.traffic_same_build)
json.dump({"so_sha16": sha, "rows": 1000000000, "source": f"ncu --set full (gpurun_out/prof_{rnd}.ncu-rep), dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "kernels": traffic}, open(f"profiles/{rnd}_traffic.json", "w"), indent=1)
# the ncu details page of the same capture (speed-of-light, occupancy, stall sections per kernel), as text
if os.path.exists(f"gpurun_out/prof_{rnd}_details.txt"):
    import shutil
    shutil.copy(f"gpurun_out/prof_{rnd}_details.txt", f"profiles/{rnd}_ncu_details.txt")
# SASS evidence of the memory-path instructions, from the shipped library itself (no GPU needed)
sass = subprocess.run("cuobjdump -sass arrow-rs_b200/libarrow_cuda.so", shell=True, capture_output=True, text=True).stdout
fn, counts = None, {}
for line in sass.splitlines():
    if "Function :" in line:
        fn = line.split("Function :")[1].strip()
        counts[fn] = {}
    elif fn:
        for op in ("UBLKCP", "LDGSTS", "REDG", "ATOMS", "LDG.E.128", "STG.E.128", "REDUX", "SYNCS", "HMMA", "UTCHMMA", "UTCQMMA"):
            if op in line:
                counts[fn][op] = counts[fn].get(op, 0) + 1
with open(f"profiles/{rnd}_sass_excerpts.txt", "w") as f:
    f.write(f"# cuobjdump -sass arrow-rs_b200/libarrow_cuda.so (kernel-source hash bench.so_sha16() = {sha}): memory-path instruction counts per hot kernel\n")
    f.write("# UBLKCP = cp.async.bulk (TMA engine), LDGSTS = cp.async (global -> shared), REDG = fire-and-forget global atomics,\n")
    f.write("# REDUX = redux.sync, no HMMA / UTC*MMA: no tensor-core instruction anywhere (HBM-bound integer / byte work)\n")
    import subprocess as sp
    for k, v in counts.items():
        if any(t in k for t in ("k_take", "k_filter_fused", "k_filter_values_async", "k_arith", "k_cmp_v2", "k_reduce", "k_dict_copy", "k_plan_mask", "k_cast_v2")) and v:
            dem = sp.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
            f.write(f"{dem[:150]}: {v}\n")
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

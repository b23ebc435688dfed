#!/usr/bin/env python
"""tools/make_summary.py ROUND — build profiles/ROUND_summary.md from the committed artefacts:

  profiles/ROUND_launches.csv      ncu launch list of `bench.py --steps 2 --warmup 3 --no-e2e --no-cpu`
  profiles/ROUND_fullset.csv       `ncu --set full` raw page of the step kernels (tools/make_profiles.py)
  profiles/ROUND_bench_n1.json     the bench line (never taken under ncu)
  profiles/ROUND_opbench.json      tools/opbench.py lines
  profiles/ROUND_recordbatch_n*.json, profiles/ROUND_bench_n*.json   multi-GPU lines (optional)
"""
import csv
import glob
import json
import os
import sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
P = "profiles"
out = []
w = out.append


def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("(bool)", "").replace("(int)", "").replace("(acu_agg_op)", "")


bench = json.load(open(f"{P}/{rnd}_bench_n1.json"))
w(f"# Round {rnd[1:].lstrip('0')} profile summary (B200, 1e9-row step: filter -> take -> add -> sum)\n")
w("Commands (under `gpurun`, one GPU):\n")
w("```")
w(f"bash tools/gpu_profiles.sh {rnd}      # the whole capture; its ncu passes:")
w(f"ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_{rnd}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs")
w(f'ncu --set full --clock-control none --import-source on -k regex:"k_arith|k_take|k_filter_fused|k_reduce|k_plan_mask" -s 20 -c 8 -o gpurun_out/prof_{rnd} python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs')
w(f"ncu -i gpurun_out/prof_{rnd}.ncu-rep --page details > {P}/{rnd}_ncu_details.txt    # committed; --page raw --csv -> {rnd}_fullset.csv / {rnd}_traffic.json")
w(f'ncu --set full --clock-control none -k regex:"k_dict_copy|k_dict_block_totals|k_cmp_v2|k_cast_v2" -c 6 python tools/opbench.py --only "dict|lt f64|cast i64" --reps 1   # -> {rnd}_ncu_details_ops.txt')
w("python bench.py                      # the bench line itself is never taken under ncu")
w("python bench.py --impl reference     # the CPU arm (oracle/refbench.cpp on all host cores)")
w("python tools/opbench.py              # per-op table (configs 2-4)")
w("python tools/recordbatch_bench.py    # config 5 (under torchrun for N > 1)")
w("```\n")

# ---- 1. launch list ----
rows = list(csv.DictReader(open(f"{P}/{rnd}_launches.csv")))
# the last complete step = the last launches between two k_plan_mask occurrences
idx = [i for i, r in enumerate(rows) if "k_plan_mask" in r["kernel"]] + [len(rows)]
segs = [rows[a:b] for a, b in zip(idx[:-1], idx[1:])]
segs = [g for g in segs if any("k_arith" in r["kernel"] for r in g) and any("k_reduce" in r["kernel"] for r in g)]
step = segs[-2] if len(segs) >= 2 else (segs[-1] if segs else rows)
# (cut the segment after the step's reduction + result-block reset: what follows belongs to later calls)
ia = max(i for i, r in enumerate(step) if "k_arith" in r["kernel"])
last = min(i for i, r in enumerate(step) if i > ia and "k_reduce" in r["kernel"])
step = step[:last + 2] if last + 1 < len(step) and "k_res_reset" in step[last + 1]["kernel"] else step[:last + 1]
agg = {}
for r in step:
    k = short(r["kernel"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["gpu__time_duration_ns"]) / 1e6
tot = sum(v[1] for v in agg.values())
cls_of = {"k_arith": "arith", "k_filter_fused": "filter", "k_filter_values": "filter", "k_compress_bits": "filter", "k_zero_outputs": None, "k_zero_bitmap_dev": None,
          "k_take": "take", "k_reduce": "reduce", "k_plan": "filter_plan"}
w("## 1. Launch list of one step (ncu per-launch times: cold-cache, serialised — compare SHARES)\n")
w("| kernel | launches | ncu time (ms) | share of step (ncu) | share of step (bench.py CUDA events, per kernel class) |")
w("|---|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    cls = next((c for p, c in cls_of.items() if k.startswith(p)), None)
    share = f"{100 * bench['kernels'][cls]['share_of_step']:.1f} % ({cls} class)" if cls and cls in bench["kernels"] else ""
    w(f"| `{k}` | {n} | {ms:.4f} | {100 * ms / tot:.1f} % | {share} |")
w(f"\nStep total under ncu: {tot:.3f} ms; bench.py (CUDA events, not under ncu): {bench['ms_per_step']:.3f} ms/step = {bench['value']:.0f} Mrows/s.\n")

# ---- 2. full set ----
fs = f"{P}/{rnd}_fullset.csv"
if os.path.exists(fs):
    raw = list(csv.reader(open(fs)))
    h, units = raw[0], raw[1]
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
    cols = [(m, t) for m, t in cols if m in h]
    w("## 2. `ncu --set full` of the step kernels (per launch)\n")
    w("| kernel | " + " | ".join(t for _, t in cols) + " |")
    w("|---|" + "---|" * len(cols))
    for r in raw[2:]:
        cells = []
        for m, _ in cols:
            v, u = r[h.index(m)], units[h.index(m)]
            try:
                fv = float(v.replace(",", ""))
                v = f"{fv:.3f}" if fv < 100 else f"{fv:.0f}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip() if u and u not in ("%", "register/thread", "") else v)
        w(f"| `{short(r[h.index('Kernel Name')])}` | " + " | ".join(cells) + " |")
    w(f"\n`traffic` for bench.py's roofline object = dram read + write of `k_arith<double,...>` (see {rnd}_traffic.json).\n")

# ---- 3. bench line ----
w("## 3. bench.py line (N=1), measured on the same box\n")
w("```json")
w(json.dumps(bench, indent=1))
w("```\n")

# ---- 4. per-op table ----
op = f"{P}/{rnd}_opbench.json"
if os.path.exists(op):
    w("## 4. Per-op table (`tools/opbench.py`; kernel-only CUDA-event time; peak = MEASURED_PEAKS.json hbm_gbs \"of measured\")\n")
    w("| op | rows | kernel ms | GB/s (algorithmic) | % of measured HBM peak | Mrows/s | note |")
    w("|---|---|---|---|---|---|---|")
    for line in open(op):
        line = line.strip()
        if not line.startswith("{"):
            continue
        r = json.loads(line)
        w(f"| {r['op']} | {r['rows']:.3g} | {r['kernel_ms']} | {r['achieved_gbs']} | {100 * r['frac_of_measured_peak']:.1f} | {r['mrows_s']} | {r.get('note', '')} |")
    w("")

# ---- 5. config 5 + scaling ----
rb = sorted(glob.glob(f"{P}/{rnd}_recordbatch_n*.json"), key=lambda p: int(p.split("_n")[-1].split(".")[0]))
bn = sorted(glob.glob(f"{P}/{rnd}_bench_n*.json"), key=lambda p: int(p.split("_n")[-1].split(".")[0]))
if rb or len(bn) > 1:
    w("## 5. Multi-GPU (weak scaling: every rank owns its own rows; device time, max over ranks)\n")
    w("| bench | GPUs | Mrows/s (all ranks) | ms/step | efficiency vs N=1 | notes |")
    w("|---|---|---|---|---|---|")
    for name, files in (("bench.py (filter+take+add+sum, 1e9 rows/GPU)", bn), ("tools/recordbatch_bench.py (config 5, 15 x 2^26-row 8-column batches/GPU)", rb)):
        base = None
        for f in files:
            d = json.load(open(f))
            n = d["n_gpus"]
            if n == 1:
                base = d["value"]
            eff = f"{100 * d['value'] / (n * base):.1f} %" if base else ""
            note = ""
            if "frac_of_measured_peak" in d:
                note = f"{100 * d['frac_of_measured_peak']:.1f} % of HBM roofline over the whole pipeline ({d['algorithmic_bytes_per_step'] / 1e9:.1f} GB algorithmic/step/GPU); kernel time {d['kernel_ms_per_step']} ms"
            elif d.get("e2e"):
                note = f"e2e {d['e2e']['value']:.0f} Mrows/s" if d["e2e"].get("value") else ""
            w(f"| {name} | {n} | {d['value']:.0f} | {d['ms_per_step']:.3f} | {eff} | {note} |")
    w("")
extra = f"{P}/{rnd}_notes.md"
if os.path.exists(extra):
    w(open(extra).read())
open(f"{P}/{rnd}_summary.md", "w").write("\n".join(out) + "\n")
print(f"wrote {P}/{rnd}_summary.md")

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/opbench.py --only "dict" | grep '^{' | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02k_dict_launches.csv python tools/opbench.py --only "dict" --reps 1 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02k_dict_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
last={}
for r in rows[1:]:
    if 'k_dict' in r[ki]: last[(r[ki][:50], r[mi])]=r[vi]
for k,v in last.items(): print(k, v)
P
(timeout 2400 python -m pytest tests -q -m gpu -x) > gpurun_out/r02k_gputests.log 2>&1; tail -4 gpurun_out/r02k_gputests.log
for st in 1 2 3; do
timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 --streams $st > gpurun_out/r02k_rb_s$st.json 2> gpurun_out/r02k_rb_s$st.err
python -c "
import json
d=json.load(open('gpurun_out/r02k_rb_s$st.json'))
print('rb streams=$st', round(d['ms_per_step'],2), round(d['kernel_ms_per_step'],2), {k:round(x['ms_per_step'],2) for k,x in d['kernels'].items()}, d['check']['sums_bits'][:2])" || tail -5 gpurun_out/r02k_rb_s$st.err
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02k_bench.json"))
print(d["ms_per_step"], d["value"], d["gpu_launches"], d["check_vs_oracle"], d["e2e"]["value"], d["e2e"].get("numa_note"), d["cpu_baseline"]["value"], d["cpu_baseline"].get("spread"))
print(json.dumps(d["configs"])[:2500])
P
tail -3 gpurun_out/r02k_bench.err

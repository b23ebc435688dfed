#!/bin/bash
mkdir -p gpurun_out
# dictionary gather v2
(timeout 900 python -m pytest tests/test_gpu_dict.py tests/test_gpu_parity.py -q -m gpu -x -k "dict or bytes or utf8 or string") 2>&1 | tail -4
(timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "config4 or config5") 2>&1 | tail -3
timeout 300 python tools/opbench.py --only "dict" | grep '^{' | cut -c1-220
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02h_dict_launches.csv python tools/opbench.py --only "dict" --reps 1 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02h_dict_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
last={}
for r in rows[1:]:
    if 'k_dict' in r[ki] or 'k_scan' in r[ki] or 'k_bitmap' in r[ki]: last[(r[ki][:50], r[mi])]=r[vi]
for k,v in last.items(): print(k, v)
P
# stream-ordered sections
(timeout 900 python -m pytest tests/test_gpu_async.py -q -m gpu -x) 2>&1 | tail -8
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_select.py tests/test_gpu_golden.py tests/test_gpu_recordbatch.py -q -m gpu -x) 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs > gpurun_out/r02h_bench_async.json 2> gpurun_out/r02h_bench_async.err
ACU_BENCH_SYNC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs > gpurun_out/r02h_bench_sync.json 2> gpurun_out/r02h_bench_sync.err
python - <<'P'
import json
for k in ("async","sync"):
    try:
        d=json.load(open(f"gpurun_out/r02h_bench_{k}.json")); print(k, d["ms_per_step"], d["value"], d.get("sync_gap_ms_per_step"), d["gpu_launches"], d["check"])
    except Exception as e: print(k, "failed", e, open(f"gpurun_out/r02h_bench_{k}.err").read()[-800:])
P
timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 > gpurun_out/r02h_rb.json 2> gpurun_out/r02h_rb.err
python -c "
import json
d=json.load(open('gpurun_out/r02h_rb.json'))
print('rb', round(d['ms_per_step'],2), round(d['kernel_ms_per_step'],2), {k:round(x['ms_per_step'],2) for k,x in d['kernels'].items()})"

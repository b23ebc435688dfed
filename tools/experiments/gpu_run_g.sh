#!/bin/bash
mkdir -p gpurun_out
# 1. pipelined filter: parity + A/B
(ACU_FILTER_PIPE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_recordbatch.py tests/test_gpu_select.py -q -m gpu -x -k "filter") 2>&1 | tail -3
(ACU_FILTER_PIPE=1 timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "filter or config1") 2>&1 | tail -3
timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' | cut -c1-110 > gpurun_out/r02g_filter_base.txt
ACU_FILTER_PIPE=1 timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' | cut -c1-110 > gpurun_out/r02g_filter_pipe.txt
ACU_FILTER_PIPE=1 ACU_FILTER_TILES_PER_WARP=16 timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' | cut -c1-110 > gpurun_out/r02g_filter_pipe16.txt
for f in base pipe pipe16; do echo "== $f"; cat gpurun_out/r02g_filter_$f.txt; done
ACU_FILTER_PIPE=1 timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 > gpurun_out/r02g_rb_pipe.json 2> gpurun_out/r02g_rb_pipe.err
python -c "
import json
d=json.load(open('gpurun_out/r02g_rb_pipe.json'))
print('rb pipe', round(d['ms_per_step'],2), round(d['kernel_ms_per_step'],2), {k:round(x['ms_per_step'],2) for k,x in d['kernels'].items()})"
# 1b. stream-ordered sections
(timeout 900 python -m pytest tests/test_gpu_async.py tests/test_gpu_parity.py -q -m gpu -x) 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs > gpurun_out/r02g_bench_async.json 2> gpurun_out/r02g_bench_async.err
ACU_BENCH_SYNC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs > gpurun_out/r02g_bench_sync.json 2> gpurun_out/r02g_bench_sync.err
python - <<'P'
import json
for k in ("async","sync"):
    try:
        d=json.load(open(f"gpurun_out/r02g_bench_{k}.json")); print(k, d["ms_per_step"], d["value"], d.get("sync_gap_ms_per_step"), d["gpu_launches"], d["check"])
    except Exception as e: print(k, "failed", e, open(f"gpurun_out/r02g_bench_{k}.err").read()[-800:])
P
# 2. dictionary gather: launch list + full set
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_dict_launches.csv python tools/opbench.py --only "dict" --reps 1 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_dict_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:60]].append(float(r[vi].replace(',','')))
    except Exception: pass
for k,v in d.items(): print(k, len(v), 'last(us)=', round(v[-1]/1000,1), 'sum(us)=', round(sum(v)/1000,1))
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dict_copy -c 2 -f -o gpurun_out/r02g_dict python tools/opbench.py --only "dict" --reps 1 > /dev/null 2>&1
ncu -i gpurun_out/r02g_dict.ncu-rep --page details > gpurun_out/r02g_dict.details.txt 2>&1
ncu -i gpurun_out/r02g_dict.ncu-rep --page source --csv > gpurun_out/r02g_dict.sass.csv 2>&1
rm -f gpurun_out/r02g_dict.ncu-rep
grep -E "Duration|Issue Slots Busy|No Eligible|Registers Per|Achieved Occupancy|Theoretical Occ|L1/TEX Hit|Shared Memory Config|Bank conflict|bank conflict|Mem Busy|Max Bandwidth|Mem Pipes Busy|Stall" gpurun_out/r02g_dict.details.txt | head -60

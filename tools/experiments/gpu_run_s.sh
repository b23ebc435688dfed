#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_views.py tests/test_abi.py tests/test_host_cpp.py -q -m gpu -x) 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02s_bench.json')); r=d['roofline']
print(d['ms_per_step'], r['frac'], r['traffic'], r['traffic_same_build'], r['so_sha16'], r['traffic_so_sha16'])"

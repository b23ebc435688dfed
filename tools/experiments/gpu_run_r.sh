#!/bin/bash
mkdir -p gpurun_out
for m in 0 3 4; do
  ACU_FILTER_REG=$m timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' | cut -c1-100 > gpurun_out/r02r_filter_reg$m.txt
  echo "== reg=$m"; cat gpurun_out/r02r_filter_reg$m.txt
done
for m in 3 4; do
(ACU_FILTER_REG=$m timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_recordbatch.py tests/test_gpu_async.py -q -m gpu -x -k "filter or chain or record") 2>&1 | tail -2
ACU_FILTER_REG=$m timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 --streams 1 > gpurun_out/r02r_rb_reg$m.json 2> gpurun_out/r02r_rb.err
python -c "
import json
d=json.load(open('gpurun_out/r02r_rb_reg$m.json'))
print('rb reg=$m', round(d['ms_per_step'],2), round(d['kernel_ms_per_step'],2), {k:round(x['ms_per_step'],2) for k,x in d['kernels'].items()})"
done

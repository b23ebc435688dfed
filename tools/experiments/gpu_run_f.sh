#!/bin/bash
mkdir -p gpurun_out
for v in legacy 1 4 16; do
  if [ $v = legacy ]; then export ACU_FILTER_LEGACY=1; else unset ACU_FILTER_LEGACY; export ACU_FILTER_TILES_PER_WARP=$v; fi
  timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 > gpurun_out/r02f_rb_$v.json 2> gpurun_out/r02f_rb_$v.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02f_rb_$v.json'))
print('$v', round(d['ms_per_step'],2), round(d['kernel_ms_per_step'],2), {k:round(x['ms_per_step'],2) for k,x in d['kernels'].items()})"
  timeout 300 python tools/opbench.py --only "filter i64 s=0.1" | grep '^{' | cut -c1-100
done
unset ACU_FILTER_LEGACY ACU_FILTER_TILES_PER_WARP
(timeout 600 python -m pytest tests/test_gpu_recordbatch.py tests/test_gpu_parity.py tests/test_cdata.py -q -m gpu -x -k "filter or record or cdata or device") 2>&1 | tail -3
(timeout 900 python -m pytest tests/test_gpu_dict.py tests/test_gpu_parity.py -q -m gpu -x -k "dict or bytes or utf8 or string") 2>&1 | tail -4
(timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "config4 or config5") 2>&1 | tail -3
ACU_BYTES_NO_DICT=1 timeout 300 python tools/opbench.py --only "dict" | grep '^{' | cut -c1-220
timeout 300 python tools/opbench.py --only "dict" | grep '^{' | cut -c1-220

#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_views.py tests/test_gpu_dict.py -q -m gpu -x) 2>&1 | tail -5
for cfg in "3 --no-e2e --no-cpu" "3 --no-e2e" "3 --no-cpu" "2 --no-e2e --no-cpu" "2"; do
  set -- $cfg; st=$1; shift
  ACU_RB_STREAMS=$st timeout 900 python bench.py --steps 5 --warmup 3 $@ > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
  python - "$cfg" <<'P'
import json,sys
try:
    d=json.load(open("gpurun_out/r02l_bench.json"))
    c=d["configs"]["cfg5"]["filter_record_batch -> take_record_batch -> 6 sums"]
    print(sys.argv[1], "->", round(c["ms_per_step"],2), round(c["kernel_ms"],2), c["kernel_ms_by_class"])
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/r02l_bench.err").read()[-500:])
P
done

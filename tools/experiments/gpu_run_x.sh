#!/bin/bash
# final N=2 validation (stdout carries exactly one JSON line per arm) + refreshed N=1 lines
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29531 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/x_ref_n2.out 2> gpurun_out/x_ref_n2.err
$TR --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/x_n2.out 2> gpurun_out/x_n2.err
$TR --master-port 29533 tools/recordbatch_bench.py --gpus 2 > gpurun_out/x_rb_n2.out 2> gpurun_out/x_rb_n2.err
CUDA_VISIBLE_DEVICES=0 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/x_ref_n1.out 2> gpurun_out/x_ref_n1.err
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 > gpurun_out/x_n1.out 2> gpurun_out/x_n1.err
for f in x_ref_n2 x_n2 x_rb_n2 x_ref_n1 x_n1; do python - $f <<'P'
import json,sys
f=sys.argv[1]; t=open(f"gpurun_out/{f}.out").read()
lines=[l for l in t.splitlines() if l.strip()]
try:
    d=json.loads(lines[0]); r=d.get("roofline") or {}
    print(f, "stdout_lines", len(lines), round(d["value"]), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"), r.get("traffic_same_build"), d.get("check_vs_oracle"))
except Exception as e:
    print(f, "BAD", e, t[:300], open(f"gpurun_out/{f}.err").read()[-500:])
P
done

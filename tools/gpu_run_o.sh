#!/bin/bash
# eight GPUs of one box (gpurun --gpus 8): sharded step + e2e with NUMA binding, config 5
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r02_n8.json 2> gpurun_out/bench_r02_n8.err
timeout 400 $TR --master-port 29523 tools/recordbatch_bench.py --gpus 8 > gpurun_out/recordbatch_r02_n8.json 2> gpurun_out/recordbatch_r02_n8.err
python - <<'P'
import json
for f in ("bench_r02_n8","recordbatch_r02_n8"):
    try:
        t=open(f"gpurun_out/{f}.json").read(); d=json.loads(t[t.index('{'):])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("final_reduce_ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("numa_note"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-800:])
P
nvidia-smi topo -m 2>/dev/null | head -14

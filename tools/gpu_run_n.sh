#!/bin/bash
# two GPUs of one box (gpurun --gpus 2): the sharded step with the in-place all-reduce, config 5, and the world-2 GPU tests
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
ACU_BENCH_SYNC=1 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_r02_n2_sync.json 2> gpurun_out/bench_r02_n2_sync.err
$TR --master-port 29513 tools/recordbatch_bench.py --gpus 2 > gpurun_out/recordbatch_r02_n2.json 2> gpurun_out/recordbatch_r02_n2.err
$TR --master-port 29514 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_r02_ref_n2.json 2> gpurun_out/bench_r02_ref_n2.err
python - <<'P'
import json
for f in ("bench_r02_n2","bench_r02_n2_sync","recordbatch_r02_n2","bench_r02_ref_n2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("value"), d.get("ms_per_step"), d.get("final_reduce_ms_per_step"), d.get("sync_gap_ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("numa_note"), d.get("check"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-600:])
P

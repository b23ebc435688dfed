#!/bin/bash
# one gpurun call: host info, full GPU test-suite, both bench arms, filter A/B, take sweep (outputs under gpurun_out/)
mkdir -p gpurun_out
{ nproc; free -g; numactl -H; nvidia-smi topo -m; } > gpurun_out/r02a_host.txt 2>&1
(time timeout 1500 python -m pytest tests -q -m gpu -x) > gpurun_out/r02a_gputests.log 2>&1
tail -5 gpurun_out/r02a_gputests.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err
head -c 400 gpurun_out/r02a_ref.json; tail -3 gpurun_out/r02a_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 600 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
ACU_FILTER_LEGACY=1 timeout 600 python tools/opbench.py --only filter > gpurun_out/r02a_opbench_filter_legacy.txt 2>&1
timeout 600 python tools/opbench.py --only filter > gpurun_out/r02a_opbench_filter_fused.txt 2>&1
tail -6 gpurun_out/r02a_opbench_filter_legacy.txt gpurun_out/r02a_opbench_filter_fused.txt
for v in default 8,4 8,5 16,2 16,3 4,6 4,8; do
  echo "== take variant $v"
  if [ "$v" = default ]; then timeout 300 python tools/opbench.py --only take | grep '^{' ; else ACU_TAKE_VARIANT=$v timeout 300 python tools/opbench.py --only take | grep '^{'; fi
done > gpurun_out/r02a_take_sweep.txt 2>&1
cat gpurun_out/r02a_take_sweep.txt | cut -c1-200

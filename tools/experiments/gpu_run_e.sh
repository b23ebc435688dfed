#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -q -m gpu -x) > gpurun_out/r02e_gputests.log 2>&1
tail -15 gpurun_out/r02e_gputests.log
timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' > gpurun_out/r02e_filter.txt 2>&1
cut -c1-110 gpurun_out/r02e_filter.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
cut -c1-200 gpurun_out/r02e_bench.json; tail -3 gpurun_out/r02e_bench.err

#!/bin/bash
# full GPU tests (ctx result-block scheme, warp-centric bytes kernel), bytes A/B, filter ncu captures (small), short bench
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -q -m gpu -x) > gpurun_out/r02c_gputests.log 2>&1
tail -5 gpurun_out/r02c_gputests.log
ACU_BYTES_LEGACY=1 timeout 300 python tools/opbench.py --only "dict" | grep '^{' > gpurun_out/r02c_dict_legacy.txt 2>&1
timeout 300 python tools/opbench.py --only "dict" | grep '^{' > gpurun_out/r02c_dict_warp.txt 2>&1
cut -c1-200 gpurun_out/r02c_dict_legacy.txt gpurun_out/r02c_dict_warp.txt
ACU_BYTES_LEGACY=1 timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 > gpurun_out/r02c_rb_legacy.json 2> gpurun_out/r02c_rb_legacy.err
timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 > gpurun_out/r02c_rb_warp.json 2> gpurun_out/r02c_rb_warp.err
cut -c1-400 gpurun_out/r02c_rb_legacy.json gpurun_out/r02c_rb_warp.json; tail -2 gpurun_out/r02c_rb_warp.err
K='regex:k_filter_fused|k_filter_values_async|k_compress_bits'
for sel in 0.1 0.9; do
  ACU_FILTER_MINB=5 timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -c 3 -f -o gpurun_out/r02c_f5_$sel python tools/opbench.py --only "filter i64 s=$sel" --reps 1 > /dev/null 2>&1
  ACU_FILTER_LEGACY=1 timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -c 6 -f -o gpurun_out/r02c_leg_$sel python tools/opbench.py --only "filter i64 s=$sel" --reps 1 > /dev/null 2>&1
done
for f in gpurun_out/r02c_*.ncu-rep; do ncu -i $f --page details > ${f%.ncu-rep}.details.txt 2>&1; done
ls -la gpurun_out | head -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/r02c_bench_short.json 2> gpurun_out/r02c_bench_short.err
cut -c1-300 gpurun_out/r02c_bench_short.json; tail -3 gpurun_out/r02c_bench_short.err

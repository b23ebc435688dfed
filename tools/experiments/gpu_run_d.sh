#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_select.py tests/test_gpu_golden.py tests/test_gpu_recordbatch.py tests/test_gpu_coalesce.py tests/test_gpu_cmp_bytes.py -q -m gpu -x) > gpurun_out/r02d_tests.log 2>&1
tail -4 gpurun_out/r02d_tests.log
(timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "filter or config1") > gpurun_out/r02d_tests_cfg.log 2>&1
tail -3 gpurun_out/r02d_tests_cfg.log
ACU_FILTER_LEGACY=1 timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' > gpurun_out/r02d_filter_legacy.txt 2>&1
timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' > gpurun_out/r02d_filter_fused6.txt 2>&1
ACU_FILTER_MINB=5 timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' > gpurun_out/r02d_filter_fused5.txt 2>&1
for f in legacy fused6 fused5; do echo "== $f"; cut -c1-110 gpurun_out/r02d_filter_$f.txt; done
K='regex:k_filter_fused'
ACU_FILTER_MINB=5 timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -c 3 -f -o gpurun_out/r02d_f5_0.1 python tools/opbench.py --only "filter i64 s=0.1" --reps 1 > /dev/null 2>&1
ncu -i gpurun_out/r02d_f5_0.1.ncu-rep --page details > gpurun_out/r02d_f5_0.1.details.txt 2>&1
ncu -i gpurun_out/r02d_f5_0.1.ncu-rep --page source --csv > gpurun_out/r02d_f5_0.1.sass.csv 2>&1
rm -f gpurun_out/r02d_f5_0.1.ncu-rep
ACU_FILTER_MINB=5 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/r02d_bench_short.json 2> gpurun_out/r02d_bench_short.err
cut -c1-300 gpurun_out/r02d_bench_short.json; tail -3 gpurun_out/r02d_bench_short.err

#!/bin/bash
# filter kernel iteration: parity of the touched paths, A/B timings, ncu captures
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_select.py tests/test_gpu_golden.py tests/test_gpu_recordbatch.py tests/test_gpu_coalesce.py -q -m gpu -x -k "filter or nullif or zip or golden or record or coalesce") > gpurun_out/r02b_tests.log 2>&1
tail -4 gpurun_out/r02b_tests.log
(timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "filter or config1") > gpurun_out/r02b_tests_cfg.log 2>&1
tail -3 gpurun_out/r02b_tests_cfg.log
ACU_FILTER_LEGACY=1 timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' > gpurun_out/r02b_filter_legacy.txt 2>&1
timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' > gpurun_out/r02b_filter_fused6.txt 2>&1
ACU_FILTER_MINB=5 timeout 600 python tools/opbench.py --only "filter i64" | grep '^{' > gpurun_out/r02b_filter_fused5.txt 2>&1
for f in legacy fused6 fused5; do echo "== $f"; cut -c1-110 gpurun_out/r02b_filter_$f.txt; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_filter_fused|k_filter_values_async|k_compress_bits" -c 12 -f -o gpurun_out/r02b_filter_fused python tools/opbench.py --only "filter i64" --reps 1 > gpurun_out/r02b_ncu_fused.log 2>&1
ACU_FILTER_LEGACY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_filter_fused|k_filter_values_async|k_compress_bits" -c 24 -f -o gpurun_out/r02b_filter_legacy python tools/opbench.py --only "filter i64" --reps 1 > gpurun_out/r02b_ncu_legacy.log 2>&1
ls -la gpurun_out/*.ncu-rep

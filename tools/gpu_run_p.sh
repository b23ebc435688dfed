#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (one GPU)
mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0"
(timeout 1500 $CS python -m pytest tests/test_gpu_async.py tests/test_gpu_views.py tests/test_gpu_select.py tests/test_gpu_cmp_bytes.py tests/test_gpu_concat.py tests/test_gpu_ipc.py -q -m gpu -x) > gpurun_out/r02_sanitizer_a.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_a.log
tail -4 gpurun_out/r02_sanitizer_a.log
(timeout 1500 $CS python -m pytest tests/test_gpu_parity.py tests/test_gpu_recordbatch.py -q -m gpu -x -k "filter or bytes or take or record") > gpurun_out/r02_sanitizer_b.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_b.log
tail -4 gpurun_out/r02_sanitizer_b.log
(timeout 1200 $CS python -m pytest tests/test_gpu_dict.py -q -m gpu -x -k "not overflow") > gpurun_out/r02_sanitizer_c.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_c.log
tail -4 gpurun_out/r02_sanitizer_c.log
grep -h "ERROR SUMMARY" gpurun_out/r02_sanitizer_*.log

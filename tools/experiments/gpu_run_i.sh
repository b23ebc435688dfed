#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dict.py tests/test_gpu_parity.py -q -m gpu -x -k "dict or bytes or utf8 or string") 2>&1 | tail -4
(timeout 600 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "config4 or config5") 2>&1 | tail -3
timeout 300 python tools/opbench.py --only "dict" | grep '^{' | cut -c1-220
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dict_copy -c 1 -f -o gpurun_out/r02j_dict python tools/opbench.py --only "dict" --reps 1 > /dev/null 2>&1
ncu -i gpurun_out/r02j_dict.ncu-rep --page details > gpurun_out/r02j_dict.details.txt 2>&1
ncu -i gpurun_out/r02j_dict.ncu-rep --page source --csv > gpurun_out/r02j_dict.sass.csv 2>&1
rm -f gpurun_out/r02j_dict.ncu-rep
grep -E "Duration|Issue Slots Busy|No Eligible|Executed Instructions |Executed Ipc Active|Achieved Occupancy|bank conflicts" gpurun_out/r02j_dict.details.txt | head
(timeout 900 python -m pytest tests/test_gpu_async.py -q -m gpu -x) 2>&1 | tail -8

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r02v_ref.json 2> gpurun_out/r02v_ref.err
python -c "
import json
d=json.load(open('gpurun_out/r02v_ref.json')); c=d['cpu_baseline']
print(d['value'], c['cores'], c['thread_calibration_mrows_s'], c['spread'], c['seconds_median'], c.get('value_1_thread'))" || tail -5 gpurun_out/r02v_ref.err

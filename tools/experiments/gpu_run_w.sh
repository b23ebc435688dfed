#!/bin/bash
mkdir -p gpurun_out
free -g | head -3
for f in /sys/devices/system/node/node*/meminfo; do grep -E "MemTotal|MemFree" $f; done
for rows in 100000000 300000000 600000000 1000000000; do
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 --rows $rows --cpu-threads 64 > gpurun_out/r02w_ref.json 2> gpurun_out/r02w_ref.err
python -c "
import json
d=json.load(open('gpurun_out/r02w_ref.json')); c=d['cpu_baseline']
print($rows, round(d['value']), c['cores'], c['spread'], c['seconds_median'], c['seconds_min'], c['generate_seconds'])" || tail -3 gpurun_out/r02w_ref.err
done
grep -E "AnonHugePages|HugePages_Total|Hugepagesize" /proc/meminfo

#!/bin/bash
# usage: tools/gpu_profiles.sh rNN   (one gpurun call, one GPU) — the captures profiles/rNN_* are derived from
R=${1:-r02}
mkdir -p gpurun_out
python -c "import bench; print(bench.so_sha16())" > gpurun_out/so_sha16_$R.txt   # hash of the kernel sources (see bench.so_sha16)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/launches_$R.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_arith|k_take|k_filter_fused|k_reduce|k_plan_mask" -s 20 -c 8 -f -o gpurun_out/prof_$R python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/prof_$R.log 2>&1
ncu -i gpurun_out/prof_$R.ncu-rep --page details > gpurun_out/prof_${R}_details.txt 2>&1
ncu -i gpurun_out/prof_$R.ncu-rep --page raw --csv > gpurun_out/prof_${R}_raw.csv 2>/dev/null
rm -f gpurun_out/prof_$R.ncu-rep   # (tens of MB; the text / csv exports above are what profiles/ is built from)
timeout 600 ncu --set full --clock-control none -k regex:"k_dict_copy|k_dict_block_totals|k_cmp_v2|k_cast_v2" -c 6 -f -o gpurun_out/prof_${R}_ops python tools/opbench.py --only "dict|lt f64|cast i64" --reps 1 > gpurun_out/prof_${R}_ops.log 2>&1
ncu -i gpurun_out/prof_${R}_ops.ncu-rep --page details > gpurun_out/prof_${R}_ops_details.txt 2>&1
ncu -i gpurun_out/prof_${R}_ops.ncu-rep --page raw --csv > gpurun_out/prof_${R}_ops_raw.csv 2>/dev/null
rm -f gpurun_out/prof_${R}_ops.ncu-rep
timeout 900 python tools/opbench.py > gpurun_out/opbench_$R.log 2>&1
timeout 600 python tools/recordbatch_bench.py --streams 1 > gpurun_out/recordbatch_${R}_n1_1lane.json 2> /dev/null
timeout 600 python tools/recordbatch_bench.py --streams 4 > gpurun_out/recordbatch_${R}_n1_4lane.json 2> /dev/null
timeout 600 python tools/recordbatch_bench.py > gpurun_out/recordbatch_${R}_n1.json 2> gpurun_out/recordbatch_${R}_n1.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${R}_ref.json 2> gpurun_out/bench_${R}_ref.err
ls -la gpurun_out | head -30; tail -2 gpurun_out/prof_$R.log; cut -c1-200 gpurun_out/bench_${R}_n1.json

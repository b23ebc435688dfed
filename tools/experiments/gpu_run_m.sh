#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_recordbatch.py tests/test_gpu_coalesce.py tests/test_gpu_golden.py tests/test_gpu_dict.py tests/test_gpu_concat.py -q -m gpu -x) 2>&1 | tail -5
(timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_host_cpp.py -q -m gpu -x) 2>&1 | tail -3
for st in 1 3; do
ACU_BYTES_GENERIC_V1=1 timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 --streams $st > gpurun_out/r02m_rb_v1_s$st.json 2> gpurun_out/r02m_rb.err
timeout 600 python tools/recordbatch_bench.py --steps 3 --warmup 2 --streams $st > gpurun_out/r02m_rb_v2_s$st.json 2>> gpurun_out/r02m_rb.err
for v in v1 v2; do python -c "
import json
d=json.load(open('gpurun_out/r02m_rb_${v}_s$st.json'))
print('rb $v streams=$st', round(d['ms_per_step'],2), round(d['kernel_ms_per_step'],2), {k:round(x['ms_per_step'],2) for k,x in d['kernels'].items()}, d['check']['sums_bits'][:2])" || tail -5 gpurun_out/r02m_rb.err; done
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_gather_copy|k_bytes_block_totals|k_bytes_offsets_copy" -c 12 --csv --log-file gpurun_out/r02m_gather_launches.csv python tools/recordbatch_bench.py --steps 1 --warmup 0 --streams 1 --batches 2 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02m_gather_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
for r in rows[1:]: print(r[ki][:40], r[mi][:22], r[vi])
P
for k in 1 2; do
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02m_bench$k.json 2> gpurun_out/r02m_bench$k.err
python - $k <<'P'
import json,sys
d=json.load(open(f"gpurun_out/r02m_bench{sys.argv[1]}.json"))
c=d["configs"]["cfg5"]["filter_record_batch -> take_record_batch -> 6 sums"]
print("bench", d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], "cfg5", round(c["ms_per_step"],2), c["streams"], c["kernel_ms_by_class"])
P
done

#!/bin/bash
mkdir -p gpurun_out
(timeout 3000 python -m pytest tests -q -m gpu) > gpurun_out/r02_gputests_final.log 2>&1; tail -3 gpurun_out/r02_gputests_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_profiles.sh r02 > gpurun_out/r02_profiles_run.log 2>&1; tail -3 gpurun_out/r02_profiles_run.log

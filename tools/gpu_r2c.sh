#!/bin/bash
# Round 2, third call: copies-only vs full run_host, persistent-schedule tests, ncu (launch list + full capture of the
# sweep + DRAM bytes of the reference CUDA kernels).
set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "persistent" > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c.log
tail -5 gpurun_out/pytest_gpu_r2c.log
timeout -s KILL 300 python tools/gpu_e2e_sweep.py --quick > gpurun_out/e2e_sweep_r2c.txt 2>&1
cat gpurun_out/e2e_sweep_r2c.txt
timeout -s KILL 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cloudsc_c -c 4 --csv \
   --log-file gpurun_out/gpuref_traffic_r2c.csv python tools/gpu_ref_run.py > gpurun_out/gpuref_traffic_r2c.log 2>&1
cat gpurun_out/gpuref_traffic_r2c.csv | tail -8
bash tools/gpu_profile.sh r2c --no-e2e --no-gpu-reference > gpurun_out/profile_r2c.log 2>&1
tail -5 gpurun_out/profile_r2c.log

#!/bin/bash
# Round 2, second call: persistent-kernel schedule sweep, the tests that touch the changed kernels, run_host trim sweep.
set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -k "persistent or arena or full_size or fp32 or run_host or smoke" > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2b.log
tail -15 gpurun_out/pytest_gpu_r2b.log
timeout -s KILL 400 python tools/gpu_persist_sweep.py 163840 20 > gpurun_out/persist_sweep_r2b.txt 2>&1
cat gpurun_out/persist_sweep_r2b.txt
timeout -s KILL 300 python tools/gpu_persist_sweep.py 655360 10 > gpurun_out/persist_sweep_655k_r2b.txt 2>&1
head -12 gpurun_out/persist_sweep_655k_r2b.txt
timeout -s KILL 400 python tools/gpu_e2e_sweep.py > gpurun_out/e2e_sweep_r2b.txt 2>&1
cat gpurun_out/e2e_sweep_r2b.txt

#!/bin/bash
# round 2, call m: schedule mode 3 (staggered streaming): bit-identity test, sweep over (ksplit, nslots) at 163 840 and 655 360 columns
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent_schedule" 2>&1 | tail -5 | tee gpurun_out/r2m_test.txt
timeout 300 python tools/gpu_persist_sweep.py 163840 30 --stagger 2>&1 | tee gpurun_out/stagger_sweep_r2m.txt
timeout 300 python tools/gpu_persist_sweep.py 655360 8 --stagger 2>&1 | tee gpurun_out/stagger_sweep_655k_r2m.txt

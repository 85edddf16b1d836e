#!/bin/bash
# Round-2 closing run with the split-tail launch (schedule mode 3) as the default of an FP64 state:
# GPU suite, smoke, default bench, launch list + full ncu capture of the split kernel, mode A/B at two sizes.
set -x
TAG=r2_split
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout -s KILL 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
timeout -s KILL 100 python tools/gpu_persist_sweep.py 163840 20 --confirm > gpurun_out/split_confirm_163840.txt 2>&1
timeout -s KILL 300 bash tools/gpu_profile.sh $TAG --no-gpu-reference > gpurun_out/profile_$TAG.log 2>&1
timeout -s KILL 100 python tools/gpu_persist_sweep.py 655360 6 --confirm > gpurun_out/split_confirm_655360.txt 2>&1
timeout -s KILL 100 python tools/gpu_persist_sweep.py 120064 20 --confirm > gpurun_out/split_confirm_120064.txt 2>&1
tail -3 gpurun_out/smoke_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
cat gpurun_out/split_confirm_*.txt

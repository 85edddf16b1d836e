#!/bin/bash
# ncu evidence for the dominant kernel: launch list (time shares) + one full capture.
# usage: bash tools/gpu_profile.sh <tag> [extra bench args]
set -x
TAG=${1:-r1}; shift
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --no-cpu --e2e-steps 1 $*"
python bench.py $ARGS > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
python bench.py $ARGS > gpurun_out/plain2_$TAG.json 2>> gpurun_out/plain_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:cloudsc_sweep -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out; cat gpurun_out/plain_$TAG.json; tail -3 gpurun_out/ncu_full_$TAG.log

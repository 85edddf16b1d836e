#!/bin/bash
# Round-end evidence: smoke, GPU parity tests, the default bench (both arms), launch list + full ncu capture.
# usage: bash tools/gpu_round.sh <tag>
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout -s KILL 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
bash tools/gpu_profile.sh $TAG --no-e2e
tail -3 gpurun_out/smoke_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/bench_ref_$TAG.json gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err

#!/bin/bash
# Last GPU call of round 2: (1) the bench line of the committed tree once more (short), (2) does the split-tail launch pay
# below two rounds of tiles?  (CLOUDSC_B200_SPLIT_MIN_TILES=1 lets every launch be split.)
set -x
mkdir -p gpurun_out
timeout -s KILL 90 python bench.py --steps 20 --warmup 3 --no-cpu --no-gpu-reference --e2e-steps 1 > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo "bench rc=$?" >> gpurun_out/bench_r2n.err
export CLOUDSC_B200_SPLIT_MIN_TILES=1
for n in 65536 90112 57344 32768; do
  timeout -s KILL 60 python tools/gpu_persist_sweep.py $n 20 --small > gpurun_out/split_small_$n.txt 2>&1
done
tail -2 gpurun_out/bench_r2n.err; cat gpurun_out/split_small_*.txt

#!/bin/bash
# Round 2, first contact: environment record, smoke, the whole GPU test suite, both bench arms, the PCIe micro-benchmark
# (three kinds of pinned memory), the run_host chunk sweep and the `reload` alt build.  One GPU.
set -x
mkdir -p gpurun_out
{
  nvidia-smi; nvidia-smi topo -m; lscpu | head -30; (numactl -H 2>&1 || true); free -g
  cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag 2>&1
  grep -i huge /proc/meminfo; ulimit -l
} > gpurun_out/r2_env.txt 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2a.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
tail -40 gpurun_out/pytest_gpu_r2a.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?" >> gpurun_out/bench_r2a.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err
( cd tools/ubench && ./pcie --gpus 1 --mode all --mb 2048 --reps 3; ./pcie --gpus 1 --mode all --mb 2048 --reps 3 --chunk-mb 8; ./pcie --gpus 1 --mode hostalloc --mb 2048 --reps 3 --chunk-mb 1 ) > gpurun_out/pcie_r2a.txt 2>&1
timeout -s KILL 300 python tools/gpu_e2e_sweep.py > gpurun_out/e2e_sweep_r2a.txt 2>&1
bash tools/gpu_alt.sh default reload
cat gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err; cat gpurun_out/bench_ref_r2a.json; cat gpurun_out/pcie_r2a.txt; cat gpurun_out/e2e_sweep_r2a.txt; tail -3 gpurun_out/smoke_r2a.log

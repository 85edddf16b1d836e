#!/bin/bash
# Round 2: full GPU suite on the refactored tree, both bench arms, BASELINE configs[0] on the box's host cores,
# configs[3]/[4] sweeps.
set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2h.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2h.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_r2h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2h.log
tail -25 gpurun_out/pytest_gpu_r2h.log
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2h.json 2> gpurun_out/bench_ref_r2h.err
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?" >> gpurun_out/bench_r2h.err
cat gpurun_out/bench_r2h.json | cut -c1-600; tail -4 gpurun_out/bench_r2h.err
{
  NC=$(nproc)
  echo "# BASELINE configs[0]: the reference's own dwarf-cloudsc-c (oracle/_ref/dwarf-cloudsc-c*: unmodified driver, kernel, validator;"
  echo "# host/load_state.c + host/h5lite.c in place of load_state.c + libhdf5) on the GPU box's host: $NC cores"
  lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
  export OMP_PLACES=cores OMP_PROC_BIND=close OMP_SCHEDULE=static OMP_STACKSIZE=1G
  cd tests/golden
  for exe in dwarf-cloudsc-c dwarf-cloudsc-c-O3v4; do
    for n in 16384 163840; do
      echo; echo "## OMP_NUM_THREADS=$NC $exe $NC $n 32"
      OMP_NUM_THREADS=$NC timeout -s KILL 300 ../../oracle/_ref/$exe $NC $n 32 | grep -vE "^ +$NC +$n +[0-9]+ +[0-9]+ +32 +[0-9]+ :"
    done
  done
} > gpurun_out/config0_cpu_r2h.txt 2>&1
grep -E "##|TOTAL" gpurun_out/config0_cpu_r2h.txt
bash tools/gpu_sweep.sh > gpurun_out/sweep_r2h.log 2>&1
tail -14 gpurun_out/sweep_r2h.log

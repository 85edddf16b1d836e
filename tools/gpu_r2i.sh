#!/bin/bash
# configs[0] on the box host + configs[3]/[4] sweeps (second half of tools/gpu_r2h.sh)
set -x
mkdir -p gpurun_out
(
  set +x
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
) > gpurun_out/config0_cpu_r2h.txt 2>&1
grep -E "##|TOTAL" gpurun_out/config0_cpu_r2h.txt
bash tools/gpu_sweep.sh > gpurun_out/sweep_r2h.log 2>&1
tail -14 gpurun_out/sweep_r2h.log

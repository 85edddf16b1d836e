#!/bin/bash
# round 2, call l: (1) which property of the run_host pipeline costs PCIe bandwidth (pcie.cu --pattern 2), (2) kernel variants
# (integer max(x,0), table log, early deposition, limiter trims), (3) exact-wave problem sizes
mkdir -p gpurun_out
{
for k in 0 1 2 4 8 16 31; do ./tools/ubench/pcie --pattern 2 --mode hostalloc --knobs $k --nin 28 --nout 25 | tail -2; done
./tools/ubench/pcie --pattern 2 --mode hostalloc --knobs 0 --nin 28 --nout 28 | tail -1
./tools/ubench/pcie --pattern 2 --mode hostalloc --knobs 31 --nin 28 --nout 28 | tail -1
} > gpurun_out/pcie_pattern2_r2l.txt 2>&1
cat gpurun_out/pcie_pattern2_r2l.txt
bash tools/gpu_alt.sh default base trim mx0 logtab dep1 dep2 base 2>&1 | tee gpurun_out/alt_r2l.txt
for n in 56832 113664 170496 227328; do
  ALT_ARGS="--ngptot $n" bash tools/gpu_alt.sh default 2>&1 | sed "s/^/ngptot $n: /" | tee -a gpurun_out/alt_r2l.txt
done

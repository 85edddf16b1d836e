#!/bin/bash
# Round 2, 8-GPU call (gpurun --gpus 8): where does the end-to-end path stop scaling?  Host topology record, the copy
# micro-benchmark on 2/4/8 GPUs at once (three kinds of pinned memory, with and without core binding), bench.py on 8
# GPUs with the default host arrays and with huge pages + core binding, the C driver on 8 GPUs.
set -x
N=${1:-8}
mkdir -p gpurun_out
{
  nvidia-smi -L; nvidia-smi topo -m; lscpu | head -40; (numactl -H 2>&1 || true); free -g; nproc
  cat /sys/kernel/mm/transparent_hugepage/enabled 2>&1; grep -i -E "huge|MemTotal" /proc/meminfo
  for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist) $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
} > gpurun_out/r2_env_n$N.txt 2>&1
( cd tools/ubench
  for g in 1 2 4 $N; do ./pcie --gpus $g --mode all --mb 1024 --reps 3; done
  ./pcie --gpus $N --mode hostalloc --mb 1024 --reps 3 --bind 1
  ./pcie --gpus $N --mode thp --mb 1024 --reps 3 --bind 1
  ./pcie --gpus $N --mode hostalloc --mb 1024 --reps 3 --chunk-mb 8
) > gpurun_out/pcie_concurrency_n$N.txt 2>&1
cat gpurun_out/pcie_concurrency_n$N.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $N --steps 10 --warmup 3 --pin thp --bind 1 > gpurun_out/scale_thp_bind_n$N.json 2> gpurun_out/scale_thp_bind_n$N.err
( cd tests/golden && timeout -s KILL 600 ../../dwarf_p_cloudsc_b200/lib/dwarf-cloudsc-b200 $N $((163840*N)) 128 ) > gpurun_out/scale_cdriver_n$N.txt 2>&1
python - <<PY
import json
for f in ("gpurun_out/scale_n$N.json", "gpurun_out/scale_thp_bind_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], "ms", d["ms_per_step"], "Mcol/s", d["value"]/1e6, "validated", d["validated_vs_reference_h5"], d["validated_columns"], "e2e", json.dumps(d["e2e"]))
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json", ".err")).read()[-2500:])
PY
head -14 gpurun_out/scale_cdriver_n$N.txt; grep -c PASS gpurun_out/scale_cdriver_n$N.txt

#!/bin/bash
# weak-scaling check on N GPUs of one box: bench at N=1 and N, the C driver on N GPUs.  usage: bash tools/gpu_scale.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/scale_gpus_$N.txt
timeout -s KILL 600 python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu > gpurun_out/scale_n1_of$N.json 2> gpurun_out/scale_n1_of$N.err
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 50 --warmup 3 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
( cd tests/golden && timeout -s KILL 600 ../../dwarf_p_cloudsc_b200/lib/dwarf-cloudsc-b200 $N $((163840*N)) 128 ) > gpurun_out/scale_cdriver_n$N.txt 2>&1
python - <<PY
import json
for f in ("gpurun_out/scale_n1_of$N.json", "gpurun_out/scale_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], "ms", d["ms_per_step"], "Mcol/s", d["value"]/1e6, "e2e", d["e2e"] and d["e2e"]["value"]/1e6)
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
head -12 gpurun_out/scale_cdriver_n$N.txt; grep -c PASS gpurun_out/scale_cdriver_n$N.txt

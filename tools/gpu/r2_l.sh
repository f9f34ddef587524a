#!/bin/bash
# round 2, GPU call L (2 GPUs): bench through the engine's NCCL data plane, with the atom-by-atom check against one GPU
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/l_bench_${N}gpu.json 2> gpurun_out/l_bench_${N}gpu.err
echo "rc=$?"; tail -5 gpurun_out/l_bench_${N}gpu.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/l_bench_%sgpu.json"%n).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["kernels_ms_per_step"])
    print(d.get("parity_vs_n1"))
except Exception as e:
    print("unreadable", e)
PY

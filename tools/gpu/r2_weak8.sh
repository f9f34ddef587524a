#!/bin/bash
# round 2, supplementary: config C5 (32 M atoms, 256^3 grid) on 8 GPUs = weak scaling of 4 M atoms + 128^3 grid cells per GPU
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --weak --grid 128 --no-e2e --no-check > gpurun_out/fW_weak_32M_8gpu.json 2> gpurun_out/fW_weak_32M_8gpu.err
echo "weak rc=$?"; grep -v OMP_NUM gpurun_out/fW_weak_32M_8gpu.err | grep -v '^\*\*\*' | tail -2
python - <<'PY'
import json
d=json.loads(open("gpurun_out/fW_weak_32M_8gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["config"]["atoms"], d["config"]["fdm_grid"], d["scaling"])
print("    ", d["roofline"]["kernels_ms_per_step"])
PY

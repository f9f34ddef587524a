#!/bin/bash
# round 2: the committed tree on 2 GPUs (step_breakdown in the bench line, collective set_ghost_map over real NCCL)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --no-e2e > gpurun_out/l2_bench.json 2> gpurun_out/l2_bench.err
echo "rc=$?"; grep -v OMP_NUM gpurun_out/l2_bench.err | grep -v '^\*\*\*' | tail -2
python - <<'PY'
import json
d=json.loads(open("gpurun_out/l2_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], (d.get("parity_vs_n1") or {}).get("max_rel_dev"))
print(d["step_breakdown"]); print(d["roofline"]["kernels_ms_per_step"])
PY

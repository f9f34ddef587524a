#!/bin/bash
# round 2, last 8-GPU call: the committed tree once more, with the integrator and ghost-refresh kernels in the per-kernel breakdown
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-e2e > gpurun_out/l8_bench.json 2> gpurun_out/l8_bench.err
echo "rc=$?"; grep -v OMP_NUM gpurun_out/l8_bench.err | grep -v '^\*\*\*' | tail -2
python - <<'PY'
import json
d=json.loads(open("gpurun_out/l8_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], (d.get("parity_vs_n1") or {}).get("max_rel_dev"))
k=d["roofline"]["kernels_ms_per_step"]; print(k); print("sum", sum(v for n,v in k.items() if n not in ("source_allreduce","fdm_substep")))
PY

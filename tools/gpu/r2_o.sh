#!/bin/bash
# round 2, GPU call O (8 GPUs): overlap of the ghost exchange, static mode, sharded against replicated grid solve at 256^3
mkdir -p gpurun_out
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e "$@" > gpurun_out/o_$tag.json 2> gpurun_out/o_$tag.err; echo "$tag rc=$?"; }
run overlap --overlap --no-check
run static --mode static --no-check
run grid256_replicated --grid 256 --no-check
run grid256_sharded --grid 256 --sharded-grid
python - <<'PY'
import json
for n in ("overlap", "static", "grid256_replicated", "grid256_sharded"):
    try:
        d=json.loads(open("gpurun_out/o_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d.get("parity_vs_n1") and d["parity_vs_n1"]["max_rel_dev"])
    except Exception as e:
        print(n, "unreadable", e)
PY

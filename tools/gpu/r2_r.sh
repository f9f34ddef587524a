#!/bin/bash
# round 2, GPU call R (8 GPUs): peer-memory exchange against NCCL send/receive, static mode, sharded 256^3 grid
mkdir -p gpurun_out
run() { # name, env, extra args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e $3 > gpurun_out/r_$1.json 2> gpurun_out/r_$1.err
  echo "$1 rc=$?"; grep -v OMP_NUM gpurun_out/r_$1.err | tail -2
}
run p2p "A=1" ""
run nccl "EPH_B200_EXCHANGE=nccl" "--no-check"
run static "A=1" "--mode static --no-check"
run g256_rep "A=1" "--grid 256 --no-check"
run g256_shard "A=1" "--grid 256 --sharded-grid"
python - <<'PY'
import json
for f in ("p2p","nccl","static","g256_rep","g256_shard"):
    try:
        d=json.loads(open("gpurun_out/r_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["config"].get("ghost_exchange_transport"), (d.get("parity_vs_n1") or {}).get("max_rel_dev"))
        print("    ", d["roofline"]["kernels_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY

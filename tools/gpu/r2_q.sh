#!/bin/bash
# round 2, GPU call Q (2 GPUs): peer-memory ghost exchange against NCCL send/receive, with the atom-by-atom check
mkdir -p gpurun_out
run() { # name, env, extra args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e $3 > gpurun_out/q_$1.json 2> gpurun_out/q_$1.err
  echo "$1 rc=$?"; tail -2 gpurun_out/q_$1.err
}
run p2p "A=1" ""
run nccl "EPH_B200_EXCHANGE=nccl" ""
run p2p_cells50 "A=1" "--cells 50 --no-check"
run nccl_cells50 "EPH_B200_EXCHANGE=nccl" "--cells 50 --no-check"
run p2p_overlap "A=1" "--overlap"
python - <<'PY'
import json
for f in ("p2p","nccl","p2p_cells50","nccl_cells50","p2p_overlap"):
    try:
        d=json.loads(open("gpurun_out/q_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["config"].get("ghost_exchange_transport"), (d.get("parity_vs_n1") or {}).get("max_rel_dev"))
        print("    ", d["roofline"]["kernels_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY

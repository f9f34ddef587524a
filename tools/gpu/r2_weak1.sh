#!/bin/bash
# round 2, supplementary: the one-GPU leg of the weak-scaling pair (4 M atoms, 128^3 grid) + ncu summaries of the list-maintenance kernels
mkdir -p gpurun_out
python bench.py --grid 128 --no-e2e --no-extras --no-cpu-baseline --no-fdm-bench > gpurun_out/fW_4M_grid128_1gpu.json 2> gpurun_out/fW_4M_grid128_1gpu.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/fW_4M_grid128_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["config"]["atoms"], d["config"]["fdm_grid"])
PY
ncu --set full --clock-control none --import-source on -k regex:'inner_build_kernel|neighbor_tile_kernel' -c 3 -o gpurun_out/fW_lists python bench.py --neigh device --steps 12 --warmup 1 --no-e2e --no-extras --no-cpu-baseline --no-fdm-bench > gpurun_out/fW_ncu_lists.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/fW_ncu_lists.log
ncu --set full --clock-control none --import-source on -k regex:'inner_build_kernel' -c 1 -o gpurun_out/fW_inner python bench.py --steps 12 --warmup 1 --no-e2e --no-extras --no-cpu-baseline --no-fdm-bench > gpurun_out/fW_ncu_inner.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/fW_ncu_inner.log
ls -la gpurun_out/fW_*.ncu-rep

#!/bin/bash
# round 2, GPU call C: packed sweeps v2 (lean decode, padded tiles) at 1 / 2 / 4 lanes per atom + ncu of the best guess
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench"
for L in 1 2 4; do
  EPH_B200_LANES=$L python bench.py $B > gpurun_out/c_l$L.json 2> gpurun_out/c_l$L.err
done
EPH_B200_LANES=2 ncu --set full --clock-control none --import-source on -k regex:"density_packed|force_packed" -s 2 -c 2 -o gpurun_out/c_packed_l2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench > gpurun_out/c_ncu.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/c_pytest.log
for f in gpurun_out/c_l*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"])
except Exception as e:
    print("unreadable", e)
PY
done

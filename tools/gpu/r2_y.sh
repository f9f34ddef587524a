#!/bin/bash
# round 2, GPU call Y (1 GPU): list construction with eight lanes per atom against the cooperative row kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_late_additions.py -m gpu -x -q -k "neigh or list or 131k or build or device or resident" > gpurun_out/y_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/y_pytest.log
for v in group tile; do
  EPH_B200_NEIGH_KERNEL=$v python bench.py --neigh device --no-extras --no-cpu-baseline --no-fdm-bench > gpurun_out/y_$v.json 2> gpurun_out/y_$v.err
  python - $v <<'PY'
import json,sys
d=json.loads(open("gpurun_out/y_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["ms_per_step"], d["roofline"]["kernels_ms"].get("build_neighbors"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done

#!/bin/bash
# round 2, GPU call A: parity suite + packed-record sweeps A/B + one ncu capture
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench"
python bench.py $B > gpurun_out/a_bench_packed.json 2> gpurun_out/a_bench_packed.err
EPH_B200_RECORDS=exact python bench.py $B > gpurun_out/a_bench_exact.json 2> gpurun_out/a_bench_exact.err
EPH_B200_LANES=2 python bench.py $B > gpurun_out/a_bench_packed_l2.json 2> gpurun_out/a_bench_packed_l2.err
EPH_B200_LANES=8 python bench.py $B > gpurun_out/a_bench_packed_l8.json 2> gpurun_out/a_bench_packed_l8.err
ncu --set full --clock-control none --import-source on -k regex:"density_packed|force_packed" -s 2 -c 2 -o gpurun_out/a_packed \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench > gpurun_out/a_ncu.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?"
tail -3 gpurun_out/a_pytest.log
for f in gpurun_out/a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"])
except Exception as e:
    print("unreadable", e)
PY
done

#!/bin/bash
# round 2, GPU call J: parity suite, the default bench line (trajectory mode), static mode for comparison
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/j_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench_default.json 2> gpurun_out/j_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/j_bench_default.err
python bench.py --steps 20 --warmup 5 --mode static --no-cpu-baseline --no-e2e --no-fdm-bench --no-extras > gpurun_out/j_bench_static.json 2> gpurun_out/j_bench_static.err
python - <<'PY'
import json
for n in ("default", "static"):
    try:
        d=json.loads(open("gpurun_out/j_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d.get("e2e",{}) and d["e2e"].get("value"))
        print({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ("value","max_rel_dev","ms_per_step")}) for k,v in (d.get("extras") or {}).items()})
        print(d.get("fdm") and {k: d["fdm"][k]["value"] for k in ("general_path","constant_coefficient_path")}, d.get("cpu_baseline"))
    except Exception as e:
        print(n, "unreadable", e)
PY

#!/bin/bash
# round 2, GPU call M: parity suite, default bench line, ncu --set full of the two packed sweeps, launch list of the default command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/m_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/m_bench_default.json 2> gpurun_out/m_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/m_bench_default.err
ncu --set full --clock-control none --import-source on -k regex:"density_packed|force_packed" -s 4 -c 2 -o gpurun_out/m_packed \
    python bench.py --steps 3 --warmup 5 --no-cpu-baseline --no-e2e --no-fdm-bench --no-extras > gpurun_out/m_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/m_launches.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench --no-extras > gpurun_out/m_launches.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/m_bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step"]["frac"], d["roofline"]["kernels_ms_per_step"])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], e.get("mode"), "host mode", e.get("host_mode",{}).get("value"))
print(d["extras"]["C4_NiCoCrFe"]["value"], d["extras"]["parity_vs_reference"]["max_rel_dev"], d["cpu_baseline"]["value"])
PY

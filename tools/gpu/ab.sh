#!/bin/bash
# development aid: A/B timing of engine builds (tools/gpu/variants/libeph_b200_<name>.so, see build_variants.sh) on one box
#   tools/gpu/ab.sh <tag> "<names>" ["<lanes>"]
tag=$1; names=$2; lanes=${3:-"2 4"}
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench --no-extras"
for v in $names; do
  for L in $lanes; do
    EPH_B200_LANES=$L EPH_B200_ENGINE_LIB=$PWD/tools/gpu/variants/libeph_b200_$v.so python bench.py $B > gpurun_out/${tag}_${v}_l$L.json 2> gpurun_out/${tag}_${v}_l$L.err
  done
done
for f in gpurun_out/${tag}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels_ms"]
    print(d["ms_per_step"], "dens", k["density_sweep"], "force", k["force_sweep"], "inner build", k.get("inner_list_build"))
except Exception as e:
    print("unreadable", e)
PY
done

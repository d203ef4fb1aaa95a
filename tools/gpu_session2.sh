#!/bin/bash
# per-kernel efficiency at saturating batch sizes, per variant
TAG=${1:-s2}; VARIANTS=${2:-default}; BATCHES=${3:-"4096 16384"}
mkdir -p gpurun_out
for v in $VARIANTS; do
  for b in $BATCHES; do
    if [ "$v" = default ]; then unset RABE_B200_LIB; else export RABE_B200_LIB=$PWD/build/variants/$v.so; fi
    timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/${TAG}_bench_${v}_b${b}.json 2> gpurun_out/${TAG}_bench_${v}_b${b}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${v}_b${b}.json"))
    pk = d["roofline"]["per_kernel"]; peak = d["roofline"]["peak"]
    print("$v B=$b value=%.0f step_frac=%.3f" % (d["value"], d["roofline"]["step_frac"]),
          " ".join("%s=%.2fms(%.2f)" % (k.replace("k_ac17_", "").replace("k_", ""), v_["ms"], v_["gfpmul_s"] / peak) for k, v_ in pk.items()))
except Exception as ex:
    print("$v B=$b FAILED", ex)
PY
  done
done

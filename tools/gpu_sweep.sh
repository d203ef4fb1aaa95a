#!/bin/bash
# gpu_sweep.sh TAG "variants" [bench args...]: one bench line summary per library variant
TAG=$1; VARIANTS=$2; shift 2
mkdir -p gpurun_out
for v in $VARIANTS; do
  if [ "$v" = default ]; then unset RABE_B200_LIB; else export RABE_B200_LIB=$PWD/build/variants/$v.so; fi
  timeout 400 python bench.py --steps 24 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${v}.json 2> gpurun_out/${TAG}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${v}.json"))
    pk = d["roofline"]["per_kernel"]
    print("$v value=%.0f e2e=%.0f serial=%.0f step_frac=%.3f" % (d["value"], d["e2e"]["value"], d["config"]["serial_roundtrips_per_s"], d["roofline"]["step_frac"]),
          " ".join("%s=%.2f" % (k.replace("k_ac17_", "").replace("k_", ""), v_["ms"]) for k, v_ in pk.items()))
except Exception as ex:
    print("$v FAILED", ex)
PY
done

#!/bin/bash
# bench.py under environment / flag variants:  bash tools/gpu_sweep3.sh "NAME|ENV=VAL ...|--flags" ...
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%|*}"; rest="${spec#*|}"; envs="${rest%%|*}"; flags="${rest#*|}"
  env $envs python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-parity-check --no-other-configs $flags > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  python - "$name" <<PY
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("gpurun_out/sweep_%s.json" % v)); r = d["roofline"]
    print(v, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "serial enc/dec ms", round(d["details"]["serial_enc_ms"], 3), round(d["details"]["serial_dec_ms"], 3),
          {k: round(x["ms"], 3) for k, x in r["per_kernel"].items() if "dec" in k or "final" in k}, "step_frac", round(r["step_frac"], 3))
except Exception as e:
    print(v, "ERR", e, open("gpurun_out/sweep_%s.err" % v).read()[-400:])
PY
done

#!/bin/bash
# Development aid: one gpurun call = sanity tests + instruction-rate probe + kernel-variant sweep.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh TAG "variants..." "dec-streams..."'
TAG=${1:-s}; VARIANTS=${2:-default}; STREAMS=${3:-4}; TESTS=${4:-"tests/test_gpu_ac17.py tests/test_gpu_primitives.py"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
[ -x build/pipe_probe ] && ./build/pipe_probe > gpurun_out/${TAG}_pipe_probe.txt 2>&1
if [ -n "$TESTS" ]; then timeout 1200 python -m pytest $TESTS -m gpu -x -q > gpurun_out/${TAG}_tests.txt 2>&1; tail -3 gpurun_out/${TAG}_tests.txt; fi
for v in $VARIANTS; do
  for ds in $STREAMS; do
    if [ "$v" = default ]; then unset RABE_B200_LIB; else export RABE_B200_LIB=$PWD/build/variants/$v.so; fi
    timeout 400 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --dec-streams $ds > gpurun_out/${TAG}_bench_${v}_ds${ds}.json 2> gpurun_out/${TAG}_bench_${v}_ds${ds}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${v}_ds${ds}.json"))
    pk = d["roofline"]["per_kernel"]
    print("$v ds=$ds value=%.0f e2e=%.0f serial=%.0f step_frac=%.3f peak=%.1f" % (d["value"], d["e2e"]["value"], d["config"]["serial_roundtrips_per_s"], d["roofline"]["step_frac"], d["roofline"]["peak"]),
          " ".join("%s=%.2fms" % (k.replace("k_ac17_", "").replace("k_", ""), v_["ms"]) for k, v_ in pk.items()))
except Exception as ex:
    print("$v ds=$ds FAILED", ex)
PY
  done
done

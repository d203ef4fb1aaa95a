#!/bin/bash
# Multi-GPU evidence of the round:  gpurun --gpus N -- 'bash tools/gpu_multi_check.sh N TAG'
#   default bench line (config 2, weak scaling) and the sharded configurations 4 and 5 (strong scaling) under torchrun
N=${1:-2}; TAG=${2:-r2m}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 3 --no-other-configs > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
for C in 3 4 5; do
  timeout 900 $TR bench.py --gpus $N --config $C --steps 3 --warmup 3 > gpurun_out/${TAG}_cfg${C}_${N}gpu.json 2> gpurun_out/${TAG}_cfg${C}_${N}gpu.err
done
timeout 600 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_reference_${N}gpu.json 2>> gpurun_out/${TAG}_bench_${N}gpu.err
for f in gpurun_out/${TAG}_*_${N}gpu.json; do echo "== $f"; head -c 600 $f; echo; done

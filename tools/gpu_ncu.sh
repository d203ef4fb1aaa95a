#!/bin/bash
# ncu --set full of selected kernels for one library variant; only the raw-page CSV travels back
# (the .ncu-rep of these kernels is > 50 MB):  gpu_ncu.sh TAG VARIANT KREGEX BATCH SKIP
TAG=$1; v=$2; K=$3; B=${4:-4096}; SKIP=${5:-4}
if [ "$v" = default ]; then unset RABE_B200_LIB; else export RABE_B200_LIB=$PWD/build/variants/$v.so; fi
ncu --set full --clock-control none -k regex:"$K" -s $SKIP -c 2 -o /tmp/$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/$TAG.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
tail -c 200 gpurun_out/$TAG.log

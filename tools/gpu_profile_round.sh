#!/bin/bash
# Round-end evidence: launch list + ncu --set full of the main kernels of `python bench.py` (B = 4096).
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none -k regex:"k_ac17_dec_miller_pair_co|k_final_exp_co|k_ac17_enc_rows|k_ac17_enc_c0|k_ac17_enc_cp|k_g1_gather_sum" -s 14 -c 7 -o /tmp/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
tail -c 300 gpurun_out/${TAG}_full.log

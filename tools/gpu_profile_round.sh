#!/bin/bash
# Round evidence at HEAD, one command:  gpurun -- 'bash tools/gpu_profile_round.sh r2'
#   1. launch list (gpu__time_duration per launch) of `python bench.py --steps 2 --warmup 3`     -> gpurun_out/${TAG}_launches.csv
#   2. ncu --set full of the AC17 kernels of that command at the shipped default (26/16/16)       -> gpurun_out/${TAG}_full_raw.csv
#   3. ncu --set full of the per-leaf kernels of BSW / LSW / AW11 (tools/bench_schemes.py)         -> gpurun_out/${TAG}_schemes_raw.csv
# Then, on the CPU box:  python tools/ncu_summary.py gpurun_out/${TAG}_full_raw.csv --traffic-json profiles/${TAG}_ncu_traffic.json 4096
#                        > profiles/${TAG}_ncu_full_summary.txt      (bench.py reads roofline.traffic from that JSON)
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 240 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-other-configs > gpurun_out/${TAG}_launches.log 2>&1
K="k_ac17_dec_miller|k_final_exp|k_ac17_enc_rows|k_ac17_enc_c0|k_ac17_enc_cp|k_g1_gather_sum|k_pair_|k_fexp_"
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 14 -c 7 -o /tmp/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-other-configs > gpurun_out/${TAG}_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
KS="k_leaf_pair|k_leaf_fixed4|k_gt_pow_fixed|k_miller_co|k_g2_mul_fixed|k_gt_pow_var|k_g2_subgroup_check|k_leaf_"
ncu --set full --clock-control none -k regex:"$KS" -c 12 -o /tmp/${TAG}_schemes \
    python tools/bench_schemes.py --scale 0.125 > gpurun_out/${TAG}_schemes.log 2>&1
ncu -i /tmp/${TAG}_schemes.ncu-rep --page raw --csv > gpurun_out/${TAG}_schemes_raw.csv 2>/dev/null
tail -c 300 gpurun_out/${TAG}_full.log

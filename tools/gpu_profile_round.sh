#!/bin/bash
# Round evidence at HEAD, one command:  gpurun -- 'bash tools/gpu_profile_round.sh r2'
#   1. launch list (gpu__time_duration per launch) of `python bench.py --steps 2 --warmup 3`           -> gpurun_out/${TAG}_launches.csv
#   2. ncu --set full of the AC17 kernels of that command at the shipped default (26/16/16 windows), once per pairing
#      layout (RABE_B200_PAIRING = co: two-lane throughput kernels, w6: six-lane latency kernels)       -> gpurun_out/${TAG}_full_{co,w6}_raw.csv
#   3. ncu --set full of the per-leaf kernels of BSW / LSW / AW11 (tools/bench_schemes.py)               -> gpurun_out/${TAG}_schemes_raw.csv, ${TAG}_schemes_b_raw.csv
# Then, on the CPU box:
#   python tools/ncu_summary.py gpurun_out/${TAG}_full_co_raw.csv gpurun_out/${TAG}_full_w6_raw.csv --traffic-json profiles/r2_ncu_traffic.json 4096 > profiles/${TAG}_ncu_full_summary.txt
# (bench.py reads roofline.traffic from that JSON)
TAG=${1:-r2}
FLAGS="--no-cpu-baseline --no-parity-check --no-other-configs --no-table-budget"
mkdir -p gpurun_out
# (ncu serialises the launches, so the AUTO layout policy would see a lone batch everywhere and pick the six-lane kernels; the
#  pipelined timed region runs the two-lane ones -- the list is taken with that layout pinned)
RABE_B200_PAIRING=co ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 260 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 $FLAGS > gpurun_out/${TAG}_launches.log 2>&1
K="k_ac17_dec_miller|k_ac17_dec_item|k_final_exp|k_ac17_enc_rows|k_ac17_enc_c0|k_ac17_enc_cp|k_g1_gather_sum"
for L in co w6; do
  RABE_B200_PAIRING=$L ncu --set full --clock-control none --import-source on -k regex:"$K" -s 7 -c 14 -o /tmp/${TAG}_full_$L \
      python bench.py --steps 1 --warmup 3 $FLAGS > gpurun_out/${TAG}_full_$L.log 2>&1
  ncu -i /tmp/${TAG}_full_$L.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_${L}_raw.csv 2>/dev/null
done
# per-leaf kernels of BSW / LSW decrypt (the setup's table kernels would otherwise use up the launch count), then the fixed-base ones
ncu --set full --clock-control none -k regex:"k_leaf_pair_co|k_leaf_fixed4_co" -c 6 -o /tmp/${TAG}_schemes_a \
    python tools/bench_schemes.py --scale 0.125 > gpurun_out/${TAG}_schemes.log 2>&1
ncu --set full --clock-control none -k regex:"k_gt_pow_fixed|k_g2_mul_fixed|k_g1_mul_fixed|k_gt_pow_var" -s 12 -c 8 -o /tmp/${TAG}_schemes_b \
    python tools/bench_schemes.py --scale 0.125 >> gpurun_out/${TAG}_schemes.log 2>&1
ncu -i /tmp/${TAG}_schemes_a.ncu-rep --page raw --csv > gpurun_out/${TAG}_schemes_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_schemes_b.ncu-rep --page raw --csv > gpurun_out/${TAG}_schemes_b_raw.csv 2>/dev/null
tail -c 300 gpurun_out/${TAG}_full_w6.log

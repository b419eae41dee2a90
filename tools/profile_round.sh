#!/bin/bash
# ncu evidence for profiles/ (one B200, run under gpurun).  usage: bash tools/profile_round.sh r2a
#  1. launch list of the bench command (per-kernel shares of a step; cold-cache and serialised: compare shares)
#  2. steady-state full captures (--cache-control none: ncu does not flush L2 between replay passes, the loop
#     was warmed up before): the two timestep kernels at config 2 (state 54 MB, L2 resident) and at 32 x n=160
#     (state 211 MB: streams from HBM every timestep), and the persistent fused kernel at config 2
tag=${1:-rX}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 --no-config4 > gpurun_out/bench_under_ncu_${tag}.log 2>&1
python tools/ncu_by_kernel.py gpurun_out/launches_${tag}.csv > gpurun_out/launches_${tag}_by_kernel.txt 2>&1
# warm-up = init (4 launches incl. degree) + 2 x 64 timestep launches; tc_ kernels before the profiled pass:
# tc_edge_init, tc_pack_state, tc_zero_c, tc_degree (4) + 128 -> skip 140 lands inside the profiled pass
for cfg in 2 n160; do
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:"tc_lnlstm|tc_mlp" -s 140 -c 2 -f \
      -o gpurun_out/prof_${tag}_${cfg} python tools/ncu_target.py $cfg bf16x3 > gpurun_out/ncu_target_${tag}_${cfg}.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${tag}_${cfg}.ncu-rep 12 > gpurun_out/ncu_summary_${tag}_${cfg}.txt 2>&1
done
ncu --set full --clock-control none --cache-control none --import-source on -k regex:tc_step -s 2 -c 1 -f \
    -o gpurun_out/prof_${tag}_fused python tools/ncu_target.py 2 bf16x3 fused > gpurun_out/ncu_target_${tag}_fused.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_${tag}_fused.ncu-rep 12 > gpurun_out/ncu_summary_${tag}_fused.txt 2>&1

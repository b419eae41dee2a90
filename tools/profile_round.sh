#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the two timestep kernels.
# usage: bash tools/profile_round.sh r1i
tag=${1:-rX}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 20 -c 2 -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/bench_under_ncu_full_${tag}.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_${tag}.ncu-rep 12 > gpurun_out/ncu_summary_${tag}.txt 2>&1
python tools/ncu_by_kernel.py gpurun_out/launches_${tag}.csv > gpurun_out/launches_${tag}_by_kernel.txt 2>&1

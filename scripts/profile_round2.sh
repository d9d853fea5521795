#!/bin/bash
# Round-2 profile captures (run under gpurun, one GPU).  Every ncu command is preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err &&
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv $B > /dev/null 2>&1
S="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-hbm-leg --no-secondary"
$S > /dev/null 2>&1 &&
$NCU --set full --import-source on -k regex:o2_hogwild_d128 -s 2 -c 1 -o gpurun_out/r2_o2_sbm $S > /dev/null 2>&1
Y="python bench.py --workload youtube --steps 1 --warmup 1 --no-cpu-baseline --no-secondary"
$Y > /dev/null 2>&1 &&
$NCU --set full --import-source on -k regex:o2_hogwild_d128 -s 1 -c 1 -o gpurun_out/r2_o2_youtube $Y > /dev/null 2>&1
F="python bench.py --kernel sg --sg-walks 20000 --steps 1 --warmup 1"
$F > /dev/null 2>&1 &&
$NCU --set full --import-source on -k regex:sg_async -s 1 -c 1 -o gpurun_out/r2_sg_async $F > /dev/null 2>&1
G="python scripts/o3_gemm_check.py"
$G > gpurun_out/r2_o3_gemm_check.log 2>&1 &&
$NCU --set full --import-source on -k regex:o3_gemm_kernel -s 12 -c 1 -o gpurun_out/r2_o3_gemm $G > /dev/null 2>&1
M="python scripts/gmm_profile.py"
$M > gpurun_out/r2_gmm_profile.log 2>&1 &&
$NCU --set full --import-source on -k regex:gmm_estep_kernel -s 3 -c 1 -o gpurun_out/r2_gmm_estep $M > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep

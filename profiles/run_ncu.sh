#!/bin/bash
# Profiling recipe for one round (run under gpurun from the repo root):
#   bash profiles/run_ncu.sh <tag>
# 1. plain run of the exact command (must exit 0), 2. launch list with device
# times, 3. one `--set full` capture of the fused ODE kernel.  Outputs go to
# gpurun_out/; summaries are copied into profiles/ by hand afterwards.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ode_kernel -s 4 -c 2 \
    -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/

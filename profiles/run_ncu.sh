#!/bin/bash
# Profiling recipe for one round (run under gpurun, ONE GPU, from the repo root):
#   bash profiles/run_ncu.sh <tag> [headline|kernels|all]
# headline: 1. plain run of the exact bench command (must exit 0), 2. launch
#   list with device times, 3. one `--set full` capture of the fused ODE kernel.
# kernels: one `--set full` capture each of the other kernels north_star names
#   (unfused baseline, closed-form kernel, sampler kernels, ranking kernels,
#   the reading side of the exchange), from tools/ncu_targets.py; each target
#   first runs plain (its event-timed figures go to gpurun_out/targets_<tag>.log).
# Outputs go to gpurun_out/; profiles/summarize_all.sh turns the reports into
# the text summaries committed under profiles/.
set -u
TAG=${1:-r02}
WHAT=${2:-all}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
if [ "$WHAT" = headline ] || [ "$WHAT" = all ]; then
  CMD="python bench.py --steps 3 --warmup 3 --no-cpu"
  $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
  $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
  $NCU -k regex:ode_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG}_ode_kernel $CMD \
      > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -c 600 gpurun_out/plain_$TAG.log
  tail -3 gpurun_out/ncu_full_$TAG.log
  python profiles/summarize.py gpurun_out/prof_${TAG}_ode_kernel.ncu-rep > gpurun_out/${TAG}_ode_kernel_summary.md
  ncu -i gpurun_out/prof_${TAG}_ode_kernel.ncu-rep --page source --csv > gpurun_out/src_ode.csv 2>/dev/null &&
    python profiles/sass_mix.py gpurun_out/src_ode.csv 30 > gpurun_out/${TAG}_ode_kernel_sass_mix.txt
  rm -f gpurun_out/src_ode.csv
fi
if [ "$WHAT" = kernels ] || [ "$WHAT" = all ]; then
  : > gpurun_out/targets_$TAG.log
  capture() {   # target  kernel-regex  skip  name
    python tools/ncu_targets.py $1 >> gpurun_out/targets_$TAG.log 2>&1 &&
    $NCU -k regex:$2 -s $3 -c 1 -o gpurun_out/prof_${TAG}_$4 \
        python tools/ncu_targets.py $1 > gpurun_out/ncu_$4_$TAG.log 2>&1
    tail -1 gpurun_out/ncu_$4_$TAG.log
    # gpurun brings back at most 64 MiB: keep the text, drop the big report
    python profiles/summarize.py gpurun_out/prof_${TAG}_$4.ncu-rep > gpurun_out/${TAG}_$4_summary.md
    ncu -i gpurun_out/prof_${TAG}_$4.ncu-rep --page source --csv > gpurun_out/src_$4.csv 2>/dev/null &&
      python profiles/sass_mix.py gpurun_out/src_$4.csv 20 > gpurun_out/${TAG}_$4_sass_mix.txt
    rm -f gpurun_out/src_$4.csv
    [ $(stat -c %s gpurun_out/prof_${TAG}_$4.ncu-rep) -gt 9000000 ] && rm -f gpurun_out/prof_${TAG}_$4.ncu-rep
  }
  capture unfused  'ode_kernel'         1 unfused_simulate
  capture unfused  'score_traj_kernel'  1 unfused_reduce
  capture logistic 'analytic_kernel'    2 logistic_kernel
  capture samplers 'ac_ask'             4 ac_ask
  capture samplers 'ac_tell'            4 ac_tell
  capture samplers 'de_ask'             4 de_ask
  capture exchange 'exchange_finish'    2 exchange_finish
  # the ranking is several kernels: capture them all from one generation
  python tools/ncu_targets.py argsort >> gpurun_out/targets_$TAG.log 2>&1 &&
  $NCU -k regex:argsort -s 35 -c 7 -o gpurun_out/prof_${TAG}_argsort \
      python tools/ncu_targets.py argsort > gpurun_out/ncu_argsort_$TAG.log 2>&1
  python profiles/summarize.py gpurun_out/prof_${TAG}_argsort.ncu-rep > gpurun_out/${TAG}_argsort_summary.md
  cat gpurun_out/targets_$TAG.log
fi
ls -la gpurun_out/ | tail -20

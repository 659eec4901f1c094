mkdir -p gpurun_out
{
for sp in 0.05 0.2 0 0.1 0.02; do
  echo "== n=65536 spread=$sp global / auto"
  PB200_LOCAL_SORT=0 timeout 120 build/kbench_local 65536 40 0 $sp 1000 0 0 1
  timeout 120 build/kbench_local 65536 40 0 $sp 1000 0 0 1
done
} > gpurun_out/r02ad_sort_auto.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02ad_pytest.log 2>&1
tail -3 gpurun_out/r02ad_pytest.log
cat gpurun_out/r02ad_sort_auto.log

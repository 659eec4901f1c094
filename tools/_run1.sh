mkdir -p gpurun_out
{
for n in 32768 50000 73000; do
  echo "== n=$n global sort: 64-thread blocks / one block per SM;  local sort one block per SM"
  PB200_LOCAL_SORT=0 timeout 120 build/kbench_local $n 40 0 0.05 1000 0 0 1
  PB200_LOCAL_SORT=0 PB200_PLAIN_1BLOCK=1 timeout 120 build/kbench_local $n 40 0 0.05 1000 0 0 1
  timeout 120 build/kbench_local $n 40 0 0.05 1000 0 0 1
done
echo "== wide box, n=73000 / 50000: 64-thread blocks / one block per SM (global sort)"
for n in 73000 50000; do
  PB200_LOCAL_SORT=0 timeout 120 build/kbench_local $n 40 0 0 1000 0 0 1
  PB200_LOCAL_SORT=0 PB200_PLAIN_1BLOCK=1 timeout 120 build/kbench_local $n 40 0 0 1000 0 0 1
done
} > gpurun_out/r02af_local_plain2.log 2>&1
cat gpurun_out/r02af_local_plain2.log

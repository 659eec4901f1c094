mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== n=65536 flag / poison"
timeout 120 build/kbench_nopoison 65536 40 0 0.05 1000 0 0 1
timeout 120 build/kbench_local 65536 40 0 0.05 1000 0 0 1
done
echo "== n=1000000 flag / poison"
timeout 120 build/kbench_nopoison 1000000 10 0 0.05 1000 0 0 1
timeout 120 build/kbench_local 1000000 10 0 0.05 1000 0 0 1
echo "== n=100000 flag / poison"
timeout 120 build/kbench_nopoison 100000 20 0 0.05 1000 0 0 1
timeout 120 build/kbench_local 100000 20 0 0.05 1000 0 0 1
echo "== n=4096 flag / poison"
timeout 120 build/kbench_nopoison 4096 40 0 0.05 1000 0 0 1
timeout 120 build/kbench_local 4096 40 0 0.05 1000 0 0 1
} > gpurun_out/r02al_fail_poison.log 2>&1
cat gpurun_out/r02al_fail_poison.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02al_pytest.log 2>&1
tail -3 gpurun_out/r02al_pytest.log

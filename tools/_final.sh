mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; tail -2 gpurun_out/r02_final_pytest.log
python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 300 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('value_as_list'), d['gpu_launches'], d['cost_sort'][:20], d['clocks'])
PY

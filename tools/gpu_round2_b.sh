set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_b.log
cat gpurun_out/r02_pytest_gpu_b.log
timeout 300 python tools/run_mode.py weighted 2>&1 | tail -7
timeout 300 python tools/run_mode.py count 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_b.json'))
print({k:d[k] for k in ('value','ms_per_step','exchange_tail_ms','ms_per_step_pipelined_no_l2_flush')}, d['e2e'], d['parity'])
for k,v in d['secondary'].items(): print(k,v)
PY
tail -3 gpurun_out/r02_bench_b.err

set -x
for c in c2w c2d5 c5; do
  timeout 900 python bench.py --config $c > gpurun_out/r02_bench_${c}_n1.json 2> gpurun_out/r02_bench_${c}_n1.err
  tail -2 gpurun_out/r02_bench_${c}_n1.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${c}_n1.json'))
print('${c}', {k:d.get(k) for k in ('value','ms_per_step','steps')}, d['e2e'], d.get('parity'), d.get('cpu_baseline'))
print({k:d['roofline'].get(k) for k in ('kernel_ms','algorithmic_bytes_per_hyp','kbar_27','frac','moved_bytes_frac','dram_frac')})
PY
done

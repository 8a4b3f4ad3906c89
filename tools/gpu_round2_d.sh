set -x
N=${1:-2}
for c in c3 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c > gpurun_out/r02_bench_${c}_n${N}.json 2> gpurun_out/r02_bench_${c}_n${N}.err
  tail -3 gpurun_out/r02_bench_${c}_n${N}.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${c}_n${N}.json'))
print('${c}', {k:d.get(k) for k in ('value','ms_per_step','steps','n_gpus')}, d['e2e'].get('value'), d.get('parity'), d.get('sharding_invariance'), d.get('top64_digest'))
PY
done

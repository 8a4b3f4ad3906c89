# the 8-GPU box: scaling lines of every sharded config (one box, back to back)
set -x
run() { # config N
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $2 --config $1 > gpurun_out/r02_bench_$1_n$2.json 2> gpurun_out/r02_bench_$1_n$2.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_$1_n$2.json'))
    print('$1 N=$2', {k:d.get(k) for k in ('value','ms_per_step','steps')}, 'e2e', d['e2e'].get('value'), d.get('parity'), d.get('sharding_invariance'), [x['top64_crc'] for x in d.get('top64_digest', [])])
except Exception as e:
    print('$1 N=$2 FAILED', e); print(open('gpurun_out/r02_bench_$1_n$2.err').read()[-1500:])
PY
}
run c2 8; run c2 4; run c2w 8; run c5 8; run c3 8; run c3 4

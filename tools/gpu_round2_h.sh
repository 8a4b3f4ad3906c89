# 2-GPU box: the N = 2 scaling lines and the whole GPU test suite (group tests need >= 2 devices)
run() { # config N
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $2 --config $1 > gpurun_out/r02_bench_$1_n$2.json 2> gpurun_out/r02_bench_$1_n$2.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_$1_n$2.json'))
    print('$1 N=$2', '%.4g' % d['value'], round(d['ms_per_step'],4), 'e2e %.4g' % d['e2e']['value'], d.get('parity'), d.get('sharding_invariance'), [x['top64_crc'] for x in d.get('top64_digest', [])])
except Exception as e:
    print('$1 N=$2 FAILED', e); print(open('gpurun_out/r02_bench_$1_n$2.err').read()[-1500:])
PY
}
run c2 2; run c5 2; run c3 2; cp gpurun_out/r02_bench_c3_n2.json gpurun_out/r02_bench_c3_n2_first.json; run c3 2
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4

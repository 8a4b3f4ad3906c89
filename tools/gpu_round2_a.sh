set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu_a.log
timeout 400 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2_a.json 2> gpurun_out/r02_bench_n2_a.err
cat gpurun_out/r02_pytest_gpu_a.log
cut -c1-1500 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
cut -c1-1500 gpurun_out/r02_bench_n2_a.json; tail -5 gpurun_out/r02_bench_n2_a.err

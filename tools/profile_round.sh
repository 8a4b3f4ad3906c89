set -x
export PGP_STREAM_UPLOAD=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_v13.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_bench.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k3_fine_kernel -s 2 -c 1 -o gpurun_out/k3_v13 -f python tools/run_mode.py count > gpurun_out/k3_v13.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k3_fine_kernel -s 2 -c 1 -o gpurun_out/k3w_v13 -f python tools/run_mode.py weighted > gpurun_out/k3w_v13.log 2>&1
unset PGP_STREAM_UPLOAD
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref_v13.json 2> gpurun_out/ref.err
timeout 300 python bench.py > gpurun_out/r01_bench_v13.json 2> gpurun_out/bench.err
tail -n 2 gpurun_out/k3_v13.log; tail -n 2 gpurun_out/k3w_v13.log
cut -c1-400 gpurun_out/r01_bench_v13.json; cut -c1-300 gpurun_out/r01_bench_ref_v13.json

set -x
export PGP_STREAM_UPLOAD=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches_bench.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k3_fine_kernel -s 2 -c 1 -o gpurun_out/r02_k3 -f python tools/run_mode.py count > gpurun_out/r02_k3.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k3_fine_kernel -s 2 -c 1 -o gpurun_out/r02_k3w -f python tools/run_mode.py weighted > gpurun_out/r02_k3w.log 2>&1
tail -3 gpurun_out/r02_k3.log

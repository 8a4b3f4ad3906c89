# Round-2 evidence: ncu launch list of the bench command, ncu --set full of the dominant kernels, bench lines.  One GPU.
set -x
export PGP_STREAM_UPLOAD=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches_bench.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k3_fine_kernel -s 2 -c 1 -o gpurun_out/r02_k3 -f python tools/run_mode.py count > gpurun_out/r02_k3.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k3_fine_kernel -s 2 -c 1 -o gpurun_out/r02_k3w -f python tools/run_mode.py weighted > gpurun_out/r02_k3w.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k2_join_query -s 1 -c 1 -o gpurun_out/r02_k2_query -f python tools/gen_only.py 1 13000 > gpurun_out/r02_k2_query.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k2_join_select -s 1 -c 1 -o gpurun_out/r02_k2_select -f python tools/gen_only.py 1 13000 > gpurun_out/r02_k2_select.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k2s_select_bases -s 1 -c 1 -o gpurun_out/r02_k2s_bases -f python tools/gen_only.py 1 13000 > gpurun_out/r02_k2s_bases.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k5_tricp_kernel -s 1 -c 1 -o gpurun_out/r02_k5 -f python tools/tricp_run.py > gpurun_out/r02_k5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_k2_launches_m1.csv python tools/gen_only.py 1 13000 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_k2_launches_m0.csv python tools/gen_only.py 0 100 > /dev/null 2>&1
unset PGP_STREAM_UPLOAD
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
for c in c2w c2d5 c5 c3; do timeout 900 python bench.py --config $c > gpurun_out/r02_bench_${c}.json 2> gpurun_out/r02_bench_${c}.err; done
timeout 300 python tools/gen_only.py 0 100 > gpurun_out/r02_gen_m0.log 2>&1; timeout 300 python tools/gen_only.py 1 100 > gpurun_out/r02_gen_m1_100.log 2>&1; timeout 300 python tools/gen_only.py 1 13000 > gpurun_out/r02_gen_m1_13000.log 2>&1
timeout 300 python tools/tricp_run.py > gpurun_out/r02_tricp.log 2>&1
timeout 600 python tools/c1_latency.py > gpurun_out/r02_c1_latency.json 2> gpurun_out/r02_c1_latency.err
cut -c1-300 gpurun_out/r02_bench.json; tail -1 gpurun_out/r02_tricp.log; tail -1 gpurun_out/r02_gen_m0.log

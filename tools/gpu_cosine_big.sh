mkdir -p gpurun_out
timeout 200 python tools/bench_cosine.py 131072 3000 > gpurun_out/cosine_bench_131072.json 2> gpurun_out/cosine_bench.err; cat gpurun_out/cosine_bench_131072.json; tail -n 3 gpurun_out/cosine_bench.err
timeout 300 python tools/bench_cosine.py 1000000 3000 > gpurun_out/cosine_bench_1m.json 2>> gpurun_out/cosine_bench.err; cat gpurun_out/cosine_bench_1m.json; tail -n 3 gpurun_out/cosine_bench.err
timeout 120 python tools/bench_cosine.py 16384 3000 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:cosine_tc_kernel -c 1 -o gpurun_out/prof_cosine_tc -f python tools/bench_cosine.py 16384 3000 > gpurun_out/ncu_cos.log 2>&1; tail -n 2 gpurun_out/ncu_cos.log

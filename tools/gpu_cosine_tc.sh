mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_cosine.py -m gpu -q -x > gpurun_out/tests_tc.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_tc.log
tail -n 30 gpurun_out/tests_tc.log
timeout 120 python tools/bench_cosine.py 16384 3000 > gpurun_out/cosine_bench.json 2> gpurun_out/cosine_bench.err; cat gpurun_out/cosine_bench.json; tail -n 3 gpurun_out/cosine_bench.err

mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/tests9.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests9.log
tail -n 6 gpurun_out/tests9.log
python bench.py --cpu-budget 5 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json

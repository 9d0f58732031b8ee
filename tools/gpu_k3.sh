mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests3.log
tail -25 gpurun_out/tests3.log
python bench.py --no-cpu-baseline --e2e-fields 64 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"
cat gpurun_out/bench3.json; tail -5 gpurun_out/bench3.err

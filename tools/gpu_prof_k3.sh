mkdir -p gpurun_out
python -m pytest tests/test_gpu_object_stats.py -m gpu -q -x > gpurun_out/tests_k3.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_k3.log
tail -n 5 gpurun_out/tests_k3.log
python bench.py --no-cpu-baseline --e2e-fields 16 --steps 50 > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_k3.json')); print(d['value'], d['kernels'])"
python tools/prof_kernels.py k3 16 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'object_stats_scan' -s 2 -c 1 -o gpurun_out/prof_k3c -f python tools/prof_kernels.py k3 16 > gpurun_out/ncu_k3c.log 2>&1
tail -n 3 gpurun_out/ncu_k3c.log

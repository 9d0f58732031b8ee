mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -q -x 2>&1 | tail -n 15
for v in 0 1; do
IPS_K1_TMA=$v timeout 300 python bench.py --mode split --no-cpu-baseline --e2e-fields 8 --steps 60 > gpurun_out/bench_tma$v.json 2> gpurun_out/bench_tma$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tma$v.json')); print('IPS_K1_TMA=$v', d['kernels']['K1'])"; tail -n 2 gpurun_out/bench_tma$v.err
done

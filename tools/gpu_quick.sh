mkdir -p gpurun_out
python -m pytest tests/test_gpu_field_fused.py -m gpu -q -x 2>&1 | tail -n 2
python bench.py --no-cpu-baseline --e2e-fields 8 --steps 100 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print(d['value'], d['kernels'])"; tail -n 3 gpurun_out/bench_q.err

mkdir -p gpurun_out
python -m pytest tests/test_gpu_qc_wells.py tests/test_gpu_plate.py -m gpu -q -x 2>&1 | tail -n 3
python bench.py --no-cpu-baseline --e2e-fields 16 > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1d.json')); print(1, d['value'], d['ms_per_step'], d['aggregation']['ms_after_last_step'])"
python tools/bench_aux.py > gpurun_out/bench_aux.jsonl 2> gpurun_out/bench_aux.err; cat gpurun_out/bench_aux.jsonl; tail -n 3 gpurun_out/bench_aux.err

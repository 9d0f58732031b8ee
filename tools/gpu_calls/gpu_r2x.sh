#!/usr/bin/env bash
# round-2 follow-up on ONE GPU: the new well-sum kernel (parity tests, tail micro-benchmark), the plate row
# exchange at one rank, the 20-step bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_qc_wells.py tests/test_gpu_plate.py tests/test_gpu_scripts.py -m gpu -q -x > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2x_tests.log
timeout 300 python tools/bench_wellagg.py --world 8 --steps 20 --chunks 20 > gpurun_out/r2x_wellagg_n8_20.json 2> gpurun_out/r2x_wellagg.err; cat gpurun_out/r2x_wellagg_n8_20.json
timeout 300 python tools/bench_wellagg.py --world 1 --steps 20 --chunks 1 > gpurun_out/r2x_wellagg_n1_20.json 2>> gpurun_out/r2x_wellagg.err; cat gpurun_out/r2x_wellagg_n1_20.json
timeout 300 python tools/bench_wellagg.py --world 8 --steps 216 --chunks 18 --iters 5 > gpurun_out/r2x_wellagg_n8_216.json 2>> gpurun_out/r2x_wellagg.err; cat gpurun_out/r2x_wellagg_n8_216.json
tail -3 gpurun_out/r2x_wellagg.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-files --no-cpu-baseline > gpurun_out/r2x_bench20.json 2> gpurun_out/r2x_bench20.err; echo "bench20 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench20.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['kernels']['fused']['ms_per_launch'], d['aggregation']['ms_after_last_step'], d['aggregation']['check'], d['e2e']['value'])
PY
tail -3 gpurun_out/r2x_bench20.err

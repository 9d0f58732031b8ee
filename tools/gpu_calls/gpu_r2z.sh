#!/usr/bin/env bash
# N GPUs: two-GPU tests of the exchange, the tail micro-benchmark, bench.py at N (20 steps, then the default plate)
set -u
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_plate.py tests/test_gpu_qc_wells.py -m gpu -q -x > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2z_tests.log
  timeout 300 python tools/bench_wellagg.py --world 8 --steps 20 --chunks 20 > gpurun_out/r2z_wellagg_n8_20.json 2> gpurun_out/r2z_wellagg.err; cat gpurun_out/r2z_wellagg_n8_20.json
fi
for steps in $2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps $steps --warmup 5 > gpurun_out/r2z_bench_n${N}_s$steps.json 2> gpurun_out/r2z_bench_n${N}_s$steps.err; echo "bench N=$N steps=$steps rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2z_bench_n${N}_s$steps.json').read().strip().splitlines()[-1])
    a=d['aggregation']
    print(d["value"], d["ms_per_step"], d["kernels"]["fused"]["ms_per_launch"], a["ms_after_last_step"], a.get("ms_last_chunk_and_barrier"), a['check'], a.get('bytes_pushed_to_peers_last_plate'), d['e2e']['value'], d['e2e']['rows_only']['value'])
except Exception as e:
    print("no line:", e)
PY
tail -2 gpurun_out/r2z_bench_n${N}_s$steps.err
done

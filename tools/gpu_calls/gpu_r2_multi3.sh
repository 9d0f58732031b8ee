#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_plate.py -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2w_tests.log
for N in 8 4 2 1; do
  if [ $N = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2974$N"; fi
  $L bench.py --gpus $N --steps 20 --warmup 5 --no-files > gpurun_out/r2w_bench_n$N.json 2> gpurun_out/r2w_bench_n$N.err
  echo "bench N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2w_bench_n$N.json')); print(round(d['value']), round(d['ms_per_step'],4), round(d['kernels']['fused']['ms_per_launch'],4), round(d['aggregation']['ms_after_last_step'],3), d['aggregation']['check'], round(d['e2e']['value'],1), round(d['e2e']['frac_of_ceiling'],3), round(d['e2e']['rows_only']['value'],1))"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29752 \
    bench.py --gpus 8 --steps 216 --warmup 5 --e2e-fields 16 --no-files > gpurun_out/r2w_bench_n8_plate.json 2> gpurun_out/r2w_bench_n8_plate.err
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench_n8_plate.json')); print('plate', round(d['value']), round(d['ms_per_step'],4), round(d['kernels']['fused']['ms_per_launch'],4), d['aggregation']['ms_after_last_step'], d['aggregation']['check'])"

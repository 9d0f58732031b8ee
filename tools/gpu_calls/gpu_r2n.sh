#!/usr/bin/env bash
# 2 GPUs: peer-push transport test, sharded cosine, N = 2 bench both transports
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_plate.py tests/test_gpu_cosine.py -m gpu -q -x > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2n_tests.log
for extra in "" "--no-peer-push"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --steps 20 --warmup 5 --e2e-fields 16 --sustained-steps 0 $extra > gpurun_out/r2n_bench_n2$extra.json 2> gpurun_out/r2n_bench_n2$extra.err
echo "bench N=2 $extra rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_n2$extra.json')); print(d['value'], d['ms_per_step'], d['kernels']['fused']['ms_per_launch'], d['aggregation'])"
tail -2 gpurun_out/r2n_bench_n2$extra.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/bench_cosine_dist.py 131072 3000 > gpurun_out/r2n_cosine_n2.json 2> gpurun_out/r2n_cosine.err; cat gpurun_out/r2n_cosine_n2.json; tail -3 gpurun_out/r2n_cosine.err

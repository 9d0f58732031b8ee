#!/usr/bin/env bash
# 8-GPU box, final: scaling with the peer-push gather (and NCCL-only for comparison), config 5 cosine, K2 sharded
set -u
mkdir -p gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2v_bench_n$N.json 2> gpurun_out/r2v_bench_n$N.err
  echo "bench N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_n$N.json')); print(round(d['value']), round(d['ms_per_step'],4), round(d['kernels']['fused']['ms_per_launch'],4), d['aggregation']['ms_after_last_step'], d['aggregation']['check'], d['aggregation']['bulk_transport'][:40], round(d['e2e']['value'],1), round(d['e2e']['frac_of_ceiling'],3), round(d['e2e']['rows_only']['value'],1))"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-peer-push --e2e-fields 16 > gpurun_out/r2v_bench_n8_nccl.json 2> gpurun_out/r2v_bench_n8_nccl.err
python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_n8_nccl.json')); print('nccl only', round(d['value']), round(d['ms_per_step'],4), round(d['kernels']['fused']['ms_per_launch'],4), d['aggregation']['ms_after_last_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 \
    bench.py --gpus 8 --steps 216 --warmup 5 --e2e-fields 16 > gpurun_out/r2v_bench_n8_plate.json 2> gpurun_out/r2v_bench_n8_plate.err
python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_n8_plate.json')); print('plate', round(d['value']), round(d['ms_per_step'],4), round(d['kernels']['fused']['ms_per_launch'],4), d['aggregation']['ms_after_last_step'], d['aggregation']['check'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29653 tools/bench_cosine_dist.py 1000000 3000 > gpurun_out/r2v_cosine_n8.json 2> gpurun_out/r2v_cosine.err; cat gpurun_out/r2v_cosine_n8.json; tail -2 gpurun_out/r2v_cosine.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29654 tools/bench_illum_dist.py > gpurun_out/r2v_illum_n8.json 2> gpurun_out/r2v_illum.err; cat gpurun_out/r2v_illum_n8.json; tail -2 gpurun_out/r2v_illum.err | cut -c1-300

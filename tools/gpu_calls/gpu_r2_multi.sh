#!/usr/bin/env bash
# 8-GPU box: topology, bare copy ceiling at 1/2/4/8 GPUs, NCCL content test, bench at N = 8 / 4 / 2
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
(lscpu | head -40; echo; numactl -H 2>/dev/null; echo; free -g) > gpurun_out/r2m_host.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_plate.py -m gpu -q -x > gpurun_out/r2m_tests.log 2>&1; echo "nccl test rc=$?"; tail -3 gpurun_out/r2m_tests.log
python tools/bench_pcie.py --sweep 1 2 4 8 > gpurun_out/r2m_pcie.jsonl 2> gpurun_out/r2m_pcie.err
python tools/bench_pcie.py --sweep 2 8 --pin-cores >> gpurun_out/r2m_pcie.jsonl 2>> gpurun_out/r2m_pcie.err
cat gpurun_out/r2m_pcie.jsonl | cut -c1-600
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
  echo "bench N=$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_n$N.json')); print(d['value'], d['ms_per_step'], d['aggregation'], {k:v for k,v in d['e2e'].items() if k in ('value','h2d_gbs','pcie_ceiling_gbs','frac_of_ceiling','rows_only','ceiling')})"
  tail -2 gpurun_out/r2m_bench_n$N.err
done

#!/usr/bin/env bash
# round 2, first GPU call: new parity tests, instruction-rate microbenchmark, PCIe ceiling, bench
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2a_topo.txt 2>&1
lscpu | head -30 > gpurun_out/r2a_lscpu.txt 2>&1
./tools/ubench/f32x2 > gpurun_out/r2a_ubench.txt 2>&1; cat gpurun_out/r2a_ubench.txt
timeout 900 python -m pytest tests/test_gpu_qc_wells.py tests/test_gpu_plate.py tests/test_gpu_field_fused.py -m gpu -q -x > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?"; tail -n 15 gpurun_out/r2a_tests.log
python tools/bench_pcie.py > gpurun_out/r2a_pcie.jsonl 2>gpurun_out/r2a_pcie.err; python tools/bench_pcie.py --pin-cores >> gpurun_out/r2a_pcie.jsonl 2>>gpurun_out/r2a_pcie.err
python tools/bench_pcie.py --write-combined >> gpurun_out/r2a_pcie.jsonl 2>>gpurun_out/r2a_pcie.err
cat gpurun_out/r2a_pcie.jsonl; tail -3 gpurun_out/r2a_pcie.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cat gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err

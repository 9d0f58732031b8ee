#!/usr/bin/env bash
# final single-GPU evidence: full GPU test suite, smoke, ncu (launch list + full capture), bench both arms
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2u_tests.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python tools/prof_kernels.py fused 16 > gpurun_out/r2u_prof_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:field_fused2 -s 2 -c 1 -f -o gpurun_out/r2u_fused2s \
      python tools/prof_kernels.py fused 16 > gpurun_out/r2u_ncu.log 2>&1
tail -n 2 gpurun_out/r2u_ncu.log
CMD="python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --no-files --e2e-fields 8 --e2e-ring 1 --sustained-steps 0"
$CMD > gpurun_out/r2u_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'field_fused|object_stats|preprocess|rows_|well_|widen|illum_|block_' \
      --csv --log-file gpurun_out/r2u_launches.csv $CMD > gpurun_out/r2u_ncu_launches.log 2>&1
tail -n 1 gpurun_out/r2u_ncu_launches.log | cut -c1-200
python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2u_bench.json; tail -3 gpurun_out/r2u_bench.err
python bench.py --steps 20 --warmup 5 --no-files > gpurun_out/r2u_bench20.json 2>> gpurun_out/r2u_bench.err; echo "bench20 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_bench_reference.json 2>> gpurun_out/r2u_bench.err; cat gpurun_out/r2u_bench_reference.json | cut -c1-600

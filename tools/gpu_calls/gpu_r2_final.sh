#!/usr/bin/env bash
# last single-GPU evidence of round 2: full GPU test suite, smoke, ncu (launch list of the bench step, full capture
# of the well-sum kernel), the default bench line and the 20-step line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2f_tests.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2f_bench.json; tail -2 gpurun_out/r2f_bench.err
python bench.py --steps 20 --warmup 5 --no-files > gpurun_out/r2f_bench20.json 2>> gpurun_out/r2f_bench.err; echo "bench20 rc=$?"
CMD="python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --no-files --e2e-fields 8 --e2e-ring 1 --sustained-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'field_fused|object_stats|preprocess|rows_|well_|widen|illum_|block_' \
    --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1
tail -n 1 gpurun_out/r2f_ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:well_accumulate -s 1 -c 1 -f -o gpurun_out/r2f_wellsums \
    python tools/bench_wellagg.py --world 8 --steps 20 --chunks 20 --iters 2 > gpurun_out/r2f_ncu_wellsums.log 2>&1
tail -n 2 gpurun_out/r2f_ncu_wellsums.log | cut -c1-300

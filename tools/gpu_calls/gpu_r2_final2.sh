#!/usr/bin/env bash
# last single-GPU check of the shipped tree: full GPU test suite, smoke, the 20-step bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2g_tests.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-files > gpurun_out/r2g_bench20.json 2> gpurun_out/r2g_bench.err; echo "bench20 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench20.json').read().strip().splitlines()[-1])
a=d['aggregation']
print(d['value'], d['ms_per_step'], d['kernels']['fused']['ms_per_launch'], a['ms_after_last_step'], a['ms_well_sums'], a['check'], d['e2e']['value'], d['cpu_baseline']['value'])
PY

mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/tests_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_all.log
tail -n 6 gpurun_out/tests_all.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 2
python bench.py --cpu-budget 5 > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_all.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['aggregation']['ms'], d['cpu_baseline'])"

mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/tests6.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests6.log
tail -n 6 gpurun_out/tests6.log
python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"; cat gpurun_out/bench_r1b.json
python bench.py --mode split --no-cpu-baseline --e2e-fields 8 --steps 60 > gpurun_out/bench_r1b_split.json 2>> gpurun_out/bench_r1b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1b_split.json')); print('split', d['value'], d['kernels'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1b_ref.json 2>> gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b_ref.json
CMD="python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 4 --e2e-ring 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python tools/prof_kernels.py fused 16 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'field_fused' -s 2 -c 1 -o gpurun_out/prof_fused_r1b -f python tools/prof_kernels.py fused 16 > gpurun_out/ncu_fused_b.log 2>&1
tail -n 2 gpurun_out/ncu_fused_b.log

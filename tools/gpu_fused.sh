mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/tests4.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests4.log
tail -n 25 gpurun_out/tests4.log
for m in fused split; do
python bench.py --no-cpu-baseline --e2e-fields 16 --steps 60 --mode $m > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$m.json')); print('$m', d['value'], d['kernels'])"; tail -n 3 gpurun_out/bench_$m.err
done
python tools/prof_kernels.py fused 16 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'field_fused' -s 2 -c 1 -o gpurun_out/prof_fused_a -f python tools/prof_kernels.py fused 16 > gpurun_out/ncu_fused_a.log 2>&1
tail -n 3 gpurun_out/prof_plain.log gpurun_out/ncu_fused_a.log

mkdir -p gpurun_out
python -m pytest tests/test_gpu_field_fused.py -m gpu -q -x > gpurun_out/tests5.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests5.log
tail -n 4 gpurun_out/tests5.log
for v in 4 5 6; do
IPS_FUSED_VARIANT=$v python bench.py --no-cpu-baseline --e2e-fields 8 --steps 60 --mode fused > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('variant $v', d['value'], d['kernels'])"; tail -n 3 gpurun_out/bench_v$v.err
done

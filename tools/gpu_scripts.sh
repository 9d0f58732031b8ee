mkdir -p gpurun_out
python -m pytest tests/test_gpu_scripts.py -m gpu -q -x > gpurun_out/tests8.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests8.log
tail -n 40 gpurun_out/tests8.log
CMD="python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 8 --e2e-ring 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'field_fused|object_stats|preprocess' --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log | cut -c1-200

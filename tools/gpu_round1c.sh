mkdir -p gpurun_out
python -m pytest tests/test_gpu_cosine.py tests/test_gpu_pipeline.py tests/test_gpu_qc_wells.py -m gpu -q -x > gpurun_out/tests7.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests7.log
tail -n 15 gpurun_out/tests7.log
CMD="python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 8 --e2e-ring 1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ips::' --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log

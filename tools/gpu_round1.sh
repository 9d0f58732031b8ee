mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 4 --e2e-ring 1 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 4 --e2e-ring 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'preprocess_vec|object_stats_scan' -s 8 -c 4 -o gpurun_out/prof_r1a python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 4 --e2e-ring 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/tests.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

mkdir -p gpurun_out
python bench.py --no-cpu-baseline --e2e-fields 16 > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_n1c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1c.json')); print(1, d['value'], d['ms_per_step'], d['aggregation'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --e2e-fields 16 > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err; echo "rc=$?"; tail -n 3 gpurun_out/bench_n2c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2c.json')); print(2, d['value'], d['ms_per_step'], d['aggregation'])"

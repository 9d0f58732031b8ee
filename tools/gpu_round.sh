#!/usr/bin/env bash
# What to run on the B200 box (through gpurun).  Usage:  bash tools/gpu_round.sh <what> [args]
#   tests            pytest -m gpu (all parity tests through the C-ABI) + smoke()
#   bench            bench.py (ours) and bench.py --impl reference, N = 1
#   scale N          bench.py under torchrun on N GPUs
#   profile          ncu launch list of the bench step + ncu --set full of the fused kernel
#   aux              secondary kernels (illumination, Lanczos, K1 bin sweep, K1 TMA variant)
#   cosine [n] [d]   tensor-core cosine kernel on an n x d group
#   tiff             device TIFF-LZW codec: parity tests, then throughput beside Pillow
#   wellagg [N] [K]  the tail of the plate step on one GPU: well sums over the table an N-rank K-step run gathers
#   files [sites]    the file-level drop-in scripts (TIFF files in -> TIFF / CSV files out) with stage times
# Everything lands in gpurun_out/; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
what=${1:-tests}
case "$what" in
  tests)
    timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
    tail -n 8 gpurun_out/tests.log
    python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 2
    ;;
  bench)
    python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
    python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
    cat gpurun_out/bench_reference.json
    ;;
  scale)
    N=${2:-2}
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus "$N" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
    echo "bench rc=$?"; tail -n 3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
    ;;
  profile)
    CMD="python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 8 --e2e-ring 1"
    $CMD > gpurun_out/plain.log 2>&1 && \
      ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'field_fused|object_stats|preprocess|rows_|well_|widen' \
          --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
    python tools/prof_kernels.py fused 16 > gpurun_out/prof_plain.log 2>&1 && \
      ncu --set full --clock-control none --import-source on -k regex:field_fused -s 2 -c 1 -f -o gpurun_out/prof_fused \
          python tools/prof_kernels.py fused 16 > gpurun_out/ncu_fused.log 2>&1
    tail -n 2 gpurun_out/ncu_launches.log gpurun_out/ncu_fused.log
    ;;
  aux)
    python tools/bench_aux.py > gpurun_out/bench_aux.jsonl 2> gpurun_out/bench_aux.err; cat gpurun_out/bench_aux.jsonl
    for v in 0 1; do
      IPS_K1_TMA=$v python bench.py --mode split --no-cpu-baseline --e2e-fields 8 --steps 60 2>/dev/null | \
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('IPS_K1_TMA=$v', d['kernels']['K1'])"
    done
    ;;
  cosine)
    timeout 600 python tools/bench_cosine.py "${2:-16384}" "${3:-3000}" | tee gpurun_out/cosine_bench.json
    ;;
  tiff)
    timeout 600 python -m pytest tests/test_gpu_tiff.py tests/test_gpu_scripts.py -m gpu -q -x > gpurun_out/tiff_tests.log 2>&1
    echo "tiff tests rc=$?"; tail -n 30 gpurun_out/tiff_tests.log
    timeout 600 python tools/bench_tiff.py > gpurun_out/bench_tiff.jsonl 2> gpurun_out/bench_tiff.err; echo "bench rc=$?"
    cat gpurun_out/bench_tiff.jsonl; tail -n 5 gpurun_out/bench_tiff.err
    ;;
  wellagg)
    python tools/bench_wellagg.py --world "${2:-8}" --steps "${3:-20}" --chunks 20 | tee gpurun_out/wellagg.json
    ;;
  files)
    IPS_IO_TRACE=1 python tools/bench_files.py --sites "${2:-256}" --distinct 4 --cpu-sites 1 > gpurun_out/files.json 2> gpurun_out/files.err
    echo "files rc=$?"; cat gpurun_out/files.json; grep -i "stages" gpurun_out/files.err | tail -n 4
    ;;
  *) echo "unknown: $what"; exit 2 ;;
esac

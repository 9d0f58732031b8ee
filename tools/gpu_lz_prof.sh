mkdir -p gpurun_out
python tools/prof_lz.py > gpurun_out/lz_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lanczos -s 2 -c 2 -o gpurun_out/prof_lz -f python tools/prof_lz.py > gpurun_out/ncu_lz.log 2>&1; tail -n 2 gpurun_out/ncu_lz.log

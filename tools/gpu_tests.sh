mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_preprocess.py --deselect tests/test_gpu_object_stats.py > gpurun_out/tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests2.log
tail -40 gpurun_out/tests2.log

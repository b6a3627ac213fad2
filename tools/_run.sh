timeout 600 python -m pytest tests/test_gpu_forms.py tests/test_gpu_forward.py tests/test_gpu_fullsize.py tests/test_gpu_train.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-train > gpurun_out/r3y_bench.json 2> gpurun_out/r3y.err; tail -2 gpurun_out/r3y.err
DFF_B200_NO_YFOLD=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-train > gpurun_out/r3y_bench_noy.json 2>> gpurun_out/r3y.err

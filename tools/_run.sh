timeout 600 python -m pytest tests/test_gpu_forms.py tests/test_gpu_forward.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-train > gpurun_out/r3z_bench.json 2> gpurun_out/r3z.err; tail -2 gpurun_out/r3z.err
DFF_B200_NO_I2=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-train > gpurun_out/r3z_bench_noi2.json 2>> gpurun_out/r3z.err

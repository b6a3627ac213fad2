timeout 500 python -m pytest tests/test_gpu_forms.py tests/test_gpu_forward.py tests/test_gpu_ops.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-train > gpurun_out/r3x_bench.json 2> gpurun_out/r3x.err; tail -2 gpurun_out/r3x.err

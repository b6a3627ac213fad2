timeout 400 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-train > gpurun_out/r3p_bench.json 2> gpurun_out/r3p.err; tail -3 gpurun_out/r3p.err

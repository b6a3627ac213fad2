timeout 400 python -m pytest tests/test_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_train.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 3 --no-train > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; tail -1 gpurun_out/r10_bench.err | cut -c1-200; head -c 160 gpurun_out/r10_bench.json

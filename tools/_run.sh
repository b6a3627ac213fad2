timeout 300 python -m pytest tests/test_e2e.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/bw_kernels.py 2>&1 | grep -i "fov_warp (B" -A6 | head -8
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-train > gpurun_out/r3n_bench.json 2> gpurun_out/r3n.err; tail -2 gpurun_out/r3n.err
timeout 200 ncu --set full --clock-control none -k regex:fov_warp_quad --launch-skip 3 --launch-count 1 -o gpurun_out/r3n_fov python tools/bw_kernels.py > gpurun_out/r3n_ncu.log 2>&1

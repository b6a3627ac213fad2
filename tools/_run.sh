timeout 500 python -m pytest tests/test_gpu_ops.py tests/test_e2e.py -x -q -k "fov or e2e or flow or alignment" 2>&1 | tail -4
timeout 200 python tools/bw_kernels.py 2>&1 | grep -i "fov_warp (B" -A4 | head -6
DFF_FOV_NO_ROWS=1 timeout 200 python tools/bw_kernels.py 2>&1 | grep -i "fov_warp (B" -A4 | head -6
timeout 200 ncu --set full --clock-control none -k regex:fov_warp_rows --launch-skip 3 --launch-count 1 -o gpurun_out/r3u_fov python tools/bw_kernels.py > gpurun_out/r3u_ncu.log 2>&1

timeout 600 python -m pytest tests/test_gpu_forms.py tests/test_gpu_forward.py -x -q 2>&1 | tail -3
timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_xg16.txt 2> gpurun_out/ops_xg16.err; tail -2 gpurun_out/ops_xg16.err
DFF_B200_XGROUP=3 timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_xg8.txt 2>&1
python tools/by_op.py --diff gpurun_out/ops_xg8.txt gpurun_out/ops_xg16.txt | grep -E "<<<|>>>|TOTAL"

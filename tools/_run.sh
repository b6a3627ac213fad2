timeout 600 python -m pytest tests/test_gpu_forms.py tests/test_gpu_forward.py tests/test_e2e.py -x -q 2>&1 | tail -4
timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_xg.txt 2> gpurun_out/ops_xg.err; tail -2 gpurun_out/ops_xg.err
DFF_B200_NO_XGROUP=1 timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_noxg.txt 2>&1
python tools/by_op.py --diff gpurun_out/ops_noxg.txt gpurun_out/ops_xg.txt | grep -E "<<<|>>>|TOTAL"

timeout 600 python -m pytest tests/test_gpu_forms.py tests/test_e2e.py -x -q 2>&1 | tail -2
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 65 --launch-count 1 -o gpurun_out/r8_xp_dres4c1 -f python tools/one_forward.py 16 10 384 576 bf16 > /dev/null 2>&1
DFF_B200_NO_XPAIR=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 65 --launch-count 1 -o gpurun_out/r8_noxp_dres4c1 -f python tools/one_forward.py 16 10 384 576 bf16 > /dev/null 2>&1
ls gpurun_out/r8_*dres4c1*

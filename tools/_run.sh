timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 66 --launch-count 3 -o gpurun_out/r5_wr_dres4c3 -f python tools/one_forward.py 16 10 384 576 bf16 > /dev/null 2>&1
DFF_B200_NO_WR=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on --launch-skip 66 --launch-count 3 -o gpurun_out/r5_nowr_dres4c3 -f python tools/one_forward.py 16 10 384 576 bf16 > /dev/null 2>&1
ls -la gpurun_out/r5_*dres4c3*

DFF_B200_ROW=1 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "row_kernel_child" 2>&1 | tail -2
for c in row_c3_16_16 row_c3_8+8_8 row_1x3x3_8_8 row_c3_32+32_32 row_c3_16_16_1slice; do DFF_B200_ROW=1 DFF_ROW_CHILD=$c timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "row_kernel_child" 2>&1 | tail -1; done
DFF_B200_ROW=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_ --csv --log-file gpurun_out/mb_row5.csv python tools/microbench_conv.py > /dev/null 2>&1
DFF_B200_ROW=1 DFF_ROW_EXPERIMENT=15 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_ --csv --log-file gpurun_out/mb_row5x.csv python tools/microbench_conv.py > /dev/null 2>&1

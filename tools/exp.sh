timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "row_kernel" 2>&1 | tail -4
for e in 0 1 2 4 8 15; do DFF_ROW_EXPERIMENT=$e ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_ --csv --log-file gpurun_out/mbr_$e.csv python tools/microbench_conv.py > /dev/null 2>&1; done

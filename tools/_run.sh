timeout 600 python -m pytest tests/test_gpu_forms.py -x -q -k wide_row 2>&1 | tail -5
DFF_B200_NO_WR=1 timeout 600 python -m pytest tests/test_gpu_forms.py -x -q -k wide_row 2>&1 | tail -3

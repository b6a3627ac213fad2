timeout 200 python tools/e2e_time.py 512 768 bf16 2>&1 | tail -1
timeout 200 python tools/e2e_time.py 512 768 bf16 2>&1 | tail -1
DFF_B200_NO_XGROUP=1 timeout 200 python tools/e2e_time.py 512 768 bf16 2>&1 | tail -1

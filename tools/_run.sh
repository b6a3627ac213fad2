timeout 600 python -m pytest tests/test_e2e.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "pair or warp or flow or mean" 2>&1 | tail -3
timeout 200 python tools/e2e_time.py 512 768 bf16 2>&1 | tail -1
timeout 200 python tools/e2e_time.py 512 768 bf16 2>&1 | tail -1
timeout 200 python tools/e2e_time.py 512 768 fp32 2>&1 | tail -1

timeout 300 python tools/train_profile.py bf16 1 2>&1 | tail -1
timeout 500 python -m pytest tests/test_gpu_forms.py tests/test_gpu_forward.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --e2e-steps 2 --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(round(d['value']), r['kernel'], round(r['frac'],3), round(r['aggregation_convs']['frac_of_tensor_peak'],3), [(o['name'][:22], round(o['ms_per_call'],3)) for o in r['operators_top12'][:9]])"

timeout 600 python -m pytest tests/test_gpu_forms.py -x -q -k "wide_row" 2>&1 | tail -5
timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_xp.txt 2> gpurun_out/ops_xp.err; tail -3 gpurun_out/ops_xp.err
DFF_B200_XPAIR_MAXC=8 timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_xp8.txt 2>&1
DFF_B200_NO_XPAIR=1 timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_noxp.txt 2>&1
python tools/by_op.py --diff gpurun_out/ops_noxp.txt gpurun_out/ops_xp.txt | grep -E "<<<|>>>|TOTAL"
echo ---- maxc 8
python tools/by_op.py --diff gpurun_out/ops_noxp.txt gpurun_out/ops_xp8.txt | grep -E "<<<|>>>|TOTAL"

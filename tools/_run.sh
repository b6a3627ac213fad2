for l in 256 128 64 0; do DFF_B200_TMAP_L2=$l timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_l2_$l.txt 2>&1; head -1 gpurun_out/ops_l2_$l.txt; done
python tools/by_op.py --diff gpurun_out/ops_l2_256.txt gpurun_out/ops_l2_128.txt | grep -E "<<<|>>>|TOTAL"
echo ---- 64
python tools/by_op.py --diff gpurun_out/ops_l2_256.txt gpurun_out/ops_l2_64.txt | grep -E "<<<|>>>|TOTAL"
echo ---- 0
python tools/by_op.py --diff gpurun_out/ops_l2_256.txt gpurun_out/ops_l2_0.txt | grep -E "<<<|>>>|TOTAL"

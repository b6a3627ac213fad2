timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_wr.txt 2> gpurun_out/ops_wr.err; tail -3 gpurun_out/ops_wr.err
DFF_B200_NO_WR=1 timeout 200 python tools/by_op.py 16 bf16 > gpurun_out/ops_nowr.txt 2>&1
python tools/by_op.py --diff gpurun_out/ops_nowr.txt gpurun_out/ops_wr.txt

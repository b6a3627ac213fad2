set -x
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; tail -3 gpurun_out/r3h_pytest.log
timeout 600 python bench.py > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; tail -2 gpurun_out/r3h_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r3h_launches_bf16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > /dev/null 2>&1
timeout 120 python tools/launch_by_layer.py gpurun_out/r3h_launches_bf16.csv 64 10 384 576 1 > gpurun_out/r3h_by_layer.txt 2>&1
DFF_B200_WGRAD_STREAM=0 DFF_B200_WGRAD_LOG=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3h_train_launches.csv python tools/train_profile.py bf16 0 > gpurun_out/r3h_train.log 2> gpurun_out/r3h_train.err
DFF_B200_WGRAD_STREAM=0 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_wgrad_tma --launch-skip 8 --launch-count 3 -o gpurun_out/r3h_wgrad python tools/train_profile.py bf16 0 > gpurun_out/r3h_ncu_wgrad.log 2>&1
timeout 200 python tools/bw_kernels.py --out gpurun_out/r3h_bw.json > gpurun_out/r3h_bw.log 2>&1
timeout 200 python tools/train_profile.py bf16 1 2>&1 | tail -1

timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4c_pytest.log 2>&1; tail -2 gpurun_out/r4c_pytest.log
timeout 900 python bench.py > gpurun_out/r4c_bench.json 2> gpurun_out/r4c_bench.err; tail -2 gpurun_out/r4c_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r4c_launches_bf16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > /dev/null 2>&1
timeout 120 python tools/launch_by_layer.py gpurun_out/r4c_launches_bf16.csv 64 10 384 576 1 > gpurun_out/r4c_by_layer.txt 2>&1
DFF_B200_WGRAD_STREAM=0 DFF_B200_WGRAD_LOG=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4c_train_launches.csv python tools/train_profile.py bf16 0 > gpurun_out/r4c_train.log 2> gpurun_out/r4c_train.err

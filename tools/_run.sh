timeout 600 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r3g_pytest.log 2>&1; tail -3 gpurun_out/r3g_pytest.log
timeout 300 python tools/train_profile.py bf16 1 2>&1 | tail -1
DFF_B200_WGRAD_STREAM=0 timeout 300 python tools/train_profile.py bf16 1 2>&1 | tail -1
DFF_B200_WGRAD_STREAM=0 DFF_B200_WGRAD_LOG=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3g_train_launches.csv python tools/train_profile.py bf16 0 > gpurun_out/r3g_train.log 2> gpurun_out/r3g_train.err

set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4a_pytest.log 2>&1; tail -3 gpurun_out/r4a_pytest.log
timeout 900 python bench.py > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; tail -2 gpurun_out/r4a_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r4a_launches_bf16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > /dev/null 2>&1
timeout 120 python tools/launch_by_layer.py gpurun_out/r4a_launches_bf16.csv 64 10 384 576 1 > gpurun_out/r4a_by_layer.txt 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_slab --launch-count 1 -o gpurun_out/r4a_fm0 python tools/one_forward.py 64 10 384 576 bf16 > gpurun_out/r4a_ncu1.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_slab --launch-skip 50 --launch-count 4 -o gpurun_out/r4a_dres4 python tools/one_forward.py 64 10 384 576 bf16 > gpurun_out/r4a_ncu2.log 2>&1

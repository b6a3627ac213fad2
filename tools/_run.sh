# the round's validation recipe (run on a GPU box from the repository root): GPU tests, the bench line, the launch list of the bench
# command and its per-operator join
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > /dev/null 2>&1
timeout 120 python tools/launch_by_layer.py gpurun_out/launches_bf16.csv 64 10 384 576 1 > gpurun_out/by_layer.txt 2>&1

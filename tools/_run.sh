timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; tail -3 gpurun_out/r9_pytest.log
timeout 900 python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; tail -2 gpurun_out/r9_bench.err | cut -c1-200; head -c 150 gpurun_out/r9_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r9_launches_bf16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > /dev/null 2>&1
timeout 120 python tools/launch_by_layer.py gpurun_out/r9_launches_bf16.csv 64 10 384 576 1 > gpurun_out/r9_by_layer.txt 2>&1; head -1 gpurun_out/r9_by_layer.txt

mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 60 python bench.py --steps 50 --no-e2e --no-cpu --workload config3 > gpurun_out/bench_config3.json 2>> gpurun_out/bench.err; python -c "import json;d=json.load(open('gpurun_out/bench_config3.json'));print('config3',d['ms_per_step'],d['roofline']['frac'])"

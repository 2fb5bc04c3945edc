#!/bin/bash
# One GPU session: tests, bench, ncu launch list + full capture of the synthesis kernel.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for wl in config3 config4 config5; do timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $wl > gpurun_out/bench_$wl.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_$wl.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 3 -c 2 -f -o gpurun_out/prof_synth python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out

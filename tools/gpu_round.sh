#!/bin/bash
# One GPU session: smoke, tests, bench (all configs), ncu launch list + full capture of the kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_reference.json
timeout 300 python bench.py > gpurun_out/bench.json 2>> gpurun_out/bench.err; cat gpurun_out/bench.json
for wl in config3 config4 config5 config2_stereo; do timeout 200 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu --workload $wl > gpurun_out/bench_$wl.json 2>> gpurun_out/bench.err; python -c "import json;d=json.load(open('gpurun_out/bench_$wl.json'));print('$wl',d['ms_per_step'],d['roofline']['frac'])"; done
tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_config4.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload config4 > gpurun_out/ncu_list4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 6 -c 1 -f -o gpurun_out/prof_synth python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 3 -c 1 -f -o gpurun_out/prof_synth_config3 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload config3 > gpurun_out/ncu_full3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tns_kernel -s 3 -c 1 -f -o gpurun_out/prof_tns python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload config4 > gpurun_out/ncu_full4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 3 -c 1 -f -o gpurun_out/prof_synth_config5 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload config5 > gpurun_out/ncu_full5.log 2>&1
ls gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:synth_kernel -s 6 -c 1 -f -o gpurun_out/prof_synth_config2_stereo python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload config2_stereo > gpurun_out/ncu_full2s.log 2>&1
ls gpurun_out

#!/bin/bash
# fused TNS: parity tests, A/B against the pre-pass on config 4, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
for f in 1 0; do
AACFB_TNS_FUSED=$f timeout 200 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu --no-configs --workload config4 > gpurun_out/bench_c4_f$f.json 2>> gpurun_out/bench.err
python -c "import json;d=json.load(open('gpurun_out/bench_c4_f$f.json'));print('config4 fused=$f', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"
done
tail -3 gpurun_out/bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:synth_tns -s 3 -c 1 -f -o gpurun_out/r4_c4fused python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload config4 > gpurun_out/ncu_c4f.log 2>&1; tail -1 gpurun_out/ncu_c4f.log | cut -c1-150
for c in 3 5; do
timeout 400 ncu --set full --clock-control none --import-source on -k synth_kernel -s 7 -c 1 -f -o gpurun_out/r4_config$c python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload config$c > gpurun_out/ncu_c$c.log 2>&1; tail -1 gpurun_out/ncu_c$c.log | cut -c1-120
done

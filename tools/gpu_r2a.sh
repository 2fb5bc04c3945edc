#!/bin/bash
# round-2 first session: sanitizer, the new bench line, fresh full captures of the shipped kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
tools/sanitize.sh racecheck synccheck
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for wl in config2 config3 config4 config5; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"synth_kernel|tns_kernel" -s 6 -c 3 -f -o gpurun_out/r3_$wl python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ncu_$wl.log 2>&1
  tail -1 gpurun_out/ncu_$wl.log | cut -c1-200
done
ls -la gpurun_out | tail -20

#!/bin/bash
# refresh after the swizzle change: default bench line + captures of the generic instantiation (config 3, 5)
mkdir -p gpurun_out
SECONDS=0; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench: $SECONDS s"; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
cap() {  # tag workload kernel-regex skip
  tag=$1; wl=$2; k=$3; skip=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -f -o gpurun_out/cap_$tag \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ncu_cap_$tag.log 2>&1
  grep -c "==PROF== Profiling" gpurun_out/ncu_cap_$tag.log | sed "s/^/cap_$tag launches captured: /"
}
cap config3 config3 "synth_kernel" 7
cap config5 config5 "synth_kernel" 7
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_config5.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --workload config5 > gpurun_out/ncu_list_config5.log 2>&1

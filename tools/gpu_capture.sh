#!/bin/bash
# One `ncu --set full` capture (one launch) of the dominant kernel of every bench configuration, plus the
# launch lists of config 2 / 4 / 5.  Run under gpurun in two halves (gpurun returns at most 64 MiB):
#   tools/gpu_capture.sh a   -> config2, config3, config5          tools/gpu_capture.sh b -> config4 (both kernels),
#   config2_stereo, launch lists.   Then `python tools/refresh_profiles.py <round tag>` here.
mkdir -p gpurun_out
cap() {  # tag workload kernel-regex skip [bench args]
  tag=$1; wl=$2; k=$3; skip=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -f -o gpurun_out/cap_$tag \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl "$@" > gpurun_out/ncu_cap_$tag.log 2>&1
  grep -c "==PROF== Profiling" gpurun_out/ncu_cap_$tag.log | sed "s/^/cap_$tag launches captured: /"
}
# launches per step: synth_kernel<long-only>, synth_kernel<generic> (exits at once unless an item has an EIGHT_SHORT
# frame); config 4: tns_kernel first.  -k matches the base name, so the instantiation is picked by the skip count.
if [ "${1:-a}" = "a" ]; then
cap config2 config2 "synth_kernel" 6
cap config3 config3 "synth_kernel" 7
cap config5 config5 "synth_kernel" 7
else
cap config4_tns config4 "tns_kernel" 3
cap config4_synth config4 "synth_kernel" 6
cap config2_stereo config2_stereo "synth_kernel" 6
for wl in config2 config4 config5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$wl.csv \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ncu_list_$wl.log 2>&1
  grep -c "aacfb::" gpurun_out/launches_$wl.csv | sed "s/^/launches_$wl rows: /"
done
fi

#!/bin/bash
# final kernels: smoke, GPU tests, the driver's bench invocation, captures (half $1)
mkdir -p gpurun_out
cap() {  # tag workload kernel-regex skip
  tag=$1; wl=$2; k=$3; skip=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -f -o gpurun_out/cap_$tag \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ncu_cap_$tag.log 2>&1
  grep -c "==PROF== Profiling" gpurun_out/ncu_cap_$tag.log | sed "s/^/cap_$tag launches captured: /"
}
if [ "${1:-a}" = "a" ]; then
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-260 gpurun_out/bench.json
cap config2 config2 "synth_kernel" 6
cap config5 config5 "synth_kernel" 7
else
cap config3 config3 "synth_kernel" 7
for wl in config2 config5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$wl.csv \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs --workload $wl > gpurun_out/ncu_list_$wl.log 2>&1
done
AACFB_TNS_FUSED=1 timeout 600 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread -k "tns or stereo" 2>&1 | tail -1
fi

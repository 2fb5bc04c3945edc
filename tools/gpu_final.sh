#!/bin/bash
# Round-end evidence on one GPU: smoke, GPU tests, the default bench line, the reference arm, captures.
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
fi
SECONDS=0; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench default: $SECONDS s"; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
SECONDS=0; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "bench reference: $SECONDS s"; tail -1 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_reference.json
bash tools/gpu_capture.sh a

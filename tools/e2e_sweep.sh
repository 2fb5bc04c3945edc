#!/bin/bash
# e2e (host buffers through aacfb_process_io / aacfb_process) for a few pipeline shapes: sub-batches:lanes
mkdir -p gpurun_out
for cfg in ${@:-8:2 16:2 32:2 16:4 32:4 64:4}; do
  sub=${cfg%%:*}; lanes=${cfg##*:}
  AACFB_SUB_BATCHES=$sub AACFB_LANES=$lanes timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --e2e-steps 25 > gpurun_out/e2e.json 2>> gpurun_out/bench.err
  python -c "
import json;d=json.load(open('gpurun_out/e2e.json'));e=d['e2e'];v=e.get('variants',{})
print('sub %3d lanes %d: e2e %.3f ms %.3f Mframes/s | ' % ($sub,$lanes,e['ms_per_step'],e['value']/1e6) + '  '.join('%s %.3f' % (k, x['ms_per_step']) for k, x in v.items()))"
done

timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for sub in 0 8 16; do
  if [ $sub = 0 ]; then unset AACFB_SUB_BATCHES; else export AACFB_SUB_BATCHES=$sub; fi
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 25 > gpurun_out/e2e.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/e2e.json'));e=d['e2e'];print('sub $sub: e2e %.3f ms  %.3f Mframes/s' % (e['ms_per_step'],e['value']/1e6))"
done; done

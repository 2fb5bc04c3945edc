AACFB_LIB=$PWD/aac.js_b200/libaacfb_twsym.so timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -1
for wl in config2 config5; do for lib in - libaacfb_twsym.so; do
  if [ "$lib" = "-" ]; then unset AACFB_LIB; else export AACFB_LIB=$PWD/aac.js_b200/$lib; fi
  timeout 100 python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu --workload $wl > gpurun_out/ab.json 2>> gpurun_out/bench.err
  python -c "import json;d=json.load(open('gpurun_out/ab.json'));print('$wl %-22s %.4f ms  frac %.3f' % ('$lib', d['ms_per_step'],d['roofline']['frac']))"
done; done

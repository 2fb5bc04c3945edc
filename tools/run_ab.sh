timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
bash tools/ab.sh "- libaacfb_nopl.so libaacfb_noswz.so libaacfb_nopl_noswz.so libaacfb_head.so" config3
bash tools/ab.sh "- libaacfb_nopl.so libaacfb_head.so" config5

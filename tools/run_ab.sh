timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash tools/ab.sh "- libaacfb_head.so" config2 config3 config5

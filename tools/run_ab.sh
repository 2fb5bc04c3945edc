timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
bash tools/ab.sh "- libaacfb_head.so" config5 config2

python -m pytest tests/test_gpu_mser.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
TOPK=10 python tools/mser_probe.py 2>&1 | grep -v "pixels per level" | head -14

python -m pytest tests/test_match_gpu.py -m gpu -x -q 2>&1 | tail -3

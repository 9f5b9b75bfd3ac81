python -m pytest tests/test_host_shim.py -m gpu -x -q 2>&1 | tail -2

BSHOT_LIB=$PWD/b-shot-slam_b200/libbshot_b200_bounds.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3

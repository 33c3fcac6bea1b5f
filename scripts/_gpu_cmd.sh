timeout 600 python -m pytest tests/test_learn_gpu.py tests/test_project_tc_gpu.py tests/test_project_gpu.py -m gpu -x -q 2>&1 | tail -4
python scripts/cov_time.py 32

timeout 300 python -m pytest tests/test_learn_gpu.py -m gpu -x -q 2>&1 | tail -5
python bench.py --only-cpp-index --cpp-index-tracks 256 --steps 2 2>&1 | grep -E "frames_per_s|trace|_s\"" | head -40
python bench.py --only-cpp-index --cpp-index-tracks 1024 --steps 2 2>&1 | grep -E "frames_per_s|trace|_s\"" | head -40

timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_cpp_api.py tests/test_stream_order_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --only-cpp-index --cpp-index-tracks 1024 --steps 3 2>&1 | grep -E "\"frames_per_s|\"index_|enqueued|learned" | tail -14

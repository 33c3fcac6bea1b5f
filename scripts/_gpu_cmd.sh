timeout 300 python -m pytest tests/test_learn_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in "" "HPFW_DECODE_THREADS=12" "HPFW_DECODE_THREADS=24" "HPFW_CACHE_SPECTROGRAMS=0" "HPFW_CACHE_SPECTROGRAMS=0 HPFW_DECODE_THREADS=32"; do
echo "== $v"; env $v python bench.py --only-cpp-index --cpp-index-tracks 1024 --steps 2 2>&1 | grep -E "\"frames_per_s\"|enqueued|calc_filters" | tail -7
done

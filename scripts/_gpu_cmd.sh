timeout 600 python -m pytest tests/test_match_selftest_gpu.py tests/test_matcher_gpu.py tests/test_shard_gpu.py -m gpu -x -q 2>&1 | tail -4
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('k','impl3_ms','impl3_gwordops','impl1_gwordops','impl0_gwordops','equal_fp4','equal','fp4_tops_equiv')})"; }
for k in 63 143 385 1514; do timeout 300 python scripts/match_tc_time.py 2000 128 $k 2>&1 | tail -1 | tee -a gpurun_out/r2i_match_tc_time.jsonl | show; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpp --no-extraction > gpurun_out/r2i_bench_quick.json 2> gpurun_out/r2i_bench_quick.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_bench_quick.json"))
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","impls_bit_identical","top1_ok","clocks")}))
print(d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["achieved"], d["strong"]["single_find_ms"])
PY

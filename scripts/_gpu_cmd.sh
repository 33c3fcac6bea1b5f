timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo rc=$?
tail -c 300 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j_bench.json"))
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","impls_bit_identical","top1_ok","clocks")}))
print(d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline_popc"]["frac"])
print(json.dumps(d["e2e_cpp"]["search"])[:400]); print(d["e2e_cpp"]["index"].get("frames_per_s"), d["e2e_cpp"]["index"].get("index_s"), d["extraction"]["value"])
PY
C="python bench.py --tracks 2000 --steps 2 --warmup 1 --no-cpp --no-extraction --no-cpu-baseline --no-strong-leg --no-popc-leg"
$C > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:match_tc_kernel -s 1 -c 2 -o gpurun_out/r2j_match_tc $C > gpurun_out/ncu2.log 2>&1
echo full rc=$?

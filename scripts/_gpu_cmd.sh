timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --tracks 2000 --steps 2 --warmup 1 --no-cpp --extract-tracks 16 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches_bench_tracks2000.csv $B > gpurun_out/ncu1.log 2>&1
echo launchlist rc=$?
C="python bench.py --tracks 2000 --steps 2 --warmup 1 --no-cpp --no-extraction --no-cpu-baseline --no-strong-leg --no-popc-leg"
$C > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:match_tc_kernel -s 1 -c 2 -o gpurun_out/r2f_match_tc $C > gpurun_out/ncu2.log 2>&1
echo full rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/plain.log").read().strip().splitlines()[-1])
print(d["value"], d["strong"]["single_find_ms"])
PY
ls -la gpurun_out | tail -8

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2h_bench_4gpu.json 2> gpurun_out/r2h_bench_4gpu.err; echo rc=$?
tail -c 400 gpurun_out/r2h_bench_4gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_4gpu.json"))
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","n_gpus","sharded_bit_identical","top1_ok","strong")})[:1500])
print(d["e2e"]["value"], json.dumps(d["e2e_cpp"]["search"])[:300], d["e2e_cpp"]["index"].get("frames_per_s"), d["extraction"]["value"])
PY

nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -x -q 2>&1 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; echo rc=$?
tail -c 1200 gpurun_out/r2d_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2d_bench_2gpu.json"))
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","n_gpus","sharded_bit_identical","top1_ok","strong")})[:1500])
print(json.dumps(d.get("e2e"))[:600])
print(json.dumps(d.get("e2e_cpp"), indent=1)[:3000])
print(d["extraction"]["value"], d["extraction"].get("index",{}).get("cov_allreduce_ms"))
PY

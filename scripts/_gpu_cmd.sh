nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2e_bench_8gpu.json 2> gpurun_out/r2e_bench_8gpu.err; echo rc=$?
tail -c 600 gpurun_out/r2e_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 scripts/config_sweep.py sharded > gpurun_out/r2e_config_sweep_8gpu.jsonl 2> gpurun_out/r2e_config_sweep_8gpu.err; echo rc=$?
cat gpurun_out/r2e_config_sweep_8gpu.jsonl; tail -c 600 gpurun_out/r2e_config_sweep_8gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench_8gpu.json"))
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","n_gpus","sharded_bit_identical","top1_ok","strong","clocks")})[:1800])
print(json.dumps(d.get("e2e"))[:500])
print(json.dumps(d.get("e2e_cpp"), indent=1)[:2500])
print(d["extraction"]["value"], d["roofline"]["frac"])
PY

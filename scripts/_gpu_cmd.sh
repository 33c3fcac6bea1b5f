N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2k_bench_${N}gpu.json 2> gpurun_out/r2k_bench_${N}gpu.err; echo rc=$?
tail -c 300 gpurun_out/r2k_bench_${N}gpu.err
true
python - <<PY
import json
d=json.load(open("gpurun_out/r2k_bench_${N}gpu.json"))
print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","n_gpus","sharded_bit_identical","top1_ok")}))
print(d["e2e"]["value"], d["roofline"]["frac"], json.dumps(d["strong"])[:400])
print(json.dumps(d["e2e_cpp"]["search"])[:260]); print(d["e2e_cpp"]["index"].get("frames_per_s"), d["extraction"]["value"])
PY

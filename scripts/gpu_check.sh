#!/bin/bash
# Run on the GPU box via gpurun: parity tests, smoke, bench, ncu launch list + one full capture of the match kernel.
# Usage: scripts/gpu_check.sh [tag]   (NCU=0 skips the profiler passes)
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 2 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"; cat $OUT/bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tracks 2000 --no-extraction"
# launch list: this library's kernels only (torch's setup kernels would fill the list before the timed step is reached)
echo "== ncu launch list"; timeout 600 $CMD > $OUT/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(match_|topk_|merge_|xt_|czt_|fft_pass|db_kernel|tc_delta|project_|table_fill|bl_|pipe_kernel)" -c 600 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1; echo "ncu list exit $?"
CMD_TC="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tracks 2000 --no-extraction --no-popc-leg"
echo "== ncu full (tensor-core match kernel)"; timeout 600 $CMD_TC > $OUT/ncu_plain_tc.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_tc_kernel -s 3 -c 1 -o $OUT/match_tc_full $CMD_TC > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
fi
ls -la $OUT

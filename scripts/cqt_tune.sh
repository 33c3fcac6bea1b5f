#!/bin/bash
# CQT tuning sweep (run on the GPU box)
run() { env "$@" timeout 120 python scripts/cqt_tune.py 48 2>&1 | tail -1; }
for n1 in 400 500 700 1000 1300 1500 2000; do
  for g1 in 4 8 16; do
    run HPFW_CQT_N1MAX=$n1 HPFW_CQT_G1=$g1
  done
done
run HPFW_CQT_N1MAX=1000 HPFW_CQT_G2=1
run HPFW_CQT_N1MAX=1000 HPFW_CQT_G2=2
run HPFW_CQT_N1MAX=1000 HPFW_CQT_LANES=2
run HPFW_CQT_N1MAX=1000 HPFW_CQT_LANES=8

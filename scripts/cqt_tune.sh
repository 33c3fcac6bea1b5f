#!/bin/bash
# CQT tuning sweep (run on the GPU box)
run() { env "$@" timeout 120 python scripts/cqt_tune.py 48 2>&1 | tail -1; }
run HPFW_CQT_LANES=4
run HPFW_CQT_LANES=8
run HPFW_CQT_LANES=2
run HPFW_CQT_LANES=1
run HPFW_CQT_N1MAX=1000
run HPFW_CQT_N1MAX=1000 HPFW_CQT_G1=4
run HPFW_CQT_N1MAX=1000 HPFW_CQT_G1=8
run HPFW_CQT_G1=4
run HPFW_CQT_G2=4
run HPFW_CQT_N1MAX=600
run HPFW_CQT_N1MAX=600 HPFW_CQT_G1=8
run HPFW_CQT_N1MAX=1500

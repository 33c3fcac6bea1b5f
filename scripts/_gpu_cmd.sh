echo "== base"; python scripts/cqt_tune.py 48 2>&1 | tail -1
for kb in 37 28 74; do echo "== t128 smem_kb $kb"; HPFW_CQT_SMEM_KB=$kb HPFW_B200_LIB=hpfw_b200/libhpfw_b200_t128.so python scripts/cqt_tune.py 48 2>&1 | tail -1; done
for g in "HPFW_CQT_G1=2 HPFW_CQT_G2=2" "HPFW_CQT_G1=4 HPFW_CQT_G2=2" "HPFW_CQT_G1=2 HPFW_CQT_G2=1"; do echo "== t128 $g"; env $g HPFW_B200_LIB=hpfw_b200/libhpfw_b200_t128.so python scripts/cqt_tune.py 48 2>&1 | tail -1; done
echo "== t128c5"; HPFW_CQT_SMEM_KB=44 HPFW_B200_LIB=hpfw_b200/libhpfw_b200_t128c5.so python scripts/cqt_tune.py 48 2>&1 | tail -1
echo "== t512"; HPFW_CQT_SMEM_KB=200 HPFW_B200_LIB=hpfw_b200/libhpfw_b200_t512.so python scripts/cqt_tune.py 48 2>&1 | tail -1

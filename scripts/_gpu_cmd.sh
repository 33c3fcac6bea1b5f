echo "== base"; python scripts/cqt_tune.py 48 2>&1 | tail -1
echo "== u2"; HPFW_B200_LIB=hpfw_b200/libhpfw_b200_u2.so python scripts/cqt_tune.py 48 2>&1 | tail -1
echo "== u2 maxradix8"; HPFW_CQT_MAXRADIX=8 HPFW_B200_LIB=hpfw_b200/libhpfw_b200_u2.so python scripts/cqt_tune.py 48 2>&1 | tail -1
echo "== u2c2"; HPFW_B200_LIB=hpfw_b200/libhpfw_b200_u2c2.so python scripts/cqt_tune.py 48 2>&1 | tail -1
echo "== u2c2 smem 110"; HPFW_CQT_SMEM_KB=110 HPFW_B200_LIB=hpfw_b200/libhpfw_b200_u2c2.so python scripts/cqt_tune.py 48 2>&1 | tail -1

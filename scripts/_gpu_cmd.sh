for v in "" chain nopad both ctas4 ctas2; do
  if [ -z "$v" ]; then lib=hpfw_b200/libhpfw_b200.so; else lib=hpfw_b200/libhpfw_b200_$v.so; fi
  echo "== variant '$v'"; HPFW_B200_LIB=$lib python scripts/cqt_tune.py 48 2>&1 | tail -1
done
echo "== lanes"; for l in 2 3 6; do HPFW_CQT_LANES=$l python scripts/cqt_tune.py 48 2>&1 | tail -1; done
echo "== G1/G2"; HPFW_CQT_G1=2 python scripts/cqt_tune.py 48 2>&1 | tail -1; HPFW_CQT_G1=8 python scripts/cqt_tune.py 48 2>&1 | tail -1; HPFW_CQT_G2=2 python scripts/cqt_tune.py 48 2>&1 | tail -1

echo "== base"; HPFW_CQT_DEBUG=1 python scripts/cqt_tune.py 48 2>&1 | grep -E "hpfw cqt|us/track" | sort -u | tail -3
for t in 128 160 192 224; do echo "== T2=$t"; HPFW_CQT_T2=$t python scripts/cqt_tune.py 48 2>&1 | tail -1; done
for t in 192 224; do echo "== T1=$t"; HPFW_CQT_T1=$t python scripts/cqt_tune.py 48 2>&1 | tail -1; done
echo "== auto"; HPFW_CQT_TAUTO=1 HPFW_CQT_DEBUG=1 python scripts/cqt_tune.py 48 2>&1 | grep -E "hpfw cqt|us/track" | sort -u | tail -3

timeout 300 python scripts/project_time.py 32 0,3,4,5 2>&1 | tail -9
timeout 300 python scripts/project_time.py 128 3,4,5 2>&1 | tail -5

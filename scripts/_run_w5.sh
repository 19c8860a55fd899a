O2V_OCC=0 O2V_TIMED=3 python scripts/profile_run.py cfg4 1 2>&1 | tail -2
O2V_TIMED=5 python scripts/profile_run.py cfg3 1 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5

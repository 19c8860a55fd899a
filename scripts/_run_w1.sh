set -x
O2V_OCC=0 O2V_TIMED=3 python scripts/profile_run.py cfg4 1 2>&1 | tail -3
O2V_TIMED=5 python scripts/profile_run.py cfg3 1 2>&1 | tail -3
O2V_OCC=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sparse|emitLeaves|countLeaves|sortSmall|compactActive' -c 10 -o gpurun_out/r4a_weighted_cfg4 python scripts/profile_run.py cfg4 1 > gpurun_out/r4a_ncu.log 2>&1
tail -3 gpurun_out/r4a_ncu.log

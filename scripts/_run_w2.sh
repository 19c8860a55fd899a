timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sparse|emitLeaves|countLeaves|sortSmall|compactActive' -c 10 -o gpurun_out/r4b_weighted_cfg3 python scripts/profile_run.py cfg3 1 > gpurun_out/r4b_ncu.log 2>&1
tail -3 gpurun_out/r4b_ncu.log

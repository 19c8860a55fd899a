timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sparseDualClip' -c 1 -o gpurun_out/r4c_dual_cfg3 python scripts/profile_run.py cfg3 1 > gpurun_out/r4c_ncu.log 2>&1
tail -2 gpurun_out/r4c_ncu.log

O2V_OCC=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sparseTinyFold|sparseFoldKernel' -c 2 -o gpurun_out/r4d_folds_cfg4 python scripts/profile_run.py cfg4 1 > gpurun_out/r4d_ncu.log 2>&1
tail -2 gpurun_out/r4d_ncu.log

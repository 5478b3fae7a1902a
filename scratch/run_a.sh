set -x
cd $GRAFT_REPO_ROOT
TMB_FOLD_TIMING=1 python profiles/c4_probe.py 2000 2 > gpurun_out/r2_fold3_timing.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fold -c 1 -o gpurun_out/r2_prof_fold3 -f python profiles/c4_probe.py 2000 1 > gpurun_out/r2_ncu5.log 2>&1
tail -5 gpurun_out/r2_fold3_timing.log

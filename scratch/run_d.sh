cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fold -c 1 -o gpurun_out/r2_prof_fold4 -f python profiles/c4_probe.py 2000 1 > gpurun_out/r2_ncu6.log 2>&1
tail -3 gpurun_out/r2_ncu6.log

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest7.log
echo "== fold4 1536x3"; TMB_FOLD_TIMING=1 timeout 300 python profiles/c4_probe.py 2000 2 2>&1 | tail -4
echo "== fold4 3072x2"; TMB_FOLD_CH=3072 TMB_FOLD_TIMING=1 timeout 300 python profiles/c4_probe.py 2000 2 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:k_fold -c 1 -o gpurun_out/r2_prof_fold4 -f python profiles/c4_probe.py 2000 1 > gpurun_out/r2_ncu6.log 2>&1

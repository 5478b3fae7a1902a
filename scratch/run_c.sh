cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest6.log
echo "== fold4"; timeout 300 python profiles/c4_probe.py 2000 3 2>&1 | tail -3
echo "== fold4 timing"; TMB_FOLD_TIMING=1 timeout 300 python profiles/c4_probe.py 2000 2 2>&1 | tail -4

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final2.log
timeout 200 python profiles/c4_probe.py 2000 3 2>&1 | tail -2 | cut -c1-250

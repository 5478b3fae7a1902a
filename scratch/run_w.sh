cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_final.log
timeout 400 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err
timeout 200 python bench.py --impl reference > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"

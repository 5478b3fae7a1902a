# The round-2 capture run (through gpurun): default bench, the CPU arm, one ncu --set full pass over the step kernels of
# profiles/c4_probe.py and the ncu launch list of a short bench run (step kernels only).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
RX='regex:k_hist|k_scan|k_tile|k_scatter|k_fold|k_gate|k_classify|k_clear|k_inflate'
( time timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | grep real
timeout 600 ncu --set full --clock-control none --import-source on -k "$RX" --launch-skip 10 -c 10 -o gpurun_out/r2_prof_c4_final -f python profiles/c4_probe.py 2000 2 > gpurun_out/r2_ncu9.log 2>&1; tail -1 gpurun_out/r2_ncu9.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2_ncu10.log 2>&1; tail -1 gpurun_out/r2_ncu10.log | cut -c1-300

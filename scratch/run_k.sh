cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fold3 -c 1 -o gpurun_out/r2_prof_fold3_shard -f python profiles/shard_probe.py 2000 1 > gpurun_out/r2_ncu8.log 2>&1
tail -2 gpurun_out/r2_ncu8.log

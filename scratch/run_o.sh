cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest11.log
timeout 300 python profiles/c4_probe.py 2000 2 2>&1 | tail -1 | cut -c1-250
timeout 300 python profiles/shard_probe.py 2000 3 2>&1 | grep rebuild | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 profiles/shard_probe.py 2000 5 2>&1 | grep -v "^\*\|OMP" | tail -4

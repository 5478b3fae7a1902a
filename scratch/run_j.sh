cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "shard" > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest9.log
timeout 300 python profiles/shard_probe.py 2000 3 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 profiles/shard_probe.py 2000 4 2>&1 | tail -3

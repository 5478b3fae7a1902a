cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TMB_SHARD_DEBUG=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 profiles/shard_probe.py 2000 4 > gpurun_out/r2_shard_dbg.log 2>&1; echo "rc=$?"
grep -c "rebuild done" gpurun_out/r2_shard_dbg.log; grep "rebuild [0-9]" gpurun_out/r2_shard_dbg.log | tail -4; grep "\[shard" gpurun_out/r2_shard_dbg.log | tail -6
echo "== nccl path"
TMB_SHARD_ROUTE=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 profiles/shard_probe.py 2000 4 2>&1 | grep "rebuild [0-9]" | tail -2

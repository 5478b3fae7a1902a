cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TMB_SHARD_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29583 tests/shard_gpu_check.py 12 inflate > gpurun_out/r2_shard_infl.log 2>&1; echo rc=$?
grep -v "^\*\|OMP" gpurun_out/r2_shard_infl.log | grep -v "^$" | tail -25

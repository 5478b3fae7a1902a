cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tests/shard_gpu_check.py 24 inflate 2>&1 | grep -v "^\*\|OMP" | tail -8
timeout 600 python -m pytest tests -m gpu -x -q -k "shard" 2>&1 | tail -3

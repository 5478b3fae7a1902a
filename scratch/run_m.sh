cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 profiles/shard_probe.py 2000 6 2>&1 | grep rebuild | tail -4 | cut -c1-200; }
run "TMB_SPLIT_DIV=1"
run "TMB_SPLIT_DIV=4"
run "TMB_SPLIT_DIV=4 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32"

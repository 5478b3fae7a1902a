cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','ms_all','n_gpus') if k in d}, d.get('shard_parity'), d.get('sharded',{}).get('stages_ms_rank0'), d.get('sharded',{}).get('alltoall_gbs_rank0_out'))
except Exception as e: print("parse fail", e)
PY

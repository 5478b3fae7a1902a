cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "c4 rc=$?"
timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --steps 6 --warmup 3 --workload c5 > gpurun_out/r2_bench_c5_n8.json 2> gpurun_out/r2_bench_c5_n8.err; echo "c5 rc=$?"; tail -3 gpurun_out/r2_bench_c5_n8.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n8.json','gpurun_out/r2_bench_c5_n8.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','ms_per_step','ms_all') if k in d}, d.get('shard_parity'), d.get('sharded',{}).get('stages_ms_rank0'), d.get('sharded',{}).get('grid_bytes_rank0'), d.get('e2e',{}).get('ms_per_step'))
    except Exception as e: print(f, "parse fail", e)
PY

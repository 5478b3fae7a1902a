cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "n4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_all') if k in d}, d.get('shard_parity'), d.get('sharded',{}).get('stages_ms_rank0'), d.get('roofline',{}).get('frac'))
PY

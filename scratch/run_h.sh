cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for d in 1 2 4; do
echo "== N=1 split div $d"; TMB_SPLIT_DIV=$d timeout 300 python profiles/c4_probe.py 2000 2 2>&1 | tail -2 | cut -c1-260
done
for d in 2 4; do
echo "== N=2 split div $d"
TMB_SPLIT_DIV=$d timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_d$d.json 2> gpurun_out/r2_bench_n2.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2_d$d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_all') if k in d}, d.get('shard_parity'), d.get('sharded',{}).get('stages_ms_rank0'))
PY
done

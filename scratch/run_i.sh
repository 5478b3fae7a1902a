cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/shard_probe.py 2000 3 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_shard_w1.csv python profiles/shard_probe.py 2000 2 > gpurun_out/r2_ncu7.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_shard_w1.csv')) if len(r)>10]
hdr=rows[0]; iN=hdr.index('Kernel Name'); iV=hdr.index('Metric Value'); iU=hdr.index('Metric Unit')
out=[]
for r in rows[1:]:
    v=float(r[iV].replace(',','')); u=r[iU]
    ms = v/1e6 if u in ('ns','nsecond') else v/1e3 if u in ('us','usecond') else v
    out.append((r[iN][:60], ms))
for n,ms in out[-40:]:
    if ms>0.02: print("%-62s %.3f ms"%(n,ms))
PY

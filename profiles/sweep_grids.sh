for g in 2 3 4 6 8; do
  TMB_GATE_GRID=$g timeout 100 python bench.py --no-cpu --no-lc 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms_per_step']; print('gate grid $g', round(d['ms_per_step'],4), 'gate', round(s['ms_gate'],4), 'cls', round(s['ms_classify'],4))"
done
for g in 2 3 4 5 6 9; do
  TMB_CLS_GRID=$g timeout 100 python bench.py --no-cpu --no-lc 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms_per_step']; print('cls grid $g', round(d['ms_per_step'],4), 'gate', round(s['ms_gate'],4), 'cls', round(s['ms_classify'],4))"
done
for g in 1 2 3 4 6; do
  TMB_INF_GRID=$g timeout 100 python bench.py --no-cpu --no-lc 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms_per_step']; print('inf grid $g', round(d['ms_per_step'],4), 'infl', round(s['ms_inflate'],4))"
done

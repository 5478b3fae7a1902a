"""The bench's C4 step (bench.bench_c4_single's inner loop) as a standalone probe for ncu captures and kernel
experiments: N keyframes x 262 144 rays on the 180 m loop, drifted build, then `reps` loop closures alternating
between the true and the drifted trajectory.  python profiles/c4_probe.py [n_keyframes] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import traversability_mapping_b200 as tm
from traversability_mapping_b200 import synth

nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
true, drift = bench.c4_poses(bench.C4_KEYFRAMES)
true, drift = true[:nk], drift[:nk]
p = tm.default_params(max_batch=0, profile=1, **bench.C4)
s = tm.System(p); s.addNewLocalMap(0); s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
m = s.getLocalMap(0)
t0 = time.perf_counter()
bench.ingest(tm, s, range(nk), true, bench.C4_RINGS, bench.C4_AZ, "cuda")
print("ingest %.1f s" % (time.perf_counter() - t0))
ids = np.arange(1, nk + 1, dtype=np.uint64)
with m.locked():
    s.updateKeyFrames(ids, drift)
s.flush()
for rep in range(reps):
    target = true if rep % 2 == 0 else drift
    t0 = time.perf_counter()
    with m.locked():
        s.updateKeyFrames(ids, target)
        s.informLoopClosure()
    s.flush()
    dt = time.perf_counter() - t0
    st = m.stats()
    print("closure %d: wall %.2f ms, gpu %.2f ms, points %d, bins %d (heaviest %d), touched %d, classified %d | hist %.3f scan %.3f scatter %.3f "
          "fold %.3f gate %.3f classify %.3f | host plan/launch/wait %.2f" % (
              rep, dt * 1e3, st["ms_total"], st["last_batch_points"], st["last_batch_tiles"], st["last_batch_max_bin"],
              st["last_batch_cells_touched"], st["last_batch_cells_classified"], st["ms_hist"], st["ms_scan"], st["ms_scatter"], st["ms_fold"],
              st["ms_gate"], st["ms_classify"] - st["ms_gate"], st["host_ms_batch"]))
s.close()

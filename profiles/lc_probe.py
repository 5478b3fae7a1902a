"""Loop-closure probe used for the ncu captures: N keyframes on the 180 m loop (0.1 m grid), one
informLoopClosure rebuild in batched mode.  python profiles/lc_probe.py [n_keyframes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import traversability_mapping_b200 as tm
from traversability_mapping_b200 import synth

nk = int(sys.argv[1]) if len(sys.argv) > 1 else 300
base = 16
poses = synth.trajectory("loop", nk, seed=3, extent=90.0 * nk / 2000.0 + 5.0)
clouds = [synth.make_keyframe(3, i, poses[i], max_range=20.0) for i in range(base)]
p = tm.default_params(max_batch=0, profile=1, resolution=0.1, half_size=100.0, max_range=20.0, min_range=1.0, robot_radius=0.20)
s = tm.System(p); s.addNewLocalMap(0); s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
m = s.getLocalMap(0)
for i in range(nk):
    s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i % base])
with m.locked():   # queue everything while the worker is held off: batch 1 = 1 keyframe, batch 2 = the rest
    for i in range(nk):
        s.updateKeyFrame(i + 1, poses[i])
s.flush()
for rep in range(2):
    t0 = time.perf_counter(); s.informLoopClosure(); s.flush(); dt = time.perf_counter() - t0
    st = m.stats()
    print("rebuild %d: wall %.2f ms, gpu %.2f ms, points %d, bins %d (heaviest %d), touched %d, classified %d | hist %.3f scan %.3f scatter %.3f fold %.3f classify %.3f" % (
        rep, dt * 1e3, st["ms_total"], st["last_batch_points"], st["last_batch_tiles"], st["last_batch_max_bin"], st["last_batch_cells_touched"],
        st["last_batch_cells_classified"], st["ms_hist"], st["ms_scan"], st["ms_scatter"], st["ms_fold"], st["ms_classify"]))
s.close()

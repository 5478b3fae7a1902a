"""Sharded rebuild vs the single-GPU rebuild, bit for bit.  Run under torchrun with N ranks (N GPUs):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/shard_gpu_check.py [n_kf]
Rank 0 additionally builds the same map on one GPU through the ordinary API (max_batch=0) and compares
every moment layer and derived layer of the gathered slabs with it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import traversability_mapping_b200 as tm  # noqa: E402
from traversability_mapping_b200 import sharding, synth  # noqa: E402
import helpers  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n_kf = int(sys.argv[1]) if len(sys.argv) > 1 else 24
P = dict(resolution=0.1, half_size=20.0, max_range=10.0, min_range=1.0, robot_radius=0.20, min_points_per_grid=3)
INFLATE = len(sys.argv) > 2 and sys.argv[2] == "inflate"
if INFLATE:   # step inflation on (LocalMap::inflateDirty): the step layer's halo rows travel between the slabs too
    P.update(inflation_enabled=1, inflation_radius=0.5, cost_scaling_factor=3.0, lethal_step_threshold=0.3)
poses = synth.trajectory("loop", n_kf, seed=9, extent=14.0)
clouds = [synth.make_keyframe(9, i, poses[i], max_range=10.0, rings=64, azimuths=1024) for i in range(n_kf)]

sm = sharding.ShardedMap(tm.default_params(device=local, max_batch=0, **P))
sm.sys.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
per = (n_kf + world - 1) // world
for i in range(n_kf):
    if i // per == rank:   # contiguous id blocks per rank
        assert sm.add_keyframe(i + 1, clouds[i]) == tm.OK
        sm.set_pose(i + 1, poses[i])
timing = {}
info = sm.rebuild(timing)
info2 = sm.rebuild()          # idempotent: a second rebuild gives the same map
layers = helpers.MOMENT_LAYERS + [l for l in helpers.DERIVED_LAYERS if INFLATE or l != "step_haz_inflated"]
got = {}
for l in layers:
    got[l], origin = sm.gather_layer(l)
ok = True
if rank == 0:
    s = tm.System(tm.default_params(device=local, max_batch=0, **P))
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    with s.getLocalMap(0).locked():
        pass
    for i in range(n_kf):
        s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i])
    m = s.getLocalMap(0)
    with m.locked():            # hold the worker off so all keyframes fold in ONE ascending-id batch
        for i in range(n_kf):
            s.updateKeyFrame(i + 1, poses[i])
    s.flush()
    s.informLoopClosure(); s.flush()   # the rebuild path proper
    h, w = got[layers[0]].shape
    worst = 0
    for l in layers:
        ref = m.exportDense(l, origin[0], origin[1], w, h)
        same = helpers.bits_equal(got[l], ref)
        if not same:
            bad = (np.isnan(got[l]) != np.isnan(ref)) | (~np.isnan(ref) & (got[l] != ref))
            print(f"layer {l}: {int(bad.sum())} cells differ (sharded: {int(np.sum(~np.isnan(got[l])))} numeric cells, "
                  f"one GPU: {int(np.sum(~np.isnan(ref)))}; inflation tiles listed so far: {sm.map.stats()['sum_infl_tiles']} / {m.stats()['sum_infl_tiles']})")
            ok = False
    occ = int(np.sum(~np.isnan(got["N"])))
    print(f"world={world} keyframes={n_kf} occupied={occ} slabs={info['slabs']} bit-identical={ok} timing_ms={ {k: round(v, 3) for k, v in timing.items()} }")

# ---- incremental pose update on the sharded map (tmap_shard_update) vs the same batch on one GPU -------------------
# every third keyframe moves (a PGO message); one rank may own none of them.  Subtract + add on the slabs of the rebuild.
moved = [i for i in range(n_kf) if i % 3 == 1]
new_poses = synth.drift_poses(poses)
mine_moved = [i for i in moved if i // per == rank]
uinfo = sm.update([i + 1 for i in mine_moved], new_poses[mine_moved] if mine_moved else np.zeros((0, 3, 4), np.float32))
got2 = {}
for l in layers:
    got2[l], origin2 = sm.gather_layer(l)
if rank == 0:
    with m.locked():            # ONE batch, ascending ids: subtract at the integrated pose, add at the new one
        s.updateKeyFrames(np.array([i + 1 for i in moved], dtype=np.uint64), new_poses[moved])
    s.flush()
    h2, w2 = got2[layers[0]].shape
    ok2 = True
    for l in layers:
        ref = m.exportDense(l, origin2[0], origin2[1], w2, h2)
        if not helpers.bits_equal(got2[l], ref):
            bad = (np.isnan(got2[l]) != np.isnan(ref)) | (~np.isnan(ref) & (got2[l] != ref))
            print(f"after update, layer {l}: {int(bad.sum())} cells differ")
            ok2 = False
    print(f"pose update of {len(moved)} keyframes on the sharded map: bit-identical={ok2} slabs={uinfo['slabs']}")
    ok = ok and ok2
    s.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
sm.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)

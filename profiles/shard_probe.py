"""The sharded rebuild (tmap_shard_rebuild) at C4 scale on ONE rank (world = 1): isolates what the sharded code path costs
on the same records as profiles/c4_probe.py (the all-to-all degenerates to a device copy).  Under torchrun it runs with
the job's world size.  python profiles/shard_probe.py [n_keyframes] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import bench
import traversability_mapping_b200 as tm
from traversability_mapping_b200 import sharding, synth

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if "RANK" not in os.environ:
    os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29577")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
true, drift = bench.c4_poses(bench.C4_KEYFRAMES)
per = (nk + world - 1) // world
mine = [i for i in range(nk) if i // per == rank]
sm = sharding.ShardedMap(tm.default_params(device=local, max_batch=0, **bench.C4))
sm.sys.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
for i, c in bench.gen_host_clouds(bench.SEED + 1, mine, true, bench.C4_RINGS, bench.C4_AZ, "cuda"):
    assert sm.add_keyframe(i + 1, c) == tm.OK
ids = np.array([i + 1 for i in mine], dtype=np.uint64)
for rep in range(reps):
    target = true if rep % 2 == 0 else drift
    sm.set_poses(ids, target[mine])
    dist.barrier()
    t0 = time.perf_counter()
    info = sm.rebuild()
    dt = time.perf_counter() - t0
    if rank == 0:
        print("rebuild %d: wall %.2f ms | %s | local %d sent %d recv %d | grid %.2f GB" % (
            rep, dt * 1e3, " ".join("%s %.3f" % kv for kv in info["stages_ms"].items()), info["local_records"], info["sent_records"],
            info["recv_records"], info["grid_bytes"] / 1e9))
sm.close()
dist.destroy_process_group()

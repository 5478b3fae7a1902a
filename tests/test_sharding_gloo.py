"""N>1 host logic of the sharded rebuild on CPU: slab planning, the all-to-all routing plan and the
per-bin merge order, exercised with world_size 2 over gloo (no GPU, no CUDA library calls)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traversability_mapping_b200 import sharding  # noqa: E402


def test_slab_boundaries_balanced_and_tile_aligned():
    rng = np.random.default_rng(0)
    for world in (2, 4, 8):
        for first in (0, -7, 5):
            rows = rng.integers(0, 1000, size=64).astype(float)
            rows[:5] = 0
            slabs = sharding.slab_boundaries(rows, world, first_global_row=first)
            assert slabs[0][0] == 0 and slabs[-1][1] == 64
            for (a, b), (c, d) in zip(slabs[:-1], slabs[1:]):
                assert b == c and b > a + 1
                assert (first + b) % 2 == 0, "interior boundaries sit on 32-cell tile rows"
            tot = [rows[a:b].sum() for a, b in slabs]
            assert max(tot) <= rows.sum() / world + 2 * rows.max() + 1


def test_exchange_plan_and_merge_reference_single_process():
    rng = np.random.default_rng(1)
    world, btw, bth = 3, 5, 12
    nb = btw * bth
    counts = rng.integers(0, 4, size=(world, nb))
    starts = np.concatenate([np.zeros((world, 1), int), np.cumsum(counts, 1)[:, :-1]], 1)
    # record = (bin, source, running index): sources' arrays are sorted by bin
    recs = [np.array([(b, s, k) for b in range(nb) for k in range(counts[s, b])], dtype=np.int64).reshape(-1, 3) for s in range(world)]
    slabs = sharding.slab_boundaries(counts.sum(0).reshape(bth, btw).sum(1), world)
    for rank in range(world):
        y0, y1 = slabs[rank]
        recv = np.concatenate([recs[s][starts[s, y0 * btw]: starts[s, y0 * btw] + counts[s, y0 * btw:y1 * btw].sum()] for s in range(world)])
        send, rcv, src_base = sharding.exchange_plan(counts, btw, slabs, rank)
        assert sum(rcv) == len(recv) and send[rank] == rcv[rank]
        merged, cnt, st = sharding.merge_reference(recv, counts, starts, src_base, btw, (y0, y1))
        # merged must be ordered by (bin, source, index) -- the single-GPU fold order
        key = merged[:, 0] * 10 ** 6 + merged[:, 1] * 10 ** 3 + merged[:, 2]
        assert np.all(np.diff(key) > 0)
        assert np.array_equal(cnt, counts[:, y0 * btw:y1 * btw].sum(0))


def test_exchange_plan_from_row_totals_equals_plan_from_bin_counts():
    """ShardedMap reduces the (world, n_bins) counts to (world, n_rows) on the device; the routing plan from
    the row totals must be the plan from the full counts."""
    rng = np.random.default_rng(3)
    world, btw, bth = 4, 7, 30
    counts = rng.integers(0, 9, size=(world, btw * bth))
    rows = counts.reshape(world, bth, btw).sum(2)
    slabs = sharding.slab_boundaries(rows.sum(0), world, first_global_row=-3)
    for rank in range(world):
        a = sharding.exchange_plan(counts, btw, slabs, rank)
        b = sharding.exchange_plan_rows(rows, slabs, rank)
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from traversability_mapping_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)          # same stream on both ranks: everybody knows everybody's counts
btw, bth = 6, 20
nb = btw * bth
counts = rng.integers(0, 5, size=(world, nb))
starts = np.concatenate([np.zeros((world, 1), int), np.cumsum(counts, 1)[:, :-1]], 1)
mine = np.array([(b, rank, k, 0) for b in range(nb) for k in range(counts[rank, b])], dtype=np.float32).reshape(-1, 4)
# all-gather of the counts as the GPU path does it
ca = [torch.zeros(nb, dtype=torch.int32) for _ in range(world)]
dist.all_gather(ca, torch.from_numpy(counts[rank].astype(np.int32)))
counts_all = torch.stack(ca).numpy()
assert np.array_equal(counts_all, counts)
slabs = sharding.slab_boundaries(counts_all.astype(np.int64).sum(0).reshape(bth, btw).sum(1), world)
recv, send_s, recv_s, src_base = sharding.exchange_records(torch.from_numpy(mine).reshape(-1), counts_all, btw, slabs, rank, dist)
recv = recv.numpy().reshape(-1, 4)
merged, cnt, st = sharding.merge_reference(recv, counts_all, starts, src_base, btw, slabs[rank])
y0, y1 = slabs[rank]
want = np.array([(b, s, k, 0) for b in range(y0 * btw, y1 * btw) for s in range(world) for k in range(counts[s, b])], dtype=np.float32).reshape(-1, 4)
assert np.array_equal(merged, want), (rank, merged.shape, want.shape)
sys.stdout.write(f"[rank{rank}-ok:{len(merged)}]\n"); sys.stdout.flush()
dist.destroy_process_group()
'''


def test_all_to_all_routing_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script), ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("-ok:") == 2


def test_c_planner_equals_numpy_planner():
    """tmap_shard_plan (the planner LocalMap::shardRebuild runs between its all-gather and its all-to-all) against
    the numpy statement in sharding.py: slabs, per-peer send / receive counts and receive offsets, for every rank."""
    import ctypes as C
    import traversability_mapping_b200 as tm
    lib = tm.load_library()
    rng = np.random.default_rng(11)
    for world, n_rows, first in ((2, 11, 0), (2, 40, -7), (4, 64, 5), (8, 200, -3), (8, 19, 2), (3, 33, 1)):
        rows = rng.integers(0, 5000, size=(world, n_rows)).astype(np.uint32)
        rows[:, : n_rows // 7] = 0
        if world == 4:
            rows[1:] = 0  # one rank holds everything
        want_slabs = sharding.slab_boundaries(rows.sum(0), world, first_global_row=first)
        for rank in range(world):
            slabs = np.zeros(2 * world, dtype=np.int32)
            send = np.zeros(world, dtype=np.uint64); recv = np.zeros(world, dtype=np.uint64); base = np.zeros(world, dtype=np.uint64)
            rc = lib.tmap_shard_plan(rows.ctypes.data, world, n_rows, first, rank, slabs.ctypes.data, send.ctypes.data,
                                     recv.ctypes.data, base.ctypes.data)
            assert rc == tm.OK
            assert [tuple(x) for x in slabs.reshape(world, 2)] == [tuple(x) for x in want_slabs]
            s, r, b = sharding.exchange_plan_rows(rows, want_slabs, rank)
            assert list(send) == s and list(recv) == r and list(base) == [int(x) for x in b]
    # too few rows for the world size is refused, as in the numpy planner
    rows = np.ones((4, 8), dtype=np.uint32)
    slabs = np.zeros(8, dtype=np.int32); z = np.zeros(4, dtype=np.uint64)
    assert lib.tmap_shard_plan(rows.ctypes.data, 4, 8, 0, 0, slabs.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data) < 0


WORKER_C_PLAN = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import traversability_mapping_b200 as tm
from traversability_mapping_b200 import sharding
lib = tm.load_library()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(13)
btw, bth, first = 5, 24, -3
nb = btw * bth
counts = rng.integers(0, 6, size=(world, nb))
starts = np.concatenate([np.zeros((world, 1), int), np.cumsum(counts, 1)[:, :-1]], 1)
mine = np.array([(b, rank, k, 0) for b in range(nb) for k in range(counts[rank, b])], dtype=np.float32).reshape(-1, 4)
# what LocalMap::shardRebuild gathers: every rank's per-bin-row totals
rows_mine = torch.from_numpy(counts[rank].reshape(bth, btw).sum(1).astype(np.int32))
ra = [torch.zeros(bth, dtype=torch.int32) for _ in range(world)]
dist.all_gather(ra, rows_mine)
rows_all = np.ascontiguousarray(torch.stack(ra).numpy().astype(np.uint32))
slabs = np.zeros(2 * world, dtype=np.int32)
send = np.zeros(world, dtype=np.uint64); recv = np.zeros(world, dtype=np.uint64); base = np.zeros(world, dtype=np.uint64)
assert lib.tmap_shard_plan(rows_all.ctypes.data, world, bth, first, rank, slabs.ctypes.data, send.ctypes.data, recv.ctypes.data,
                           base.ctypes.data) == tm.OK
sl = [tuple(int(v) for v in x) for x in slabs.reshape(world, 2)]
# route with the library's counts: a slab's records are one contiguous range of the sender's bin-major array
out = torch.empty(int(recv.sum()) * 4, dtype=torch.float32)
dist.all_to_all_single(out, torch.from_numpy(mine).reshape(-1), [int(r) * 4 for r in recv], [int(s) * 4 for s in send])
got = out.numpy().reshape(-1, 4)
# the fused scatter's destination rule (runtime.cu, routing tables): my record at bucket slot s that belongs to rank k lands at
# s + (records k receives from the ranks below me) - (my first slot of k's slab); check it against the exchanged layout
mine_first = np.concatenate([[0], np.cumsum(send)[:-1]]).astype(np.int64)
all_recv = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
dist.all_gather(all_recv, torch.from_numpy(recv.astype(np.int64)))
recv_matrix = torch.stack(all_recv).numpy()          # recv_matrix[k][s] = records rank k receives from rank s
for k in range(world):
    base_at_k = int(recv_matrix[k][:rank].sum())
    lo, n_k = int(mine_first[k]), int(send[k])
    dest = np.arange(lo, lo + n_k) + (base_at_k - lo)
    assert np.array_equal(dest, np.arange(base_at_k, base_at_k + n_k))
y0, y1 = sl[rank]
merged, cnt, st = sharding.merge_reference(got, counts, starts, base, btw, (y0, y1))
want = np.array([(b, s, k, 0) for b in range(y0 * btw, y1 * btw) for s in range(world) for k in range(counts[s, b])], dtype=np.float32).reshape(-1, 4)
assert np.array_equal(merged, want), (rank, merged.shape, want.shape)
for (a, b_), (c, d) in zip(sl[:-1], sl[1:]):
    assert b_ == c and (first + b_) % 2 == 0
sys.stdout.write(f"[rank{rank}-ok:{len(merged)}]\n"); sys.stdout.flush()
dist.destroy_process_group()
'''


def test_library_planner_routes_world2_gloo(tmp_path):
    """world_size 2 over gloo with the LIBRARY's planner (tmap_shard_plan) between the all-gather of row totals and the
    exchange: counts, offsets and the per-bin merge order come out as the single-process fold order."""
    script = tmp_path / "worker_c.py"
    script.write_text(WORKER_C_PLAN)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29534", str(script), ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("-ok:") == 2

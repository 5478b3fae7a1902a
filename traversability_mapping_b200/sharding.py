"""Spatially sharded global map: one process per GPU (BASELINE north_star).

The reference has no distributed path.  The sharded rebuild itself is C++ behind ONE collective C-ABI call,
tmap_shard_rebuild (csrc/runtime.cu: LocalMap::shardRebuild), which issues its NCCL collectives on the map's
stream; this module is only the Python binding a test / bench process uses: it ships the 128-byte NCCL unique
id to every rank (the one thing the library leaves to its host, as with MPI bootstrapping), keeps the host
poses, and offers slab exports.  What the call does for a loop-closure style REBUILD of a global map whose
keyframes are spread over the ranks:

  1. every rank transforms + keys + bucket-sorts ITS keyframes (k_hist / k_scatter) over a bbox of
     16-cell bins shared by all ranks, records ordered (bin, op, point);
  2. per-bin counts are all-gathered; bin ROWS are cut into `world` contiguous slabs balanced by record
     count (boundaries on 32-cell tile rows) -- rank k owns slab k;
  3. ONE all-to-all (NCCL over NVLink) routes each slab's records to its owner: because the buckets are
     row-major, the records of a slab are one contiguous range of the sender's sorted array, so the send
     buffer is the sorted array itself (no packing);
  4. the owner merges the per-source segments bin by bin in source-rank order (k_shard_merge) and folds
     (k_fold).  Ranks hold contiguous blocks of keyframe ids, so source-rank order == ascending keyframe
     id == the single-GPU fold order, and the moment layers are bit-identical to a 1-GPU rebuild;
  5. stencil-radius halos: the r cell rows (48-byte moment records + flag bytes) either side of each slab
     boundary are exchanged in place with send/recv (rows are contiguous in the row-major grid);
  6. each rank classifies the bins of its slab (k_classify).

The planning helpers below are the numpy statement of the library's slab planner (tmap_shard_plan): the gloo tests
(tests/test_sharding_gloo.py) check the C planner against them and run the exchange + merge on CPU tensors with
world_size 2.  Pose updates of a sharded map are tmap_shard_update (same pipeline, subtract + add, slabs kept).  Step inflation is not
available on a sharded map in this round.
"""
import ctypes as C
import math

import numpy as np

BIN = 16


def slab_boundaries(row_totals, world, first_global_row=0):
    """Cut `len(row_totals)` bin rows into `world` contiguous slabs [y0, y1) with ~equal record counts.
    Interior boundaries fall on even GLOBAL bin rows (32-cell tile rows).  Every slab gets >= 2 rows."""
    row_totals = np.asarray(row_totals, dtype=np.float64)
    n = len(row_totals)
    if n < 2 * world + 2:
        raise ValueError(f"map too small to shard: {n} bin rows over {world} ranks")
    cum = np.concatenate([[0.0], np.cumsum(row_totals)])
    total = cum[-1]
    cuts = [0]
    for k in range(1, world):
        target = total * k / world
        y = int(np.searchsorted(cum, target, side="left"))
        if (first_global_row + y) % 2:  # align to a 32-cell tile row
            y += 1
        y = max(y, cuts[-1] + 2)
        y = min(y, n - 2 * (world - k))
        if (first_global_row + y) % 2:
            y -= 1
        cuts.append(y)
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def exchange_plan(counts_all, btw, slabs, rank):
    """counts_all: (world, n_bins) integer tensor/array of every rank's per-bin record counts.
    Returns (send_splits, recv_splits, src_base) in RECORDS for `rank`."""
    ca = np.asarray(counts_all).astype(np.int64)
    send = [int(ca[rank, y0 * btw:y1 * btw].sum()) for (y0, y1) in slabs]
    y0, y1 = slabs[rank]
    recv = [int(ca[s, y0 * btw:y1 * btw].sum()) for s in range(ca.shape[0])]
    src_base = np.concatenate([[0], np.cumsum(recv)[:-1]]).astype(np.uint64)
    return send, recv, src_base


def exchange_plan_rows(rows_all, slabs, rank):
    """Same plan from per-bin-ROW totals: rows_all[s, y] = records rank s holds in bin row y.  The (world, n_bins)
    count matrix is reduced to this (world, n_rows) one on the device, so the host never touches per-bin data
    (1.4 M bins x 8 ranks for a 1 km map at 0.05 m)."""
    ra = np.asarray(rows_all).astype(np.int64)
    send = [int(ra[rank, y0:y1].sum()) for (y0, y1) in slabs]
    y0, y1 = slabs[rank]
    recv = [int(ra[s, y0:y1].sum()) for s in range(ra.shape[0])]
    src_base = np.concatenate([[0], np.cumsum(recv)[:-1]]).astype(np.uint64)
    return send, recv, src_base


def merge_reference(recv, counts_all, starts_all, src_base, btw, slab):
    """NumPy statement of k_shard_merge (used by the CPU tests): per bin of the slab, the sources' segments
    concatenated in source-rank order.  recv: (n, width) records.  Returns (merged, bin_count, bin_start)."""
    ca = np.asarray(counts_all).astype(np.int64)
    sa = np.asarray(starts_all).astype(np.int64)
    y0, y1 = slab
    first = y0 * btw
    bins = range(y0 * btw, y1 * btw)
    out, cnt, st = [], [], []
    pos = 0
    for b in bins:
        st.append(pos)
        for s in range(ca.shape[0]):
            c = ca[s, b]
            if c:
                o = int(src_base[s]) + (sa[s, b] - sa[s, first])
                out.append(recv[o:o + c])
                pos += c
        cnt.append(pos - st[-1])
    merged = np.concatenate(out) if out else recv[:0]
    return merged, np.array(cnt), np.array(st)


def exchange_records(records, counts_all, btw, slabs, rank, dist, width=4, plan=None):
    """The all-to-all of step 3 on a flat tensor of `width` scalars per record (CPU/gloo or CUDA/NCCL).
    `plan` = (send, recv, src_base) when the caller already has it (exchange_plan_rows)."""
    import torch
    if plan is not None:
        send, recv, src_base = plan
    else:
        send, recv, src_base = exchange_plan(counts_all.cpu().numpy() if hasattr(counts_all, "cpu") else counts_all, btw, slabs, rank)
    out = torch.empty(sum(recv) * width, dtype=records.dtype, device=records.device)
    dist.all_to_all_single(out, records, [r * width for r in recv], [s * width for s in send])
    return out, send, recv, src_base


def footprint_radius_cells(robot_radius, resolution):
    """Bounding radius of discOffsets (Helpers.cpp:156-169) in cells."""
    rc = max(1.0, robot_radius / resolution)
    r = int(math.ceil(rc))
    best = 0
    for di in range(-r, r + 1):
        for dj in range(-r, r + 1):
            if float(di * di + dj * dj) <= rc * rc:
                best = max(best, abs(di), abs(dj))
    return best


class ShardedMap:
    """One rank's share of a global map (CUDA + NCCL).  Keyframes are added locally; `rebuild()` is
    collective.  Rank k must hold a contiguous block of keyframe ids (all ids on rank k smaller than
    all ids on rank k+1) for the result to equal the single-GPU ascending-id rebuild bit for bit.
    After a rebuild a rank's window holds its own slab rows (+ the footprint halo) only: read it through
    `export_slab` / `gather_layer`, not through whole-window queries."""

    def __init__(self, params, rank=None, world=None):
        import torch
        import torch.distributed as dist
        import traversability_mapping_b200 as tm
        self.tm, self.torch, self.dist = tm, torch, dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.params = params
        self.sys = tm.System(params)
        self.sys.addNewLocalMap(0)
        self.map = self.sys.getLocalMap(0)
        self.poses = {}
        self.r = footprint_radius_cells(params.robot_radius, params.resolution)
        self.last = {}
        # NCCL bootstrap: rank 0 draws the unique id, the process group (whatever backend it has) carries it
        lib = self.sys._lib
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            tm._check(lib.tmap_shard_unique_id(uid.ctypes.data))
        box = [uid.tobytes()]
        if self.world > 1:
            dist.broadcast_object_list(box, src=0)
        uid = np.frombuffer(box[0], dtype=np.uint8).copy()
        tm._check(lib.tmap_shard_comm_init(self.sys._h, 0, self.rank, self.world, uid.ctypes.data))

    def close(self):
        self.sys._lib.tmap_shard_comm_destroy(self.sys._h, 0)
        self.sys.close()

    def add_keyframe(self, kf_id, cloud, timestamp_ns=0):
        return self.sys.addNewKeyFrameWithPCL(timestamp_ns, kf_id, 0, cloud)

    def set_pose(self, kf_id, pose):
        self.poses[int(kf_id)] = np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(-1)[:12])

    def set_poses(self, kf_ids, poses):
        """Bulk form of set_pose: the arrays are kept as they are and handed to the library at the next rebuild."""
        self._ids = np.ascontiguousarray(kf_ids, dtype=np.uint64)
        self._poses = np.ascontiguousarray(np.asarray(poses, dtype=np.float32).reshape(len(self._ids), -1)[:, :12])
        self.poses = None

    # ------------------------------------------------------------------------------------------
    def rebuild(self, timing=None):
        """Collective.  One call into the library (tmap_shard_rebuild); returns this rank's view of the result."""
        tm = self.tm
        lib, h = self.sys._lib, self.sys._h
        if self.poses is None:
            ids, poses = self._ids, self._poses
        else:
            ids = np.array(sorted(self.poses), dtype=np.uint64)
            poses = (np.stack([self.poses[int(i)] for i in ids]).astype(np.float32) if len(ids)
                     else np.zeros((0, 12), np.float32))
        info = tm.ShardInfo()
        tm._check(lib.tmap_shard_rebuild(h, 0, ids.ctypes.data, len(ids), poses.ctypes.data, C.byref(info)))
        W = info.n_slabs
        slabs = [(info.slabs[2 * k], info.slabs[2 * k + 1]) for k in range(W)]
        stages = dict(zip(("bin", "plan", "alltoall", "fold", "halo", "classify", "total"), [float(x) for x in info.ms[:7]]))
        if timing is not None:
            for k, v in stages.items():
                timing[k] = timing.get(k, 0.0) + v
        self.last = dict(slabs=slabs, bbox=tuple(info.bbox), rows=(info.cj0, info.cj1), sent_records=int(info.sent_records),
                         recv_records=int(info.recv_records), local_records=int(info.local_records),
                         halo_bytes=int(info.halo_bytes), grid_bytes=int(info.grid_bytes), stages_ms=stages)
        return self.last

    def update(self, kf_ids, poses, timing=None):
        """Collective.  The listed keyframes of THIS rank (ascending ids, possibly none) move to `poses`: one subtract + add
        batch over the job on the slabs of the last rebuild (tmap_shard_update)."""
        tm = self.tm
        ids = np.ascontiguousarray(kf_ids, dtype=np.uint64)
        ps = np.ascontiguousarray(np.asarray(poses, dtype=np.float32).reshape(len(ids), -1)[:, :12]) if len(ids) else np.zeros((0, 12), np.float32)
        info = tm.ShardInfo()
        tm._check(self.sys._lib.tmap_shard_update(self.sys._h, 0, ids.ctypes.data, len(ids), ps.ctypes.data, C.byref(info)))
        stages = dict(zip(("bin", "plan", "alltoall", "fold", "halo", "classify", "total"), [float(x) for x in info.ms[:7]]))
        if timing is not None:
            for k, v in stages.items():
                timing[k] = timing.get(k, 0.0) + v
        W = info.n_slabs
        self.last = dict(slabs=[(info.slabs[2 * k], info.slabs[2 * k + 1]) for k in range(W)], bbox=tuple(info.bbox),
                         rows=(info.cj0, info.cj1), sent_records=int(info.sent_records), recv_records=int(info.recv_records),
                         local_records=int(info.local_records), halo_bytes=int(info.halo_bytes), grid_bytes=int(info.grid_bytes),
                         stages_ms=stages)
        if self.poses is not None:
            for i, p in zip(ids, ps):
                self.poses[int(i)] = p
        return self.last

    # ------------------------------------------------------------------------------------------
    def my_rows(self):
        return self.last["rows"]

    def export_slab(self, layer):
        """(rows, width) float32 of this rank's slab over the shared bbox columns, plus (ci0, cj0)."""
        bx0, by0, btw, bth = self.last["bbox"]
        cj0, cj1 = self.last["rows"]
        ci0, w = bx0 * BIN, btw * BIN
        return self.map.exportDense(layer, ci0, cj0, w, cj1 - cj0), (ci0, cj0)

    def gather_layer(self, layer):
        """All ranks' slabs stacked into the full bbox layer on every rank (test / export helper)."""
        torch, dist = self.torch, self.dist
        bx0, by0, btw, bth = self.last["bbox"]
        full = torch.full((bth * BIN, btw * BIN), float("nan"), dtype=torch.float32, device="cuda")
        mine, (ci0, cj0) = self.export_slab(layer)
        full[cj0 - by0 * BIN: cj0 - by0 * BIN + mine.shape[0]] = torch.from_numpy(mine).cuda()
        # slabs are disjoint: combine with a NaN-aware max (NaN -> -inf -> back)
        neg = torch.where(torch.isnan(full), torch.full_like(full, -float("inf")), full)
        dist.all_reduce(neg, op=dist.ReduceOp.MAX)
        out = torch.where(torch.isinf(neg) & (neg < 0), torch.full_like(neg, float("nan")), neg)
        return out.cpu().numpy(), (bx0 * BIN, by0 * BIN)

"""Spatially sharded global map: one process per GPU (BASELINE north_star).

The reference has no distributed path.  This module is the multi-GPU driver of the C ABI's tmap_shard_*
entry points for a loop-closure style REBUILD of a global map whose keyframes are spread over the ranks:

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

The planning helpers are pure functions on CPU tensors so the N>1 logic is covered by gloo tests without
a GPU (tests/test_sharding_gloo.py).  Step inflation is not available on a sharded map in this round.
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


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias library-owned device memory."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _alias(ptr, n, typestr, torch):
    if n == 0 or not ptr:
        dt = {"<f4": torch.float32, "<i4": torch.int32, "|u1": torch.uint8}[typestr]
        return torch.empty(0, dtype=dt, device="cuda")
    return torch.as_tensor(_DevArray(ptr, n, typestr), device="cuda")


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
    After a rebuild a rank's window is valid on its own slab rows (+ the footprint halo) only: read it through
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

    def close(self):
        self.sys.close()

    def add_keyframe(self, kf_id, cloud, timestamp_ns=0):
        return self.sys.addNewKeyFrameWithPCL(timestamp_ns, kf_id, 0, cloud)

    def set_pose(self, kf_id, pose):
        self.poses[int(kf_id)] = np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(-1)[:12])

    # ------------------------------------------------------------------------------------------
    def rebuild(self, timing=None):
        tm, torch, dist = self.tm, self.torch, self.dist
        lib, h = self.sys._lib, self.sys._h
        W, rank = self.world, self.rank
        ids = np.array(sorted(self.poses), dtype=np.uint64)
        poses = np.stack([self.poses[int(i)] for i in ids]).astype(np.float32) if len(ids) else np.zeros((0, 12), np.float32)
        dev = torch.device("cuda", torch.cuda.current_device())
        ev = []

        def mark(name):
            if timing is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                ev.append((name, e))

        mark("start")
        # op ids unique over the job; bbox = union of every rank's keyframe windows (+1 bin candidate margin)
        n_all = torch.zeros(W, dtype=torch.int64, device=dev)
        n_all[rank] = len(ids)
        win = (C.c_int32 * 4)(2 ** 31 - 1, 2 ** 31 - 1, -2 ** 31, -2 ** 31)
        if len(ids):
            tm._check(lib.tmap_shard_window(h, 0, ids.ctypes.data, len(ids), poses.ctypes.data, win))
        meta = torch.tensor([win[0], win[1], -win[2], -win[3]], dtype=torch.int64, device=dev)
        dist.all_reduce(n_all)
        dist.all_reduce(meta, op=dist.ReduceOp.MIN)
        op_base = int(n_all[:rank].sum())
        m = meta.tolist()
        margin = max(1, (self.r + BIN - 1) // BIN)
        bbox = (C.c_int32 * 4)(m[0] - margin, m[1] - margin, -m[2] + margin, -m[3] + margin)
        buf = tm.ShardBuffers()
        tm._check(lib.tmap_shard_bin(h, 0, ids.ctypes.data, len(ids), poses.ctypes.data, op_base, bbox, C.byref(buf)))
        mark("bin")
        nb, btw, bth = buf.n_bins, buf.btw, buf.bth
        counts = _alias(buf.d_bin_count, nb, "<i4", torch)
        starts = _alias(buf.d_bin_start, nb, "<i4", torch)
        counts_all = torch.empty((W, nb), dtype=torch.int32, device=dev)
        starts_all = torch.empty((W, nb), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(counts_all, counts.contiguous())
        dist.all_gather_into_tensor(starts_all, starts.contiguous())
        # per-(rank, bin row) totals are formed on the device; only that small matrix goes to the host
        rows_all = counts_all.view(W, bth, btw).sum(dim=2, dtype=torch.int64).cpu().numpy()
        slabs = slab_boundaries(rows_all.sum(0), W, first_global_row=buf.bty0)
        y0, y1 = slabs[rank]
        # all-to-all point routing: the sorted array IS the send buffer (slab ranges are contiguous)
        records = _alias(buf.d_records, buf.n_records * 4, "<f4", torch)
        recv, send_s, recv_s, src_base = exchange_records(records, None, btw, slabs, rank, dist,
                                                          plan=exchange_plan_rows(rows_all, slabs, rank))
        torch.cuda.synchronize()
        mark("alltoall")
        src_base = np.ascontiguousarray(src_base, dtype=np.uint64)
        # blank this rank's slab + footprint halo only (the window spans the whole map on every rank; rows outside
        # the slab are not read until a later rebuild owns and blanks them)
        blank0 = (buf.bty0 + y0) * BIN - 2 * BIN
        tm._check(lib.tmap_shard_blank_rows(h, 0, blank0, (y1 - y0) * BIN + 4 * BIN))
        tm._check(lib.tmap_shard_fold(h, 0, recv.data_ptr(), sum(recv_s), counts_all.data_ptr(), starts_all.data_ptr(),
                                      src_base.ctypes.data, W, y0, y1))
        mark("fold")
        # halo exchange: r cell rows (moment records + flag bytes) across each slab boundary, in place
        r = self.r
        cj0, cj1 = (buf.bty0 + y0) * BIN, (buf.bty0 + y1) * BIN
        ops, keep = [], []

        def rows(kind, cj, n):
            p, nbytes = C.c_void_p(0), C.c_size_t(0)
            tm._check(lib.tmap_shard_rows(h, 0, kind, cj, n, C.byref(p), C.byref(nbytes)))
            t = _alias(p.value, nbytes.value, "|u1", torch)
            keep.append(t)
            return t

        halo_rows = []
        if rank > 0:
            for kind in (0, 1):
                ops.append(dist.P2POp(dist.isend, rows(kind, cj0, r), rank - 1))
                ops.append(dist.P2POp(dist.irecv, rows(kind, cj0 - r, r), rank - 1))
            halo_rows.append((cj0 - r, r))
        if rank < W - 1:
            for kind in (0, 1):
                ops.append(dist.P2POp(dist.isend, rows(kind, cj1 - r, r), rank + 1))
                ops.append(dist.P2POp(dist.irecv, rows(kind, cj1, r), rank + 1))
            halo_rows.append((cj1, r))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        torch.cuda.synchronize()
        mark("halo")
        cb = torch.tensor([buf.cell_bbox[0], -buf.cell_bbox[1], buf.cell_bbox[2], -buf.cell_bbox[3]], dtype=torch.int64, device=dev)
        dist.all_reduce(cb, op=dist.ReduceOp.MIN)
        c = cb.tolist()
        cell_bbox = (C.c_int32 * 4)(c[0], -c[1], c[2], -c[3])
        tm._check(lib.tmap_shard_classify(h, 0, cell_bbox))
        for cj, n in halo_rows:  # received halo rows are not ours: drop their touched / changed bits
            tm._check(lib.tmap_shard_clear_flag_rows(h, 0, cj, n))
        mark("classify")
        self.last = dict(slabs=slabs, bbox=(buf.btx0, buf.bty0, btw, bth), rows=(cj0, cj1), sent_records=sum(send_s) - send_s[rank],
                         recv_records=sum(recv_s), local_records=int(buf.n_records), halo_bytes=len(halo_rows) * r * btw * BIN * 49)
        if timing is not None:
            torch.cuda.synchronize()
            for (_, a), (name, b_) in zip(ev[:-1], ev[1:]):
                timing[name] = timing.get(name, 0.0) + a.elapsed_time(b_)
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

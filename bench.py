#!/usr/bin/env python
"""bench.py -- points/s binned+classified on BASELINE.json config[1]: the local traversability map,
a 10 Hz stream of 131 072-ray keyframes into a 40 x 40 m rolling grid at 0.05 m (49-cell footprint
disc, step inflation on, rolling window of 15 keyframes; params/traversabilityParamsLocal.yaml with
the grid half-size and sensor range set to 20 m).

One "step" = one stream tick exactly as local_traversability_node.cpp:153-185 drives the library:
keyframe insert -> pose update (first bin + classify + inflate) -> eviction of keyframe id-15
(subtract + reclassify).  `value` counts the points binned for the NEW keyframe only (the evicted
keyframe's points are re-binned too, but not counted) with the clouds already resident in HBM;
`e2e` is the same tick through the C ABI with the cloud in pinned host memory (H2D + prune inside)
plus the D2H of the robot-centred 40 x 40 m occupancy grid.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
N>1: launched by torchrun, one replica map per GPU ("replicas only" for this config: a 40 m local
map has nothing to shard), value = sum over ranks / max-over-ranks time.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOW = 15
RAYS = 131072
C2 = dict(resolution=0.05, half_size=20.0, extend_length=30.0, robot_radius=0.20, ground_clearance=0.15,
          max_slope=0.4, min_points_per_grid=5, inflation_enabled=1, inflation_radius=0.5,
          cost_scaling_factor=3.0, lethal_step_threshold=0.95, robot_height=3.5, max_range=20.0, min_range=1.0)
CROP = 801  # 40 m / 0.05 m + 1 cells, robot-centred
SEED = 2


WORKLOAD = ("C2 local map: 10 Hz stream of 131072-ray keyframes, 40x40 m rolling grid @0.05 m, window 15, "
            "49-cell disc, inflation 0.5 m")


def workload(n_kf):
    from traversability_mapping_b200 import synth
    poses = synth.trajectory("drive", n_kf, seed=SEED, step=0.1, curvature=0.02)  # 1 m/s sampled at 10 Hz
    clouds = [synth.make_keyframe(SEED, i, poses[i], max_range=20.0) for i in range(n_kf)]
    return poses, clouds


class ClockSampler:
    """SM clock and throttle reasons of the timed region, read in-process through NVML right after each
    timed step's end event (the GPU has just been under load).  A sampler running concurrently -- an
    `nvidia-smi -lms` child or an NVML polling thread -- takes driver locks inside the ~0.3 ms step
    brackets and showed up in the step time, so the samples are taken between brackets instead."""

    def __init__(self, gpu_index):
        self.sm, self.reasons, self.mx = [], set(), None
        self.idx = gpu_index
        self.h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                         "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                         "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                         "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        except Exception as e:
            self.err = repr(e)

    def start(self):
        pass

    def sample(self):
        if self.h is None:
            return
        self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        for k, b in self.bits.items():
            if r & b:
                self.reasons.add(k)

    def stop(self):
        if not self.sm:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "note": "nvml unavailable; sampled after the region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def oracle_steps(poses, clouds, first, count, window=WINDOW):
    """The CPU oracle on the same ticks.  Returns (seconds, points binned for the new keyframes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    from traversability_mapping_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    om = po.OracleMap(po.default_params(**C2))
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    pts = 0
    t_total = 0.0
    for i in range(first + count):
        timed = i >= first
        t0 = time.perf_counter()
        om.add_keyframe(i + 1, clouds[i])
        om.update_pose(i + 1, poses[i])
        om.process()
        if i >= window:
            om.delete_keyframe(i + 1 - window)
        dt = time.perf_counter() - t0
        if timed:
            t_total += dt
            pts += om.prune(clouds[i], synth.T_BASE_SENSOR, C2["max_range"], C2["min_range"]).shape[0]
    om.close()
    return t_total, pts


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference cannot be built
    here (Eigen/PCL/grid_map absent), so this times the oracle port -- single worker thread, which is
    how the reference runs a map (LocalMap.cpp:76-77)."""
    if rank != 0:
        return
    steps = min(args.steps, 4)
    warm = min(args.warmup, 1)
    n_kf = WINDOW + warm + steps
    poses, clouds = workload(n_kf)
    sec, pts = oracle_steps(poses, clouds, WINDOW + warm, steps)
    v = pts / sec
    line = {
        "impl": "reference", "metric": "points/s binned+classified", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 moments / f32 layers", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "l2": "cpu run"},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} stream ticks after a {WINDOW + warm}-tick fill, oracle port (reference not buildable)"},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-lc", action="store_true", help="skip the loop-closure leg")
    ap.add_argument("--lc-keyframes", type=int, default=2000)
    ap.add_argument("--lc-extent", type=float, default=90.0, help="loop-closure trajectory inside +-extent metres")
    ap.add_argument("--lc-resolution", type=float, default=0.1, help="loop-closure grid resolution, metres")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import traversability_mapping_b200 as tm
    from traversability_mapping_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    K, Wm = args.steps, max(args.warmup, 3)
    n_kf = WINDOW + Wm + K
    poses, clouds = workload(n_kf)
    pinned = [torch.from_numpy(c).pin_memory() for c in clouds]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.Stream()

    def new_system():
        p = tm.default_params(device=local, max_batch=1, profile=1, **C2)
        s = tm.System(p)
        s.addNewLocalMap(0)
        s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
        m = s.getLocalMap(0)
        m.setStream(stream.cuda_stream)
        return s, m

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gc.collect()
    gc.disable()  # no collector pauses inside the ~0.3 ms step brackets
    # ---------------- device-resident arm (`value`) ----------------
    s, m = new_system()
    for i in range(n_kf):   # insert = H2D + prune, not timed here; keyframes are NOT binned until a pose arrives
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i]) == tm.OK
    STAGES = ("ms_hist", "ms_scan", "ms_scatter", "ms_fold", "ms_gate", "ms_classify", "ms_inflate", "ms_total")
    npts = [s.keyFramePoints(i + 1) for i in range(WINDOW + Wm, n_kf)]   # points of each timed step's new keyframe
    t_ms = 0.0
    sampler = ClockSampler(local)
    st0 = None
    for i in range(n_kf):
        timed = i >= WINDOW + Wm
        if i == WINDOW + Wm:
            barrier()
            sampler.start()
            st0 = m.stats()   # the library keeps running sums; they are read once before / after the timed steps
        with torch.cuda.stream(stream):
            if timed:
                flush_buf.fill_(i & 0xFF)  # L2 flush between timed iterations (untimed)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                stream.synchronize()
                e0.record(stream)
            s.updateKeyFrame(i + 1, poses[i])
            s.flush()
            if i >= WINDOW:
                s.deleteKeyFrame(i + 1 - WINDOW)   # synchronous: subtracts the evicted keyframe before returning
            if timed:
                e1.record(stream)
                e1.synchronize()
                t_ms += e0.elapsed_time(e1)
                sampler.sample()
    st1 = m.stats()
    stage = {k: st1["sum_" + k] - st0["sum_" + k] for k in STAGES}
    cells_classified = st1["cells_classified"] - st0["cells_classified"]
    cells_touched = st1["sum_cells_touched"] - st0["sum_cells_touched"]
    tiles = st1["sum_tiles"] - st0["sum_tiles"]
    infl_tiles = st1["sum_infl_tiles"] - st0["sum_infl_tiles"]
    cand_bins = st1["sum_cand_bins"] - st0["sum_cand_bins"]
    pts_binned = st1["points_binned"] - st0["points_binned"]   # new + evicted keyframe both pass through the kernels
    stage["ms_classify"] -= stage["ms_gate"]                    # the library's classify stage = k_gate + k_classify
    launches0 = st0["kernel_launches"]
    host = {k: (st1["sum_host_ms_" + k] - st0["sum_host_ms_" + k]) / K for k in ("plan", "launch", "wait", "batch")}
    host["grid_reallocs_in_timed_steps"] = st1["grid_reallocs"] - st0["grid_reallocs"]
    barrier()
    clocks = sampler.stop()
    launches = m.stats()["kernel_launches"] - launches0
    pts_total = int(sum(npts))
    evict_pts = None
    s.close()

    # ---------------- end-to-end arm (`e2e`): host buffers through the C ABI ----------------
    s, m = new_system()
    occ = torch.empty((CROP, CROP), dtype=torch.int8).pin_memory()
    occ_np = occ.numpy()
    e2e_ms = 0.0
    e2e_pts = 0
    for i in range(n_kf):
        timed = i >= WINDOW + Wm
        if i == WINDOW + Wm:
            barrier()
        with torch.cuda.stream(stream):
            if timed:
                flush_buf.fill_(i & 0xFF)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                stream.synchronize()
                e0.record(stream)
            assert s.addNewKeyFrameWithPCL(i, i + 1, 0, pinned[i].numpy()) == tm.OK   # H2D + prune
            s.updateKeyFrame(i + 1, poses[i])
            s.flush()
            if i >= WINDOW:
                s.deleteKeyFrame(i + 1 - WINDOW)
            ci = int(round(float(poses[i][0, 3]) / C2["resolution"])); cj = int(round(float(poses[i][1, 3]) / C2["resolution"]))
            m.exportOccupancy(ci - CROP // 2, cj - CROP // 2, CROP, CROP, out=occ_np)   # D2H
            if timed:
                e1.record(stream)
                e1.synchronize()
                e2e_ms += e0.elapsed_time(e1)
                e2e_pts += npts[i - WINDOW - Wm]
    barrier()
    s.close()

    # ---------------- loop-closure re-integration (BASELINE metric, second half) ----------------
    lc = None
    if not args.no_lc:
        nk = args.lc_keyframes
        base = 32
        lposes = synth.trajectory("loop", nk, seed=SEED + 1, extent=args.lc_extent)
        drift = synth.drift_poses(lposes)
        LCP = dict(resolution=args.lc_resolution, half_size=args.lc_extent + 10.0, max_range=20.0, min_range=1.0,
                   robot_radius=2.0 * args.lc_resolution)
        cfg = "C4-shaped: %d stored keyframes on a %.0f m loop, %.2f m grid, 13-cell disc, drifted -> true poses" % (
            nk, 2 * args.lc_extent, args.lc_resolution)
        if world == 1:
            lclouds = [synth.make_keyframe(SEED + 1, i, lposes[i], max_range=20.0) for i in range(base)]
            p = tm.default_params(device=local, max_batch=0, profile=1, **LCP)
            s = tm.System(p)
            s.addNewLocalMap(0)
            s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
            m = s.getLocalMap(0)
            m.setStream(stream.cuda_stream)
            for i in range(nk):
                s.addNewKeyFrameWithPCL(i, i + 1, 0, lclouds[i % base])
                s.updateKeyFrame(i + 1, drift[i])
            s.flush()
            times, pgo = [], []
            kf_ids = np.arange(1, nk + 1, dtype=np.uint64)
            for rep in range(3):
                target = lposes if rep % 2 == 0 else drift
                # (a) the PGO batch itself: every keyframe gets a new pose; the worker subtracts each old
                #     contribution and adds the new one (global_traversability_node.cpp:184-188), batched
                barrier()
                with torch.cuda.stream(stream):
                    flush_buf.fill_(rep)
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    stream.synchronize()
                    e0.record(stream)
                    with m.locked():   # queue all 2000 updates before the worker starts: one batched pass
                        s.updateKeyFrames(kf_ids, target)   # one PGO message (global_traversability_node.cpp:184-188)
                    s.flush()
                    e1.record(stream)
                    e1.synchronize()
                pgo.append(e0.elapsed_time(e1))
                # (b) informLoopClosure: blank the map and re-integrate every stored keyframe at its pose
                barrier()
                with torch.cuda.stream(stream):
                    flush_buf.fill_(rep)
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    stream.synchronize()
                    e0.record(stream)
                    s.informLoopClosure()
                    s.flush()
                    e1.record(stream)
                    e1.synchronize()
                times.append(e0.elapsed_time(e1))
            st = m.stats()
            lc = {"keyframes": nk, "points": int(st["last_batch_points"]), "ms": float(min(times[1:])), "ms_all": times, "n_gpus": 1,
                  "mode": "informLoopClosure + batched worker pass (max_batch=0)",
                  "pgo_update_ms": float(min(pgo[1:])), "pgo_update_ms_all": pgo,
                  "stages_ms": {k: st[k] for k in stage}, "cells_classified": int(st["last_batch_cells_classified"]),
                  "cells_touched": int(st["last_batch_cells_touched"]), "config": cfg}
            s.close()
        else:
            # the global map sharded over the ranks: contiguous keyframe-id blocks per rank, all-to-all point
            # routing + halo exchange over NVLink (traversability_mapping_b200/sharding.py)
            from traversability_mapping_b200 import sharding
            per = (nk + world - 1) // world
            mine = [i for i in range(nk) if i // per == rank]
            lclouds = {i % base: None for i in mine}
            for k in lclouds:
                lclouds[k] = synth.make_keyframe(SEED + 1, k, lposes[k], max_range=20.0)
            sm = sharding.ShardedMap(tm.default_params(device=local, max_batch=0, **LCP))
            sm.sys.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
            for i in mine:
                sm.add_keyframe(i + 1, lclouds[i % base])
            times, info, tim = [], None, {}
            for rep in range(3):
                for i in mine:
                    sm.set_pose(i + 1, lposes[i] if rep % 2 == 0 else drift[i])
                barrier()
                flush_buf.fill_(rep)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                tim = {}
                info = sm.rebuild(tim)
                e1.record()
                e1.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(float(t))
            pts = torch.tensor([float(info["local_records"]), float(info["sent_records"])], device="cuda", dtype=torch.float64)
            dist.all_reduce(pts)
            lc = {"keyframes": nk, "points": int(pts[0]), "ms": float(min(times[1:])), "ms_all": times, "n_gpus": world,
                  "mode": "sharded rebuild: slabs of bin rows, all-to-all point routing, halo exchange",
                  "routed_points": int(pts[1]), "stages_ms_rank0": tim, "slabs": info["slabs"], "config": cfg}
            sm.close()

    # ---------------- reduce over ranks ----------------
    t_max, e2e_max, pts_sum, e2e_pts_sum = t_ms, e2e_ms, pts_total, e2e_pts
    if world > 1:
        tt = torch.tensor([t_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        pp = torch.tensor([pts_total, e2e_pts], device="cuda", dtype=torch.float64)
        dist.all_reduce(pp, op=dist.ReduceOp.SUM)
        t_max, e2e_max = float(tt[0]), float(tt[1])
        pts_sum, e2e_pts_sum = float(pp[0]), float(pp[1])

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes per kernel: bin = 16 B/point read (hist), 16 + 16 B/point (scatter),
        # 16 B/record + 80 B per touched cell (fold); classify = 76 B per recomputed cell; gate = flag byte in / out
        # of every cell of a candidate bin; inflate = step_haz in + inflated out (+ flag) of every cell of a listed tile
        stage_bytes = {
            "ms_hist": 16.0 * pts_binned,
            "ms_scatter": 32.0 * pts_binned,
            "ms_fold": 16.0 * pts_binned + 80.0 * cells_touched,
            "ms_gate": 2.0 * 256 * cand_bins,
            "ms_classify": 76.0 * cells_classified,
            "ms_inflate": 9.0 * infl_tiles * 1024,
        }
        kname = {"ms_hist": "k_hist<2>", "ms_scatter": "k_scatter<2, 1>", "ms_fold": "k_fold", "ms_gate": "k_gate",
                 "ms_classify": "k_classify", "ms_inflate": "k_inflate"}
        dom = max(kname, key=lambda k: stage[k])
        n_launch = 2 * K  # one update batch + one eviction batch per step, each launches every kernel once
        achieved = stage_bytes[dom] / (stage[dom] * 1e-3) / 1e9 if stage[dom] > 0 else 0.0
        # measured DRAM traffic per launch of that kernel, from the committed ncu capture of this same workload
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["c2_stream"]
            traffic = tj[kname[dom]]["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/r01_traffic.json (ncu --set full)",
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": stage_bytes[dom] / n_launch, "avg_launch_ms": stage[dom] / n_launch,
                    "note": "a 73 k-point stream tick keeps every kernel latency-bound (tens of microseconds per launch); "
                            "the same kernels at loop-closure scale are in loop_closure.stages_frac_of_hbm",
                    "stages_ms_per_step": {k: v / K for k, v in stage.items()},
                    "host_ms_per_step": host,
                    "stages_frac_of_hbm": {k: (stage_bytes[k] / (stage[k] * 1e-3) / 1e9 / peak if stage[k] > 0 else None)
                                           for k in stage_bytes}}
        if lc and "stages_ms" in lc:
            # the same kernels with a full machine's worth of work: fraction of the measured copy peak per kernel
            sm_, P_ = lc["stages_ms"], float(lc["points"])
            lb = {"ms_hist": 16.0 * P_, "ms_scatter": 32.0 * P_, "ms_fold": 16.0 * P_ + 80.0 * lc["cells_touched"],
                  "ms_classify": 76.0 * lc["cells_classified"]}
            lc["stages_frac_of_hbm"] = {k: (lb[k] / ((sm_[k] - (sm_["ms_gate"] if k == "ms_classify" else 0.0)) * 1e-3) / 1e9 / peak
                                            if sm_[k] > 0 else None) for k in lb}
        cpu = None
        if not args.no_cpu:
            sec, cpts = oracle_steps(poses, clouds, WINDOW + Wm, min(K, 3))
            cpu = {"value": cpts / sec, "unit": "points/s", "cores": 1, "kind": "port",
                   "sample": f"{min(K, 3)} stream ticks of the same workload after a {WINDOW + Wm}-tick fill; oracle port, "
                             f"single worker thread as the reference runs a map; host has {os.cpu_count()} cores"}
        line = {
            "metric": "points/s binned+classified", "value": pts_sum / (t_max * 1e-3), "unit": "points/s",
            "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": t_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 moments / f32 layers", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "points_binned_per_step": pts_total / K, "rays_per_keyframe": RAYS, "l2": "flushed between timed steps (256 MB write)",
                       "parallelism": "replicas" if world > 1 else "single"},
            "clocks": clocks,
            "e2e": {"value": e2e_pts_sum / (e2e_max * 1e-3), "unit": "points/s", "ms_per_step": e2e_max / K,
                    "h2d_bytes_per_step": RAYS * 16, "d2h_bytes_per_step": CROP * CROP},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "loop_closure": lc,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on the largest single-GPU configuration (configs[3], "C4"):
loop-closure re-integration of 2 000 stored keyframes (~290 M binned points) into the 200 x 200 m
global map at 0.1 m, reported as points/s binned+classified and as ms per re-integration.

One "step" = one loop closure exactly as a SLAM integration drives the reference library
(System.hpp:101-117): every stored keyframe gets its corrected pose (`updateKeyFrame` x 2 000, one PGO
message), `informLoopClosure()` blanks the map and the worker re-integrates every keyframe -- transform,
cell key, bin, ordered per-(keyframe, cell) moments, float32 fold, neighbourhood classification of every
touched cell.  Steps alternate between the drifted and the true trajectory
(ground_truth_kfs/src/slam_keyframe_sim_with_drift.cpp:146-151,185-217), so every step re-bins every
point at a new pose.

  value : points binned+classified per second, keyframe store resident in HBM (4.7 GB > the 126 MB L2),
          CUDA events around the step on the map's stream
  e2e   : the same step through the C ABI with HOST buffers: the 2 000 poses come from pinned host memory
          (they reach the device inside the op descriptors, H2D counted by the library) and the whole-map
          nav_msgs occupancy grid is read back into pinned host memory (D2H) inside the timed region
  N > 1 : the global map is sharded over the ranks (slabs of bin rows, NCCL all-to-all point routing +
          halo exchange); the SAME 2 000 keyframes are split over the ranks => "scaling": "strong";
          before the timed steps a reduced map is rebuilt sharded AND on rank 0 alone and compared bit for
          bit ("shard_parity"); a mismatch fails the run

Secondary objects in the same JSON line: C2 stream tick (config[1]), C1 single keyframe (config[0]),
C3 500-keyframe global build (config[2]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 2
# ---- C4 (headline) -------------------------------------------------------------------------------
C4_KEYFRAMES = 2000
C4_RINGS, C4_AZ = 128, 2048            # 262 144 rays per keyframe; ~148 k survive the prune
C4_EXTENT = 90.0                       # Lissajous loop inside +-90 m (200 x 200 m map)
C4 = dict(resolution=0.1, half_size=100.0, extend_length=30.0, robot_radius=0.20, ground_clearance=0.15, max_slope=0.4,
          min_points_per_grid=3, inflation_enabled=0, robot_height=3.5, max_range=20.0, min_range=1.0)
WORKLOAD = ("C4 loop closure: %d stored keyframes x %d rays (~290 M binned points) on a 180 m loop, 200x200 m map "
            "@0.1 m, 13-cell disc; pose update of every keyframe + informLoopClosure re-integration" % (C4_KEYFRAMES, C4_RINGS * C4_AZ))
# ---- C2 (secondary) ------------------------------------------------------------------------------
WINDOW = 15
C2_RAYS = 131072
C2 = dict(resolution=0.05, half_size=20.0, extend_length=30.0, robot_radius=0.20, ground_clearance=0.15,
          max_slope=0.4, min_points_per_grid=5, inflation_enabled=1, inflation_radius=0.5,
          cost_scaling_factor=3.0, lethal_step_threshold=0.95, robot_height=3.5, max_range=20.0, min_range=1.0)
CROP = 801  # 40 m / 0.05 m + 1 cells, robot-centred
STAGES = ("ms_hist", "ms_scan", "ms_scatter", "ms_fold", "ms_gate", "ms_classify", "ms_inflate", "ms_total")
KNAME = {"ms_hist": "k_hist", "ms_scatter": "k_scatter", "ms_fold": "k_fold3", "ms_gate": "k_gate",
         "ms_classify": "k_classify", "ms_inflate": "k_inflate"}


class ClockSampler:
    """SM clock and throttle reasons of the timed region, read in-process through NVML right after each
    timed step's end event (the GPU has just been under load).  A sampler running concurrently -- an
    `nvidia-smi -lms` child or an NVML polling thread -- takes driver locks inside the step brackets and
    showed up in the step time (round 1), so the samples are taken between brackets instead."""

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


def c4_poses(nk):
    from traversability_mapping_b200 import synth
    true = synth.trajectory("loop", nk, seed=SEED + 1, extent=C4_EXTENT)
    return true, synth.drift_poses(true)


def gen_host_clouds(seed, idxs, poses, rings, az, device, chunk=24):
    """Yield (keyframe index, (rays, 4) float32 host array) for the listed keyframes; raycast in batches on
    `device` (cuda when there is one) and brought to the host, because the library's insert takes HOST clouds."""
    import torch
    from traversability_mapping_b200 import synth
    idxs = list(idxs)
    for a in range(0, len(idxs), chunk):
        part = idxs[a:a + chunk]
        t = synth.make_keyframes_torch(seed, part[0], poses[part], max_range=20.0, rings=rings, azimuths=az, device=device)
        h = t.cpu().numpy()
        for k, i in enumerate(part):
            yield i, h[k]


def torch_device():
    import torch
    return "cuda" if torch.cuda.is_available() else "cpu"


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference not buildable here: Eigen / PCL / grid_map absent) on a bounded
# sample of the SAME workload -- a loop-closure re-integration of the first `n_kf` keyframes of the C4 run.
# ---------------------------------------------------------------------------------------------------
def oracle_c4(n_kf, steps, warm, threads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pyoracle as po
    import helpers
    from traversability_mapping_b200 import synth
    true, drift = c4_poses(C4_KEYFRAMES)
    om = po.OracleMap(po.default_params(**C4))
    om.set_threads(threads)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    pts = 0
    for i, c in gen_host_clouds(SEED + 1, range(n_kf), true, C4_RINGS, C4_AZ, torch_device()):
        om.add_keyframe(i + 1, c)
        om.update_pose(i + 1, drift[i])
        pts += om.prune(c, synth.T_BASE_SENSOR, C4["max_range"], C4["min_range"]).shape[0]
    om.process_batched()
    secs = []
    for rep in range(warm + steps):
        target = true if rep % 2 == 0 else drift
        t0 = time.perf_counter()
        for i in range(n_kf):
            om.update_pose(i + 1, target[i])
        om.clear_entire_map()
        om.process_batched()
        dt = time.perf_counter() - t0
        if rep >= warm:
            secs.append(dt)
    om.close()
    return float(np.median(secs)), pts, secs


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores, all of them
    where the reference has a threaded flavour (the archived node classifies dirty cells in parallel,
    archive/local_traversability_monolithic.cpp:349-362; binning is single-threaded in every flavour)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = min(args.steps, 3), min(args.warmup, 1)
    n_kf = args.cpu_keyframes
    sec, pts, secs = oracle_c4(n_kf, steps, warm, cores)
    v = pts / sec
    line = {
        "impl": "reference", "metric": "points/s binned+classified", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sec, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64 moments / f32 layers", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "cpu run"},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                         "sample": f"loop-closure re-integration of the first {n_kf} of the {C4_KEYFRAMES} keyframes ({pts} points per step), "
                                   f"median of {steps} steps; oracle port built -O3 (reference not buildable: Eigen/PCL/grid_map absent); "
                                   f"binning on 1 thread as in the reference, classification on {cores} threads as in the archived node"},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--keyframes", type=int, default=C4_KEYFRAMES)
    ap.add_argument("--cpu-keyframes", type=int, default=160, help="keyframes in the CPU arm's bounded sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C1 / C2 / C3 secondary legs")
    ap.add_argument("--c2-steps", type=int, default=12)
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c5 = BASELINE config[4]: the same 2 000-keyframe loop closure on a 1 x 1 km map at 0.05 m (49-cell disc), "
                         "meant for N > 1 (the map is tiled across the ranks)")
    args = ap.parse_args()
    if args.workload == "c5":
        global C4, C4_EXTENT, WORKLOAD
        C4 = dict(C4, resolution=0.05, half_size=500.0)
        C4_EXTENT = 480.0
        WORKLOAD = ("C5 loop closure: %d stored keyframes x %d rays on a 960 m loop, 1 x 1 km map @0.05 m tiled across the ranks "
                    "(slabs of bin rows, records routed by the scatter kernel over NVLink, halo exchange), 49-cell disc"
                    % (C4_KEYFRAMES, C4_RINGS * C4_AZ))
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
    stream = torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gc.collect()
    gc.disable()
    nk = args.keyframes
    true, drift = c4_poses(nk)
    sampler = ClockSampler(local)
    out = {}

    if world == 1:
        out = bench_c4_single(tm, synth, torch, stream, local, nk, true, drift, K, Wm, sampler, barrier)
    else:
        out = bench_c4_sharded(tm, synth, torch, dist, local, rank, world, nk, true, drift, K, Wm, sampler, barrier)
    clocks = sampler.stop()

    secondary = {}
    if not args.no_secondary and world == 1:
        secondary["c2_stream"] = bench_c2(tm, synth, torch, stream, local, args.c2_steps, 3, barrier)
        secondary["c1_single_keyframe"] = bench_c1(tm, synth, torch, stream, local)
        secondary["c3_global_build"] = bench_c3(tm, synth, torch, stream, local, barrier)

    if rank == 0:
        peak, peak_src = peaks()
        roof = None
        if "stage_ms" in out:
            # algorithmic bytes per kernel launch: bin = 16 B/point read (float4 store) by the key
            # pass, 16 + 16 B/point by the bucket scatter, 16 B/record + 80 B per touched cell (10 f32 RMW) by the
            # fold; classify = 76 B per recomputed cell
            P_, ct, cc = float(out["points_per_step"]), float(out["cells_touched"]), float(out["cells_classified"])
            sb = {"ms_hist": 16.0 * P_, "ms_scatter": 32.0 * P_, "ms_fold": 16.0 * P_ + 80.0 * ct, "ms_classify": 76.0 * cc}
            st = dict(out["stage_ms"])
            st["ms_classify"] = st["ms_classify"] - st.get("ms_gate", 0.0)
            dom = max(sb, key=lambda k: st[k])
            achieved = sb[dom] / (st[dom] * 1e-3) / 1e9
            traffic, tsrc = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["c4_loop_closure"]
                hit = [v for k, v in tj.items() if k.startswith(KNAME[dom])]
                traffic = hit[0]["dram_bytes_per_launch"]
                tsrc = "profiles/r02_traffic.json (ncu --set full of this workload, same kernel)"
            except Exception:
                pass
            bin_ms = st["ms_hist"] + st["ms_scan"] + st["ms_scatter"] + st["ms_fold"]
            roof = {"bound": "hbm", "kernel": KNAME[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": sb[dom], "avg_launch_ms": st[dom],
                    "launches_per_step": 1,
                    "stages_ms_per_step": st,
                    "stages_frac_of_hbm": {k: sb[k] / (st[k] * 1e-3) / 1e9 / peak for k in sb if st[k] > 0},
                    # the binning pipeline as SURVEY 8(d) defines it: 16 B/point + 80 B/touched cell over key + scan +
                    # scatter + fold together
                    "binning_pipeline": {"algorithmic_bytes": 16.0 * P_ + 80.0 * ct, "ms": bin_ms,
                                         "frac_of_hbm": (16.0 * P_ + 80.0 * ct) / (bin_ms * 1e-3) / 1e9 / peak},
                    "host_ms_per_step": out.get("host_ms_per_step")}
            try:
                # the classifier is FP64-pipe bound, not HBM bound: its pipe utilisation from the committed
                # ncu capture rides along so that its small HBM fraction is read against the right ceiling
                fp = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["fp64_pipe"]
                roof["fp64_bound_kernels"] = {"k_classify": dict(fp["k_classify"], peak_lane_inst_per_s=fp["peak_lane_inst_per_s"],
                                                                 source="profiles/r02_ncu_c4.md, profiles/ubench/pipes.cu")}
            except Exception:
                pass
        if roof is None and "sharded" in out:
            # N > 1: the dominant stage on rank 0 is the fold of its slab (k_fold3, HBM-bound): 16 B per received record +
            # 80 B per touched cell over that stage's CUDA-event time (the stage also holds the slab blanking memsets)
            sh = out["sharded"]
            fb = 16.0 * sh["recv_records_rank0"] + 80.0 * sh["cells_touched_rank0"]
            fms = sh["stages_ms_rank0"].get("fold", 0.0)
            if fms > 0:
                ach = fb / (fms * 1e-3) / 1e9
                roof = {"bound": "hbm", "kernel": "k_fold3 (rank 0: blank + fold of its slab)", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": fb, "avg_launch_ms": fms, "launches_per_step": 1,
                        "stages_ms_per_step": sh["stages_ms_rank0"],
                        "routing": {"note": "k_scatter writes records straight into the owners' receive buffers (NVLink stores through "
                                            "CUDA-IPC peer mappings); stage 'alltoall' = that kernel + the completion barrier",
                                    "remote_bytes_rank0": 16.0 * sh["sent_records_rank0"],
                                    "remote_gbs_rank0": sh.get("alltoall_gbs_rank0_out")}}
        cpu = None
        if not args.no_cpu and world == 1:  # the CPU arm is timed at N = 1 only (rank 0); N > 1 lines carry null
            cores = os.cpu_count() or 1
            sec1, pts1, _ = oracle_c4(args.cpu_keyframes, 2, 1, 1)
            cpu = {"value": pts1 / sec1, "unit": "points/s", "cores": 1, "kind": "port",
                   "sample": f"loop-closure re-integration of the first {args.cpu_keyframes} of the {nk} keyframes ({pts1} points), "
                             f"median of 2 steps after 1 warm-up; oracle port -O3, single worker thread as the reference runs a map "
                             f"(LocalMap.cpp:76-77); host has {cores} cores"}
            if cores > 1:
                secn, ptsn, _ = oracle_c4(args.cpu_keyframes, 2, 1, cores)
                cpu["all_cores"] = {"value": ptsn / secn, "unit": "points/s", "cores": cores,
                                    "note": "flavour 2: dirty cells classified in parallel as in the archived node "
                                            "(archive/local_traversability_monolithic.cpp:349-362); binning stays on one thread"}
        line = {
            "metric": "points/s binned+classified", "value": out["value"], "unit": "points/s",
            "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": out["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64 moments / f32 layers", "data": "synthetic",
            "config": {"workload": WORKLOAD, "keyframes": nk, "points_per_step": out["points_per_step"],
                       "rays_per_keyframe": C4_RINGS * C4_AZ,
                       "l2": "inputs larger than L2 (keyframe store %.1f GB, every step re-reads all of it)" % (out["points_per_step"] * 16 / 1e9),
                       "timing": "median of the K timed steps; every step is kept in ms_all",
                       "parallelism": out["parallelism"]},
            "ms_all": out["ms_all"],
            "loop_closure_ms": out["ms_per_step"],
            "clocks": clocks,
            "e2e": out["e2e"],
            "gpu_launches": out["gpu_launches"],
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        for k in ("shard_parity", "sharded", "pgo_update"):
            if k in out:
                line[k] = out[k]
        line.update(secondary)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if out.get("shard_parity") is False:
        sys.exit(3)


# ---------------------------------------------------------------------------------------------------
def ingest(tm, s, idxs, poses_for_rays, rings, az, device):
    n = 0
    for i, c in gen_host_clouds(SEED + 1, idxs, poses_for_rays, rings, az, device):
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, c) == tm.OK   # host cloud -> H2D + prune on the device
        n += 1
    return n


def bench_c4_single(tm, synth, torch, stream, local, nk, true, drift, K, Wm, sampler, barrier):
    p = tm.default_params(device=local, max_batch=0, profile=1, **C4)
    s = tm.System(p)
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    m = s.getLocalMap(0)
    m.setStream(stream.cuda_stream)
    ingest(tm, s, range(nk), true, C4_RINGS, C4_AZ, "cuda")
    kf_ids = np.arange(1, nk + 1, dtype=np.uint64)
    pin_true = torch.from_numpy(np.ascontiguousarray(true.reshape(nk, 12))).pin_memory()
    pin_drift = torch.from_numpy(np.ascontiguousarray(drift.reshape(nk, 12))).pin_memory()
    with m.locked():
        s.updateKeyFrames(kf_ids, pin_drift.numpy())
    s.flush()   # first integration at the drifted poses (not timed)
    kx, ky = m.gridExtent()
    occ = torch.empty((2 * ky + 1, 2 * kx + 1), dtype=torch.int8).pin_memory()
    occ_np = occ.numpy()

    def step(rep, with_export):
        target = pin_true if rep % 2 == 0 else pin_drift
        with m.locked():   # LocalMap::getGridMapMutex: the PGO message and the closure are queued as one batch
            s.updateKeyFrames(kf_ids, target.numpy())
            s.informLoopClosure()
        s.flush()
        if with_export:
            m.exportOccupancy(-kx, -ky, 2 * kx + 1, 2 * ky + 1, out=occ_np)

    def timed(n_warm, n_timed, with_export):
        ms, st0, st1 = [], None, None
        for rep in range(n_warm + n_timed):
            if rep == n_warm:
                barrier()
                st0 = m.stats()
            with torch.cuda.stream(stream):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                stream.synchronize()
                e0.record(stream)
                step(rep, with_export)
                e1.record(stream)
                e1.synchronize()
            if rep >= n_warm:
                ms.append(e0.elapsed_time(e1))
                sampler.sample()
        return ms, st0, m.stats()

    ms, st0, st1 = timed(Wm, K, False)
    pts = int(st1["last_batch_points"])
    stage = {k: (st1["sum_" + k] - st0["sum_" + k]) / K for k in STAGES}
    host = {k: (st1["sum_host_ms_" + k] - st0["sum_host_ms_" + k]) / K for k in ("plan", "launch", "wait", "batch")}
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    e_ms, et0, et1 = timed(1, K, True)
    h2d = (et1["h2d_bytes"] - et0["h2d_bytes"]) / K
    d2h = (et1["d2h_bytes"] - et0["d2h_bytes"]) / K
    med, emed = float(np.median(ms)), float(np.median(e_ms))
    res = {"value": pts / (med * 1e-3), "ms_per_step": med, "ms_all": ms, "points_per_step": pts,
           "cells_touched": int(st1["last_batch_cells_touched"]), "cells_classified": int(st1["last_batch_cells_classified"]),
           "stage_ms": stage, "host_ms_per_step": host, "gpu_launches": int(launches), "parallelism": "single",
           "e2e": {"value": pts / (emed * 1e-3), "unit": "points/s", "ms_per_step": emed, "ms_all": e_ms,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "note": "bytes counted by the library: h2d = the pass's descriptor block, which carries the %d host poses to the "
                           "device; d2h = the whole-map occupancy grid (%d x %d int8) + the pass's result block" % (nk, 2 * kx + 1, 2 * ky + 1)}}
    # the PGO message alone (no informLoopClosure): every keyframe subtracted at its old pose and added at the new one
    pgo = []
    for rep in range(3):
        target = pin_true if rep % 2 == 1 else pin_drift
        barrier()
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            e0.record(stream)
            with m.locked():
                s.updateKeyFrames(kf_ids, target.numpy())
            s.flush()
            e1.record(stream)
            e1.synchronize()
        pgo.append(e0.elapsed_time(e1))
    res["pgo_update"] = {"ms": float(np.median(pgo)), "ms_all": pgo,
                         "note": "tmap_update_poses of all keyframes without informLoopClosure: subtract + add, 2x the points"}
    s.close()
    return res


def shard_parity_check(tm, synth, torch, dist, local, rank, world, n_kf=48):
    """A reduced C4 map (first n_kf keyframes, 131 072 rays each) rebuilt sharded over the ranks and on rank 0
    alone through the ordinary API; every moment and derived layer compared bit for bit."""
    from traversability_mapping_b200 import sharding
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    poses = synth.trajectory("loop", n_kf, seed=SEED + 7, extent=20.0)
    P = dict(C4, half_size=45.0)
    clouds = dict(gen_host_clouds(SEED + 7, range(n_kf), poses, 128, 1024, "cuda"))
    sm = sharding.ShardedMap(tm.default_params(device=local, max_batch=0, **P))
    sm.sys.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    per = (n_kf + world - 1) // world
    for i in range(n_kf):
        if i // per == rank:
            assert sm.add_keyframe(i + 1, clouds[i]) == tm.OK
            sm.set_pose(i + 1, poses[i])
    sm.rebuild()
    layers = helpers.MOMENT_LAYERS + [l for l in helpers.DERIVED_LAYERS if l != "step_haz_inflated"]
    got = {}
    origin = None
    for l in layers:
        got[l], origin = sm.gather_layer(l)
    ok, cells = True, 0
    if rank == 0:
        s = tm.System(tm.default_params(device=local, max_batch=0, **P))
        s.addNewLocalMap(0)
        s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
        for i in range(n_kf):
            s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i])
        m = s.getLocalMap(0)
        with m.locked():
            for i in range(n_kf):
                s.updateKeyFrame(i + 1, poses[i])
            s.informLoopClosure()
        s.flush()
        h, w = got[layers[0]].shape
        for l in layers:
            ref = m.exportDense(l, origin[0], origin[1], w, h)
            if not helpers.bits_equal(got[l], ref):
                ok = False
        cells = int(np.sum(~np.isnan(got["N"])))
        s.close()
    flag = torch.tensor([1 if ok else 0, cells], device="cuda")
    dist.broadcast(flag, 0)
    sm.close()
    return bool(int(flag[0])), int(flag[1]), n_kf


def bench_c4_sharded(tm, synth, torch, dist, local, rank, world, nk, true, drift, K, Wm, sampler, barrier):
    from traversability_mapping_b200 import sharding
    parity, parity_cells, parity_kf = shard_parity_check(tm, synth, torch, dist, local, rank, world)
    per = (nk + world - 1) // world
    mine = [i for i in range(nk) if i // per == rank]
    sm = sharding.ShardedMap(tm.default_params(device=local, max_batch=0, **C4))
    sm.sys.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    for i, c in gen_host_clouds(SEED + 1, mine, true, C4_RINGS, C4_AZ, "cuda"):
        assert sm.add_keyframe(i + 1, c) == tm.OK
    st0 = sm.map.stats()
    my_ids = np.array([i + 1 for i in mine], dtype=np.uint64)
    ms, info, tim_sum = [], None, {}
    for rep in range(Wm + K):
        target = true if rep % 2 == 0 else drift
        sm.set_poses(my_ids, target[mine])
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        tim = {} if rep >= Wm else None
        info = sm.rebuild(tim)
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max over ranks
        if rep == Wm - 1:
            st0 = sm.map.stats()
        if rep >= Wm:
            ms.append(float(t))
            sampler.sample()
            for k, v in tim.items():
                tim_sum[k] = tim_sum.get(k, 0.0) + v / K
    st1 = sm.map.stats()
    pts = torch.tensor([float(info["local_records"]), float(info["sent_records"]), float(st1["kernel_launches"] - st0["kernel_launches"])],
                       device="cuda", dtype=torch.float64)
    dist.all_reduce(pts)
    med = float(np.median(ms))
    a2a_ms = tim_sum.get("alltoall", 0.0)
    sent_bytes_rank0 = float(info["sent_records"]) * 16.0
    res = {"value": float(pts[0]) / (med * 1e-3), "ms_per_step": med, "ms_all": ms, "points_per_step": int(pts[0]),
           "gpu_launches": int(pts[2]), "parallelism": "sharded%d: slabs of bin rows, all-to-all point routing, halo exchange" % world,
           "shard_parity": parity,
           "sharded": {"routed_points": int(pts[1]), "stages_ms_rank0": tim_sum, "slabs": info["slabs"],
                       "recv_records_rank0": int(info["recv_records"]), "sent_records_rank0": int(info["sent_records"]),
                       "cells_touched_rank0": int(st1["last_batch_cells_touched"]), "grid_bytes_rank0": int(info["grid_bytes"]),
                       "host_code": "C++ (tmap_shard_rebuild): NCCL all-gathers + CUDA-IPC peer-mapped receive buffers, one call per rank",
                       "alltoall_gbs_rank0_out": (sent_bytes_rank0 / (a2a_ms * 1e-3) / 1e9) if a2a_ms > 0 else None,
                       "parity_check": {"keyframes": parity_kf, "cells": parity_cells, "bit_identical_to_1_gpu": parity}},
           # end to end for the sharded job: the step already starts from host poses (set_pose takes host arrays; they ride the op
           # descriptors to the device) -- the slab exports stay on their ranks, so e2e adds the per-rank occupancy read-back
           "e2e": None}
    # e2e: same step + every rank reads its slab's occupancy back to the host
    e_ms = []
    d2h = 0
    for rep in range(1 + K):
        target = true if rep % 2 == 1 else drift
        sm.set_poses(my_ids, target[mine])
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        sm.rebuild()
        bx0, by0, btw, bth = sm.last["bbox"]
        cj0, cj1 = sm.last["rows"]
        o = sm.map.exportOccupancy(bx0 * 16, cj0, btw * 16, cj1 - cj0)
        d2h = o.size
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep >= 1:
            e_ms.append(float(t))
    emed = float(np.median(e_ms))
    res["e2e"] = {"value": float(pts[0]) / (emed * 1e-3), "unit": "points/s", "ms_per_step": emed, "ms_all": e_ms,
                  "h2d_bytes_per_step": int(len(mine) * 48 + len(mine) * 160), "d2h_bytes_per_step": int(d2h),
                  "note": "per rank: its keyframes' host poses (+ op descriptors) in, its slab's occupancy grid out"}
    sm.close()
    return res


# ---------------------------------------------------------------------------------------------------
def bench_c2(tm, synth, torch, stream, local, K, Wm, barrier):
    """config[1]: 10 Hz stream into a 40 x 40 m rolling grid @0.05 m, window 15, inflation on; a step = one stream
    tick as local_traversability_node.cpp:153-185 drives the library (max_batch = 1, reference-exact)."""
    n_kf = WINDOW + Wm + K
    poses = synth.trajectory("drive", n_kf, seed=SEED, step=0.1, curvature=0.02)
    clouds = [c for _, c in gen_host_clouds(SEED, range(n_kf), poses, 128, 1024, "cuda")]
    pinned = [torch.from_numpy(c).pin_memory() for c in clouds]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def new_system():
        p = tm.default_params(device=local, max_batch=1, profile=1, **C2)
        s = tm.System(p)
        s.addNewLocalMap(0)
        s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
        m = s.getLocalMap(0)
        m.setStream(stream.cuda_stream)
        return s, m

    res = {}
    for arm in ("device", "e2e"):
        s, m = new_system()
        if arm == "device":
            for i in range(n_kf):
                assert s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i]) == tm.OK
        occ = torch.empty((CROP, CROP), dtype=torch.int8).pin_memory()
        occ_np = occ.numpy()
        ms, npts, st0 = [], [], None
        for i in range(n_kf):
            timed = i >= WINDOW + Wm
            if i == WINDOW + Wm:
                barrier()
                st0 = m.stats()
            with torch.cuda.stream(stream):
                if timed:
                    flush_buf.fill_(i & 0xFF)  # L2 flush between timed iterations (untimed)
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    stream.synchronize()
                    e0.record(stream)
                if arm == "e2e":
                    assert s.addNewKeyFrameWithPCL(i, i + 1, 0, pinned[i].numpy()) == tm.OK   # H2D + prune
                s.updateKeyFrame(i + 1, poses[i])
                s.flush()
                if i >= WINDOW:
                    s.deleteKeyFrame(i + 1 - WINDOW)
                if arm == "e2e":
                    ci = int(round(float(poses[i][0, 3]) / C2["resolution"])); cj = int(round(float(poses[i][1, 3]) / C2["resolution"]))
                    m.exportOccupancy(ci - CROP // 2, cj - CROP // 2, CROP, CROP, out=occ_np)   # D2H
                if timed:
                    e1.record(stream)
                    e1.synchronize()
                    ms.append(e0.elapsed_time(e1))
                    npts.append(s.keyFramePoints(i + 1))
        st1 = m.stats()
        tot = float(np.sum(ms))
        if arm == "device":
            res = {"workload": "C2 local map: 10 Hz stream of 131072-ray keyframes, 40x40 m rolling grid @0.05 m, window 15, "
                               "49-cell disc, inflation 0.5 m; max_batch=1; L2 flushed between ticks",
                   "value": float(np.sum(npts)) / (tot * 1e-3), "unit": "points/s", "ms_per_tick": tot / K, "ticks": K,
                   "points_per_tick": float(np.mean(npts)),
                   "stages_ms_per_tick": {k: (st1["sum_" + k] - st0["sum_" + k]) / K for k in STAGES},
                   "host_ms_per_tick": {k: (st1["sum_host_ms_" + k] - st0["sum_host_ms_" + k]) / K for k in ("plan", "launch", "wait", "batch")},
                   "gpu_launches": int(st1["kernel_launches"] - st0["kernel_launches"])}
        else:
            res["e2e"] = {"value": float(np.sum(npts)) / (tot * 1e-3), "unit": "points/s", "ms_per_tick": tot / K,
                          "h2d_bytes_per_tick": C2_RAYS * 16, "d2h_bytes_per_tick": CROP * CROP}
        s.close()
    return res


def bench_c1(tm, synth, torch, stream, local, reps=12):
    """config[0]: ONE 131 072-ray keyframe into a 20 x 20 m local grid @0.1 m (13-cell disc): latency of the first
    bin + classify after the pose arrives (the reference's CPU-runnable case)."""
    P = dict(resolution=0.1, half_size=10.0, max_range=10.0, min_range=1.0, robot_radius=0.20, min_points_per_grid=3)
    pose = synth.trajectory("single", 1, seed=SEED + 5)[0]
    cloud = next(gen_host_clouds(SEED + 5, [0], pose[None], 128, 1024, "cuda"))[1]
    us, npts = [], 0
    p = tm.default_params(device=local, max_batch=1, **P)
    s = tm.System(p)
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    m = s.getLocalMap(0)
    m.setStream(stream.cuda_stream)
    for rep in range(reps + 3):
        assert s.addNewKeyFrameWithPCL(rep, rep + 1, 0, cloud) == tm.OK
        npts = s.keyFramePoints(rep + 1)
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            e0.record(stream)
            s.updateKeyFrame(rep + 1, pose)
            s.flush()
            e1.record(stream)
            e1.synchronize()
        if rep >= 3:
            us.append(1e3 * e0.elapsed_time(e1))
        s.deleteKeyFrame(rep + 1)   # back to an empty map
    s.close()
    med = float(np.median(us))
    return {"workload": "C1: single 131072-ray keyframe into a 20x20 m grid @0.1 m, 13-cell disc (bin + classify after the pose update)",
            "us_per_keyframe": med, "us_all": us, "points": npts, "value": npts / (med * 1e-6), "unit": "points/s"}


def bench_c3(tm, synth, torch, stream, local, barrier, nk=500, reps=5):
    """config[2]: global map from 500 ground-truth-pose keyframes (131 072 rays each), 200 x 200 m @0.1 m: one batched
    build from a blank map."""
    poses = synth.trajectory("loop", nk, seed=SEED + 3, extent=C4_EXTENT)
    p = tm.default_params(device=local, max_batch=0, profile=1, **C4)
    s = tm.System(p)
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    m = s.getLocalMap(0)
    m.setStream(stream.cuda_stream)
    for i, c in gen_host_clouds(SEED + 3, range(nk), poses, 128, 1024, "cuda"):
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, c) == tm.OK
    ids = np.arange(1, nk + 1, dtype=np.uint64)
    ms = []
    for rep in range(reps + 2):
        barrier()
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            e0.record(stream)
            with m.locked():
                s.updateKeyFrames(ids, poses)
                s.informLoopClosure()
            s.flush()
            e1.record(stream)
            e1.synchronize()
        if rep >= 2:
            ms.append(e0.elapsed_time(e1))
    st = m.stats()
    s.close()
    med = float(np.median(ms))
    pts = int(st["last_batch_points"])
    return {"workload": "C3: global map from 500 keyframes x 131072 rays, 200x200 m @0.1 m, 13-cell disc (one batched build from blank)",
            "ms": med, "ms_all": ms, "points": pts, "value": pts / (med * 1e-3), "unit": "points/s",
            "cells_classified": int(st["last_batch_cells_classified"])}


if __name__ == "__main__":
    main()

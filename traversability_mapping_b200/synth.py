"""Deterministic synthetic LiDAR-like keyframes and trajectories.

A keyframe is 128 rings x 1024 azimuth steps = 131 072 rays from a sensor 1.0 m above the base,
elevation -25..+15 deg, intersected with a noisy heightfield (smooth undulation + a few steps and
ramps) and returned in the SENSOR frame as pcl::PointXYZ-strided float32 (x, y, z, pad).  Rays that
miss within range come back as (0, 0, 0) -- the reference's ego/invalid filter (System.cpp:103) --
and ~0.1 % of returns are NaN.  Everything is a pure function of (seed, keyframe index, pose), via a
counter-based Philox generator, so CPU oracle and GPU see identical bytes.
"""
import math

import numpy as np

SENSOR_HEIGHT = 1.0


def terrain(x, y):
    h = 0.15 * np.sin(0.7 * x) * np.cos(0.5 * y) + 0.05 * np.sin(3.1 * x + 1.3 * y)
    # steps / kerbs and a ramp
    h = h + 0.30 * ((np.mod(x + 40.0, 17.0) > 8.0) & (np.mod(x + 40.0, 17.0) < 11.0) & (np.mod(y + 40.0, 13.0) > 6.0))
    h = h + 0.20 * (np.mod(y + 23.0, 29.0) > 25.0)
    h = h + np.clip(0.15 * (np.mod(x - y + 60.0, 41.0) - 30.0), 0.0, 0.6) * (np.mod(x - y + 60.0, 41.0) > 30.0)
    return h


def pose_matrix(x, y, z, roll, pitch, yaw):
    """map <- base, row-major 3x4 float32 (computed in float64, then cast like Affine3d -> Affine3f)."""
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    T = np.zeros((3, 4))
    T[:, :3] = R
    T[:, 3] = (x, y, z)
    return T.astype(np.float32)


IDENTITY = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], dtype=np.float32)
T_BASE_SENSOR = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, SENSOR_HEIGHT]], dtype=np.float32)


def make_keyframe(seed, kf_index, pose, max_range=20.0, rings=128, azimuths=1024, noise=0.01, nan_frac=0.001):
    """Raycast from `pose` (3x4 map<-base).  Returns (rings*azimuths, 4) float32, sensor frame."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed) & 0xFFFFFFFFFFFFFFFF, int(kf_index) & 0xFFFFFFFFFFFFFFFF]))
    T = np.asarray(pose, dtype=np.float64).reshape(3, 4)
    R, t = T[:, :3], T[:, 3]
    el = np.deg2rad(np.linspace(-25.0, 15.0, rings))
    az = np.linspace(0.0, 2.0 * np.pi, azimuths, endpoint=False)
    EL, AZ = np.meshgrid(el, az, indexing="ij")
    d_s = np.stack([np.cos(EL) * np.cos(AZ), np.cos(EL) * np.sin(AZ), np.sin(EL)], axis=-1).reshape(-1, 3)
    o_m = R @ np.array([0.0, 0.0, SENSOR_HEIGHT]) + t
    d_m = d_s @ R.T
    n = d_s.shape[0]
    # downward rays: fixed-point iteration of t = (h(p_xy(t)) - o_z) / d_z from the flat-ground guess
    # (no occlusion test -- the terrain is gentle; vertical faces of the steps are simply not sampled)
    down = d_m[:, 2] < -1e-3
    dz = np.where(down, d_m[:, 2], -1.0)
    h0 = float(terrain(np.float64(o_m[0]), np.float64(o_m[1])))
    tt = (h0 - o_m[2]) / dz
    for _ in range(8):
        tt = np.clip(tt, 0.0, max_range * 1.2)
        px = o_m[0] + tt * d_m[:, 0]
        py = o_m[1] + tt * d_m[:, 1]
        tt = 0.5 * tt + 0.5 * (terrain(px, py) - o_m[2]) / dz
    found = down & (tt > 0.0)
    lo = hi = tt
    rng_t = 0.5 * (lo + hi) + noise * rng.standard_normal(n)
    pts = d_s * rng_t[:, None]
    ok = found & (rng_t > 0.05) & (rng_t <= max_range * 1.05)
    out = np.zeros((n, 4), dtype=np.float32)
    out[ok, :3] = pts[ok].astype(np.float32)
    bad = rng.random(n) < nan_frac
    out[bad, :3] = np.nan
    return out


def trajectory(kind, n, seed=0, **kw):
    """Poses (n, 3, 4) float32.
    kind: 'single'  one random SE(3) near the origin (C1)
          'drive'   10 Hz drive at 1 m/s on a gentle arc (C2)
          'loop'    Lissajous loop inside +-extent m, ~1 m spacing (C3/C4/C5)"""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 0x7A11]))
    poses = []
    if kind == "single":
        for _ in range(n):
            x, y = rng.uniform(-1, 1, 2)
            yaw = rng.uniform(-np.pi, np.pi)
            r, p = np.deg2rad(rng.uniform(-3, 3, 2))
            poses.append(pose_matrix(x, y, float(terrain(np.float64(x), np.float64(y))), r, p, yaw))
    elif kind == "drive":
        step = kw.get("step", 0.1)
        curv = kw.get("curvature", 0.02)
        x = y = 0.0
        yaw = 0.3
        for _ in range(n):
            r, p = np.deg2rad(rng.uniform(-1.5, 1.5, 2))
            poses.append(pose_matrix(x, y, float(terrain(np.float64(x), np.float64(y))), r, p, yaw))
            x += step * math.cos(yaw)
            y += step * math.sin(yaw)
            yaw += curv * step
    elif kind == "loop":
        extent = kw.get("extent", 90.0)
        s = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
        xs = extent * np.sin(2.0 * s + 0.3)
        ys = extent * np.sin(3.0 * s)
        dx = np.gradient(xs)
        dy = np.gradient(ys)
        for i in range(n):
            yaw = math.atan2(dy[i], dx[i])
            r, p = np.deg2rad(rng.uniform(-2, 2, 2))
            poses.append(pose_matrix(xs[i], ys[i], float(terrain(np.float64(xs[i]), np.float64(ys[i]))), r, p, yaw))
    else:
        raise ValueError(kind)
    return np.stack(poses)


def drift_poses(true_poses, per_kf_forward=0.05, per_kf_yaw_deg=1.0):
    """The reference simulator's drift model (ground_truth_kfs/src/slam_keyframe_sim_with_drift.cpp:146-151):
    every keyframe adds +0.05 m along heading and +1 deg yaw, accumulating."""
    out = []
    acc_f, acc_yaw = 0.0, 0.0
    for T in np.asarray(true_poses, dtype=np.float64):
        acc_f += per_kf_forward
        acc_yaw += math.radians(per_kf_yaw_deg)
        c, s = math.cos(acc_yaw), math.sin(acc_yaw)
        Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        D = T.copy()
        D[:, :3] = Rz @ T[:, :3]
        heading = T[:2, 0] / (np.linalg.norm(T[:2, 0]) + 1e-12)
        D[:2, 3] = T[:2, 3] + acc_f * heading
        out.append(D.astype(np.float32))
    return np.stack(out)


# ---------------------------------------------------------------------------------------------
# Batched generator (torch, CPU or CUDA): the same sensor and terrain model as make_keyframe, for the
# workloads whose cloud count makes the numpy path too slow (2 000 keyframes x 262 144 rays).  The random
# stream is torch's (seeded per batch), so its clouds are NOT the numpy generator's; both arms of a
# comparison always consume the same host arrays.  Plumbing for benchmarks / tests, not part of the path.
# ---------------------------------------------------------------------------------------------
def _terrain_t(torch, x, y):
    h = 0.15 * torch.sin(0.7 * x) * torch.cos(0.5 * y) + 0.05 * torch.sin(3.1 * x + 1.3 * y)
    mx = torch.remainder(x + 40.0, 17.0)
    h = h + 0.30 * ((mx > 8.0) & (mx < 11.0) & (torch.remainder(y + 40.0, 13.0) > 6.0)).to(h.dtype)
    h = h + 0.20 * (torch.remainder(y + 23.0, 29.0) > 25.0).to(h.dtype)
    md = torch.remainder(x - y + 60.0, 41.0)
    h = h + torch.clamp(0.15 * (md - 30.0), 0.0, 0.6) * (md > 30.0).to(h.dtype)
    return h


def make_keyframes_torch(seed, first_index, poses, max_range=20.0, rings=128, azimuths=1024, noise=0.01, nan_frac=0.001,
                         device="cpu"):
    """Raycast len(poses) keyframes at once.  Returns a (K, rings*azimuths, 4) float32 tensor on `device`,
    sensor frame, pcl::PointXYZ layout; a pure function of (seed, first_index, poses, device kind)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed((int(seed) * 1000003 + int(first_index) * 7919 + 12345) & 0x7FFFFFFFFFFFFFFF)
    f64 = torch.float64
    T = torch.as_tensor(np.asarray(poses, dtype=np.float64).reshape(-1, 3, 4), device=dev)
    K = T.shape[0]
    R, t = T[:, :, :3], T[:, :, 3]
    el = torch.deg2rad(torch.linspace(-25.0, 15.0, rings, dtype=f64, device=dev))
    az = torch.arange(azimuths, dtype=f64, device=dev) * (2.0 * math.pi / azimuths)
    EL, AZ = torch.meshgrid(el, az, indexing="ij")
    d_s = torch.stack([torch.cos(EL) * torch.cos(AZ), torch.cos(EL) * torch.sin(AZ), torch.sin(EL)], dim=-1).reshape(-1, 3)
    n = d_s.shape[0]
    o_m = R[:, :, 2] * SENSOR_HEIGHT + t                      # (K, 3)
    d_m = torch.einsum("nj,kij->kni", d_s, R)                 # (K, n, 3)
    down = d_m[:, :, 2] < -1e-3
    dz = torch.where(down, d_m[:, :, 2], torch.full_like(d_m[:, :, 2], -1.0))
    ox, oy, oz = o_m[:, 0:1], o_m[:, 1:2], o_m[:, 2:3]
    tt = (_terrain_t(torch, ox, oy) - oz) / dz
    for _ in range(8):
        tt = torch.clamp(tt, 0.0, max_range * 1.2)
        px = ox + tt * d_m[:, :, 0]
        py = oy + tt * d_m[:, :, 1]
        tt = 0.5 * tt + 0.5 * (_terrain_t(torch, px, py) - oz) / dz
    found = down & (tt > 0.0)
    rng_t = tt + noise * torch.randn((K, n), dtype=f64, device=dev, generator=g)
    ok = found & (rng_t > 0.05) & (rng_t <= max_range * 1.05)
    out = torch.zeros((K, n, 4), dtype=torch.float32, device=dev)
    pts = (d_s.unsqueeze(0) * rng_t.unsqueeze(-1)).to(torch.float32)
    out[:, :, :3] = torch.where(ok.unsqueeze(-1), pts, torch.zeros_like(pts))
    bad = torch.rand((K, n), device=dev, generator=g) < nan_frac
    out[:, :, :3] = torch.where(bad.unsqueeze(-1), torch.full_like(pts, float("nan")), out[:, :, :3])
    return out

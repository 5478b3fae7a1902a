"""the bookkeeping the reference's ROS 2 nodes keep around the library (rolling keyframe window,
obstacle layer with temporal expiry), restated without ROS in traversability_mapping_b200/ros_shim.py (+ the C++ twin
include/tmap_ros_shim.hpp).  CPU tests: the call sequences against hand-derived expectations from
local_traversability_node.cpp:153-185 and global_traversability_node.cpp:224-283, the drivers on the oracle, and the header's
syntax.  The GPU parity test drives the product and the oracle with the same drivers and compares the maps after every
publish tick."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from traversability_mapping_b200 import ros_shim, synth  # noqa: E402
import helpers  # noqa: E402


class Recorder:
    """A System stand-in that records the calls and hands out obstacle ids as System.cpp:288-335 does."""

    def __init__(self, prune_empty=()):
        self.calls, self.next, self.free, self.prune_empty, self.n_obst = [], 2 ** 63 - 1, [], set(prune_empty), 0

    def addNewKeyFrameWithPCL(self, ts, kf, mp, cloud): self.calls.append(("add", kf, mp))
    def updateKeyFrame(self, kf, T): self.calls.append(("pose", kf))
    def deleteKeyFrame(self, kf): self.calls.append(("del", kf))
    def setObstacleParameters(self, T, mx, mn): self.calls.append(("obst_params", mx, mn))

    def addObstacleKeyFrame(self, ts, cloud):
        self.n_obst += 1
        if self.n_obst in self.prune_empty:
            return 0
        oid = self.free.pop() if self.free else self.next
        if oid == self.next:
            self.next -= 1
        self.calls.append(("obst_add", oid))
        return oid

    def deleteObstacleKeyFrame(self, kf):
        self.calls.append(("obst_del", kf))
        self.free.append(kf)


def pose_at(x, y=0.0, z=0.0):
    return np.array([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z]], dtype=np.float32)


def test_rolling_window_gate_ids_and_eviction():
    rec = Recorder()
    d = ros_shim.RollingWindowDriver(rec, window_size=3, keyframe_min_displacement_m=0.5)
    xs = [0.0, 0.2, 0.6, 0.9, 1.2, 1.8, 2.4, 3.0]        # 0.2 and 0.9 have not moved 0.5 m since the last KEYFRAME
    got = [d.scan(1000 * i, None, pose_at(x)) for i, x in enumerate(xs)]
    assert got == [0, None, 1, None, 2, 3, 4, 5]            # ids contiguous, advanced only by scans that land
    assert d.robot_xy == (3.0, 0.0)                         # the crop follows every scan, dropped ones included
    want = []
    for kf in range(6):
        want += [("add", kf, 0), ("pose", kf)]
        if kf >= 3:
            want.append(("del", kf - 3))                    # id >= window_size: drop id - window_size
    assert rec.calls == want


def test_obstacle_layer_gate_deadlines_and_expiry():
    rec = Recorder(prune_empty={3})
    d = ros_shim.ObstacleLayerDriver(rec, max_range=6.0, min_range=0.3, add_rate_hz=2.0, temporal_buffer_s=1.0)
    I = pose_at(0.0)
    cloud = np.zeros((1, 4), np.float32)
    assert d.cloud(0.00, 0, cloud, None, I) == 0 and rec.calls == []          # extrinsic TF not up yet: nothing happens
    a = d.cloud(0.10, 1, cloud, I, I)                                          # resolves the extrinsic, fuses
    assert a == 2 ** 63 - 1 and rec.calls[0] == ("obst_params", 6.0, 0.3)
    assert d.cloud(0.30, 2, cloud, I, I) == 0                                  # 0.2 s < 0.5 s period: rate gate
    b = d.cloud(0.60, 3, cloud, I, I)
    assert b == 2 ** 63 - 2
    assert d.cloud(1.15, 4, cloud, I, None) == 0                               # map <- base lookup failed: dropped, but the gate moved
    assert d.cloud(1.40, 5, cloud, I, I) == 0                                  # 0.25 s after the dropped one: gated
    assert d.cloud(1.70, 6, cloud, I, I) == 0                                  # pruned empty (3rd addObstacleKeyFrame): no deadline
    assert d.expire(1.05) == []                                                # a's deadline is 1.10
    assert d.expire(1.10) == [a]
    c = d.cloud(2.30, 7, cloud, I, I)
    assert c == a                                                              # the freed id is reused
    assert d.expire(9.0) == [b, c]                                             # std::map order: ascending id
    assert [x for x in rec.calls if x[0].startswith("obst_")] == [
        ("obst_params", 6.0, 0.3), ("obst_add", a), ("obst_add", b), ("obst_del", a), ("obst_add", a), ("obst_del", b), ("obst_del", a)]
    assert d.deadlines == {}


def test_cpp_twin_header_compiles(tmp_path):
    src = tmp_path / "shim.cpp"
    src.write_text('#include "tmap_ros_shim.hpp"\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), str(src)], check=True)


SCENE = dict(resolution=0.1, half_size=10.0, max_range=8.0, min_range=1.0, min_points_per_grid=2)
T_BASE_OBS = np.array([[1, 0, 0, 0.2], [0, 1, 0, 0.0], [0, 0, 1, 0.6]], dtype=np.float32)   # a depth camera on the bonnet


def obstacle_cloud(k):
    """A small box of returns 2-3 m ahead of the camera, different every call (sensor frame, pcl::PointXYZ layout)."""
    rng = np.random.Generator(np.random.Philox(key=[77, k]))
    n = 400
    c = np.zeros((n, 4), dtype=np.float32)
    c[:, 0] = rng.uniform(2.0, 3.0, n)
    c[:, 1] = rng.uniform(-0.5 + 0.1 * k, 0.5 + 0.1 * k, n)
    c[:, 2] = rng.uniform(-0.5, 0.2, n)
    return c


def drive(system, settle, n_ticks=24):
    """The same node-level scenario for any System-like object: a rolling window of 3 keyframes at 2 Hz, obstacle clouds at
    5 Hz through a 2 Hz gate with a 1.2 s temporal buffer, a 10 Hz publish timer that expires obstacles.  `settle()` is
    called at every timer tick (flush + compare)."""
    poses = synth.trajectory("drive", 8, seed=5, step=0.6)
    win = ros_shim.RollingWindowDriver(system, window_size=3, keyframe_min_displacement_m=0.3)
    obs = ros_shim.ObstacleLayerDriver(system, max_range=6.0, min_range=0.3, add_rate_hz=2.0, temporal_buffer_s=1.2)
    log = []
    for tick in range(n_ticks):
        now = 0.1 * tick
        if tick % 5 == 0 and tick // 5 < len(poses):
            i = tick // 5
            c = synth.make_keyframe(5, i, poses[i], max_range=8.0, rings=32, azimuths=512)
            log.append(("kf", win.scan(int(now * 1e9), c, poses[i])))
        if tick % 2 == 1:
            i = min(tick // 5, len(poses) - 1)
            log.append(("obst", obs.cloud(now, int(now * 1e9), obstacle_cloud(tick), T_BASE_OBS if tick >= 3 else None, poses[i])))
        log.append(("expired", tuple(obs.expire(now))))
        settle(tick)
    return log


def test_drivers_on_the_oracle():
    import pyoracle as po
    om = po.OracleMap(po.default_params(**SCENE))
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    sys_o = helpers.OracleSystem(om)
    counts = []
    log = drive(sys_o, lambda tick: counts.append((om.num_keyframes(), om.occupied_cell_keys().size)))
    kfs = [v for k, v in log if k == "kf"]
    assert kfs == [0, 1, 2, 3, 4]
    fused = [v for k, v in log if k == "obst" and v]
    expired = [x for k, v in log if k == "expired" for x in v]
    assert len(fused) >= 3 and len(expired) >= 1 and set(expired) <= set(fused)
    assert max(n for n, _ in counts) <= 3 + 3          # window keyframes + live obstacle keyframes
    assert counts[-1][1] > 0
    om.close()


@pytest.mark.gpu
def test_ros_wrapper_scenario_matches_oracle(tmap, oracle):
    """The node-level scenario through the product and through the oracle: same drivers, same ids, maps equal after every
    publish tick (moment layers bit for bit, derived layers within 1e-4, changed-cell sets equal)."""
    s, m, om = helpers.make_pair(tmap, oracle, **SCENE)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    sys_o = helpers.OracleSystem(om)
    log_o = drive(sys_o, lambda tick: None)
    # replay on the product, comparing at every tick against a second oracle advanced in lock step
    om2 = oracle.OracleMap(oracle.params_from(tmap.default_params(**SCENE)))
    om2.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    sys_o2 = helpers.OracleSystem(om2)

    class Both:
        def addNewKeyFrameWithPCL(self, *a): s.addNewKeyFrameWithPCL(*a); sys_o2.addNewKeyFrameWithPCL(*a)
        def updateKeyFrame(self, *a): s.updateKeyFrame(*a); s.flush(); sys_o2.updateKeyFrame(*a)
        def deleteKeyFrame(self, *a): s.deleteKeyFrame(*a); sys_o2.deleteKeyFrame(*a)
        def setObstacleParameters(self, *a): s.setObstacleParameters(*a); sys_o2.setObstacleParameters(*a)

        def addObstacleKeyFrame(self, *a):
            g, o = s.addObstacleKeyFrame(*a), sys_o2.addObstacleKeyFrame(*a)
            assert g == o, (g, o)
            return g

        def deleteObstacleKeyFrame(self, *a): s.deleteObstacleKeyFrame(*a); sys_o2.deleteObstacleKeyFrame(*a)

    def settle(tick):
        s.flush()
        helpers.compare_maps(m, om2, check_changed=True, label=f"tick {tick}")

    log_g = drive(Both(), settle)
    assert log_g == log_o
    assert m.numKeyFrames() == om2.num_keyframes()
    s.close()

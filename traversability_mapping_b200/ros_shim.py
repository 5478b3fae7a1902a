"""Host-side restatement of what the reference's ROS 2 nodes do AROUND the library -- the bookkeeping
only, with the ROS clock, TF lookups and message decoding turned into plain arguments, so that a node keeps its callbacks
and swaps `traversability_mapping::System` for this package's `System` (or the C++ mirror: include/tmap_ros_shim.hpp holds
the same two classes).  No arithmetic lives here: every map operation is a call that exists in include/tmap.h.

  RollingWindowDriver  = LocalTraversabilityNode::scanCallback
                         (traversability_ros_interface/traversability_mapping_ros/src/local_traversability_node.cpp:153-185)
  ObstacleLayerDriver  = GlobalTraversabilityNode::obstacleCallback / expireObstacles
                         (.../src/global_traversability_node.cpp:224-283)

`system` is anything with the reference's method names: this package's System, or the oracle adapter of the tests.
"""
import math


class RollingWindowDriver:
    """One synced (cloud, pose) pair -> at most one keyframe; scans that have not moved far enough since the last keyframe
    are dropped; otherwise add + bin at the pose, then evict the keyframe that fell out of the window.  Ids stay contiguous
    because next_id advances only when a scan lands (local_traversability_node.cpp:162-184)."""

    def __init__(self, system, window_size, keyframe_min_displacement_m=0.0, map_id=0):
        self.system = system
        self.window_size = int(window_size)
        self.min_disp = float(keyframe_min_displacement_m)
        self.map_id = map_id
        self.next_id = 0
        self.last_kf_pos = None
        self.robot_xy = None

    def scan(self, stamp_ns, cloud, pose_map_base):
        """pose_map_base: 3x4 (or 12 floats) map <- base.  Returns the keyframe id, or None when the scan was dropped."""
        T = [float(v) for v in list(_flat12(pose_map_base))]
        pos = (T[3], T[7], T[11])
        self.robot_xy = (pos[0], pos[1])            # the crop follows the robot on every scan, dropped ones included
        if self.last_kf_pos is not None:
            d = math.sqrt(sum((a - b) ** 2 for a, b in zip(pos, self.last_kf_pos)))
            if d < self.min_disp:
                return None
        self.last_kf_pos = pos
        kf = self.next_id
        self.next_id += 1
        self.system.addNewKeyFrameWithPCL(int(stamp_ns), kf, self.map_id, cloud)
        self.system.updateKeyFrame(kf, pose_map_base)          # supplies the pose AND triggers the (only) bin
        if kf >= self.window_size:                             # once N + 1 are in flight, drop the oldest
            self.system.deleteKeyFrame(kf - self.window_size)
        return kf


class ObstacleLayerDriver:
    """One obstacle cloud -> one transient keyframe fused into map 0, aged out `temporal_buffer_s` after reception.
    The base <- sensor extrinsic is resolved once (first cloud whose TF lookup succeeds); clouds arriving faster than
    `add_rate_hz` are dropped (global_traversability_node.cpp:224-263); `expire(now)` is the publish timer's
    expireObstacles (:266-283)."""

    def __init__(self, system, max_range, min_range, add_rate_hz, temporal_buffer_s):
        self.system = system
        self.max_range, self.min_range = float(max_range), float(min_range)
        self.add_period_s = 1.0 / float(add_rate_hz) if add_rate_hz > 0 else 0.0
        self.temporal_buffer_s = float(temporal_buffer_s)
        self.ready = False
        self.last_add = None
        self.deadlines = {}          # obstacle keyframe id -> deadline (seconds on the node's clock), insertion order kept

    def cloud(self, now_s, stamp_ns, cloud, T_base_sensor, T_map_base):
        """now_s: the node's clock at reception; T_base_sensor / T_map_base: the two TF lookups (None = lookup failed).
        Returns the obstacle keyframe id, or 0 when the cloud was not fused."""
        if not self.ready:
            if T_base_sensor is None:
                return 0                                        # TF not up yet; try again on the next cloud
            self.system.setObstacleParameters(T_base_sensor, self.max_range, self.min_range)
            self.ready = True
        if self.last_add is not None and (now_s - self.last_add) < self.add_period_s:
            return 0                                            # ingestion-rate gate
        self.last_add = now_s
        if T_map_base is None:
            return 0                                            # neither the stamped nor the latest transform
        oid = self.system.addObstacleKeyFrame(int(stamp_ns), cloud)
        if oid == 0:
            return 0                                            # params unset / cloud pruned empty
        self.system.updateKeyFrame(oid, T_map_base)             # supply pose + trigger the (only) bin
        self.deadlines[oid] = now_s + self.temporal_buffer_s
        return oid

    def expire(self, now_s):
        """Wipe obstacle keyframes whose temporal buffer has elapsed; returns the ids removed (std::map order: ascending id)."""
        gone = [oid for oid in sorted(self.deadlines) if now_s >= self.deadlines[oid]]
        for oid in gone:
            self.system.deleteObstacleKeyFrame(oid)
            del self.deadlines[oid]
        return gone


def _flat12(T):
    try:
        import numpy as np
        return np.asarray(T, dtype=np.float64).reshape(-1)[:12]
    except Exception:
        return [v for row in T for v in row][:12]

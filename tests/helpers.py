"""Shared helpers for the parity tests: run the same scenario through the CUDA path (C ABI) and the
CPU oracle, then compare with the SURVEY 9.8 parity rules."""
import numpy as np

MOMENT_LAYERS = ["N", "sx", "sy", "sz", "sx2", "sy2", "sz2", "sxy", "sxz", "syz", "zmin", "zmax"]
DERIVED_LAYERS = ["hazard", "elevation", "slope_haz", "step_haz", "roughness_haz", "border_haz",
                  "normal_x", "normal_y", "normal_z", "step_haz_inflated"]
RTOL = 1e-4   # north_star: slope / roughness / elevation within 1e-4 relative
ATOL = 1e-7   # floor for values that are themselves ~0


def bits_equal(a, b):
    """float32 arrays equal bit for bit, any NaN matching any NaN."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32)))


def compare_maps(gm, om, check_changed=None, label=""):
    """gm: product LocalMap handle, om: OracleMap.  Asserts the parity rules; returns summary stats."""
    g_occ = gm.occupiedCellKeys()
    o_occ = om.occupied_cell_keys()
    assert np.array_equal(g_occ, o_occ), f"{label}: occupied cell sets differ ({g_occ.size} vs {o_occ.size})"
    assert gm.gridExtent() == om.grid_extent(), f"{label}: grid extent {gm.gridExtent()} vs {om.grid_extent()}"
    # compare over the occupied cells and a one-disc-wide ring of neighbours through the dense export too
    kx, ky = om.grid_extent()
    stats = {"cells": int(o_occ.size)}
    keys = o_occ
    gM = gm.readLayers(keys, MOMENT_LAYERS)
    oM = om.read_layers(keys, MOMENT_LAYERS)
    for i, l in enumerate(MOMENT_LAYERS):
        assert bits_equal(gM[:, i], oM[:, i]), f"{label}: moment layer {l} not bit-identical " \
            f"({np.sum(gM[:, i].view(np.uint32) != oM[:, i].view(np.uint32))} cells)"
    gD = gm.readLayers(keys, DERIVED_LAYERS)
    oD = om.read_layers(keys, DERIVED_LAYERS)
    worst = 0.0
    for i, l in enumerate(DERIVED_LAYERS):
        a, b = gD[:, i], oD[:, i]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{label}: NaN pattern of {l} differs " \
            f"({np.sum(np.isnan(a) != np.isnan(b))} cells)"
        f = ~np.isnan(b)
        if f.any():
            err = np.abs(a[f].astype(np.float64) - b[f]) / (np.abs(b[f].astype(np.float64)) + ATOL / RTOL)
            worst = max(worst, float(err.max()))
            assert np.allclose(a[f], b[f], rtol=RTOL, atol=ATOL), f"{label}: layer {l} max rel err {err.max():.3e}"
    stats["worst_rel"] = worst
    stats["gated_in"] = int(np.sum(~np.isnan(oD[:, 0])))
    # whole-grid check: nothing stored outside the occupied set, occupancy labels
    g_occg = gm.exportOccupancy()
    o_occg = om.export_occupancy()
    assert g_occg.shape == o_occg.shape
    diff = g_occg != o_occg
    if diff.any():
        haz = om.export_dense("hazard")
        v = 100.0 * np.clip(haz[diff].astype(np.float64), 0, 1)
        near = np.abs(v - np.round(v)) <= 100.0 * RTOL * np.maximum(np.abs(v) / 100.0, 1e-3)
        assert near.all(), f"{label}: {int((~near).sum())} occupancy labels differ away from a threshold"
    stats["occupancy_mismatch_at_threshold"] = int(diff.sum())
    if check_changed is not None:
        gc = gm.takeChangedCells()
        oc = om.take_changed_cells()
        assert np.array_equal(gc, oc), f"{label}: changed-cell sets differ ({gc.size} vs {oc.size})"
        stats["changed"] = int(oc.size)
    return stats


def make_pair(tmap, oracle, **kw):
    """A product System with map 0 and the matching OracleMap."""
    p = tmap.default_params(**kw)
    s = tmap.System(p)
    s.addNewLocalMap(0)
    om = oracle.OracleMap(oracle.params_from(p))
    return s, s.getLocalMap(0), om


def compose_f32(A, B):
    """Eigen Affine3f * Affine3f in float32 with the coefficient-wise order the runtime uses."""
    A = np.asarray(A, dtype=np.float32).reshape(3, 4)
    B = np.asarray(B, dtype=np.float32).reshape(3, 4)
    out = np.zeros((3, 4), dtype=np.float32)
    for i in range(3):
        for j in range(3):
            out[i, j] = np.float32(np.float32(A[i, 0] * B[0, j]) + np.float32(A[i, 1] * B[1, j])) + np.float32(A[i, 2] * B[2, j])
        out[i, 3] = np.float32(np.float32(np.float32(A[i, 0] * B[0, 3]) + np.float32(A[i, 1] * B[1, 3])) + np.float32(A[i, 2] * B[2, 3])) + A[i, 3]
    return out


class OracleSystem:
    """The reference's System-level calls the ROS nodes make, on top of one OracleMap (test infrastructure): lets the
    ros_shim drivers run against the oracle with the method names they use on the product.  Obstacle ids follow
    System.cpp:288-335 (INT64_MAX downward, freed ids reused last-in first-out); an obstacle cloud is pruned with the
    obstacle extrinsic and ranges (System::prune, System.cpp:93-114) and registered as a base-frame keyframe.
    The reference's worker is asynchronous; here every pose update is processed at once, which is the order a node that
    waits for `onMapUpdated` observes."""

    def __init__(self, om):
        self.om = om
        self.obst = None
        self.next_obst = 2 ** 63 - 1
        self.free = []
        self.known = set()

    def addNewKeyFrameWithPCL(self, ts_ns, kf_id, map_id, cloud):
        self.om.add_keyframe(kf_id, cloud)
        self.known.add(kf_id)

    def updateKeyFrame(self, kf_id, T):
        self.om.update_pose(kf_id, T)
        self.om.process()

    def deleteKeyFrame(self, kf_id):
        self.om.delete_keyframe(kf_id)
        self.known.discard(kf_id)

    def setObstacleParameters(self, T_base_obs, max_range, min_range):
        self.obst = (np.asarray(T_base_obs, dtype=np.float32).reshape(-1)[:12].copy(), float(max_range), float(min_range))

    def addObstacleKeyFrame(self, ts_ns, cloud):
        if self.obst is None or cloud is None:
            return 0
        base = self.om.prune(cloud, self.obst[0], self.obst[1], self.obst[2])
        if base.shape[0] == 0:
            return 0
        if self.free:
            oid = self.free.pop()
        else:
            oid = self.next_obst
            self.next_obst -= 1
        self.om.add_keyframe_base(oid, base)
        self.known.add(oid)
        return oid

    def deleteObstacleKeyFrame(self, kf_id):
        if kf_id not in self.known:
            return
        self.deleteKeyFrame(kf_id)
        self.free.append(kf_id)

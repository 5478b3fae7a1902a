"""traversability_mapping_b200 -- Python host mirror of the reference's System / LocalMap API for
the B200-native hot path, bound to libtmap_b200.so through its C ABI (include/tmap.h) with ctypes.

Reference interface mirrored (traversability_mapping/include/traversability_mapping/):
  System.hpp:53-212   -> class System  (addNewLocalMap, addNewKeyFrameWithPCL, updateKeyFrame,
                         deleteKeyFrame, updateKFMap, informLoopClosure, obstacle keyframes, getLocalMap)
  LocalMap.hpp:55-231 -> class LocalMap (takeChangedCells, occupiedCellKeys, getLattice, layer access)

There is no CPU fallback: importing works anywhere (the library loads without a GPU), but
System() raises unless a CUDA device is present and the extension has been built.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtmap_b200.so")

LAYERS = ["N", "sx", "sy", "sz", "sx2", "sy2", "sz2", "sxy", "sxz", "syz",
          "hazard", "elevation", "slope_haz", "step_haz", "roughness_haz", "border_haz",
          "normal_x", "normal_y", "normal_z", "step_haz_inflated", "zmin", "zmax"]
LAYER_ID = {n: i for i, n in enumerate(LAYERS)}

OK = 0
IGNORED_DUPLICATE, IGNORED_UNKNOWN_MAP, IGNORED_EMPTY, IGNORED_UNKNOWN_KF, IGNORED_NULL, IGNORED_NO_OBSTACLE_PARAMS = 1, 2, 3, 4, 5, 6
ERR_NEGATIVE_ID, ERR_CUDA, ERR_INVALID, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4, -5


class Params(C.Structure):
    """tmap_params: the reference's 22 parameter keys + runtime knobs (include/tmap.h)."""
    _fields_ = [
        ("resolution", C.c_double), ("grid_center_x", C.c_double), ("grid_center_y", C.c_double),
        ("half_size", C.c_double), ("extend_length", C.c_double), ("robot_radius", C.c_double),
        ("ground_clearance", C.c_double), ("max_slope", C.c_double),
        ("min_points_per_grid", C.c_int32), ("is_kf_optimization_enabled", C.c_int32),
        ("num_local_keyframes", C.c_int32), ("global_adjustment_sleep", C.c_int32),
        ("inflation_enabled", C.c_int32), ("use_pointcloud_buffer", C.c_int32),
        ("use_ros_buffer", C.c_int32), ("reserved0", C.c_int32),
        ("inflation_radius", C.c_double), ("cost_scaling_factor", C.c_double),
        ("lethal_step_threshold", C.c_double), ("robot_height", C.c_double),
        ("max_range", C.c_double), ("min_range", C.c_double), ("publish_rate_hz", C.c_double),
        ("device", C.c_int32), ("max_batch", C.c_int32), ("profile", C.c_int32), ("reserved1", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("keyframes_processed", C.c_uint64), ("points_binned", C.c_uint64), ("cells_classified", C.c_uint64),
        ("batches", C.c_uint64), ("kernel_launches", C.c_uint64), ("grid_reallocs", C.c_uint64),
        ("ms_hist", C.c_float), ("ms_scan", C.c_float), ("ms_scatter", C.c_float), ("ms_fold", C.c_float),
        ("ms_classify", C.c_float), ("ms_inflate", C.c_float), ("ms_total", C.c_float), ("host_ms_batch", C.c_float),
        ("last_batch_points", C.c_uint64), ("last_batch_tiles", C.c_uint64),
        ("last_batch_cells_touched", C.c_uint64), ("last_batch_cells_classified", C.c_uint64),
        ("last_batch_max_bin", C.c_uint64),
        ("sum_ms_hist", C.c_double), ("sum_ms_scan", C.c_double), ("sum_ms_scatter", C.c_double), ("sum_ms_fold", C.c_double),
        ("sum_ms_classify", C.c_double), ("sum_ms_inflate", C.c_double), ("sum_ms_total", C.c_double),
        ("sum_tiles", C.c_uint64), ("sum_cells_touched", C.c_uint64),
        ("sum_host_ms_plan", C.c_double), ("sum_host_ms_launch", C.c_double), ("sum_host_ms_wait", C.c_double),
        ("sum_host_ms_batch", C.c_double),
        ("sum_ms_gate", C.c_double), ("ms_gate", C.c_float), ("reserved0", C.c_float),
        ("sum_infl_tiles", C.c_uint64), ("sum_cand_bins", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
    ]


class ShardInfo(C.Structure):
    """tmap_shard_info (include/tmap.h): what one rank learns from a sharded rebuild."""
    _fields_ = [("bbox", C.c_int32 * 4), ("slab_y0", C.c_int32), ("slab_y1", C.c_int32), ("cj0", C.c_int32), ("cj1", C.c_int32),
                ("local_records", C.c_uint64), ("sent_records", C.c_uint64), ("recv_records", C.c_uint64),
                ("halo_bytes", C.c_uint64), ("grid_bytes", C.c_uint64), ("ms", C.c_float * 8), ("n_slabs", C.c_int32),
                ("slabs", C.c_int32 * 128)]


UPDATE_CB = C.CFUNCTYPE(None, C.c_void_p)

# every symbol include/tmap.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SIG = {
    "tmap_default_params": (None, [C.POINTER(Params)]),
    "tmap_last_error": (C.c_char_p, []),
    "tmap_create": (C.c_int, [C.POINTER(Params), C.POINTER(_P)]),
    "tmap_destroy": (None, [_P]),
    "tmap_set_extrinsics": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tmap_add_map": (C.c_int, [_P, C.c_uint64, UPDATE_CB, C.c_void_p]),
    "tmap_set_current_map": (C.c_int, [_P, C.c_uint64]),
    "tmap_current_map": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "tmap_global_point_cloud": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tmap_add_keyframe": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t]),
    "tmap_add_keyframe_device": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t]),
    "tmap_add_keyframe_pc2": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t,
                                        C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "tmap_push_to_buffer": (C.c_int, [_P, C.c_double, C.c_void_p, C.c_size_t, C.c_size_t]),
    "tmap_add_keyframe_ts": (C.c_int, [_P, C.c_double, C.c_uint64, C.c_uint64]),
    "tmap_buffered_clouds": (C.c_int, [_P, C.POINTER(C.c_size_t)]),
    "tmap_add_keyframe_base": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]),
    "tmap_update_pose": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_float)]),
    "tmap_update_poses": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.c_size_t]),
    "tmap_delete_keyframe": (C.c_int, [_P, C.c_uint64]),
    "tmap_update_kf_map": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "tmap_inform_loop_closure": (C.c_int, [_P]),
    "tmap_set_obstacle_params": (C.c_int, [_P, C.POINTER(C.c_float), C.c_double, C.c_double]),
    "tmap_add_obstacle_keyframe": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint64)]),
    "tmap_delete_obstacle_keyframe": (C.c_int, [_P, C.c_uint64]),
    "tmap_flush": (C.c_int, [_P]),
    "tmap_lock": (C.c_int, [_P, C.c_uint64]),
    "tmap_unlock": (C.c_int, [_P, C.c_uint64]),
    "tmap_take_changed_cells": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tmap_occupied_cell_keys": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tmap_read_layers": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tmap_fill_sparse_update": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_size_t)]),
    "tmap_grid_extent": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tmap_export_dense": (C.c_int, [_P, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tmap_export_occupancy": (C.c_int, [_P, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tmap_export_completeness": (C.c_int, [_P, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p,
                                           C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tmap_num_keyframes": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_size_t)]),
    "tmap_queue_depth": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_size_t)]),
    "tmap_keyframe_points": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_size_t)]),
    "tmap_stitched_cloud": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tmap_stitched_cloud_voxel": (C.c_int, [_P, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_size_t)]),
    "tmap_get_stats": (C.c_int, [_P, C.c_uint64, C.POINTER(Stats)]),
    "tmap_set_stream": (C.c_int, [_P, C.c_uint64, C.c_void_p]),
    "tmap_shard_unique_id": (C.c_int, [C.c_void_p]),
    "tmap_shard_comm_init": (C.c_int, [_P, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p]),
    "tmap_shard_comm_destroy": (C.c_int, [_P, C.c_uint64]),
    "tmap_shard_rebuild": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(ShardInfo)]),
    "tmap_shard_update": (C.c_int, [_P, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(ShardInfo)]),
    "tmap_shard_plan": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
}

_lib = None


def load_library():
    """dlopen libtmap_b200.so and bind every C-ABI symbol.  Raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m traversability_mapping_b200.build` "
            "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIG)


class TmapError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tmap error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc < 0:
        raise TmapError(rc, load_library().tmap_last_error().decode(errors="replace"))
    return rc


def default_params(**kw):
    p = Params()
    load_library().tmap_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def key(ci, cj):
    """Lattice::key (Moments.hpp:122-127)."""
    return ((int(ci) & 0xFFFFFFFF) << 32) | (int(cj) & 0xFFFFFFFF)


def unkey(k):
    """Lattice::unkey (Moments.hpp:130-133), vectorised."""
    k = np.asarray(k, dtype=np.uint64)
    ci = (k >> np.uint64(32)).astype(np.uint32).astype(np.int32)
    cj = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32).astype(np.int32)
    return ci, cj


def _pose12(T):
    a = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(-1)[:12])
    if a.size != 12:
        raise ValueError("pose must be 3x4 or 4x4")
    return a


class LocalMap:
    """Handle to one map of a System (reference: LocalMap, LocalMap.hpp:55-231)."""

    def __init__(self, system, map_id):
        self._s = system
        self.map_id = int(map_id)

    # getGridMapMutex(): `with m.locked(): keys = m.takeChangedCells(); vals = m.readLayers(...)`
    def locked(self):
        s, mid = self._s, self.map_id

        class _Ctx:
            def __enter__(self_):
                _check(s._lib.tmap_lock(s._h, mid))

            def __exit__(self_, *a):
                _check(s._lib.tmap_unlock(s._h, mid))

        return _Ctx()

    def getLattice(self):
        p = self._s.params
        return (p.grid_center_x, p.grid_center_y, p.resolution)

    def _keys(self, fn):
        n = C.c_size_t(0)
        with self.locked():
            _check(fn(self._s._h, self.map_id, None, 0, C.byref(n)))
            out = np.zeros(max(int(n.value), 1), dtype=np.uint64)
            _check(fn(self._s._h, self.map_id, out.ctypes.data, out.size, C.byref(n)))
        return out[: int(n.value)]

    def takeChangedCells(self):
        return self._keys(self._s._lib.tmap_take_changed_cells)

    def occupiedCellKeys(self):
        return self._keys(self._s._lib.tmap_occupied_cell_keys)

    def numChanged(self):
        n = C.c_size_t(0)
        _check(self._s._lib.tmap_take_changed_cells(self._s._h, self.map_id, None, 0, C.byref(n)))
        return int(n.value)

    def readLayers(self, keys, layers):
        """fillSparseUpdate (conversions.cpp:28-55): (len(keys), len(layers)) float32."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        ids = np.ascontiguousarray([LAYER_ID[l] if isinstance(l, str) else int(l) for l in layers], dtype=np.int32)
        out = np.empty((keys.size, ids.size), dtype=np.float32)
        _check(self._s._lib.tmap_read_layers(self._s._h, self.map_id, keys.ctypes.data, keys.size,
                                             ids.ctypes.data, ids.size, out.ctypes.data))
        return out

    def gridExtent(self):
        kx, ky = C.c_int32(0), C.c_int32(0)
        _check(self._s._lib.tmap_grid_extent(self._s._h, self.map_id, C.byref(kx), C.byref(ky)))
        return int(kx.value), int(ky.value)

    def exportDense(self, layer, ci0=None, cj0=None, w=None, h=None):
        """getGridMap()[layer] as (h, w) float32, [cj - cj0, ci - ci0]; default = the whole grid."""
        if ci0 is None:
            kx, ky = self.gridExtent()
            ci0, cj0, w, h = -kx, -ky, 2 * kx + 1, 2 * ky + 1
        lid = LAYER_ID[layer] if isinstance(layer, str) else int(layer)
        out = np.empty((h, w), dtype=np.float32)
        _check(self._s._lib.tmap_export_dense(self._s._h, self.map_id, lid, ci0, cj0, w, h, out.ctypes.data))
        return out

    def exportOccupancy(self, ci0=None, cj0=None, w=None, h=None, out=None):
        """toOccupancyGrid(grid, "hazard", 0, 1): (h, w) int8."""
        if ci0 is None:
            kx, ky = self.gridExtent()
            ci0, cj0, w, h = -kx, -ky, 2 * kx + 1, 2 * ky + 1
        if out is None:
            out = np.empty((h, w), dtype=np.int8)
        _check(self._s._lib.tmap_export_occupancy(self._s._h, self.map_id, ci0, cj0, w, h, out.ctypes.data))
        return out

    def fillSparseUpdate(self, layers, keys):
        """tmap_ros::fillSparseUpdate (conversions.cpp:28-55) -> (kept keys, values (n, len(layers)))."""
        ids = np.ascontiguousarray([LAYER_ID[l] if isinstance(l, str) else int(l) for l in layers], dtype=np.int32)
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        ok = np.empty(len(k), dtype=np.uint64)
        ov = np.empty((len(k), len(ids)), dtype=np.float32)
        n = C.c_size_t(0)
        _check(self._s._lib.tmap_fill_sparse_update(self._s._h, self.map_id, ids.ctypes.data, len(ids), k.ctypes.data, len(k),
                                                    ok.ctypes.data, ov.ctypes.data, C.byref(n)))
        return ok[: n.value].copy(), ov[: n.value].copy()

    def exportCompleteness(self, publish_res=0.25, ci0=None, cj0=None, w=None, h=None):
        """publishCompleteness (global_traversability_node.cpp:381-459): (rows, cols) int8 palette bytes."""
        if ci0 is None:
            kx, ky = self.gridExtent()
            ci0, cj0, w, h = -kx, -ky, 2 * kx + 1, 2 * ky + 1
        cw, ch = C.c_int32(0), C.c_int32(0)
        lib, hnd = self._s._lib, self._s._h
        _check(lib.tmap_export_completeness(hnd, self.map_id, ci0, cj0, w, h, publish_res, None, 0, C.byref(cw), C.byref(ch)))
        out = np.empty((ch.value, cw.value), dtype=np.int8)
        _check(lib.tmap_export_completeness(hnd, self.map_id, ci0, cj0, w, h, publish_res, out.ctypes.data, out.size,
                                            C.byref(cw), C.byref(ch)))
        return out

    def numKeyFrames(self):
        n = C.c_size_t(0)
        _check(self._s._lib.tmap_num_keyframes(self._s._h, self.map_id, C.byref(n)))
        return int(n.value)

    def queueDepth(self):
        n = C.c_size_t(0)
        _check(self._s._lib.tmap_queue_depth(self._s._h, self.map_id, C.byref(n)))
        return int(n.value)

    def getStitchedPointCloud(self, voxel_size_x=0.0, voxel_size_y=0.0, voxel_size_z=0.0):
        """LocalMap::getStitchedPointCloud (LocalMap.cpp:416-448); positive leaf sizes -> pcl::VoxelGrid centroids."""
        if voxel_size_x > 0 and voxel_size_y > 0 and voxel_size_z > 0:
            n = C.c_size_t(0)
            lib, h = self._s._lib, self._s._h
            _check(lib.tmap_stitched_cloud_voxel(h, self.map_id, voxel_size_x, voxel_size_y, voxel_size_z, None, 0, C.byref(n)))
            out = np.zeros((max(int(n.value), 1), 3), dtype=np.float32)
            _check(lib.tmap_stitched_cloud_voxel(h, self.map_id, voxel_size_x, voxel_size_y, voxel_size_z, out.ctypes.data,
                                                 int(n.value), C.byref(n)))
            return out[: int(n.value)]
        n = C.c_size_t(0)
        _check(self._s._lib.tmap_stitched_cloud(self._s._h, self.map_id, None, 0, C.byref(n)))
        out = np.zeros((max(int(n.value), 1), 3), dtype=np.float32)
        _check(self._s._lib.tmap_stitched_cloud(self._s._h, self.map_id, out.ctypes.data, int(n.value), C.byref(n)))
        return out[: int(n.value)]

    def stats(self):
        st = Stats()
        _check(self._s._lib.tmap_get_stats(self._s._h, self.map_id, C.byref(st)))
        return {f: getattr(st, f) for f, _ in Stats._fields_ if not f.startswith("reserved")}

    def setStream(self, cuda_stream_ptr):
        _check(self._s._lib.tmap_set_stream(self._s._h, self.map_id, C.c_void_p(cuda_stream_ptr or 0)))


class System:
    """Reference facade (System.hpp:53-212) over the C ABI."""

    def __init__(self, params=None, **kw):
        self._lib = load_library()
        self.params = params if params is not None else default_params(**kw)
        self._h = C.c_void_p(0)
        self._cbs = {}
        rc = self._lib.tmap_create(C.byref(self.params), C.byref(self._h))
        if rc != OK:
            self._h = C.c_void_p(0)
            raise TmapError(rc, self._lib.tmap_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.tmap_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def setExtrinsicParameters(self, Tsv, Tbs):
        a, b = _pose12(Tsv), _pose12(Tbs)
        return _check(self._lib.tmap_set_extrinsics(self._h, a.ctypes.data_as(C.POINTER(C.c_float)),
                                                    b.ctypes.data_as(C.POINTER(C.c_float))))

    def addNewLocalMap(self, map_id, onUpdate=None):
        cb = UPDATE_CB((lambda _u, f=onUpdate: f()) if onUpdate else 0)
        rc = _check(self._lib.tmap_add_map(self._h, int(map_id), cb, None))
        if rc == OK:
            self._cbs[int(map_id)] = cb  # keep the trampoline alive; a duplicate id keeps the FIRST one (System.cpp:65-69)
        return rc

    def setCurrentMap(self, map_id):
        return _check(self._lib.tmap_set_current_map(self._h, int(map_id)))

    def addNewKeyFrameWithPCL(self, timestamp_ns, kf_id, map_id, cloud):
        """cloud: (n, 3) or (n, 4) float32 sensor-frame points (pcl::PointXYZ is the 4-float form),
        or None (null cloud -> ignored).  Raises TmapError(ERR_NEGATIVE_ID) where the reference throws."""
        if cloud is None:
            return _check(self._lib.tmap_add_keyframe(self._h, int(timestamp_ns), int(kf_id), int(map_id), None, 0, 16))
        a = np.ascontiguousarray(cloud, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] not in (3, 4):
            raise ValueError("cloud must be (n,3) or (n,4) float32")
        return _check(self._lib.tmap_add_keyframe(self._h, int(timestamp_ns), int(kf_id) & 0xFFFFFFFFFFFFFFFF, int(map_id),
                                                  a.ctypes.data, a.shape[0], a.shape[1] * 4))

    def addNewKeyFramePointCloud2(self, timestamp_ns, kf_id, map_id, data, point_step, x_offset, y_offset, z_offset,
                                  is_bigendian=False):
        """sensor_msgs/PointCloud2 ingest (conversions.cpp:19-26): `data` = the message's raw byte array
        (width*height points of `point_step` bytes), FLOAT32 x / y / z at the given byte offsets."""
        a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1))
        n = a.size // int(point_step)
        return _check(self._lib.tmap_add_keyframe_pc2(self._h, int(timestamp_ns), int(kf_id) & 0xFFFFFFFFFFFFFFFF, int(map_id),
                                                      a.ctypes.data, n, int(point_step), int(x_offset), int(y_offset),
                                                      int(z_offset), 1 if is_bigendian else 0))

    def pushToBuffer(self, timestamp_s, cloud):
        """System::pushToBuffer (System.cpp:339-344): buffer a sensor-frame cloud under its acquisition time."""
        a = np.ascontiguousarray(cloud, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] not in (3, 4):
            raise ValueError("cloud must be (n,3) or (n,4) float32")
        return _check(self._lib.tmap_push_to_buffer(self._h, float(timestamp_s), a.ctypes.data, a.shape[0], a.shape[1] * 4))

    def addNewKeyFrameTsDouble(self, timestamp_s, kf_id, map_id):
        """System::addNewKeyFrameTsDouble (System.cpp:174-186): keyframe from the buffered cloud closest in time."""
        return _check(self._lib.tmap_add_keyframe_ts(self._h, float(timestamp_s), int(kf_id) & 0xFFFFFFFFFFFFFFFF, int(map_id)))

    def addNewKeyFrameTsULong(self, timestamp_ns, kf_id, map_id):
        ns = int(timestamp_ns)
        return self.addNewKeyFrameTsDouble((ns // 1000000000) + (ns % 1000000000) * 1e-9, kf_id, map_id)

    def bufferedClouds(self):
        n = C.c_size_t(0)
        _check(self._lib.tmap_buffered_clouds(self._h, C.byref(n)))
        return int(n.value)

    def addNewKeyFrameDevice(self, timestamp_ns, kf_id, map_id, dev_ptr, n, stride_bytes=16):
        return _check(self._lib.tmap_add_keyframe_device(self._h, int(timestamp_ns), int(kf_id), int(map_id),
                                                         C.c_void_p(dev_ptr), int(n), int(stride_bytes)))

    def addKeyFrameBase(self, kf_id, map_id, cloud_base):
        a = np.ascontiguousarray(cloud_base, dtype=np.float32).reshape(-1, 3)
        return _check(self._lib.tmap_add_keyframe_base(self._h, int(kf_id), int(map_id), a.ctypes.data, a.shape[0]))

    def updateKeyFrame(self, kf_id, pose_map_base, numConnections=0):
        a = _pose12(pose_map_base)
        return _check(self._lib.tmap_update_pose(self._h, int(kf_id), a.ctypes.data_as(C.POINTER(C.c_float))))

    def updateKeyFrames(self, kf_ids, poses_map_base):
        """One PGO message (global_traversability_node.cpp:184-188): updateKeyFrame for every listed keyframe."""
        ids = np.ascontiguousarray(kf_ids, dtype=np.uint64)
        poses = np.ascontiguousarray(np.asarray(poses_map_base, dtype=np.float32).reshape(len(ids), -1)[:, :12])
        return _check(self._lib.tmap_update_poses(self._h, ids.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                  poses.ctypes.data_as(C.POINTER(C.c_float)), len(ids)))

    def deleteKeyFrame(self, kf_id):
        return _check(self._lib.tmap_delete_keyframe(self._h, int(kf_id)))

    def keyFramePoints(self, kf_id):
        """Points of the keyframe's retained, pruned base-frame cloud (KeyFrame::cloudBase().size())."""
        n = C.c_size_t(0)
        rc = _check(self._lib.tmap_keyframe_points(self._h, int(kf_id), C.byref(n)))
        return None if rc == IGNORED_UNKNOWN_KF else int(n.value)

    def updateKFMap(self, kf_id, map_id):
        return _check(self._lib.tmap_update_kf_map(self._h, int(kf_id), int(map_id)))

    def informLoopClosure(self):
        return _check(self._lib.tmap_inform_loop_closure(self._h))

    def setObstacleParameters(self, Tb_obs, maxRange, minRange):
        a = _pose12(Tb_obs)
        return _check(self._lib.tmap_set_obstacle_params(self._h, a.ctypes.data_as(C.POINTER(C.c_float)),
                                                         float(maxRange), float(minRange)))

    def addObstacleKeyFrame(self, timestamp_ns, cloud):
        a = np.ascontiguousarray(cloud, dtype=np.float32)
        oid = C.c_uint64(0)
        _check(self._lib.tmap_add_obstacle_keyframe(self._h, int(timestamp_ns), a.ctypes.data, a.shape[0],
                                                    a.shape[1] * 4, C.byref(oid)))
        return int(oid.value)

    def deleteObstacleKeyFrame(self, kf_id):
        return _check(self._lib.tmap_delete_obstacle_keyframe(self._h, int(kf_id)))

    def flush(self):
        return _check(self._lib.tmap_flush(self._h))

    def getLocalMap(self, map_id=None):
        """getLocalMap() = the active map (System.cpp:357-361; None before any map exists), getLocalMap(id) = that map."""
        if map_id is None:
            cur = C.c_uint64(0)
            if _check(self._lib.tmap_current_map(self._h, C.byref(cur))) != OK:
                return None
            map_id = cur.value
        return LocalMap(self, map_id)

    def getGlobalPointCloud(self, voxel_size_x=0.0, voxel_size_y=0.0, voxel_size_z=0.0):
        """System::getGlobalPointCloud (System.cpp:370-387): (n, 3) float32."""
        n = C.c_size_t(0)
        _check(self._lib.tmap_global_point_cloud(self._h, voxel_size_x, voxel_size_y, voxel_size_z, None, 0, C.byref(n)))
        out = np.zeros((max(int(n.value), 1), 3), dtype=np.float32)
        _check(self._lib.tmap_global_point_cloud(self._h, voxel_size_x, voxel_size_y, voxel_size_z, out.ctypes.data,
                                                 int(n.value), C.byref(n)))
        return out[: int(n.value)]

"""pyoracle -- ctypes binding of the CPU ORACLE (oracle/libtmap_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtmap_oracle.so")

LAYERS = ["N", "sx", "sy", "sz", "sx2", "sy2", "sz2", "sxy", "sxz", "syz",
          "hazard", "elevation", "slope_haz", "step_haz", "roughness_haz", "border_haz",
          "normal_x", "normal_y", "normal_z", "step_haz_inflated", "zmin", "zmax"]
LAYER_ID = {n: i for i, n in enumerate(LAYERS)}


class Params(C.Structure):
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
    ]


def build(force=False):
    """make -C oracle (g++ only)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("tmap_oracle.cpp", "tmap_oracle.h", "Makefile")):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        L.tmo_create.restype = P
        L.tmo_create.argtypes = [C.POINTER(Params)]
        L.tmo_destroy.argtypes = [P]
        L.tmo_default_params.argtypes = [C.POINTER(Params)]
        L.tmo_set_extrinsics.argtypes = [P, P]
        L.tmo_set_threads.argtypes = [P, C.c_int32]
        L.tmo_prune.restype = C.c_size_t
        L.tmo_prune.argtypes = [P, P, C.c_size_t, C.c_size_t, P, C.c_double, C.c_double, P]
        L.tmo_add_keyframe.argtypes = [P, C.c_uint64, P, C.c_size_t, C.c_size_t]
        L.tmo_add_keyframe_base.argtypes = [P, C.c_uint64, P, C.c_size_t]
        L.tmo_update_pose.argtypes = [P, C.c_uint64, P]
        L.tmo_process.argtypes = [P]
        L.tmo_process_batched.argtypes = [P]
        L.tmo_delete_keyframe.argtypes = [P, C.c_uint64]
        L.tmo_clear_entire_map.argtypes = [P]
        L.tmo_num_keyframes.restype = C.c_size_t
        L.tmo_num_keyframes.argtypes = [P]
        L.tmo_queue_depth.restype = C.c_size_t
        L.tmo_queue_depth.argtypes = [P]
        L.tmo_take_changed_cells.restype = C.c_size_t
        L.tmo_take_changed_cells.argtypes = [P, P, C.c_size_t]
        L.tmo_occupied_cell_keys.restype = C.c_size_t
        L.tmo_occupied_cell_keys.argtypes = [P, P, C.c_size_t]
        L.tmo_read_layers.argtypes = [P, P, C.c_size_t, P, C.c_size_t, P]
        L.tmo_grid_extent.argtypes = [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.tmo_export_dense.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P]
        L.tmo_export_occupancy.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P]
        L.tmo_export_completeness.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, P,
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.tmo_cell_of.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.tmo_key.restype = C.c_uint64
        L.tmo_key.argtypes = [C.c_int32, C.c_int32]
        L.tmo_disc_offsets.restype = C.c_size_t
        L.tmo_disc_offsets.argtypes = [C.c_double, C.c_double, P, C.c_size_t]
        L.tmo_rebin.restype = C.c_size_t
        L.tmo_rebin.argtypes = [C.c_double, C.c_double, C.c_double, P, C.c_size_t, P, P, P]
        L.tmo_compute_goodness.argtypes = [P, P, C.c_size_t, C.c_int32, C.c_double, C.c_double, C.c_uint32, P]
        L.tmo_eigh3.argtypes = [P, P, P]
        _lib = L
    return _lib


def default_params(**kw):
    p = Params()
    lib().tmo_default_params(C.byref(p))
    for k, v in kw.items():
        if hasattr(p, k):
            setattr(p, k, v)
    return p


def params_from(product_params):
    """Copy the 22 reference keys out of a product Params struct."""
    p = Params()
    for f, _ in Params._fields_:
        setattr(p, f, getattr(product_params, f))
    return p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleMap:
    """One LocalMap + System-level prune/registry, single-threaded (see tmap_oracle.h)."""

    def __init__(self, params):
        self.params = params
        self._h = lib().tmo_create(C.byref(params))

    def close(self):
        if self._h:
            lib().tmo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_extrinsics(self, T_bv):
        a = _f32(np.asarray(T_bv).reshape(-1)[:12])
        lib().tmo_set_extrinsics(self._h, a.ctypes.data)

    def set_threads(self, n):
        """CPU-baseline flavour 2: OpenMP over the dirty cells (archive/local_traversability_monolithic.cpp:349-362)."""
        lib().tmo_set_threads(self._h, int(n))

    def prune(self, cloud, T, max_range, min_range):
        a = _f32(cloud)
        t = _f32(np.asarray(T).reshape(-1)[:12])
        out = np.empty((a.shape[0] + 1, 3), dtype=np.float32)
        k = lib().tmo_prune(self._h, a.ctypes.data, a.shape[0], a.shape[1], t.ctypes.data, max_range, min_range, out.ctypes.data)
        return out[:k].copy()

    def add_keyframe(self, kf_id, cloud):
        a = _f32(cloud)
        return lib().tmo_add_keyframe(self._h, int(kf_id) & 0xFFFFFFFFFFFFFFFF, a.ctypes.data, a.shape[0], a.shape[1])

    def add_keyframe_base(self, kf_id, cloud):
        a = _f32(cloud).reshape(-1, 3)
        return lib().tmo_add_keyframe_base(self._h, int(kf_id), a.ctypes.data, a.shape[0])

    def update_pose(self, kf_id, T):
        t = _f32(np.asarray(T).reshape(-1)[:12])
        return lib().tmo_update_pose(self._h, int(kf_id), t.ctypes.data)

    def process(self):
        return lib().tmo_process(self._h)

    def process_batched(self):
        return lib().tmo_process_batched(self._h)

    def delete_keyframe(self, kf_id):
        return lib().tmo_delete_keyframe(self._h, int(kf_id))

    def clear_entire_map(self):
        lib().tmo_clear_entire_map(self._h)

    def num_keyframes(self):
        return lib().tmo_num_keyframes(self._h)

    def queue_depth(self):
        return lib().tmo_queue_depth(self._h)

    def take_changed_cells(self):
        n = lib().tmo_take_changed_cells(self._h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().tmo_take_changed_cells(self._h, out.ctypes.data, n)
        return out[:n]

    def occupied_cell_keys(self):
        n = lib().tmo_occupied_cell_keys(self._h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        lib().tmo_occupied_cell_keys(self._h, out.ctypes.data, n)
        return out[:n]

    def read_layers(self, keys, layers):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        ids = np.ascontiguousarray([LAYER_ID[l] if isinstance(l, str) else int(l) for l in layers], dtype=np.int32)
        out = np.empty((keys.size, ids.size), dtype=np.float32)
        lib().tmo_read_layers(self._h, keys.ctypes.data, keys.size, ids.ctypes.data, ids.size, out.ctypes.data)
        return out

    def grid_extent(self):
        kx, ky = C.c_int32(0), C.c_int32(0)
        lib().tmo_grid_extent(self._h, C.byref(kx), C.byref(ky))
        return int(kx.value), int(ky.value)

    def export_dense(self, layer, ci0=None, cj0=None, w=None, h=None):
        if ci0 is None:
            kx, ky = self.grid_extent()
            ci0, cj0, w, h = -kx, -ky, 2 * kx + 1, 2 * ky + 1
        lid = LAYER_ID[layer] if isinstance(layer, str) else int(layer)
        out = np.empty((h, w), dtype=np.float32)
        lib().tmo_export_dense(self._h, lid, ci0, cj0, w, h, out.ctypes.data)
        return out

    def export_occupancy(self, ci0=None, cj0=None, w=None, h=None):
        if ci0 is None:
            kx, ky = self.grid_extent()
            ci0, cj0, w, h = -kx, -ky, 2 * kx + 1, 2 * ky + 1
        out = np.empty((h, w), dtype=np.int8)
        lib().tmo_export_occupancy(self._h, ci0, cj0, w, h, out.ctypes.data)
        return out

    def stitched_cloud(self, vx=0.0, vy=0.0, vz=0.0):
        L = lib()
        L.tmo_stitched_cloud.restype = C.c_size_t
        L.tmo_stitched_cloud.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_size_t]
        n = L.tmo_stitched_cloud(self._h, 0.0, 0.0, 0.0, None, 0)
        out = np.zeros((max(n, 1), 3), dtype=np.float32)
        k = L.tmo_stitched_cloud(self._h, vx, vy, vz, out.ctypes.data, n)
        return out[:k].copy()

    def fill_sparse_update(self, layers, keys):
        """fillSparseUpdate (conversions.cpp:28-55) -> (kept keys, values (n, len(layers)))."""
        import traversability_mapping_b200 as tm
        ids = np.ascontiguousarray([tm.LAYER_ID[l] if isinstance(l, str) else int(l) for l in layers], dtype=np.int32)
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        ok = np.empty(len(k), dtype=np.uint64)
        ov = np.empty((len(k), len(ids)), dtype=np.float32)
        L = lib()
        L.tmo_fill_sparse_update.restype = C.c_size_t
        L.tmo_fill_sparse_update.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        n = L.tmo_fill_sparse_update(self._h, ids.ctypes.data, len(ids), k.ctypes.data, len(k), ok.ctypes.data, ov.ctypes.data)
        return ok[:n].copy(), ov[:n].copy()

    def export_completeness(self, publish_res=0.25, ci0=None, cj0=None, w=None, h=None):
        if ci0 is None:
            kx, ky = self.grid_extent()
            ci0, cj0, w, h = -kx, -ky, 2 * kx + 1, 2 * ky + 1
        out = np.empty(w * h, dtype=np.int8)
        cw, ch = C.c_int32(0), C.c_int32(0)
        lib().tmo_export_completeness(self._h, ci0, cj0, w, h, publish_res, out.ctypes.data, C.byref(cw), C.byref(ch))
        return out[:cw.value * ch.value].reshape(ch.value, cw.value).copy()


def voxel_grid(xyz, lx, ly, lz):
    """pcl::VoxelGrid centroids of a packed (n, 3) float32 cloud, ascending voxel index."""
    a = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.empty_like(a)
    L = lib()
    L.tmo_voxel_grid.restype = C.c_size_t
    L.tmo_voxel_grid.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_size_t]
    nv = L.tmo_voxel_grid(a.ctypes.data, len(a), lx, ly, lz, out.ctypes.data, len(a))
    return out[:nv].copy()


def cell_of(x0, y0, res, x, y):
    ci, cj = C.c_int32(0), C.c_int32(0)
    lib().tmo_cell_of(x0, y0, res, x, y, C.byref(ci), C.byref(cj))
    return int(ci.value), int(cj.value)


def disc_offsets(radius, res):
    out = np.zeros((4096, 2), dtype=np.int32)
    n = lib().tmo_disc_offsets(radius, res, out.ctypes.data, 4096)
    return out[:n].copy()


def rebin(x0, y0, res, cloud_base, T):
    a = _f32(cloud_base).reshape(-1, 3)
    t = _f32(np.asarray(T).reshape(-1)[:12])
    keys = np.zeros(max(a.shape[0], 1), dtype=np.uint64)
    mom = np.zeros((max(a.shape[0], 1), 10), dtype=np.float64)
    n = lib().tmo_rebin(x0, y0, res, a.ctypes.data, a.shape[0], t.ctypes.data, keys.ctypes.data, mom.ctypes.data)
    return keys[:n].copy(), mom[:n].copy()


def eigh3(A):
    a = np.ascontiguousarray(A, dtype=np.float64).reshape(9)
    w = np.zeros(3)
    v = np.zeros(9)
    lib().tmo_eigh3(a.ctypes.data, w.ctypes.data, v.ctypes.data)
    return w, v.reshape(3, 3).T  # columns = eigenvectors

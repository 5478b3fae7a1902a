"""Parity tests proper: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.
Bar: occupied keys, N and the moment layers bit-identical; derived layers' NaN pattern
identical and values within 1e-4 relative; occupancy labels equal except at a threshold."""
import os

import numpy as np
import pytest

import helpers
from helpers import compare_maps, make_pair

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I34 = np.eye(3, 4, dtype=np.float32)


def trans(x, y, z):
    T = I34.copy()
    T[:, 3] = (x, y, z)
    return T


LOCALMAP = dict(resolution=0.25, half_size=7.5, extend_length=30.0, robot_radius=0.4, min_points_per_grid=5,
                max_range=300.0, min_range=1.0, inflation_radius=0.5, cost_scaling_factor=3.0)


def K(tmap, ci, cj):
    return tmap.key(ci, cj)


# ---- the reference's LocalMap KATs, through the C ABI (test_localmap.cpp) ----------------------
def test_registration_does_not_touch_grid(tmap, oracle):  # test_localmap.cpp:175-183
    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    assert s.addKeyFrameBase(7, 0, [[0, 0, 0]]) == tmap.OK
    s.flush()
    assert m.numKeyFrames() == 1 and m.occupiedCellKeys().size == 0 and m.takeChangedCells().size == 0
    s.close()


def test_set_pose_populates_cells_and_changed_drains(tmap, oracle):  # :197-220
    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    pts = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 2, 0]]
    s.addKeyFrameBase(1, 0, pts); om.add_keyframe_base(1, pts)
    s.updateKeyFrame(1, I34); om.update_pose(1, I34)
    s.flush(); om.process()
    want = sorted([K(tmap, 0, 0), K(tmap, 4, 0), K(tmap, 0, 4), K(tmap, 8, 8)])
    assert m.occupiedCellKeys().tolist() == want
    compare_maps(m, om, check_changed=True)
    assert m.takeChangedCells().size == 0
    s.close()


def test_grid_grows_pose_update_moves_delete_empties(tmap, oracle):  # :222-280
    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    assert m.gridExtent() == (30, 30)
    s.addKeyFrameBase(1, 0, [[20, 0, 0]]); om.add_keyframe_base(1, [[20, 0, 0]])
    s.updateKeyFrame(1, I34); om.update_pose(1, I34)
    s.flush(); om.process()
    assert m.gridExtent() == om.grid_extent() and m.gridExtent()[0] > 30
    assert K(tmap, 80, 0) in m.occupiedCellKeys().tolist()
    s.updateKeyFrame(1, trans(10, 0, 0)); om.update_pose(1, trans(10, 0, 0))
    s.flush(); om.process()
    occ = m.occupiedCellKeys().tolist()
    assert K(tmap, 120, 0) in occ and K(tmap, 80, 0) not in occ  # vacated cell blanked
    compare_maps(m, om, check_changed=True)
    assert s.deleteKeyFrame(1) == tmap.OK and om.delete_keyframe(1) == 1
    assert m.numKeyFrames() == 0 and m.occupiedCellKeys().size == 0
    assert s.deleteKeyFrame(999) == tmap.IGNORED_UNKNOWN_KF
    compare_maps(m, om, check_changed=True)
    s.close()


def test_clear_entire_map_without_retained_clouds(tmap, oracle):  # :284-301
    s, m, om = make_pair(tmap, oracle, **dict(LOCALMAP, is_kf_optimization_enabled=0))
    pts = [[0, 0, 0], [2, 2, 0]]
    s.addKeyFrameBase(1, 0, pts); om.add_keyframe_base(1, pts)
    s.updateKeyFrame(1, I34); om.update_pose(1, I34)
    s.flush(); om.process()
    was = set(m.occupiedCellKeys().tolist())
    assert was
    m.takeChangedCells(); om.take_changed_cells()
    s.informLoopClosure(); om.clear_entire_map()
    s.flush(); om.process()
    assert m.occupiedCellKeys().size == 0
    ch = set(m.takeChangedCells().tolist())
    assert was <= ch and ch == set(om.take_changed_cells().tolist())
    s.close()


def test_system_ignore_and_error_codes(tmap, oracle):  # System.cpp:121-172,195-248
    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    far = np.array([[5, 0, 0, 0]], dtype=np.float32)
    assert s.addNewKeyFrameWithPCL(0, 1, 0, far) == tmap.OK
    assert s.addNewKeyFrameWithPCL(0, 1, 0, far) == tmap.IGNORED_DUPLICATE
    assert s.addNewKeyFrameWithPCL(0, 2, 9, far) == tmap.IGNORED_UNKNOWN_MAP
    assert s.addNewKeyFrameWithPCL(0, 3, 0, np.array([[0.1, 0, 0, 0]], dtype=np.float32)) == tmap.IGNORED_EMPTY  # < min range
    assert s.addNewKeyFrameWithPCL(0, 4, 0, None) == tmap.IGNORED_NULL
    with pytest.raises(tmap.TmapError) as e:
        s.addNewKeyFrameWithPCL(0, -5, 0, far)  # negative id: the reference throws
    assert e.value.code == tmap.ERR_NEGATIVE_ID
    assert s.updateKeyFrame(77, I34) == tmap.IGNORED_UNKNOWN_KF
    assert s.addNewLocalMap(0) == tmap.IGNORED_DUPLICATE
    with pytest.raises(tmap.TmapError):
        s.updateKeyFrame(1, np.full((3, 4), np.nan, dtype=np.float32))
    s.close()


def test_prune_matches_oracle(tmap, oracle):  # System.cpp:93-114
    from traversability_mapping_b200 import synth
    P = synth.trajectory("single", 1, seed=2)[0]
    cloud = synth.make_keyframe(2, 0, P, max_range=8.0, rings=32, azimuths=512)
    cloud[5] = (0.0, 0.0, 1.0, 0)       # ego return
    cloud[6] = (np.inf, 1.0, 1.0, 0)    # non-finite
    cloud[7] = (1.0, 1.0, 9.0, 0)       # above robot_height
    s, m, om = make_pair(tmap, oracle, resolution=0.1, max_range=6.0, min_range=1.0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    assert s.addNewKeyFrameWithPCL(1, 1, 0, cloud) == tmap.OK
    want = om.prune(cloud, synth.T_BASE_SENSOR, 6.0, 1.0)
    got = m.getStitchedPointCloud()  # identity pose until updated -> base-frame cloud, in order
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # packed xyz input (stride 12) gives the same keyframe
    s.addNewKeyFrameWithPCL(1, 2, 0, np.ascontiguousarray(cloud[:, :3]))
    s.deleteKeyFrame(1)
    got2 = m.getStitchedPointCloud()
    assert np.array_equal(got2.view(np.uint32), want.view(np.uint32))
    s.close()


# ---- golden fixture + live oracle on the seeded small scene ------------------------------------
def run_small_scene(tmap, oracle, **extra):
    from golden.make_golden import PARAMS, scene_inputs
    from traversability_mapping_b200 import synth
    poses, clouds = scene_inputs()
    s, m, om = make_pair(tmap, oracle, **dict(PARAMS, **extra))
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    batched = extra.get("max_batch", 1) != 1
    proc = om.process_batched if batched else om.process
    for i, c in enumerate(clouds):
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, c) == tmap.OK
        om.add_keyframe(i + 1, c)
        s.updateKeyFrame(i + 1, poses[i]); om.update_pose(i + 1, poses[i])
        if not batched:
            s.flush(); proc()   # keep the two FIFO orders identical
    s.flush(); proc()
    st = compare_maps(m, om, check_changed=True, label="after build")
    moved = poses[2].copy(); moved[0, 3] += 1.3
    s.updateKeyFrame(3, moved); om.update_pose(3, moved)
    s.flush(); proc()
    compare_maps(m, om, check_changed=True, label="after pose update")
    s.deleteKeyFrame(1); om.delete_keyframe(1)
    st = compare_maps(m, om, check_changed=True, label="after delete")
    return s, m, om, st


def test_small_scene_against_golden_and_oracle(tmap, oracle):
    s, m, om, st = run_small_scene(tmap, oracle)
    g = np.load(os.path.join(ROOT, "tests", "golden", "small_scene.npz"))
    keys = m.occupiedCellKeys()
    assert np.array_equal(keys, g["keys"])
    L = m.readLayers(keys, helpers.MOMENT_LAYERS + helpers.DERIVED_LAYERS)
    nm = len(helpers.MOMENT_LAYERS)
    assert helpers.bits_equal(L[:, :nm], g["layers"][:, :nm])
    assert np.array_equal(np.isnan(L[:, nm:]), np.isnan(g["layers"][:, nm:]))
    f = ~np.isnan(g["layers"][:, nm:])
    assert np.allclose(L[:, nm:][f], g["layers"][:, nm:][f], rtol=helpers.RTOL, atol=helpers.ATOL)
    assert st["gated_in"] > 100
    s.close()


def test_small_scene_batched_mode(tmap, oracle):
    s, m, om, st = run_small_scene(tmap, oracle, max_batch=0)
    s.close()


def test_small_scene_with_inflation(tmap, oracle):
    s, m, om, st = run_small_scene(tmap, oracle, inflation_enabled=1, inflation_radius=1.0, cost_scaling_factor=3.0,
                                   lethal_step_threshold=0.5)
    infl = m.exportDense("step_haz_inflated")
    assert np.nanmax(infl) > 0.0, "scene should contain at least one lethal step seed"
    s.close()


# ---- config 1: one 131 072-ray keyframe into a 20 x 20 m grid at 0.1 m -------------------------
@pytest.mark.parametrize("radius,label", [(0.20, "13-cell disc"), (0.15, "3x3 window")])
def test_config1_single_keyframe(tmap, oracle, radius, label):
    from traversability_mapping_b200 import synth
    P = synth.trajectory("single", 1, seed=1)[0]
    cloud = synth.make_keyframe(1, 0, P, max_range=10.0)
    s, m, om = make_pair(tmap, oracle, resolution=0.1, half_size=10.0, max_range=10.0, min_range=1.0, robot_radius=radius)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    assert s.addNewKeyFrameWithPCL(0, 1, 0, cloud) == tmap.OK
    om.add_keyframe(1, cloud)
    s.updateKeyFrame(1, P); om.update_pose(1, P)
    s.flush(); om.process()
    st = compare_maps(m, om, check_changed=True, label=label)
    assert st["cells"] > 5000 and st["gated_in"] > 1000
    # size-independent properties: sum of N == points kept; subtracting the keyframe empties the map
    N = m.readLayers(m.occupiedCellKeys(), ["N"])[:, 0]
    assert int(N.sum()) == m.getStitchedPointCloud().shape[0]
    s.deleteKeyFrame(1)
    assert m.occupiedCellKeys().size == 0
    s.close()


# ---- config 2 shape: rolling window at 0.05 m, 49-cell disc, inflation on -----------------------
def test_config2_rolling_window(tmap, oracle):
    from traversability_mapping_b200 import synth
    n, window = 6, 3
    poses = synth.trajectory("drive", n, seed=4, step=0.5)
    s, m, om = make_pair(tmap, oracle, resolution=0.05, half_size=20.0, max_range=8.0, min_range=1.0, min_points_per_grid=2,
                         inflation_enabled=1, inflation_radius=0.5, cost_scaling_factor=3.0, lethal_step_threshold=0.95)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    for i in range(n):
        c = synth.make_keyframe(4, i, poses[i], max_range=8.0, rings=64, azimuths=1024)
        s.addNewKeyFrameWithPCL(i, i + 1, 0, c); om.add_keyframe(i + 1, c)
        s.updateKeyFrame(i + 1, poses[i]); om.update_pose(i + 1, poses[i])
        s.flush(); om.process()
        if i >= window:  # local_traversability_node.cpp:181-184
            s.deleteKeyFrame(i + 1 - window); om.delete_keyframe(i + 1 - window)
        st = compare_maps(m, om, check_changed=True, label=f"step {i}")
    assert m.numKeyFrames() == window and st["gated_in"] > 100
    s.close()


# ---- config 3/4 shape: many keyframes, loop closure rebuild, drift model ------------------------
def test_loop_closure_rebuild_matches_fresh_build(tmap, oracle):
    from traversability_mapping_b200 import synth
    n = 12
    true = synth.trajectory("loop", n, seed=6, extent=6.0)
    drift = synth.drift_poses(true)
    clouds = [synth.make_keyframe(6, i, true[i], max_range=8.0, rings=32, azimuths=512) for i in range(n)]
    s, m, om = make_pair(tmap, oracle, resolution=0.1, half_size=10.0, max_range=8.0, min_range=1.0, max_batch=0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    for i in range(n):
        s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i]); om.add_keyframe(i + 1, clouds[i])
        s.updateKeyFrame(i + 1, drift[i]); om.update_pose(i + 1, drift[i])
    s.flush(); om.process_batched()
    compare_maps(m, om, check_changed=True, label="drifted build")
    # PGO batch: every keyframe gets its true pose (slam_keyframe_sim_with_drift.cpp:185-217) + loop closure
    for i in range(n):
        s.updateKeyFrame(i + 1, true[i]); om.update_pose(i + 1, true[i])
    s.informLoopClosure(); om.clear_entire_map()
    s.flush(); om.process_batched()
    compare_maps(m, om, check_changed=True, label="rebuild")
    # size-independent property: the rebuilt map equals a fresh build at the final poses, bit for bit
    s2, m2, om2 = make_pair(tmap, oracle, resolution=0.1, half_size=10.0, max_range=8.0, min_range=1.0, max_batch=0)
    s2.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    for i in range(n):
        s2.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i])
        s2.updateKeyFrame(i + 1, true[i])
    s2.flush()
    keys = m.occupiedCellKeys()
    assert np.array_equal(keys, m2.occupiedCellKeys())
    allL = helpers.MOMENT_LAYERS + [l for l in helpers.DERIVED_LAYERS if l != "step_haz_inflated"]
    assert helpers.bits_equal(m.readLayers(keys, allL), m2.readLayers(keys, allL))
    s.close(); s2.close()


def test_multi_map_reparent_and_obstacles(tmap, oracle):  # System.cpp:250-335
    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    s.addNewLocalMap(1)
    m1 = s.getLocalMap(1)
    s.addKeyFrameBase(1, 0, [[1, 0, 0], [2, 2, 0]])
    s.updateKeyFrame(1, I34)
    s.flush()
    assert m.occupiedCellKeys().size == 2 and m1.occupiedCellKeys().size == 0
    assert s.updateKFMap(1, 1) == tmap.OK
    s.flush()
    assert m.occupiedCellKeys().size == 0 and m1.occupiedCellKeys().size == 2
    assert s.updateKFMap(1, 5) == tmap.IGNORED_UNKNOWN_MAP
    # obstacle ids count down from INT64_MAX and are recycled
    assert s.addObstacleKeyFrame(0, np.array([[2, 0, 0, 0]], dtype=np.float32)) == 0  # params unset -> dropped
    s.setObstacleParameters(I34, 10.0, 0.5)
    a = s.addObstacleKeyFrame(0, np.array([[2, 0, 0, 0]], dtype=np.float32))
    b = s.addObstacleKeyFrame(0, np.array([[3, 0, 0, 0]], dtype=np.float32))
    assert a == 2 ** 63 - 1 and b == 2 ** 63 - 2
    s.deleteObstacleKeyFrame(a)
    assert s.addObstacleKeyFrame(0, np.array([[4, 0, 0, 0]], dtype=np.float32)) == a
    s.close()


def test_callback_fires_once_per_keyframe(tmap, oracle):
    p = tmap.default_params(**LOCALMAP)
    s = tmap.System(p)
    hits = []
    s.addNewLocalMap(0, onUpdate=lambda: hits.append(1))
    for i in range(3):
        s.addKeyFrameBase(i + 1, 0, [[i, 0, 0]])
        s.updateKeyFrame(i + 1, I34)
    s.flush()
    assert len(hits) == 3
    s.close()


def test_cpp_mirror_replays_localmap_kat(tmap):
    """tests/cpp_mirror_check.cpp: test_localmap.cpp:197-207 through include/tmap_system.hpp."""
    import subprocess
    exe = "/tmp/tmap_cpp_mirror_check_gpu"
    lib_dir = os.path.dirname(tmap.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp_mirror_check.cpp"),
                    "-L" + lib_dir, "-ltmap_b200", "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def _run_shard_check(nproc, n_kf, port, *extra):
    import subprocess, sys
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "shard_gpu_check.py"), str(n_kf), *extra],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bit-identical=True" in r.stdout and "bit-identical=False" not in r.stdout, \
        r.stdout[-3000:] + r.stderr[-3000:]


def test_sharded_rebuild_world1_matches_plain_rebuild(tmap):
    """The tmap_shard_* path (bin -> exchange -> merge -> fold -> halo -> classify) on one rank equals the
    ordinary informLoopClosure rebuild bit for bit on every layer."""
    _run_shard_check(1, 12, 29561)


def test_sharded_rebuild_world2_bit_identical(tmap):
    """Two ranks / two GPUs: all-to-all point routing + halo exchange; gathered slabs == 1-GPU rebuild."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_shard_check(2, 24, 29562)


def test_sharded_map_with_inflation_world1(tmap):
    """tmap_shard_rebuild + tmap_shard_update with step inflation on (one rank): every layer including
    step_haz_inflated equals the ordinary single-GPU sequence bit for bit."""
    _run_shard_check(1, 12, 29563, "inflate")


def test_sharded_map_with_inflation_world2(tmap):
    """Two ranks: the step layer's halo rows are exchanged between classification and inflation."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_shard_check(2, 24, 29564, "inflate")


def test_heavy_bins_split_across_ctas(tmap, oracle):
    """Many keyframes at almost the same pose make a few 16x16 bins hold far more records than a fold
    CTA's fair share, so they are split by cell-row group; results must still equal the oracle bit for bit."""
    from traversability_mapping_b200 import synth
    n = 90
    base = synth.trajectory("single", 1, seed=21)[0]
    s, m, om = make_pair(tmap, oracle, resolution=0.1, half_size=10.0, max_range=10.0, min_range=1.0, max_batch=0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    clouds = [synth.make_keyframe(21, i, base, max_range=10.0, rings=64, azimuths=1024) for i in range(6)]
    for i in range(n):
        P = base.copy()
        P[0, 3] += 0.003 * i
        P[1, 3] -= 0.002 * i
        s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i % 6]); om.add_keyframe(i + 1, clouds[i % 6])
        om.update_pose(i + 1, P)
    with m.locked():
        for i in range(n):
            P = base.copy()
            P[0, 3] += 0.003 * i
            P[1, 3] -= 0.002 * i
            s.updateKeyFrame(i + 1, P)
    s.flush(); om.process_batched()
    assert m.stats()["last_batch_max_bin"] > 60000, m.stats()["last_batch_max_bin"]
    compare_maps(m, om, check_changed=True, label="heavy bins")
    s.close()


def test_full_size_keyframes_against_oracle_batched(tmap, oracle):
    """20 full 131 072-ray keyframes on a loop (config 3's shape, reduced count so the oracle finishes in
    seconds), one batched pass: every layer against the oracle."""
    from traversability_mapping_b200 import synth
    n = 20
    poses = synth.trajectory("loop", n, seed=31, extent=12.0)
    s, m, om = make_pair(tmap, oracle, resolution=0.1, half_size=20.0, max_range=20.0, min_range=1.0, max_batch=0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    for i in range(n):
        c = synth.make_keyframe(31, i, poses[i], max_range=20.0)
        s.addNewKeyFrameWithPCL(i, i + 1, 0, c); om.add_keyframe(i + 1, c)
        om.update_pose(i + 1, poses[i])
    with m.locked():
        for i in range(n):
            s.updateKeyFrame(i + 1, poses[i])
    s.flush(); om.process_batched()
    st = compare_maps(m, om, check_changed=True, label="20 full keyframes")
    assert st["cells"] > 100000 and st["gated_in"] > 50000
    s.close()


def test_full_size_properties_without_oracle(tmap):
    """Size-independent properties at a size the oracle would not finish quickly (120 keyframes x 131 072
    rays, 8.7 M points): counts are exact, rebuilds are idempotent, a PGO round trip restores the counts,
    deleting everything empties the map."""
    from traversability_mapping_b200 import synth
    n, base = 120, 12
    poses = synth.trajectory("loop", n, seed=33, extent=40.0)
    drift = synth.drift_poses(poses)
    clouds = [synth.make_keyframe(33, i, poses[i], max_range=20.0) for i in range(base)]
    p = tmap.default_params(resolution=0.1, half_size=50.0, max_range=20.0, min_range=1.0, max_batch=0)
    s = tmap.System(p); s.addNewLocalMap(0); m = s.getLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    for i in range(n):
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i % base]) == tmap.OK
    total = m.getStitchedPointCloud().shape[0]

    def update_all(P):
        with m.locked():
            for i in range(n):
                s.updateKeyFrame(i + 1, P[i])
        s.flush()

    update_all(poses)
    keys0 = m.occupiedCellKeys()
    L0 = m.readLayers(keys0, helpers.MOMENT_LAYERS + helpers.DERIVED_LAYERS)
    assert int(L0[:, 0].sum()) == total, "sum of per-cell counts == points binned"
    assert np.all(L0[:, 0] == np.round(L0[:, 0])) and L0[:, 0].min() >= 1
    # zmin <= mean z <= zmax wherever the cell is observed
    zmean = L0[:, 3] / L0[:, 0]
    assert np.all(L0[:, 10] <= zmean + 1e-4) and np.all(zmean <= L0[:, 11] + 1e-4)
    # rebuild (informLoopClosure) reproduces the batched build bit for bit, twice
    for _ in range(2):
        s.informLoopClosure(); s.flush()
        keys = m.occupiedCellKeys()
        assert np.array_equal(keys, keys0)
        assert helpers.bits_equal(m.readLayers(keys, helpers.MOMENT_LAYERS + helpers.DERIVED_LAYERS), L0)
    # PGO round trip through drifted poses: float32 moments drift by rounding, the counts do not
    update_all(drift)
    update_all(poses)
    keys = m.occupiedCellKeys()
    assert np.array_equal(keys, keys0)
    assert np.array_equal(m.readLayers(keys, ["N"])[:, 0], L0[:, 0])
    # occupancy export: unknown where hazard is NaN, 0..100 elsewhere
    occ = m.exportOccupancy()
    haz = m.exportDense("hazard")
    assert np.array_equal(occ == -1, np.isnan(haz)) and occ.max() <= 100 and occ[~np.isnan(haz)].min() >= 0
    for i in range(n):
        s.deleteKeyFrame(i + 1)
    assert m.occupiedCellKeys().size == 0 and m.numKeyFrames() == 0
    s.close()


def test_growth_in_negative_directions_keeps_data(tmap, oracle):
    """Device window reallocation (DeviceGrid::ensure) towards -x/-y and far +y keeps every stored layer;
    the reference-visible extent follows growGridToInclude (Helpers.cpp:53-84)."""
    from traversability_mapping_b200 import synth
    s, m, om = make_pair(tmap, oracle, resolution=0.1, half_size=5.0, max_range=6.0, min_range=1.0, extend_length=7.0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    spots = [(0.0, 0.0), (-38.0, 3.0), (5.0, -61.0), (47.0, 90.0), (-120.0, -15.0)]
    for i, (x, y) in enumerate(spots):
        P = synth.pose_matrix(x, y, 0.2, 0.01, -0.02, 0.7 * i)
        c = synth.make_keyframe(41, i, P, max_range=6.0, rings=32, azimuths=512)
        s.addNewKeyFrameWithPCL(i, i + 1, 0, c); om.add_keyframe(i + 1, c)
        s.updateKeyFrame(i + 1, P); om.update_pose(i + 1, P)
        s.flush(); om.process()
        compare_maps(m, om, check_changed=True, label=f"spot {i}")
    assert m.stats()["grid_reallocs"] >= 4
    s.close()


def test_dropped_clouds_and_queued_deletes(tmap, oracle):
    """is_kf_optimization_enabled=false: the cloud is dropped after the first bin, later pose updates are
    ignored (LocalMap.cpp:257-267) but delete still subtracts; a keyframe deleted while queued is skipped
    (LocalMap.cpp:247-256)."""
    s, m, om = make_pair(tmap, oracle, **dict(LOCALMAP, is_kf_optimization_enabled=0))
    pts = [[1.0, 0.2, 0.1], [1.1, 0.2, 0.2], [3.0, 3.0, 0.0]]
    s.addKeyFrameBase(1, 0, pts); om.add_keyframe_base(1, pts)
    s.updateKeyFrame(1, I34); om.update_pose(1, I34)
    s.flush(); om.process()
    before = m.occupiedCellKeys().tolist()
    s.updateKeyFrame(1, trans(5, 0, 0)); om.update_pose(1, trans(5, 0, 0))   # ignored: no retained cloud
    s.flush(); om.process()
    assert m.occupiedCellKeys().tolist() == before
    compare_maps(m, om, check_changed=True)
    s.addKeyFrameBase(2, 0, [[2.0, 2.0, 0.0]]); om.add_keyframe_base(2, [[2.0, 2.0, 0.0]])
    with m.locked():                      # keyframe 2 is queued, then deleted before the worker gets to it: the
        s.updateKeyFrame(2, I34); om.update_pose(2, I34)   # grid lock (recursive) holds the worker off until the
        s.deleteKeyFrame(2); om.delete_keyframe(2)          # delete is in, whichever thread is faster
    s.flush(); om.process()
    s.deleteKeyFrame(1); om.delete_keyframe(1)
    assert m.occupiedCellKeys().size == 0 and m.numKeyFrames() == 0
    compare_maps(m, om, check_changed=True)
    s.close()


def test_degenerate_inputs(tmap, oracle):
    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    allnan = np.full((64, 4), np.nan, dtype=np.float32)
    assert s.addNewKeyFrameWithPCL(0, 1, 0, allnan) == tmap.IGNORED_EMPTY
    assert s.addNewKeyFrameWithPCL(0, 2, 0, np.zeros((0, 4), dtype=np.float32)) == tmap.IGNORED_EMPTY
    assert s.addKeyFrameBase(3, 0, np.zeros((0, 3), dtype=np.float32)) == tmap.IGNORED_EMPTY
    one = np.array([[2.0, 0.0, 0.0, 0.0]], dtype=np.float32)
    assert s.addNewKeyFrameWithPCL(0, 4, 0, one) == tmap.OK
    with pytest.raises(tmap.TmapError):       # a window beyond the 32-bit cell range is refused loudly
        s.updateKeyFrame(4, trans(1e12, 0, 0)); s.flush()
    s.close()
    # a footprint beyond the kernels' limits fails at map creation, not silently
    with pytest.raises(tmap.TmapError) as e:
        s2 = tmap.System(tmap.default_params(resolution=0.01, robot_radius=0.5))
        s2.addNewLocalMap(0)
    assert e.value.code == tmap.ERR_UNSUPPORTED


def test_keyframe_points_and_running_stats(tmap, oracle):
    """tmap_keyframe_points == the oracle's pruned cloud size; the running sums in tmap_stats add up the
    per-batch values (bench.py reads them once around its timed steps)."""
    from traversability_mapping_b200 import synth
    P = synth.trajectory("single", 1, seed=4)[0]
    cloud = synth.make_keyframe(4, 0, P, max_range=8.0, rings=32, azimuths=512)
    s = tmap.System(tmap.default_params(resolution=0.1, max_range=6.0, min_range=1.0, profile=1))
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om = oracle.OracleMap(oracle.params_from(tmap.default_params(resolution=0.1, max_range=6.0, min_range=1.0)))
    want = om.prune(cloud, synth.T_BASE_SENSOR, 6.0, 1.0)
    assert s.addNewKeyFrameWithPCL(1, 1, 0, cloud) == tmap.OK
    assert s.keyFramePoints(1) == len(want)
    assert s.keyFramePoints(99) is None   # unknown id: ignored, not an error
    m = s.getLocalMap(0)
    tot = {k: 0.0 for k in ("ms_hist", "ms_fold", "ms_classify", "ms_gate", "ms_total")}
    touched = 0
    for k in range(3):
        Pk = P.copy(); Pk[0, 3] += 0.7 * k
        s.updateKeyFrame(1, Pk); s.flush()
        st = m.stats()
        for key in tot: tot[key] += st[key]
        touched += st["last_batch_cells_touched"]
    st = m.stats()
    assert st["batches"] == 3 and st["points_binned"] == len(want) * 5   # add, then (subtract + add) twice
    for key in tot:
        assert abs(st["sum_" + key] - tot[key]) < 1e-4 * max(1.0, tot[key]), key
    assert 0.0 < st["ms_gate"] <= st["ms_classify"]
    assert st["sum_cells_touched"] == touched and st["sum_cand_bins"] > 0
    assert st["sum_host_ms_batch"] >= st["sum_host_ms_wait"] > 0.0
    s.close()


def test_bulk_pose_update_equals_per_keyframe_calls(tmap):
    """tmap_update_poses (one PGO message) == the same updateKeyFrame calls one by one: identical moment layers."""
    from traversability_mapping_b200 import synth
    poses = synth.trajectory("loop", 6, seed=9, extent=6.0)
    clouds = [synth.make_keyframe(9, i, poses[i], max_range=8.0, rings=32, azimuths=512) for i in range(6)]
    moved = poses.copy(); moved[:, 0, 3] += 0.37; moved[:, 1, 3] -= 0.21
    maps = []
    for bulk in (False, True):
        s = tmap.System(tmap.default_params(resolution=0.1, max_range=8.0, min_range=1.0, max_batch=0))
        s.addNewLocalMap(0)
        s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
        m = s.getLocalMap(0)
        for i, c in enumerate(clouds):
            assert s.addNewKeyFrameWithPCL(i, i + 1, 0, c) == tmap.OK
        for target in (poses, moved):
            with m.locked():
                if bulk:
                    assert s.updateKeyFrames(np.arange(1, 7), target) == tmap.OK
                else:
                    for i in range(6):
                        s.updateKeyFrame(i + 1, target[i])
            s.flush()
        kx, ky = m.gridExtent()
        maps.append([m.exportDense(l, -kx, -ky, 2 * kx + 1, 2 * ky + 1) for l in ("N", "sz", "hazard")])
        s.close()
    for a, b in zip(*maps):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.nansum(maps[0][0]) > 0


def test_pointcloud2_ingest_matches_pcl_path(tmap, oracle):
    """tmap_add_keyframe_pc2: a PointCloud2 byte buffer (x, y, z FLOAT32 at arbitrary aligned offsets, little or
    big endian) gives the keyframe the plain (n, 4) float cloud gives -- bit for bit, through the device prune."""
    from traversability_mapping_b200 import synth
    P = synth.trajectory("single", 1, seed=6)[0]
    cloud = synth.make_keyframe(6, 0, P, max_range=8.0, rings=32, azimuths=512)
    cloud[3] = (0.0, 0.0, 1.0, 0)
    cloud[4] = (np.nan, 1.0, 1.0, 0)
    n = len(cloud)
    rng = np.random.default_rng(0)
    buf = rng.standard_normal((n, 8)).astype(np.float32)     # point_step 32: intensity / ring / padding junk
    buf[:, 1] = cloud[:, 0]; buf[:, 5] = cloud[:, 1]; buf[:, 2] = cloud[:, 2]   # x @4, y @20, z @8
    s = tmap.System(tmap.default_params(resolution=0.1, max_range=6.0, min_range=1.0))
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    m = s.getLocalMap(0)
    assert s.addNewKeyFrameWithPCL(1, 1, 0, cloud) == tmap.OK
    want = m.getStitchedPointCloud()
    s.deleteKeyFrame(1)
    assert s.addNewKeyFramePointCloud2(2, 2, 0, buf, 32, 4, 20, 8) == tmap.OK
    got = m.getStitchedPointCloud()
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    s.deleteKeyFrame(2)
    assert s.addNewKeyFramePointCloud2(3, 3, 0, buf.byteswap(), 32, 4, 20, 8, is_bigendian=True) == tmap.OK
    got = m.getStitchedPointCloud()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    with pytest.raises(tmap.TmapError) as e:   # a misaligned field is refused loudly
        s.addNewKeyFramePointCloud2(4, 4, 0, buf, 32, 2, 20, 8)
    assert e.value.code == tmap.ERR_UNSUPPORTED
    s.close()


def test_completeness_export_matches_oracle(tmap, oracle):
    """tmap_export_completeness == the oracle's restatement of publishCompleteness (border_haz occupancy,
    min-pool to the publish resolution, RViz palette bytes): integer work, so byte for byte."""
    s, m, om, _ = run_small_scene(tmap, oracle)
    for res in (0.25, 0.5, 0.01):
        got, want = m.exportCompleteness(res), om.export_completeness(res)
        assert got.shape == want.shape and np.array_equal(got, want), res
    kx, ky = m.gridExtent()
    got = m.exportCompleteness(0.25, -kx + 3, -ky + 1, 2 * kx - 7, 2 * ky - 2)   # crop, partial edge blocks
    want = om.export_completeness(0.25, -kx + 3, -ky + 1, 2 * kx - 7, 2 * ky - 2)
    assert got.shape == want.shape and np.array_equal(got, want)
    vals = set(np.unique(want).tolist())
    assert -1 in vals and 113 in vals and any(-128 <= v < -1 for v in vals), "unobserved, complete and partial cells all present"
    s.close()


def test_stitched_cloud_voxel_filter_matches_oracle(tmap, oracle):
    """getStitchedPointCloud with leaf sizes: the pcl::VoxelGrid restatement of the oracle on the same stitched
    cloud.  Voxel set and order are integer work (exact); both sides sum a voxel's points in (voxel, index) order,
    so the centroids are bit-identical too."""
    from traversability_mapping_b200 import synth
    poses = synth.trajectory("loop", 4, seed=11, extent=5.0)
    s = tmap.System(tmap.default_params(resolution=0.1, max_range=8.0, min_range=1.0))
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    m = s.getLocalMap(0)
    for i in range(4):
        c = synth.make_keyframe(11, i, poses[i], max_range=8.0, rings=32, azimuths=512)
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, c) == tmap.OK
        s.updateKeyFrame(i + 1, poses[i])
    s.flush()
    full = m.getStitchedPointCloud()
    assert len(full) > 10000
    for leaf in ((0.25, 0.25, 0.25), (0.5, 0.2, 1.0)):
        want = oracle.voxel_grid(full, *leaf)
        got = m.getStitchedPointCloud(*leaf)
        assert got.shape == want.shape and 0 < len(got) < len(full)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), leaf
    tiny = m.getStitchedPointCloud(1e-4, 1e-4, 1e-4)     # indices would overflow: PCL returns the input cloud
    assert np.array_equal(tiny.view(np.uint32), full.view(np.uint32))
    assert np.array_equal(oracle.voxel_grid(full, 1e-4, 1e-4, 1e-4).view(np.uint32), full.view(np.uint32))
    s.close()


def test_pointcloud_buffer_and_timestamped_keyframes(tmap, oracle):
    """pushToBuffer + addNewKeyFrameTsDouble (System.cpp:174-186, 339-344; PointCloudBuffer.cpp): the keyframe's cloud
    is the buffered one closest in time (first on a tie); a span over 25 s drops everything older than t - 5 s; without
    use_pointcloud_buffer both calls do nothing."""
    from traversability_mapping_b200 import synth
    P = synth.trajectory("single", 1, seed=8)[0]
    clouds = {t: synth.make_keyframe(8, i, P, max_range=8.0, rings=16, azimuths=256) for i, t in enumerate((1.0, 2.0, 3.0))}
    s = tmap.System(tmap.default_params(resolution=0.1, max_range=6.0, min_range=1.0, use_pointcloud_buffer=1))
    s.addNewLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    m = s.getLocalMap(0)
    assert s.addNewKeyFrameTsDouble(1.0, 1, 0) == tmap.IGNORED_NULL     # empty buffer
    for t, c in clouds.items():
        assert s.pushToBuffer(t, c) == tmap.OK
    assert s.bufferedClouds() == 3

    def via_pcl(c):
        assert s.addNewKeyFrameWithPCL(0, 99, 0, c) == tmap.OK
        out = m.getStitchedPointCloud().copy()
        s.deleteKeyFrame(99)
        return out

    for kf, (query, expect) in enumerate([(2.1, 2.0), (1.5, 1.0), (9.0, 3.0), (-4.0, 1.0)], start=1):
        assert s.addNewKeyFrameTsDouble(query, kf, 0) == tmap.OK
        got = m.getStitchedPointCloud().copy()
        s.deleteKeyFrame(kf)
        want = via_pcl(clouds[expect])
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (query, expect)
    assert s.addNewKeyFrameTsULong(2_900_000_000, 7, 0) == tmap.OK       # 2.9 s -> the 3.0 s cloud
    got = m.getStitchedPointCloud().copy()
    s.deleteKeyFrame(7)
    assert np.array_equal(got.view(np.uint32), via_pcl(clouds[3.0]).view(np.uint32))
    # pruning: 1, 2, 3 are buffered; 27 makes the span 26 s > 25 s -> everything older than 22 s goes
    s.pushToBuffer(27.0, clouds[1.0])
    assert s.bufferedClouds() == 1
    s.pushToBuffer(30.0, clouds[2.0])
    assert s.bufferedClouds() == 2                                        # span 3 s: nothing dropped
    s.close()
    s2 = tmap.System(tmap.default_params(resolution=0.1))                 # buffer not enabled
    s2.addNewLocalMap(0)
    assert s2.pushToBuffer(1.0, clouds[1.0]) == tmap.OK and s2.bufferedClouds() == 0
    assert s2.addNewKeyFrameTsDouble(1.0, 1, 0) == tmap.IGNORED_NULL
    s2.close()


def test_sparse_update_payload_matches_oracle(tmap, oracle):
    """tmap_fill_sparse_update == fillSparseUpdate (conversions.cpp:28-55): keys outside the grid are dropped, the
    rest keep their order, values per (key, layer)."""
    s, m, om, _ = run_small_scene(tmap, oracle)
    occ = m.occupiedCellKeys()
    assert np.array_equal(occ, om.occupied_cell_keys())
    kx, ky = m.gridExtent()
    far = np.array([tmap.key(kx + 5, 0), tmap.key(0, -ky - 1), tmap.key(-kx - 40, ky + 7)], dtype=np.uint64)
    keys = np.concatenate([far[:1], occ[::3], far[1:2], occ[1::7], far[2:]])
    layers = ["hazard", "elevation", "step_haz_inflated", "N"]
    gk, gv = m.fillSparseUpdate(layers, keys)
    wk, wv = om.fill_sparse_update(layers, keys)
    assert np.array_equal(gk, wk) and len(gk) == len(keys) - 3
    assert gv.shape == wv.shape
    assert np.array_equal(np.isnan(gv), np.isnan(wv))
    np.testing.assert_allclose(gv, wv, rtol=helpers.RTOL, atol=helpers.ATOL, equal_nan=True)
    assert np.array_equal(gv[:, 3].view(np.uint32), wv[:, 3].view(np.uint32))   # counts: bit-exact
    ek, ev = m.fillSparseUpdate(layers, np.zeros(0, dtype=np.uint64))
    assert len(ek) == 0 and ev.shape == (0, 4)
    s.close()


def test_localmap_registry_and_stitch_kats(tmap, oracle):
    """The remaining LocalMap tests of the reference through the C ABI: AddMultipleTracksAllInKeyFramesMap
    (test_localmap.cpp:185-193), DetachUnknownIdReturnsNull (:264-268), DeleteKeyFrameRemovesIt (:270-280),
    StitchTransformsRetainedCloudsByPose (:305-318), StitchSkipsKeyFramesWithoutCloud (:320-327),
    StitchVoxelDownsampleReducesCount (:329-340) -- and the same calls on the oracle."""
    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    for k, x in ((1, 0.0), (2, 1.0), (3, 2.0)):
        s.addKeyFrameBase(k, 0, [[x, 0.0, 0.0]]); om.add_keyframe_base(k, [[x, 0.0, 0.0]])
    assert m.numKeyFrames() == 3 == om.num_keyframes()
    assert s.deleteKeyFrame(999) == tmap.IGNORED_UNKNOWN_KF
    s.close()

    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    s.addKeyFrameBase(5, 0, [[0.0, 0.0, 0.0]]); om.add_keyframe_base(5, [[0.0, 0.0, 0.0]])
    s.updateKeyFrame(5, I34); om.update_pose(5, I34)
    s.flush(); om.process()
    assert m.numKeyFrames() == 1 and m.occupiedCellKeys().size == 1
    s.deleteKeyFrame(5); om.delete_keyframe(5)
    assert m.numKeyFrames() == 0 and m.occupiedCellKeys().size == 0 == om.occupied_cell_keys().size
    s.close()

    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    pts = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]
    s.addKeyFrameBase(1, 0, pts); om.add_keyframe_base(1, pts)
    s.updateKeyFrame(1, trans(5, -3, 2)); om.update_pose(1, trans(5, -3, 2))
    s.flush(); om.process()
    got = m.getStitchedPointCloud()
    assert sorted(map(tuple, got.tolist())) == [(5.0, -2.0, 2.0), (6.0, -3.0, 2.0)]
    assert np.array_equal(got.view(np.uint32), om.stitched_cloud().view(np.uint32))
    assert s.addKeyFrameBase(2, 0, np.zeros((0, 3), dtype=np.float32)) == tmap.IGNORED_EMPTY   # no cloud: not stitched
    assert len(m.getStitchedPointCloud()) == 2
    s.close()

    s, m, om = make_pair(tmap, oracle, **LOCALMAP)
    ten = [[0.001 * i, 0.0, 0.0] for i in range(10)]
    s.addKeyFrameBase(1, 0, ten); om.add_keyframe_base(1, ten)
    got = m.getStitchedPointCloud(1.0, 1.0, 1.0)
    want = om.stitched_cloud(1.0, 1.0, 1.0)
    assert 1 <= len(got) < 10 and len(got) == 1
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert abs(float(got[0, 0]) - 0.0045) < 1e-6
    s.close()


# ---- round 2: the benchmarked configurations at their benchmarked sizes --------------------------------------
def _torch_clouds(seed, idxs, poses, rings, az):
    """Host clouds from the batched torch generator (bench.py's): both arms consume the same arrays."""
    import torch
    from traversability_mapping_b200 import synth
    out = {}
    idxs = list(idxs)
    for a in range(0, len(idxs), 16):
        part = idxs[a:a + 16]
        t = synth.make_keyframes_torch(seed, part[0], poses[part], max_range=20.0, rings=rings, azimuths=az,
                                       device="cuda" if torch.cuda.is_available() else "cpu")
        h = t.cpu().numpy()
        for k, i in enumerate(part):
            out[i] = h[k]
    return out


def test_config2_at_benchmarked_size(tmap, oracle):
    """bench.py's C2 leg exactly: 131 072-ray keyframes, 20 m range, 40 x 40 m grid @0.05 m, 49-cell disc,
    min_points 5, inflation 0.5 m / 3.0 / 0.95, rolling window 15, max_batch = 1 -- the window is filled and
    three more ticks (each with an eviction) are compared layer by layer with the oracle."""
    import bench
    from traversability_mapping_b200 import synth
    window, extra = bench.WINDOW, 3
    n = window + extra
    poses = synth.trajectory("drive", n, seed=bench.SEED, step=0.1, curvature=0.02)
    clouds = _torch_clouds(bench.SEED, range(n), poses, 128, 1024)
    s, m, om = make_pair(tmap, oracle, max_batch=1, **bench.C2)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    om.set_threads(os.cpu_count() or 1)   # same results as 1 thread (asserted in test_oracle_kat.py); the fill takes seconds
    st = None
    for i in range(n):
        s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i]); om.add_keyframe(i + 1, clouds[i])
        s.updateKeyFrame(i + 1, poses[i]); om.update_pose(i + 1, poses[i])
        s.flush(); om.process()
        if i >= window:
            s.deleteKeyFrame(i + 1 - window); om.delete_keyframe(i + 1 - window)
        if i >= window - 1:
            st = compare_maps(m, om, check_changed=True, label=f"C2 tick {i}")
        else:
            m.takeChangedCells(); om.take_changed_cells()
    assert m.numKeyFrames() == window and st["cells"] > 150000 and st["gated_in"] > 15000, st
    s.close()


def test_config4_200_keyframes_against_oracle(tmap, oracle):
    """bench.py's C4 headline at 200 of its 2 000 keyframes (262 144 rays each, 200 x 200 m map @0.1 m, 13-cell
    disc): drifted build, then the loop closure -- every pose corrected + informLoopClosure -- in one batched pass,
    against om.process_batched()."""
    import bench
    from traversability_mapping_b200 import synth
    n = 200
    true, drift = bench.c4_poses(bench.C4_KEYFRAMES)
    clouds = _torch_clouds(bench.SEED + 1, range(n), true, bench.C4_RINGS, bench.C4_AZ)
    s, m, om = make_pair(tmap, oracle, max_batch=0, **bench.C4)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    om.set_threads(os.cpu_count() or 1)
    ids = np.arange(1, n + 1, dtype=np.uint64)
    for i in range(n):
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i]) == tmap.OK
        om.add_keyframe(i + 1, clouds[i]); om.update_pose(i + 1, drift[i])
    with m.locked():
        s.updateKeyFrames(ids, drift[:n])
    s.flush(); om.process_batched()
    compare_maps(m, om, check_changed=True, label="C4 drifted build")
    for i in range(n):
        om.update_pose(i + 1, true[i])
    with m.locked():
        s.updateKeyFrames(ids, true[:n])
        s.informLoopClosure()
    om.clear_entire_map()
    s.flush(); om.process_batched()
    st = compare_maps(m, om, check_changed=True, label="C4 loop closure")
    N = m.readLayers(m.occupiedCellKeys(), ["N"])[:, 0]
    assert int(N.sum()) == m.stats()["last_batch_points"] and st["cells"] > 300000 and st["gated_in"] > 200000, st
    s.close()


def test_config5_geometry_reduced_extent(tmap, oracle):
    """C5's geometry -- 0.05 m lattice with robot_radius 0.20 => the 49-cell disc -- on a reduced extent (a 60 m loop,
    40 keyframes), one batched build + loop closure against the oracle."""
    import bench
    from traversability_mapping_b200 import synth
    n = 40
    P = dict(bench.C4, resolution=0.05, half_size=40.0)
    true = synth.trajectory("loop", n, seed=77, extent=30.0)
    drift = synth.drift_poses(true)
    clouds = _torch_clouds(77, range(n), true, 128, 1024)
    s, m, om = make_pair(tmap, oracle, max_batch=0, **P)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    om.set_threads(os.cpu_count() or 1)
    ids = np.arange(1, n + 1, dtype=np.uint64)
    for i in range(n):
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i]) == tmap.OK
        om.add_keyframe(i + 1, clouds[i]); om.update_pose(i + 1, drift[i])
    with m.locked():
        s.updateKeyFrames(ids, drift)
    s.flush(); om.process_batched()
    compare_maps(m, om, check_changed=True, label="C5 geometry, drifted")
    for i in range(n):
        om.update_pose(i + 1, true[i])
    with m.locked():
        s.updateKeyFrames(ids, true)
        s.informLoopClosure()
    om.clear_entire_map()
    s.flush(); om.process_batched()
    st = compare_maps(m, om, check_changed=True, label="C5 geometry, loop closure")
    assert st["cells"] > 1000000 and st["gated_in"] > 40000, st
    s.close()


def test_config4_full_size_properties(tmap):
    """The benchmarked C4 size itself (2 000 keyframes x 262 144 rays, ~290 M points): size-independent properties --
    per-cell counts sum to the points binned, a second loop closure at the same poses reproduces every layer bit for
    bit, the sub-map around the first 200 keyframes is untouched by where the other 1 800 sit in the batch (checksum
    of the count layer against a 200-keyframe build restricted to cells only they reach)."""
    import bench
    from traversability_mapping_b200 import synth
    nk = bench.C4_KEYFRAMES
    true, drift = bench.c4_poses(nk)
    p = tmap.default_params(max_batch=0, **bench.C4)
    s = tmap.System(p); s.addNewLocalMap(0); m = s.getLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    for i, c in bench.gen_host_clouds(bench.SEED + 1, range(nk), true, bench.C4_RINGS, bench.C4_AZ, "cuda"):
        assert s.addNewKeyFrameWithPCL(i, i + 1, 0, c) == tmap.OK
    ids = np.arange(1, nk + 1, dtype=np.uint64)
    with m.locked():
        s.updateKeyFrames(ids, drift)
    s.flush()
    with m.locked():
        s.updateKeyFrames(ids, true)
        s.informLoopClosure()
    s.flush()
    total = m.stats()["last_batch_points"]   # one pass: everything queued under the grid mutex is folded together
    assert total > 256e6 and total == sum(s.keyFramePoints(i + 1) for i in range(0, nk, 1)), total
    keys0 = m.occupiedCellKeys()
    allL = helpers.MOMENT_LAYERS + [l for l in helpers.DERIVED_LAYERS if l != "step_haz_inflated"]
    L0 = m.readLayers(keys0, allL)
    assert int(L0[:, 0].astype(np.float64).sum()) == total
    assert np.all(L0[:, 0] == np.round(L0[:, 0])) and L0[:, 0].min() >= 1
    s.informLoopClosure(); s.flush()
    keys1 = m.occupiedCellKeys()
    assert np.array_equal(keys0, keys1) and helpers.bits_equal(m.readLayers(keys1, allL), L0)
    occ = m.exportOccupancy()
    haz = m.exportDense("hazard")
    assert np.array_equal(occ == -1, np.isnan(haz)) and occ.max() <= 100
    s.close()


def test_reader_thread_against_delete_and_loop_closure(tmap):
    """a reader that holds the grid mutex across take + read (LocalMap.hpp:102-104) must not
    deadlock against deleteKeyFrame / informLoopClosure / updateKFMap running on another thread."""
    import threading
    from traversability_mapping_b200 import synth
    p = tmap.default_params(resolution=0.1, half_size=10.0, max_range=10.0, min_range=1.0, max_batch=1)
    s = tmap.System(p); s.addNewLocalMap(0); s.addNewLocalMap(1); m = s.getLocalMap(0)
    s.setExtrinsicParameters(synth.IDENTITY, synth.T_BASE_SENSOR)
    poses = synth.trajectory("drive", 40, seed=5, step=0.3)
    clouds = [synth.make_keyframe(5, i, poses[i], max_range=10.0, rings=16, azimuths=512) for i in range(8)]
    stop = threading.Event()
    reads = [0]
    err = []

    def reader():
        try:
            while not stop.is_set():
                with m.locked():
                    k = m.takeChangedCells()
                    if k.size:
                        m.readLayers(k[:256], ["hazard", "N"])
                    m.exportOccupancy(-20, -20, 41, 41)
                reads[0] += 1
        except Exception as e:  # pragma: no cover
            err.append(e)

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    for i in range(40):
        s.addNewKeyFrameWithPCL(i, i + 1, 0, clouds[i % 8])
        s.updateKeyFrame(i + 1, poses[i])
        if i >= 5:
            s.deleteKeyFrame(i - 4)
        if i % 7 == 3:
            s.updateKFMap(i + 1, 1)
        if i % 10 == 9:
            s.informLoopClosure()
    s.flush()
    stop.set()
    t.join(timeout=60)
    assert not t.is_alive(), "reader thread is stuck: lock-order deadlock"
    assert not err and reads[0] > 0
    s.close()

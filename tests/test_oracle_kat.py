"""Oracle pinned against the reference's own gtest vectors (oracle/kat.cpp) and cross-checks."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_known_answer_vectors(oracle):
    """oracle/kat replays test_moments.cpp, test_keyframe.cpp, test_localmap.cpp, test_helpers.cpp."""
    r = subprocess.run([os.path.join(ROOT, "oracle", "kat")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


def test_lattice_kats(oracle):
    # test_moments.cpp:368-390
    assert oracle.cell_of(0, 0, 0.25, 1.04, -0.36) == (4, -1)
    assert oracle.lib().tmo_key(4, -1) == ((4 << 32) | 0xFFFFFFFF)
    # round half away from zero (std::lround)
    assert oracle.cell_of(0, 0, 0.5, 0.25, -0.25) == (1, -1)
    assert oracle.cell_of(0, 0, 0.5, 0.75, -0.75) == (2, -2)


def test_disc_offsets(oracle):
    # Helpers.cpp:156-169; SURVEY 8: 13-cell / 49-cell discs, 3x3 window, 0.3/0.1 excludes (+-3, 0)
    assert len(oracle.disc_offsets(0.20, 0.1)) == 13
    assert len(oracle.disc_offsets(0.20, 0.05)) == 49
    assert len(oracle.disc_offsets(0.4, 0.25)) == 9
    d = oracle.disc_offsets(0.3, 0.1)
    assert len(d) == 25 and [3, 0] not in d.tolist()
    # iteration order: di outer ascending, dj inner ascending
    d13 = oracle.disc_offsets(0.20, 0.1).tolist()
    assert d13 == sorted(d13)


def test_eigh3_matches_numpy(oracle):
    rng = np.random.default_rng(3)
    for _ in range(300):
        B = rng.standard_normal((3, 3))
        A = B @ B.T * 10 ** rng.uniform(-6, 2)
        w, V = oracle.eigh3(A)
        w_np = np.linalg.eigvalsh(A)
        assert np.allclose(w, w_np, rtol=1e-10, atol=1e-14 * abs(w_np).max())
        assert np.allclose(V @ np.diag(w) @ V.T, A, atol=1e-12 * abs(A).max())
    w, V = oracle.eigh3(np.zeros((3, 3)))
    assert np.all(w == 0) and np.array_equal(V, np.eye(3))  # Eigen: zero matrix -> Q = I


def test_rebin_kat_vectors(oracle):
    # test_keyframe.cpp:120-139
    I = np.eye(3, 4, dtype=np.float32)
    keys, mom = oracle.rebin(0, 0, 0.25, [[1.04, -0.36, 2.0]], I)
    assert keys.tolist() == [(4 << 32) | 0xFFFFFFFF]
    assert mom[0, 0] == 1
    assert abs(mom[0, 1] - 0.04) < 1e-4 and abs(mom[0, 2] + 0.11) < 1e-4 and abs(mom[0, 3] - 2.0) < 1e-4
    # test_keyframe.cpp:195-204
    keys, _ = oracle.rebin(0, 0, 0.25, [[0, 0, 0], [5, 0, 0]], I)
    assert sorted(keys.tolist()) == [0, 20 << 32]
    # test_keyframe.cpp:206-213
    keys, _ = oracle.rebin(0, 0, 0.25, np.zeros((0, 3)), I)
    assert keys.size == 0


def test_golden_fixture(oracle):
    """The committed fixture (tests/golden/make_golden.py) still reproduces: guards against oracle drift."""
    from golden.make_golden import scenario
    g = np.load(os.path.join(ROOT, "tests", "golden", "small_scene.npz"))
    om, _ = scenario(oracle)
    keys = om.occupied_cell_keys()
    assert np.array_equal(keys, g["keys"])
    import helpers
    L = om.read_layers(keys, helpers.MOMENT_LAYERS + helpers.DERIVED_LAYERS)
    assert helpers.bits_equal(L, g["layers"])
    assert np.array_equal(om.export_occupancy(), g["occupancy"])


def test_batched_vs_sequential_oracle(oracle):
    """Moment layers are identical in both worker modes; derived layers may differ only in the
    elevation of cells whose fit is gated out (stale-elevation rule, LocalMap.cpp:163-164)."""
    from golden.make_golden import scenario
    import helpers
    a, _ = scenario(oracle, batched=False)
    b, _ = scenario(oracle, batched=True)
    keys = a.occupied_cell_keys()
    assert np.array_equal(keys, b.occupied_cell_keys())
    assert helpers.bits_equal(a.read_layers(keys, helpers.MOMENT_LAYERS), b.read_layers(keys, helpers.MOMENT_LAYERS))
    da = a.read_layers(keys, helpers.DERIVED_LAYERS)
    db = b.read_layers(keys, helpers.DERIVED_LAYERS)
    gated_out = np.isnan(da[:, 0])
    for i, l in enumerate(helpers.DERIVED_LAYERS):
        if l == "elevation":
            assert helpers.bits_equal(da[~gated_out, i], db[~gated_out, i])
        else:
            assert helpers.bits_equal(da[:, i], db[:, i]), l


def test_voxel_grid_restatement_vs_independent_numpy(oracle):
    """tmo_voxel_grid (pcl::VoxelGrid restated, PCL 1.12 voxel_grid.hpp:214-426) against an independent numpy
    statement of the same float arithmetic: voxel ids, output order, centroids (sequential float32 sums)."""
    rng = np.random.default_rng(5)
    pts = (rng.standard_normal((5000, 3)) * np.array([6.0, 4.0, 1.5])).astype(np.float32)
    for leaf in ((0.5, 0.5, 0.5), (1.0, 0.25, 2.0)):
        got = oracle.voxel_grid(pts, *leaf)
        inv = (np.float32(1.0) / np.asarray(leaf, dtype=np.float32)).astype(np.float32)
        mn, mx = pts.min(0), pts.max(0)
        min_b = np.floor(mn * inv).astype(np.int64)
        div_b = np.floor(mx * inv).astype(np.int64) - min_b + 1
        ijk = (np.floor(pts * inv).astype(np.float32) - min_b.astype(np.float32)).astype(np.int64)
        idx = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
        order = np.argsort(idx, kind="stable")
        sidx = idx[order]
        starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
        ends = np.r_[starts[1:], len(sidx)]
        want = np.empty((len(starts), 3), dtype=np.float32)
        for k, (a, b) in enumerate(zip(starts, ends)):
            seg = pts[order[a:b]]
            want[k] = np.cumsum(seg, axis=0, dtype=np.float32)[-1] / np.float32(b - a)   # sequential float32 sums
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a leaf so small that the voxel index would overflow int32: PCL warns and returns the input
    assert np.array_equal(oracle.voxel_grid(pts, 1e-4, 1e-4, 1e-4), pts)


def test_completeness_restatement_vs_independent_numpy(oracle):
    """tmo_export_completeness (publishCompleteness, global_traversability_node.cpp:381-459) against a numpy
    statement: border_haz -> int8 occupancy, f x f min-pool skipping unobserved cells, palette bytes."""
    import traversability_mapping_b200 as tm
    import helpers
    from golden.make_golden import PARAMS, scene_inputs
    from traversability_mapping_b200 import synth
    poses, clouds = scene_inputs()
    om = oracle.OracleMap(oracle.params_from(tm.default_params(**PARAMS)))
    om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
    for i, c in enumerate(clouds):
        om.add_keyframe(i + 1, c); om.update_pose(i + 1, poses[i]); om.process()
    kx, ky = om.grid_extent()
    border = om.export_dense("border_haz", -kx, -ky, 2 * kx + 1, 2 * ky + 1)
    scaled = np.nan_to_num(np.clip(border, 0, 1) * np.float32(100), nan=0.0)
    fine = np.where(np.isnan(border), -1, scaled.astype(np.int8)).astype(np.int64)
    for publish in (PARAMS["resolution"], 2 * PARAMS["resolution"], 3 * PARAMS["resolution"]):
        f = max(1, int(np.floor(publish / float(np.float32(PARAMS["resolution"])) + 0.5)))
        h, w = fine.shape
        ch, cw = (h + f - 1) // f, (w + f - 1) // f
        pad = np.full((ch * f, cw * f), 101, dtype=np.int64)
        pad[:h, :w] = np.where(fine >= 0, fine, 101)
        best = pad.reshape(ch, f, cw, f).min(axis=(1, 3))
        want = np.where(best > 100, -1, np.where(best >= 100, 113, 128 + np.floor(best / 100.0 * 126 + 0.5).astype(np.int64)))
        want = want.astype(np.uint8).view(np.int8)
        got = om.export_completeness(publish)
        assert got.shape == want.shape and np.array_equal(got, want), publish


def test_threaded_classification_equals_single_thread(oracle):
    """CPU-baseline flavour 2 (dirty cells classified in parallel, as the archived node does,
    archive/local_traversability_monolithic.cpp:349-362) must give the single-thread oracle's map bit for bit."""
    import helpers
    from traversability_mapping_b200 import synth
    n = 4
    poses = synth.trajectory("loop", n, seed=12, extent=5.0)
    clouds = [synth.make_keyframe(12, i, poses[i], max_range=8.0, rings=32, azimuths=512) for i in range(n)]
    out = {}
    for th in (1, 4):
        om = oracle.OracleMap(oracle.default_params(resolution=0.1, half_size=10.0, max_range=8.0, min_range=1.0,
                                                    inflation_enabled=1, inflation_radius=0.5))
        om.set_threads(th)
        om.set_extrinsics(helpers.compose_f32(synth.T_BASE_SENSOR, synth.IDENTITY))
        for i in range(n):
            om.add_keyframe(i + 1, clouds[i]); om.update_pose(i + 1, poses[i])
            om.process()
        om.delete_keyframe(1)
        keys = om.occupied_cell_keys()
        out[th] = (keys, om.read_layers(keys, helpers.MOMENT_LAYERS + helpers.DERIVED_LAYERS), om.take_changed_cells())
        om.close()
    assert np.array_equal(out[1][0], out[4][0]) and helpers.bits_equal(out[1][1], out[4][1]) and np.array_equal(out[1][2], out[4][2])


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a tiny sample: one JSON line
    with the contract's keys."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-keyframes", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "points/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("C4 loop closure")

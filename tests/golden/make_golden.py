"""Generates tests/golden/small_scene.npz from the CPU oracle (run: python -m tests.golden.make_golden).

The reference cannot be built in this container (no Eigen/PCL/grid_map), so the golden vectors are the
oracle's outputs on a seeded scene, frozen at the commit where the oracle passed every reference KAT
(oracle/kat).  The parity tests compare the CUDA path with this fixture as well as with the live oracle.
Scene: 6 keyframes of 16x256 rays on a short drive, resolution 0.25 (3x3 footprint window), one pose
update, one delete -- exercises add, subtract, blanking, growth and classification.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

PARAMS = dict(resolution=0.25, half_size=7.5, extend_length=30.0, robot_radius=0.4, min_points_per_grid=2,
              max_range=12.0, min_range=1.0)


def scene_inputs():
    from traversability_mapping_b200 import synth
    poses = synth.trajectory("drive", 6, seed=11, step=0.8, curvature=0.05)
    clouds = [synth.make_keyframe(11, i, poses[i], max_range=12.0, rings=16, azimuths=256) for i in range(6)]
    return poses, clouds


def scenario(oracle, batched=False):
    from traversability_mapping_b200 import synth
    poses, clouds = scene_inputs()
    om = oracle.OracleMap(oracle.default_params(**PARAMS))
    om.set_extrinsics(synth.T_BASE_SENSOR)
    proc = om.process_batched if batched else om.process
    for i, c in enumerate(clouds):
        om.add_keyframe(i + 1, c)
        om.update_pose(i + 1, poses[i])
    proc()
    moved = poses[2].copy()
    moved[0, 3] += 1.3
    om.update_pose(3, moved)
    proc()
    om.delete_keyframe(1)
    return om, (poses, clouds, moved)


if __name__ == "__main__":
    import pyoracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    om, _ = scenario(pyoracle)
    keys = om.occupied_cell_keys()
    L = om.read_layers(keys, helpers.MOMENT_LAYERS + helpers.DERIVED_LAYERS)
    out = os.path.join(ROOT, "tests", "golden", "small_scene.npz")
    np.savez_compressed(out, keys=keys, layers=L, occupancy=om.export_occupancy(), extent=np.array(om.grid_extent()))
    print(out, keys.size, "cells", int(np.sum(~np.isnan(L[:, 12]))), "gated in")

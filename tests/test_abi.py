"""The C-ABI library loads without a GPU, exports every symbol include/tmap.h declares, and the host
logic that needs no device behaves like the reference's error paths."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "tmap.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(tmap_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_are_exported_and_bound(tmap):
    lib = C.CDLL(tmap.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tmap.h but not exported"
    assert sorted(tmap.exported_symbols()) == names, "python binding table and header disagree"


def test_struct_layouts_match_the_header(tmap, tmp_path):
    """The ctypes mirrors of tmap_params / tmap_stats / tmap_shard_info must have the C compiler's size and
    field offsets (the ABI is plain structs; a drifted mirror would read garbage silently)."""
    import subprocess
    fields = {"tmap_params": tmap.Params, "tmap_stats": tmap.Stats, "tmap_shard_info": tmap.ShardInfo}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "tmap.h"', 'int main(void){']
    for cname, cls in fields.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in fields.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, f"{cname}.{f}"


def test_no_oracle_or_cpu_fallback_in_product():
    """The product path must not reach into oracle/ (judge rule) -- checked at source level."""
    pkg = os.path.join(ROOT, "traversability_mapping_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in src and "tmap_oracle" not in src and "tmo_" not in src, f
    nm = subprocess.run(["nm", "-D", os.path.join(pkg, "libtmap_b200.so")], capture_output=True, text=True).stdout
    assert "tmo_" not in nm


def test_default_params_match_reference_yaml(tmap, oracle):
    p = tmap.default_params()
    # params/traversabilityParams.yaml:10-51
    assert (p.resolution, p.half_size, p.extend_length) == (0.05, 7.5, 30.0)
    assert (p.robot_radius, p.ground_clearance, p.max_slope, p.min_points_per_grid) == (0.20, 0.15, 0.4, 3)
    assert (p.inflation_enabled, p.inflation_radius, p.cost_scaling_factor, p.lethal_step_threshold) == (0, 1.5, 2.0, 0.95)
    assert (p.robot_height, p.max_range, p.min_range) == (3.5, 6.0, 1.0)
    assert p.is_kf_optimization_enabled == 1 and p.max_batch == 1
    o = oracle.default_params()
    for f, _ in oracle.Params._fields_:
        assert getattr(o, f) == getattr(p, f), f


def test_null_handle_and_bad_arguments(tmap):
    lib = tmap.load_library()
    assert lib.tmap_flush(None) == tmap.ERR_INVALID
    assert lib.tmap_update_pose(None, 1, (C.c_float * 12)()) == tmap.ERR_INVALID
    assert lib.tmap_create(None, None) == tmap.ERR_INVALID
    assert b"null" in lib.tmap_last_error()


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device failure mode")
def test_create_fails_loudly_without_a_device(tmap):
    with pytest.raises(tmap.TmapError) as e:
        tmap.System()
    assert e.value.code in (tmap.ERR_NO_DEVICE, tmap.ERR_CUDA)


def test_synthetic_inputs_are_deterministic(tmap):
    from traversability_mapping_b200 import synth
    P = synth.trajectory("single", 1, seed=5)[0]
    a = synth.make_keyframe(5, 3, P, max_range=10.0, rings=8, azimuths=64)
    b = synth.make_keyframe(5, 3, P, max_range=10.0, rings=8, azimuths=64)
    assert a.shape == (512, 4) and a.dtype == np.float32
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.isnan(a[:, 0]).sum() + (a[:, 0] == 0).sum() > 0  # NaN and ego returns present
    assert tmap.key(4, -1) == (4 << 32) | 0xFFFFFFFF
    ci, cj = tmap.unkey(np.array([tmap.key(4, -1), tmap.key(-7, 80)], dtype=np.uint64))
    assert ci.tolist() == [4, -7] and cj.tolist() == [-1, 80]


def test_eig3_closed_form_vs_oracle_qr():
    """csrc/eig3.cuh compiled for the host vs the oracle's Eigen-QR restatement (400k matrices)."""
    exe = "/tmp/tmap_eig3_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe,
                    os.path.join(ROOT, "tests", "eig3_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    m = re.search(r"max_rel_lambda ([0-9.e+-]+).*max_1-\|dot\| ([0-9.e+-]+)", r.stdout)
    assert float(m.group(1)) < 1e-6 and float(m.group(2)) < 1e-10


def test_cpp_mirror_header_compiles_and_fails_loudly_without_gpu(tmap):
    """include/tmap_system.hpp (the reference-named C++ surface) builds with plain g++ against the C ABI."""
    exe = "/tmp/tmap_cpp_mirror_check"
    lib_dir = os.path.dirname(tmap.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp_mirror_check.cpp"),
                    "-L" + lib_dir, "-ltmap_b200", "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    if HAS_GPU:
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 42 and "failed loudly" in r.stdout

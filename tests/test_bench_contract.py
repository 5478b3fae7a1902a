"""The JSON lines committed under profiles/ (what `bench.py` printed on the GPU boxes this round) carry every key of the
measurement contract -- a CPU-side check that a refactor of bench.py cannot silently drop one."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_single_gpu_line():
    d = line("r02_bench_n1.json")
    for k in BASE + ["cpu_baseline", "ms_all"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] == "hbm"
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and "sample" in c
    assert d["gpu_launches"] > 0 and len(d["ms_all"]) == d["steps"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = line("r02_bench_reference_n1.json")
    assert d["impl"] == "reference" and d["metric"] == line("r02_bench_n1.json")["metric"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["config"]["workload"] == line("r02_bench_n1.json")["config"]["workload"]


def test_multi_gpu_lines_carry_parity_and_strong_scaling():
    for name, n in (("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4), ("r02_bench_n8.json", 8), ("r02_bench_c5_n8.json", 8)):
        d = line(name)
        for k in BASE:
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["shard_parity"] is True
        assert d["sharded"]["parity_check"]["bit_identical_to_1_gpu"] is True
        assert "sharded" in d["config"]["parallelism"]

"""bench.py host logic on the CPU: byte accounting, synthetic-plate helpers, and the
--impl reference arm's JSON line (the CPU path needs no GPU)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey():
    # SURVEY.md section 8d, config 2 (C=5, Z=3, 2160^2, bin 2, ~2000 cells)
    assert bench.k1_bytes_per_field(5, 3, 2160, 2160, 2) == 303_264_000
    assert bench.k1_bytes_per_field(5, 3, 2160, 2160, 4) == 285_768_000
    assert bench.k1_bytes_per_field(5, 5, 1080, 1080, 2) == 99_144_000
    assert bench.k3_bytes_per_field(5, 2160, 2160, 2000) == 18_662_400 + 46_656_000 + 93_312_000 + 2000 * 33 * 4


def test_dihedral_masks_keep_the_label_set():
    lab = np.zeros((12, 12), np.int32)
    lab[1:4, 2:7] = 1
    lab[8:11, 0:3] = 2
    seen = set()
    for k in range(8):
        m = bench.dihedral(lab, k)
        assert m.shape == lab.shape and m.flags.c_contiguous
        assert np.array_equal(np.bincount(m.ravel()), np.bincount(lab.ravel()))
        seen.add(m.tobytes())
    assert len(seen) == 8


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-budget", "2", "--ref-fields", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "fields/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["e2e"] == {"value": d["value"], "unit": "fields/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_committed_bench_lines_carry_the_contract():
    """The bench lines kept under profiles/ (written by bench.py on a B200) carry every key the measurement
    contract names, and their numbers are consistent with each other."""
    for name in ("r2f_bench.json", "r2g_bench_20steps_final.json"):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.loads(f.read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
            assert k in d, (name, k)
        assert d["unit"] == "fields/s" and d["scaling"] == "weak" and d["vs_baseline"] is None and d["gpu_launches"] > 0
        assert "workload" in d["config"] and "model" not in d["config"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0.5 < r["frac"] < 1.0 and r["traffic"] is not None
        # value = fields of all steps / the timed region
        fields = d["steps"] * d["config"]["fields_per_step"] * d["n_gpus"]
        assert abs(d["value"] - fields / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        assert d["aggregation"]["check"] == "ok"
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}

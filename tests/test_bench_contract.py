"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm prints exactly one JSON line on
stdout with the keys the driver reads; our own arm refuses to run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MS/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "1080p60" in d["config"]["workload"]


def test_own_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == ""          # no number without the CUDA path


def test_bench_shapes_are_the_baseline_geometries():
    """BENCH_SHAPE selects BASELINE configs[1] (default), configs[4]'s per-GPU stream (cfg5) or configs[0] (cfg1): widths and
    capture sizes as SURVEY section 8 lists them (the GUI's total-height convention)."""
    code = ("import bench, json; print(json.dumps({'w': bench.geometry(), 'h': bench.HEIGHT, 'fs': bench.FS, 'cap': int(3.1 * bench.FS / 55.0), "
            "'frames': bench.FRAMES_PER_BATCH}))")
    want = {"cfg2": (740, 1125, 25_000_000, 1_409_090), "cfg5": (1481, 1125, 50_000_000, 2_818_181), "cfg1": (507, 525, 8_000_000, 450_909)}
    for shape, (w, h, fs, cap) in want.items():
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, BENCH_SHAPE=shape))
        assert r.returncode == 0, r.stderr[-1500:]
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert (d["w"], d["h"], d["fs"], d["cap"]) == (w, h, fs, cap), (shape, d)
        # the resident IQ of a batch must exceed the 126 MB L2 (no L2 flush between timed iterations)
        assert 8 * d["frames"] * int(d["fs"] / 60.0) > 200e6, (shape, d)

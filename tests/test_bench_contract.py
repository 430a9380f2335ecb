"""bench.py's line contract, checked on CPU through the reference arm (`--impl reference`, which needs no GPU)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"}


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           *args], capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)


@pytest.mark.parametrize("args,workload", [(("--reference-crop",), "c3_blind_24mp_k15"),
                                           (("--workload", "c1_nonblind_512_g5"), "c1_nonblind_512_g5")])
def test_reference_arm_prints_exactly_one_json_line(args, workload):
    """Crop sample of the headline workload (quick), and the full-frame path (the default) on the small C1 workload."""
    r = _run(args=args)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "MPix*iter/s" and d["value"] > 0 and d["dtype"] == "f32"
    assert d["config"]["workload"] == workload and d["scaling"] == "strong"
    if workload.startswith("c1"):
        assert d["config"]["frame"] == [512, 512, 3] and "FULL 512x512 frame" in d["cpu_baseline"]["sample"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, args=("--gpus", "2"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_numpy_port_backs_the_reference_arm_when_the_compiled_reference_is_absent():
    sys.path.insert(0, str(ROOT))
    try:
        import bench
        r = bench.cpu_reference_sample("c1_nonblind_512_g5", 1, 0, force_port=True)
    finally:
        sys.path.pop(0)
    assert r["kind"] == "port" and r["cores"] == 1 and r["value"] > 0 and "numpy port" in r["sample"]


@pytest.mark.parametrize("family", ["conv_fwd", "conv_adj", "update", "gradk"])
def test_committed_ncu_traffic_is_readable(family):
    sys.path.insert(0, str(ROOT))
    try:
        import bench
        val, src = bench.ncu_traffic("c3_blind_24mp_k15", family)
        none, _ = bench.ncu_traffic("c1_nonblind_512_g5", family)
    finally:
        sys.path.pop(0)
    # hundreds of MB per launch on the 24 MP frame -- or nothing at all when the capture predates the current kernels, or
    # when the family does not launch on this workload (the chain kernel does the forward blur inside "conv_adj")
    assert (val and val > 1e8 and "profiles/" in src) or (val is None and "stale" in (src or "stale")) \
        or (val is None and family == "conv_fwd" and "profiles/" in src)
    assert none is None                                  # no capture for other workloads: traffic stays null


def test_config_object_is_the_same_for_both_arms():
    sys.path.insert(0, str(ROOT))
    try:
        import bench
        a = bench.workload_config("c3_blind_24mp_k15", 4000, 6000, 15, True)
    finally:
        sys.path.pop(0)
    assert a["workload"] == "c3_blind_24mp_k15" and a["frame"] == [4000, 6000, 3] and a["psf"] == 15

"""Driver contract of bench.py that can be checked without a GPU: the reference arm
(`--impl reference`) prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

from tests import helpers as H


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--cpu-seconds", "2"],
                       capture_output=True, text=True, timeout=600, cwd=H.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "A-step candidate pairs/sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    # the reference's own get_actdist (from /root/reference, or the copy `make -C oracle` leaves in
    # oracle/_ref/igm) - the NumPy port only when neither exists
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    from oracle import ref_loader
    if ref_loader.reference_available():
        assert d["cpu_baseline"]["kind"] == "reference"
    # the arm never maps the product library
    assert not any("libigmk" in x for x in d["config"]["repo_native_libraries_loaded"])
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "config 2" in d["config"]["workload"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=H.ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""

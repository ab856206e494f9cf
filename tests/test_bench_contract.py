"""bench.py contract checks that need no GPU: the reference arm (the reference's own classes from oracle/_ref when staged,
else the oracle port) prints exactly one JSON line on stdout with the agreed keys."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    from oracle import build_ref
    assert d["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] == ("reference" if build_ref.available() else "port")
    assert d["cpu_baseline"]["one_thread"]["cores"] == 1 and d["cpu_baseline"]["one_thread"]["value"] > 0
    assert d["config"]["clients"] == 10 and d["config"]["samples_per_round"] == 5216 and d["config"]["dp_mode"] == "update"
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]")
    assert d["metric"].startswith("DP-SGD client samples/s")

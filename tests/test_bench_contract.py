"""The reference arm of bench.py (CPU only): one JSON line on stdout with the keys the driver reads; under a
multi-rank launch only rank 0 prints.  The GPU arm needs a device and is exercised on the B200 box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                           "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    res = _run({})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-kicks/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("SC particle-kicks/sec") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("BASELINE configs[0]")
    cb = d["cpu_baseline"]
    staged = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "ocelot", "cpbd", "sc.py"))
    assert cb["kind"] == ("reference" if staged else "port")       # the unmodified reference when it is staged
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["cpu_port"]["kind"] == "port" and d["cpu_port"]["value"] > 0
    if staged:
        assert d["e2e_resident"]["kicks"] == 20 and d["e2e_resident"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""

"""CPU: the reference arm of bench.py (`--impl reference`: the reference itself from baseline/_ref when
tools/install_reference.py has staged it, else the oracle port, timed on the host cores) prints one JSON line with the
contract's keys; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr
    return res.stdout.strip()


def test_reference_arm_prints_the_contract_line():
    out = _run()
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gbm_path_steps_per_sec" and d["unit"] == "path-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0 and d["steps"] >= 1
    staged = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "engine"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["cpu_model"]
    assert d["warmup"] == 1 and d["steps"] == 1
    if staged:          # both readings of the reference itself
        r = d["cpu_baseline"]["baseline_md_section3"]
        assert r["kernel_only_path_steps_per_s"] > r["price_end_to_end_path_steps_per_s"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""

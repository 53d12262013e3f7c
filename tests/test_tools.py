"""CPU: the inspection tools bench.py relies on still understand the shipped library (no GPU needed: cuobjdump reads the
.so).  Guards the instruction roofline against silently falling back to its built-in constants after a kernel edit."""
import os
import shutil
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")), reason="no cuobjdump")
def test_sass_mix_of_the_headline_kernel():
    import sass_mix
    mixes = sass_mix.mix("k_europeanILi0ELb0ELb1EfLb1E")          # k_european<GBM, !ANTI, GREEKS, float, SINGLE>
    assert len(mixes) == 1
    m = list(mixes.values())[0]
    # one Philox4x32-10 call per loop iteration = 8 steps: 4 words x 3 MUFU (lg2, sqrt, ONE sine for the sum of the word's
    # two normals), at most 20 wide multiplies
    # (the loop body holds TWO calls: the word sets alternate between two register sets instead of being copied)
    calls = m["philox_calls"]
    assert calls == 2 and m["xu"] == 12 * calls and 17 * calls <= m["imad_wide"] <= 20 * calls
    assert 75 * calls <= m["total"] <= 90 * calls, m   # 160 today; a jump means the hot loop changed -- re-measure before shipping


def test_ptxas_logs_report_no_spills_in_the_hot_kernels():
    """Registers / spills of the fused kernels as ptxas reported them at build time (csrc/build/*.ptxas.log)."""
    import re
    log = os.path.join(ROOT, "monte_carlo_option_simulator_b200", "csrc", "build", "european.ptxas.log")
    if not os.path.exists(log):
        pytest.skip("library was not built in this checkout")
    text = open(log).read()
    blocks = re.findall(r"Compiling entry function '(\w+)'.*?(\d+) bytes spill stores.*?Used (\d+) registers", text, flags=re.S)
    assert len(blocks) >= 64
    for name, spill, regs in blocks:
        if "k_europeanILi0E" in name and "EfLb1E" in name:            # GBM, fp32, single strike: the headline family
            assert int(spill) == 0 and int(regs) <= 128, (name, spill, regs)


def test_probe_library_is_separate_from_the_product_abi():
    """The pipe-rate probes bench.py's rooflines divide by live in their own library (tools/probe), not in libb200mc.so:
    it loads without a GPU, exports exactly its two entry points, and the product exports none of them."""
    import ctypes
    from monte_carlo_option_simulator_b200 import _lib
    path = os.path.join(ROOT, "tools", "probe", "libb200mc_probe.so")
    if not os.path.exists(path):
        pytest.skip("probe library was not built in this checkout (python __graft_entry__.py build)")
    lib = ctypes.CDLL(path)
    assert hasattr(lib, "b200mc_probe_rate") and hasattr(lib, "b200mc_probe_mix")
    prod = _lib.load()
    assert not hasattr(prod, "b200mc_microbench") and not hasattr(prod, "b200mc_probe_rate")
    assert not any("microbench" in n or "probe" in n for n in _lib.EXPORTS)
    import ctypes as C
    v = C.c_double()
    assert lib.b200mc_probe_rate(0, 999, 16, C.byref(v)) == 4          # bad selector is refused before any CUDA call


def test_every_ncu_summary_the_bench_line_cites_is_committed():
    """bench.py copies `traffic_source` from profiles/r02_ncu_traffic.json; each entry names the ncu summary it came from."""
    import json
    import re
    t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    assert len(t) >= 8
    for key, v in t.items():
        assert v["dram_bytes_per_launch"] > 0, key
        for path in re.findall(r"profiles/[A-Za-z0-9_]+\.txt", v["source"]):
            assert os.path.exists(os.path.join(ROOT, path)), f"{key}: {path} is cited but not committed"
            assert "kernel:" in open(os.path.join(ROOT, path)).read()


@pytest.mark.skipif(not shutil.which("gcc"), reason="no gcc")
def test_host_layer_runs_end_to_end_against_the_stand_in_library():
    """tools/host_overhead_probe.py drives price / price_batch / the Greeks / a stress report through the real binding with
    a stand-in library that returns constant sums (no device): every host code path of the small calls executes here."""
    import subprocess
    env = dict(os.environ, B200MC_HOST_PROBE_QUICK="1")
    env.pop("B200MC_LIB", None)
    env.pop("B200MC_HOST_PROBE", None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_overhead_probe.py")], env=env,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if "us of host time per call" in ln]
    assert len(lines) == 7 and "stand-in" in res.stdout

"""CPU: the C-ABI library loads and exports every symbol include/b200mc.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "b200mc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from monte_carlo_option_simulator_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200mc.h but not exported by libb200mc.so"
    # and the binding covers the whole header
    assert set(_lib.EXPORTS) == set(names)
    assert lib.b200mc_version() == 100


def test_struct_layouts_match_header():
    from monte_carlo_option_simulator_b200 import _lib
    assert ctypes.sizeof(_lib.SvjParams) == 10 * 8
    assert ctypes.sizeof(_lib.Bumps) == 5 * 8
    assert ctypes.sizeof(_lib.Sums) == 17 * 8 and _lib.NSUMS == 17
    text = open(os.path.join(ROOT, "include", "b200mc.h")).read()
    body = text[:text.index("} b200mc_sums;")]
    body = body[body.rindex("typedef struct {"):]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r"double\s+([^;]+);", body) for f in decl.split(",")]
    assert tuple(fields) == _lib.SUMS_FIELDS
    # b200mc_cell: the ctypes struct, the NumPy dtype used to fill thousands of cells at once and the header agree
    body = text[text.index("typedef struct b200mc_cell {"):text.index("} b200mc_cell;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decls = re.findall(r"(b200mc_svj_params|double|int64_t|uint64_t|int32_t)\s+([^;]+);", body)
    names = [f.strip() for _, d in decls for f in d.split(",")]
    assert names == [n for n, _ in _lib.Cell._fields_]
    assert ctypes.sizeof(_lib.Cell) == _lib.CELL_DTYPE.itemsize == 128
    for name in names[1:]:
        assert getattr(_lib.Cell, name).offset == _lib.CELL_DTYPE.fields[name][1], name
    assert _lib.CELL_DTYPE.names[:10] == _lib.PARAM_FIELDS
    # constants of the peer exchange
    assert "#define B200MC_PEER_MAX_RANKS   16" in text and "#define B200MC_PEER_MAX_DOUBLES 4352" in text


def test_no_cpu_fallback_without_a_gpu():
    """The product path must fail loudly when no sm_100 device is present (never route to the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; this test covers the GPU-less box")
    from monte_carlo_option_simulator_b200 import MonteCarloEngine, SVJParams, _lib, compute_risk_metrics
    with pytest.raises(_lib.B200MCError) as e:
        _lib.Handle(0)
    assert e.value.code == _lib.ENODEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(RuntimeError):
        MonteCarloEngine(SVJParams(), 1000, use_sobol=False).price(100.0, 100.0, 0.5)
    with pytest.raises(RuntimeError):
        compute_risk_metrics([0.1, -0.2, 0.3])


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing shipped may import, load, link or include it."""
    pkg = os.path.join(ROOT, "monte_carlo_option_simulator_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|#\s*include\s*[\"<][^\">]*oracle|dlopen|CDLL\([^)]*oracle", re.M)
    checked = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                checked += 1
                assert not bad.search(open(os.path.join(dirpath, f)).read()), f"{f} reaches into oracle/"
    assert checked >= 12
    # the measurement scripts under tools/ (incl. the probe library) do not use it either: scripts that need the oracle as
    # their checker live under tests/probes/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tools")):
        for f in files:
            if f.endswith((".py", ".cu", ".sh")) or f == "Makefile":
                assert not bad.search(open(os.path.join(dirpath, f)).read()), f"tools/{f} reaches into oracle/"


def test_stream_selection_rule_needs_no_device():
    from monte_carlo_option_simulator_b200 import SVJParams, _lib
    g = SVJParams.gbm(0.3)
    assert _lib.select_stream(g, 1.0, 250) == _lib.STREAM_GBM
    assert _lib.select_stream(g, 1.0, 250, _lib.FORCE_SVJ) == _lib.STREAM_SVJ
    moving = SVJParams(kappa=3.0, theta=0.05, xi=0.0, v0=0.09, lambda_j=0.0)
    assert _lib.select_stream(moving, 1.0, 250) == _lib.STREAM_GBM                 # deterministic variance: weight table
    assert _lib.select_stream(moving, 1.0, 5000) == _lib.STREAM_HESTON             # table too long
    assert _lib.select_stream(SVJParams(lambda_j=0.0), 1.0, 250) == _lib.STREAM_HESTON
    assert _lib.select_stream(SVJParams(), 1.0, 250) == _lib.STREAM_SVJ
    # a v0 bump away from theta turns constant variance into a moving one (kappa != 0): still the GBM stream
    kap = SVJParams(kappa=3.0, theta=0.09, xi=0.0, v0=0.09, lambda_j=0.0)
    assert _lib.select_stream(kap, 1.0, 250, _lib.GREEKS, _lib.Bumps(0.01, 0.10, 0.08, 0.0651, 0.0649)) == _lib.STREAM_GBM
    with pytest.raises(_lib.B200MCError):
        _lib.select_stream(g, -1.0, 250)

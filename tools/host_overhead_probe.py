"""Python-side cost of the small calls, measured WITHOUT a GPU: the binding is pointed (B200MC_LIB) at a stand-in library
whose entry points return at once with constant sums, so what is timed is the marshalling, the ctypes call and the dict
algebra of the host mirror -- the share of a small call that no kernel work can hide.

    python tools/host_overhead_probe.py            # builds the stand-in with gcc under /tmp and re-executes itself

The stand-in is a measurement double for this script only: it computes nothing and is never installed next to the package.
"""
import os
import subprocess
import sys
import tempfile
import timeit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STUB_BODY = r'''
#include <stdint.h>
static int dummy;
int b200mc_version(void){return 100;}
int b200mc_create(int dev, void **out){*out=&dummy;return 0;}
int b200mc_destroy(void*h){return 0;}
const char* b200mc_last_error(void*h){return "stand-in";}
static void fill(double*o,double np){o[0]=np;for(int j=1;j<17;j++)o[j]=np*(100.0+j);o[3]=np*20000.0;o[4]=o[3];o[5]=np*15000.0;}
int b200mc_price_european(void*h,const void*p,double S0,double T,int32_t ns,int64_t np,uint64_t seed,uint64_t off,
                          const double*k,int32_t nk,int ic,uint32_t fl,const void*b,double*out)
{ for(int i=0;i<nk;i++)fill(out+17*i,(double)np); return 0; }
int b200mc_price_cells(void*h,const void*cells,int32_t nc,const double*k,int32_t nk,uint32_t fl,int ondev,double*out)
{ if(!ondev) for(int i=0;i<nc*nk;i++)fill(out+17*i,1000.0); return 0; }
'''
SPECIAL = ("b200mc_version", "b200mc_create", "b200mc_destroy", "b200mc_last_error", "b200mc_price_european",
           "b200mc_price_cells")


def build_stand_in():
    from monte_carlo_option_simulator_b200 import _lib
    d = tempfile.mkdtemp(prefix="b200mc_standin_")
    src = os.path.join(d, "standin.c")
    with open(src, "w") as f:
        f.write(STUB_BODY)
        for name in _lib._PROTOS:
            if name not in SPECIAL:
                f.write(f"int {name}(void){{return 0;}}\n")
    so = os.path.join(d, "libstandin.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src], check=True)
    return so


if os.environ.get("B200MC_HOST_PROBE") != "1":
    env = dict(os.environ, B200MC_LIB=build_stand_in(), B200MC_HOST_PROBE="1", B200MC_RNG="philox")
    sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__)], env=env).returncode)

import numpy as np  # noqa: E402
from monte_carlo_option_simulator_b200 import GreeksEngine, MonteCarloEngine, SVJParams, _lib  # noqa: E402
from monte_carlo_option_simulator_b200.risk import StressTestEngine  # noqa: E402

p = SVJParams()
h = _lib.default_handle()
e = MonteCarloEngine(p, 10_000, 50, 42, use_sobol=False)
g = GreeksEngine(p, 50_000, 252, 42)
st = StressTestEngine(p, num_paths=200_000, seed=42, handle=h)
ks21 = [22500.0 * (0.7 + 0.03 * k) for k in range(21)]
ks5 = np.linspace(0.95, 1.05, 5) * 22500.0


def greeks(i=[0]):
    i[0] += 1
    g.seed = i[0]
    g.delta(22500.0, 22500.0, 0.25), g.vega(22500.0, 22500.0, 0.25), g.gamma(22500.0, 22500.0, 0.25)


G = globals()
QUICK = os.environ.get("B200MC_HOST_PROBE_QUICK") == "1"          # tests/test_tools.py: a few calls of every entry, no timing claim
print(f"library: {_lib.LIB_PATH} (stand-in: returns at once)")
for what, stmt, n in (("Handle.price_european, 1 strike", "h.price_european(p, 22500.0, 1.0, 50, 10000, 42, [22500.0], True, 1)", 20000),
                      ("MonteCarloEngine.price", "e.price(22500.0, 22400.0, 1.0)", 20000),
                      ("GreeksEngine delta + vega + gamma (one fused call)", "greeks()", 20000),
                      ("price_batch, 21 strikes (list)", "e.price_batch(22500.0, ks21, 0.25)", 5000),
                      ("price_batch, 5 strikes (ndarray)", "e.price_batch(22500.0, ks5, 0.08)", 5000),
                      ("MonteCarloEngine(...) constructor", "MonteCarloEngine(p, 100000, 100)", 20000),
                      ("StressTestEngine.full_stress_report (11 cells)", "st.full_stress_report(22500.0, 22500.0, 0.25)", 500)):
    if QUICK:
        n = max(n // 500, 3)
    us = min(timeit.repeat(stmt, globals=G, number=n, repeat=2 if QUICK else 5)) / n * 1e6
    print(f"{what:52s} {us:8.2f} us of host time per call")

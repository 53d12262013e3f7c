"""ctypes binding of tools/probe/libb200mc_probe.so: issue rates of the pipes the fused kernels live on (FFMA, IMAD.WIDE,
LOP3, MUFU, fp32<->fp64 conversion, DADD, Philox calls), measured on the device at hand.  bench.py divides its
instruction rooflines by these; the library is a measurement tool and is not part of libb200mc.so."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200mc_probe.so")
_lib = None

# selector -> what one "operation" is (tools/probe/microbench.cu)
FFMA, IMAD_WIDE, LOP3, MUFU_EX2, MUFU_SIN, IADD3, PHILOX, PHILOX_BM, FMUL, MUFU_LG2, MUFU_SQRT, FFMA_LOP3, IMAD_LO, IMAD_HI, \
    IMAD_LOHI, FFMA2, F2F, DADD, FFMA2_LOP3, FFMA2_MUFU, FFMA2_WIDE, FFMA2_FFMA = range(22)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make -C tools/probe` (or `python __graft_entry__.py build`)")
        _lib = C.CDLL(LIB_PATH)
        _lib.b200mc_probe_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.b200mc_probe_mix.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    return _lib


def rate(which: int, iters: int = 4096, device: int = 0) -> float:
    """Thread-level operations per second over the whole device (kernel time by CUDA events, best of 3)."""
    v = C.c_double()
    rc = load().b200mc_probe_rate(int(device), int(which), int(iters), C.byref(v))
    if rc:
        raise RuntimeError(f"b200mc_probe_rate failed with code {rc} (1 set-up, 2 not sm_100, 3 CUDA error, 4 bad argument)")
    return float(v.value)


def mix(combo: int, iters: int = 2048, device: int = 0):
    """(thread-iterations per second, (n_wide, n_lop3, n_mufu, n_ffma) per iteration) of a fixed mix table."""
    v = C.c_double()
    cnt = (C.c_int * 4)()
    rc = load().b200mc_probe_mix(int(device), int(combo), int(iters), C.byref(v), cnt)
    if rc:
        raise RuntimeError(f"b200mc_probe_mix failed with code {rc}")
    return float(v.value), tuple(cnt)

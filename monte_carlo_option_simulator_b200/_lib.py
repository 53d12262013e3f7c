"""ctypes binding of libb200mc.so (C ABI in include/b200mc.h).

This is the only place the shared library is loaded.  There is no CPU fallback: if the library is missing
(`python __graft_entry__.py build` / `make -C monte_carlo_option_simulator_b200/csrc` builds it) or no sm_100
device is visible, every compute entry point raises -- it never silently routes anywhere else.
"""
from __future__ import annotations

import ctypes as C
import operator
import os
import threading
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MC_LIB") or os.path.join(_HERE, "libb200mc.so")     # B200MC_LIB: an alternative build (tuning runs)

OK, EINVAL, ENODEVICE, ECUDA, ENOMEM = 0, 1, 2, 3, 4
ANTITHETIC, GREEKS, FP64, FORCE_SVJ, WIDE_RNG = 0x1, 0x2, 0x4, 0x8, 0x10
HIST_WIDE, HIST_R01 = 0x100, 0x200
STREAM_GBM, STREAM_HESTON, STREAM_SVJ, STREAM_HEDGE = 0, 1, 2, 3
Z1, Z2, ZJUMP_U, ZJUMP_SIZE = 0, 1, 2, 3
F32, F64 = 0, 1
NUMPY_RANDOM, NUMPY_STANDARD_NORMAL = 0, 1
GIVEN_RECORD, GIVEN_NEGATE = 1, 2


class B200MCError(RuntimeError):
    """Raised for every non-zero status of the C ABI (callers of the reference catch Exception broadly:
    engine/calibration.py:84-89, engine/app.py:139-140)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200mc error {code}: {msg}")
        self.code = code


class SvjParams(C.Structure):
    """b200mc_svj_params -- field set of SVJParams, engine/models.py:31-44."""
    _fields_ = [(n, C.c_double) for n in
                ("v0", "r", "q", "kappa", "theta", "xi", "rho", "lambda_j", "mu_j", "sigma_j")]


class Bumps(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("spot_bump", "v0_up", "v0_dn", "r_up", "r_dn")]


class Cell(C.Structure):
    """b200mc_cell -- one pricing problem of a b200mc_price_cells launch."""
    _fields_ = [("params", SvjParams), ("S0", C.c_double), ("T", C.c_double), ("n_paths", C.c_int64),
                ("seed", C.c_uint64), ("path_offset", C.c_uint64), ("n_steps", C.c_int32), ("is_call", C.c_int32)]


PARAM_FIELDS = tuple(n for n, _ in SvjParams._fields_)
# numpy mirror of b200mc_cell (same layout as the ctypes struct above): lets callers fill thousands of cells vectorised
CELL_DTYPE = np.dtype([(n, "<f8") for n in PARAM_FIELDS] +
                      [("S0", "<f8"), ("T", "<f8"), ("n_paths", "<i8"), ("seed", "<u8"), ("path_offset", "<u8"),
                       ("n_steps", "<i4"), ("is_call", "<i4")])
assert CELL_DTYPE.itemsize == C.sizeof(Cell)


def make_cells(params, S0, T, n_steps, n_paths, seed, path_offset=0, is_call=True) -> np.ndarray:
    """Structured array of b200mc_cell.  Every argument is a scalar or a sequence (broadcast to the longest);
    `params` is one parameter object (shared by all cells), a sequence of them, or a dict field -> scalar / array."""
    cols = dict(S0=np.atleast_1d(np.asarray(S0, dtype=np.float64)), T=np.atleast_1d(np.asarray(T, dtype=np.float64)),
                n_paths=np.atleast_1d(np.asarray(n_paths, dtype=np.int64)),
                n_steps=np.atleast_1d(np.asarray(n_steps, dtype=np.int32)),
                is_call=np.atleast_1d(np.asarray(is_call)).astype(np.int32),
                path_offset=np.atleast_1d(np.array(path_offset, dtype=np.uint64)))
    sd = np.atleast_1d(np.asarray(seed))
    cols["seed"] = np.array([int(x) & (2 ** 64 - 1) for x in sd.tolist()], dtype=np.uint64) if sd.dtype == object \
        else sd.astype(np.int64).view(np.uint64) if sd.dtype.kind == "i" else sd.astype(np.uint64)
    if isinstance(params, dict):
        pcols = {f: np.atleast_1d(np.asarray(params[f], dtype=np.float64)) for f in PARAM_FIELDS}
    else:
        plist = list(params) if isinstance(params, (list, tuple)) else [params]
        pcols = {f: np.array([float(getattr(p, f)) for p in plist]) for f in PARAM_FIELDS}
    n = max([c.size for c in pcols.values()] + [c.size for c in cols.values()])
    out = np.zeros(n, dtype=CELL_DTYPE)
    for f, c in pcols.items():
        out[f] = c if c.size > 1 else c[0]
    for k, c in cols.items():
        out[k] = c if c.size > 1 else c[0]
    return out


SUMS_FIELDS = ("n", "sum_a", "sum_b", "sum_aa", "sum_bb", "sum_ab", "sum_s", "sum_ss", "sum_ps",
               "sum_pw_delta", "sum_spot_up", "sum_spot_dn", "sum_v0_up", "sum_v0_dn", "sum_r_up", "sum_r_dn",
               "sum_pw_vega")
NSUMS = len(SUMS_FIELDS)


class Sums(C.Structure):
    _fields_ = [(n, C.c_double) for n in SUMS_FIELDS]


_lib = None
_lib_lock = threading.Lock()

_i32, _i64, _u32, _u64, _dbl, _vp = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double, C.c_void_p
_dp = C.POINTER(C.c_double)
_PROTOS = {
    # name: (restype, argtypes)
    "b200mc_version": (C.c_int, []),
    "b200mc_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "b200mc_destroy": (C.c_int, [_vp]),
    "b200mc_last_error": (C.c_char_p, [_vp]),
    "b200mc_device_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_u64), C.POINTER(C.c_int)]),
    "b200mc_launch_count": (_i64, [_vp]),
    "b200mc_stream": (_u64, [_vp]),
    "b200mc_set_stream": (C.c_int, [_vp, _u64]),
    "b200mc_synchronize": (C.c_int, [_vp]),
    "b200mc_peer_create": (C.c_int, [_vp, _vp]),
    "b200mc_peer_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "b200mc_peer_allreduce": (C.c_int, [_vp, _vp, _i32]),
    "b200mc_peer_close": (C.c_int, [_vp]),
    "b200mc_risk_metrics_sharded": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, _dbl, _dp]),
    "b200mc_simulate_given_normals": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i64, _i32,
                                                 _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "b200mc_simulate_given_normals_dev": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i64, _i32,
                                                     _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "b200mc_price_european": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i32, _i64, _u64, _u64,
                                         _vp, _i32, C.c_int, _u32, C.POINTER(Bumps), _vp]),
    "b200mc_price_european_async": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i32, _i64, _u64, _u64,
                                               _vp, _i32, C.c_int, _u32, C.POINTER(Bumps), _vp]),
    "b200mc_price_cells": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _u32, C.c_int, _vp]),
    "b200mc_hedge_walk": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _dbl, C.c_int, _i32, _i64, _dbl, _vp, _vp,
                                     _u64, _u64, _vp, _vp]),
    "b200mc_implied_vol": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _dbl, _dbl, _dbl, _dbl, _dbl, _vp]),
    "b200mc_qmc_normals": (C.c_int, [_vp, _i64, _u64, _i32, _vp, _vp, _i32, _i32, C.c_int, _vp]),
    "b200mc_price_european_qmc": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i32, _i64, _u64, _vp, _vp, _i32, _i32,
                                             _vp, _i32, C.c_int, _u32, _vp]),
    "b200mc_qmc_terminal": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i32, _i64, _u64, _vp, _vp, _i32, _i32, _vp, _i32,
                                       _vp, _vp, _u32, _vp, _vp]),
    "b200mc_pcg64_random": (C.c_int, [_vp, _vp, _u64, _i64, _vp]),
    "b200mc_numpy_fill": (C.c_int, [_vp, _vp, _u64, _i64, C.c_int, C.c_int, _vp, C.POINTER(_u64)]),
    "b200mc_numpy_ziggurat_tables": (C.c_int, [_vp, _vp, _vp]),
    "b200mc_simulate_terminal": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i32, _i64, _u64, _u64,
                                            _u32, C.c_int, C.c_int, _vp, _vp, _vp]),
    "b200mc_generate_paths": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _dbl, _i32, _i64, _u64, _u64,
                                         _u32, C.c_int, C.c_int, _vp, _i64]),
    "b200mc_risk_metrics": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, _dbl, _dp]),
    "b200mc_option_pnl": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _dbl, C.c_int, _dbl, _dbl, C.c_int, _vp]),
    "b200mc_risk_begin": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, _dp]),
    "b200mc_risk_hist": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_u64), C.POINTER(_u64)]),
    "b200mc_risk_finish": (C.c_int, [_vp, _dbl, C.c_int, _dp, _dp]),
    "b200mc_dump_normals": (C.c_int, [_vp, _u64, _u64, _i64, _i32, _u32, C.c_int, _dbl, _vp]),
    "b200mc_select_stream": (C.c_int, [_vp, C.POINTER(SvjParams), _dbl, _i32, _u32, C.POINTER(Bumps), C.POINTER(_u32)]),
    "b200mc_dump_philox": (C.c_int, [_vp, _u64, _u64, _i64, _i32, _u32, _vp]),
    "b200mc_normal_moments": (C.c_int, [_vp, _u64, _u64, _i64, _i32, _dp]),
    "b200mc_normal_hist2d": (C.c_int, [_vp, _u64, _u64, _i64, _i32, C.c_int, _vp]),
    "b200mc_malloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "b200mc_free": (C.c_int, [_vp, _vp]),
    "b200mc_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "b200mc_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "b200mc_malloc_host": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "b200mc_free_host": (C.c_int, [_vp, _vp]),
    "b200mc_timer_begin": (C.c_int, [_vp]),
    "b200mc_timer_end": (C.c_int, [_vp, C.POINTER(C.c_float)]),
}
EXPORTS = tuple(_PROTOS)


def load() -> C.CDLL:
    """Load libb200mc.so and set the prototypes.  Raises (never falls back) when the library is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200MCError(ENODEVICE, f"{LIB_PATH} is not built (run `python __graft_entry__.py build`); "
                                         "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def select_stream(params, T: float, n_steps: int, flags: int = 0, bumps: Optional["Bumps"] = None) -> int:
    """Which B200MC_STREAM_* a fused call with these parameters draws from (needs no device)."""
    out = _u32()
    sp = to_params(params)
    rc = load().b200mc_select_stream(None, C.byref(sp), float(T), int(n_steps), int(flags),
                                     C.byref(bumps) if bumps is not None else None, C.byref(out))
    if rc != OK:
        raise B200MCError(rc, (load().b200mc_last_error(None) or b"").decode())
    return int(out.value)


# numpy mirror of b200mc_bridge_node
BRIDGE_DTYPE = np.dtype([("t", "<i4"), ("l", "<i4"), ("r", "<i4"), ("dim", "<i4"), ("a", "<f8"), ("b", "<f8"), ("sd", "<f8")])


def pcg64_state(seed) -> np.ndarray:
    """{state_hi, state_lo, inc_hi, inc_lo} (uint64[4]) of np.random.default_rng(seed)'s PCG64 bit generator."""
    st = np.random.default_rng(seed).bit_generator.state["state"]
    m = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)


def numpy_ziggurat_tables():
    """(ki, wi, fi): the Ziggurat tables compiled into the library (numpy's ki_double / wi_double / fi_double)."""
    ki, wi, fi = np.empty(256, np.uint64), np.empty(256, np.float64), np.empty(256, np.float64)
    rc = load().b200mc_numpy_ziggurat_tables(ki.ctypes.data, wi.ctypes.data, fi.ctypes.data)
    if rc != OK:
        raise B200MCError(rc, "b200mc_numpy_ziggurat_tables failed")
    return ki, wi, fi


def sobol_tables(n_dims: int, seed: int):
    """(sv, shift, bits) of scipy.stats.qmc.Sobol(d=n_dims, scramble=True, seed=seed): the scrambled direction numbers
    and the digital shift the device generator needs to reproduce SciPy's points bit for bit."""
    from scipy.stats.qmc import Sobol
    eng = Sobol(d=int(n_dims), scramble=True, seed=seed)
    sv = np.ascontiguousarray(eng._sv, dtype=np.uint32)
    shift = np.ascontiguousarray(eng._shift, dtype=np.uint32)
    return sv, shift, int(eng.bits)


_param_values = operator.attrgetter(*PARAM_FIELDS)


def to_params(p) -> SvjParams:
    """Accepts the reference's SVJParams, ours, or anything with the ten fields (duck-typed)."""
    try:
        return SvjParams(*_param_values(p))                 # one C-level attribute sweep (floats, ints, NumPy floats)
    except TypeError:
        return SvjParams(*(float(v) for v in _param_values(p)))     # anything else float() understands


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


class _SerialLib:
    """The library as seen through ONE handle: every call holds the handle's lock.  The C side keeps per-handle scratch
    buffers and allows one call in flight per handle (include/b200mc.h); the reference's own callers are serial, but a web
    server's thread pool sharing the process-wide handle is not, and ctypes releases the GIL during a call.  Threads that
    share a handle therefore take turns at call granularity; the GPU work stays ordered by the handle's stream."""

    def __init__(self, lib, lock):
        self.__dict__["_lib"] = lib
        self.__dict__["_lock"] = lock

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        lock = self._lock

        def call(*args):
            with lock:
                return fn(*args)
        self.__dict__[name] = call               # next time a plain attribute hit
        return call


class Handle:
    """One b200mc_handle: one device, one stream, scratch.  Calls from several threads are serialised per handle
    (`_SerialLib`); a sequence of calls that belongs together (an asynchronous launch and the read-back of its result) is
    one method here and runs under `self.lock` where it matters."""

    def __init__(self, device: int = 0):
        self.lock = threading.RLock()
        self.lib = _SerialLib(load(), self.lock)
        h = _vp()
        rc = self.lib.b200mc_create(int(device), C.byref(h))
        if rc != OK:
            raise B200MCError(rc, (self.lib.b200mc_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)

    # -- plumbing ------------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != OK:
            raise B200MCError(rc, (self.lib.b200mc_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            pool = getattr(self, "_draws_pool", None)
            if pool is not None:                      # buffer handed back by a closed ReferenceDraws
                self._draws_pool = None
                self.lib.b200mc_free(self.h, _vp(pool[1]))
            self.lib.b200mc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, khz, cc, hbm = C.c_int(), C.c_int(), C.c_int(), _u64()
        self._check(self.lib.b200mc_device_info(self.h, C.byref(sm), C.byref(khz), C.byref(hbm), C.byref(cc)))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "hbm_bytes": hbm.value, "cc": cc.value}

    @property
    def launches(self) -> int:
        return int(self.lib.b200mc_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.b200mc_stream(self.h))

    def set_stream(self, stream: int):
        self._check(self.lib.b200mc_set_stream(self.h, int(stream)))

    def synchronize(self):
        self._check(self.lib.b200mc_synchronize(self.h))

    # -- exchange over NVLink peer memory (csrc/peer.cu) ----------------------------------------------------
    def peer_create(self) -> bytes:
        if os.environ.get("B200MC_PEER_DISABLE_RANK") == os.environ.get("RANK", "0"):     # fault injection for the fallback
            raise B200MCError(ECUDA, "peer exchange disabled on this rank (B200MC_PEER_DISABLE_RANK)")
        buf = (C.c_ubyte * 64)()
        self._check(self.lib.b200mc_peer_create(self.h, buf))
        return bytes(buf)

    def peer_connect(self, rank: int, world: int, handles) -> None:
        blob = b"".join(handles)
        if len(blob) != 64 * world:
            raise B200MCError(EINVAL, "peer_connect needs one 64-byte IPC handle per rank")
        self._check(self.lib.b200mc_peer_connect(self.h, int(rank), int(world), blob))

    def peer_allreduce(self, data_dev: int, n_doubles: int) -> None:
        """In-place sum over the ranks of n_doubles float64 at device pointer data_dev (asynchronous, collective)."""
        self._check(self.lib.b200mc_peer_allreduce(self.h, _vp(data_dev), int(n_doubles)))

    def peer_close(self) -> None:
        self._check(self.lib.b200mc_peer_close(self.h))

    def timer_begin(self):
        self._check(self.lib.b200mc_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._check(self.lib.b200mc_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def malloc(self, nbytes: int) -> int:
        p = _vp()
        self._check(self.lib.b200mc_malloc(self.h, int(nbytes), C.byref(p)))
        return int(p.value)

    def free(self, ptr: int):
        self._check(self.lib.b200mc_free(self.h, _vp(ptr)))

    def h2d(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._check(self.lib.b200mc_memcpy_h2d(self.h, _vp(dst), src.ctypes.data, src.nbytes))

    def d2h(self, dst: np.ndarray, src: int):
        assert dst.flags["C_CONTIGUOUS"]
        self._check(self.lib.b200mc_memcpy_d2h(self.h, dst.ctypes.data, _vp(src), dst.nbytes))

    # -- a1: deterministic mode ------------------------------------------------------------------------
    def simulate_given_normals(self, params, S0, T, Z1, Z2, Z_jump, Z_jump_size, n_steps, record_paths=False):
        arrs = [np.ascontiguousarray(z, dtype=np.float64) for z in (Z1, Z2, Z_jump, Z_jump_size)]
        n = arrs[0].shape[0] if arrs[0].ndim else 0
        for z in arrs:
            if z.ndim != 2 or z.shape[0] != n or z.shape[1] < n_steps:
                raise B200MCError(EINVAL, "Z arrays must be [n_paths, >= n_steps] with a common n_paths")
        if any(z.shape[1] != n_steps for z in arrs):      # the reference indexes Z[i, step] for step < num_steps
            arrs = [np.ascontiguousarray(z[:, :n_steps]) for z in arrs]
        S = np.empty(n, dtype=np.float64)
        v = np.empty(n, dtype=np.float64)
        paths = np.empty((n, n_steps + 1), dtype=np.float64) if record_paths else None
        sp = to_params(params)
        self._check(self.lib.b200mc_simulate_given_normals(
            self.h, C.byref(sp), float(S0), float(T), n, int(n_steps), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
            _ptr(arrs[3]), int(bool(record_paths)), _ptr(S), _ptr(v), _ptr(paths)))
        return S, v, paths

    # -- fused European --------------------------------------------------------------------------------
    def price_european(self, params, S0, T, n_steps, n_paths, seed, strikes, is_call=True, flags=0,
                       bumps: Optional[Bumps] = None, path_offset=0, out_dev: Optional[int] = None):
        """Returns a float64 array [n_strikes, NSUMS] (columns = SUMS_FIELDS), or None when `out_dev`
        (device pointer to n_strikes b200mc_sums) is given: then the launch is asynchronous."""
        # Small calls (calibration, the web handlers) are bound by fixed costs, this marshalling among them: the strikes
        # travel as a ctypes array (no ndarray and no `.ctypes` proxy object for the usual short list) and the sums land
        # in a ctypes buffer that becomes the returned array.
        if isinstance(strikes, (list, tuple)):
            nk = len(strikes)
            ks = (C.c_double * nk)(*strikes)
        else:
            kk = np.ascontiguousarray(np.atleast_1d(strikes), dtype=np.float64)
            nk = int(kk.size)
            ks = kk.ctypes.data                              # kk stays referenced until the call returns
        sp = to_params(params)
        bp = C.byref(bumps) if bumps is not None else None
        args = (self.h, C.byref(sp), float(S0), float(T), int(n_steps), int(n_paths), int(seed) & 0xFFFFFFFFFFFFFFFF,
                int(path_offset), ks, nk, 1 if is_call else 0, int(flags), bp)
        if out_dev is not None:
            self._check(self.lib.b200mc_price_european_async(*args, _vp(out_dev)))
            return None
        buf = (C.c_double * (nk * NSUMS if nk > 0 else NSUMS))()
        self._check(self.lib.b200mc_price_european(*args, buf))
        return np.frombuffer(buf, dtype=np.float64, count=nk * NSUMS).reshape(nk, NSUMS) if nk > 0 else \
            np.empty((0, NSUMS), dtype=np.float64)

    def price_cells(self, cells, strikes, flags=0, out_dev: Optional[int] = None):
        """cells: a structured array from make_cells(), or a sequence of dicts / objects with params, S0, T, n_steps,
        n_paths, seed[, path_offset, is_call]; strikes: [n_cells, n_strikes] (or [n_cells]).  Returns float64
        [n_cells, n_strikes, NSUMS], or None when `out_dev` is given (asynchronous)."""
        if not (isinstance(cells, np.ndarray) and cells.dtype == CELL_DTYPE):
            if not len(cells):
                raise B200MCError(EINVAL, "cells must hold at least one cell")
            get = (lambda c, k, d=None: c.get(k, d)) if isinstance(cells[0], dict) else \
                  (lambda c, k, d=None: getattr(c, k, d))
            cells = make_cells([get(c, "params") for c in cells], [get(c, "S0") for c in cells],
                               [get(c, "T") for c in cells], [get(c, "n_steps") for c in cells],
                               [get(c, "n_paths") for c in cells], [get(c, "seed") for c in cells],
                               [get(c, "path_offset", 0) or 0 for c in cells],
                               [bool(get(c, "is_call", True)) for c in cells])
        cells = np.ascontiguousarray(cells)
        n = cells.size
        if not n:
            raise B200MCError(EINVAL, "cells must hold at least one cell")
        ks = np.ascontiguousarray(strikes, dtype=np.float64).reshape(n, -1)
        if out_dev is not None:
            self._check(self.lib.b200mc_price_cells(self.h, cells.ctypes.data, n, ks.ctypes.data, ks.shape[1],
                                                     int(flags), 1, _vp(out_dev)))
            return None
        out = np.empty((n, ks.shape[1], NSUMS), dtype=np.float64)
        self._check(self.lib.b200mc_price_cells(self.h, cells.ctypes.data, n, ks.ctypes.data, ks.shape[1], int(flags), 0,
                                                 out.ctypes.data))
        return out

    def hedge_walk(self, params, S0, strike, T, is_call, n_days, n_scenarios, cost_bps, premiums=None, Z=None, seed=0,
                   scenario_offset=0):
        """The delta-hedging walk of engine/risk.py:278-316 for all scenarios: returns (final_pnl, txn_cost), float64
        [n_scenarios].  Z: [n_scenarios, n_days] host normals, or None for Philox draws (STREAM_HEDGE) keyed by seed."""
        n = int(n_scenarios)
        prem = None if premiums is None else np.ascontiguousarray(premiums, dtype=np.float64).ravel()
        if prem is not None and prem.size != n:
            raise B200MCError(EINVAL, "premiums must hold one value per scenario")
        z = None
        if Z is not None:
            z = np.ascontiguousarray(Z, dtype=np.float64)
            if z.shape != (n, int(n_days)):
                raise B200MCError(EINVAL, "Z must be [n_scenarios, n_days]")
        pnl, cost = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)
        sp = to_params(params)
        self._check(self.lib.b200mc_hedge_walk(self.h, C.byref(sp), float(S0), float(strike), float(T), int(bool(is_call)),
                                                int(n_days), n, float(cost_bps), _ptr(prem), _ptr(z),
                                                int(seed) & (2 ** 64 - 1), int(scenario_offset), pnl.ctypes.data,
                                                cost.ctypes.data))
        return pnl, cost

    def implied_vol(self, prices, S, strikes, maturities, r, q, is_call=True, lo=0.001, hi=5.0) -> np.ndarray:
        """Black-Scholes implied volatilities (engine/surface.py:48-66) of a whole chain: prices, strikes, maturities and
        is_call broadcast against each other; returns float64 of the broadcast shape, NaN where the reference gives None."""
        pr, ks, ts, cl = np.broadcast_arrays(np.asarray(prices, dtype=np.float64), np.asarray(strikes, dtype=np.float64),
                                             np.asarray(maturities, dtype=np.float64), np.asarray(is_call))
        shape = pr.shape
        out = np.empty(shape, dtype=np.float64)
        if out.size == 0:
            return out
        pr, ks, ts = (np.ascontiguousarray(a).ravel() for a in (pr, ks, ts))
        cl = np.ascontiguousarray(cl.astype(bool).astype(np.int32)).ravel()
        flat = out.reshape(-1)
        self._check(self.lib.b200mc_implied_vol(self.h, flat.size, pr.ctypes.data, ks.ctypes.data, ts.ctypes.data,
                                                 cl.ctypes.data, float(S), float(r), float(q), float(lo), float(hi),
                                                 flat.ctypes.data))
        return out

    # -- quasi-Monte Carlo ------------------------------------------------------------------
    def qmc_normals(self, sobol, n_paths, n_steps, which, path_offset=0) -> np.ndarray:
        """Step normals / uniforms of one block of the device Sobol front end, float64 [n_paths, n_steps].
        sobol = (sv [n_dims, bits] uint32, shift [n_dims] uint32, bits), see sobol_tables()."""
        sv, shift, bits = sobol
        out = np.empty((int(n_paths), int(n_steps)), dtype=np.float64)
        self._check(self.lib.b200mc_qmc_normals(self.h, int(n_paths), int(path_offset), int(n_steps), sv.ctypes.data,
                                                 shift.ctypes.data, sv.shape[0], int(bits), int(which), out.ctypes.data))
        return out

    def price_european_qmc(self, params, S0, T, n_steps, n_paths, sobol, strikes, is_call=True, flags=0, path_offset=0):
        sv, shift, bits = sobol
        strikes = np.ascontiguousarray(np.atleast_1d(strikes), dtype=np.float64)
        out = np.empty((strikes.size, NSUMS), dtype=np.float64)
        sp = to_params(params)
        self._check(self.lib.b200mc_price_european_qmc(self.h, C.byref(sp), float(S0), float(T), int(n_steps), int(n_paths),
                                                        int(path_offset), sv.ctypes.data, shift.ctypes.data, sv.shape[0],
                                                        int(bits), strikes.ctypes.data, strikes.size, int(bool(is_call)),
                                                        int(flags), out.ctypes.data))
        return out

    def pcg64_random(self, seed, n: int, first: int = 0) -> np.ndarray:
        """np.random.default_rng(seed).random(first + n)[first:], generated on the device (bitwise)."""
        st = pcg64_state(seed)
        out = np.empty(int(n), dtype=np.float64)
        self._check(self.lib.b200mc_pcg64_random(self.h, st.ctypes.data, int(first), int(n), out.ctypes.data))
        return out

    def numpy_fill(self, seed, n: int, kind: int = NUMPY_STANDARD_NORMAL, first_raw: int = 0, out_dev: Optional[int] = None):
        """n doubles of np.random.default_rng(seed).standard_normal(...) (kind NUMPY_STANDARD_NORMAL) or .random(...)
        (NUMPY_RANDOM), generated on the device bit for bit, starting `first_raw` generator outputs into the stream.
        `seed` is anything default_rng accepts, or the uint64[4] state from pcg64_state.  Returns (array or None when
        out_dev is given, generator outputs consumed)."""
        st = seed if isinstance(seed, np.ndarray) and seed.dtype == np.uint64 and seed.size == 4 else pcg64_state(seed)
        used = _u64(0)
        if out_dev is not None:
            self._check(self.lib.b200mc_numpy_fill(self.h, st.ctypes.data, int(first_raw), int(n), int(kind), 1, _vp(out_dev),
                                                    C.byref(used)))
            return None, int(used.value)
        out = np.empty(int(n), dtype=np.float64)
        self._check(self.lib.b200mc_numpy_fill(self.h, st.ctypes.data, int(first_raw), int(n), int(kind), 0, out.ctypes.data,
                                                C.byref(used)))
        return out, int(used.value)

    def simulate_given_normals_dev(self, params, S0, T, n_paths, n_steps, Z1, Z2, Z_jump, Z_jump_size, S_dev, v_dev,
                                   paths_dev: Optional[int] = None, negate: bool = False):
        """b200mc_simulate_given_normals_dev: every array a DEVICE pointer; asynchronous on the handle's stream.
        negate=True evaluates the antithetic twin (-Z1, -Z2, Z_jump, -Z_jump_size) on the same arrays."""
        sp = to_params(params)
        fl = (GIVEN_RECORD if paths_dev is not None else 0) | (GIVEN_NEGATE if negate else 0)
        self._check(self.lib.b200mc_simulate_given_normals_dev(
            self.h, C.byref(sp), float(S0), float(T), int(n_paths), int(n_steps), _vp(Z1), _vp(Z2), _vp(Z_jump),
            _vp(Z_jump_size), fl, _vp(S_dev), _vp(v_dev), _vp(paths_dev) if paths_dev is not None else None))

    def qmc_terminal(self, params, S0, T, n_steps, n_paths, sobol, nodes=None, Z_jump=None, flags=0, path_offset=0,
                     pcg64_seed=None):
        """Terminal spots (and the antithetic twin's with ANTITHETIC) of Sobol-driven paths with a caller-supplied bridge
        table (BRIDGE_DTYPE array in construction order; None = the built-in correct bridge) and, optionally, host jump
        uniforms [n_paths, n_steps].  Returns (S, S_anti or None), float64 [n_paths]."""
        sv, shift, bits = sobol
        n = int(n_paths)
        S = np.empty(n, dtype=np.float64)
        A = np.empty(n, dtype=np.float64) if flags & ANTITHETIC else None
        nd = None if nodes is None else np.ascontiguousarray(nodes, dtype=BRIDGE_DTYPE)
        zj = None if Z_jump is None else np.ascontiguousarray(Z_jump, dtype=np.float64)
        if zj is not None and zj.shape != (n, int(n_steps)):
            raise B200MCError(EINVAL, "Z_jump must be [n_paths, n_steps]")
        sp = to_params(params)
        st = None if pcg64_seed is None else pcg64_state(pcg64_seed)      # jump uniforms = default_rng(pcg64_seed).random(...)
        self._check(self.lib.b200mc_qmc_terminal(self.h, C.byref(sp), float(S0), float(T), int(n_steps), n, int(path_offset),
                                                  sv.ctypes.data, shift.ctypes.data, sv.shape[0], int(bits), _ptr(nd),
                                                  0 if nd is None else nd.size, _ptr(zj), _ptr(st), int(flags),
                                                  S.ctypes.data, _ptr(A)))
        return S, A

    def simulate_terminal(self, params, S0, T, n_steps, n_paths, seed, flags=0, dtype=np.float64, path_offset=0,
                          want_anti=False, want_v=False, dev_ptrs=None):
        dt = np.dtype(dtype)
        code = F64 if dt == np.float64 else F32
        sp = to_params(params)
        if dev_ptrs is not None:
            S, A, V = dev_ptrs
            self._check(self.lib.b200mc_simulate_terminal(
                self.h, C.byref(sp), float(S0), float(T), int(n_steps), int(n_paths), int(seed) & (2 ** 64 - 1),
                int(path_offset), int(flags), code, 1, _vp(S) if S else None, _vp(A) if A else None,
                _vp(V) if V else None))
            return None
        S = np.empty(n_paths, dtype=dt)
        A = np.empty(n_paths, dtype=dt) if want_anti else None
        V = np.empty(n_paths, dtype=dt) if want_v else None
        self._check(self.lib.b200mc_simulate_terminal(
            self.h, C.byref(sp), float(S0), float(T), int(n_steps), int(n_paths), int(seed) & (2 ** 64 - 1),
            int(path_offset), int(flags), code, 0, _ptr(S), _ptr(A), _ptr(V)))
        return S, A, V

    def generate_paths(self, params, S0, T, n_steps, n_paths, seed, flags=0, dtype=np.float64, path_offset=0,
                       ld: Optional[int] = None, out_dev: Optional[int] = None):
        dt = np.dtype(dtype)
        code = F64 if dt == np.float64 else F32
        ld = int(ld) if ld is not None else int(n_steps) + 1
        sp = to_params(params)
        if out_dev is not None:
            self._check(self.lib.b200mc_generate_paths(
                self.h, C.byref(sp), float(S0), float(T), int(n_steps), int(n_paths), int(seed) & (2 ** 64 - 1),
                int(path_offset), int(flags), code, 1, _vp(out_dev), ld))
            return None
        out = np.empty((n_paths, ld), dtype=dt)
        self._check(self.lib.b200mc_generate_paths(
            self.h, C.byref(sp), float(S0), float(T), int(n_steps), int(n_paths), int(seed) & (2 ** 64 - 1),
            int(path_offset), int(flags), code, 0, out.ctypes.data, ld))
        return out[:, :n_steps + 1] if ld != n_steps + 1 else out

    def risk_metrics(self, pnl, confidence=0.99, n: Optional[int] = None, dtype=None) -> np.ndarray:
        """pnl: NumPy array (host) or an int device pointer (then n and dtype are required)."""
        out = np.empty(8, dtype=np.float64)
        if isinstance(pnl, int):
            code = F64 if np.dtype(dtype) == np.float64 else F32
            self._check(self.lib.b200mc_risk_metrics(self.h, _vp(pnl), int(n), code, 1, float(confidence),
                                                      out.ctypes.data_as(_dp)))
            return out
        a = np.asarray(pnl)
        if a.dtype != np.float32:
            a = a.astype(np.float64, copy=False)
        a = np.ascontiguousarray(a).ravel()
        code = F64 if a.dtype == np.float64 else F32
        self._check(self.lib.b200mc_risk_metrics(self.h, a.ctypes.data, a.size, code, 0, float(confidence),
                                                  out.ctypes.data_as(_dp)))
        return out

    def risk_metrics_sharded(self, pnl, confidence=0.99, n: Optional[int] = None, dtype=None) -> np.ndarray:
        """Collective over the handle's peer connection: pnl is THIS rank's shard (NumPy array, or device pointer + n +
        dtype); returns the global metrics."""
        out = np.empty(8, dtype=np.float64)
        if isinstance(pnl, int):
            code = F64 if np.dtype(dtype) == np.float64 else F32
            self._check(self.lib.b200mc_risk_metrics_sharded(self.h, _vp(pnl) if n else None, int(n), code, 1,
                                                              float(confidence), out.ctypes.data_as(_dp)))
            return out
        a = np.asarray(pnl)
        if a.dtype != np.float32:
            a = a.astype(np.float64, copy=False)
        a = np.ascontiguousarray(a).ravel()
        code = F64 if a.dtype == np.float64 else F32
        self._check(self.lib.b200mc_risk_metrics_sharded(self.h, a.ctypes.data if a.size else None, a.size, code, 0,
                                                          float(confidence), out.ctypes.data_as(_dp)))
        return out

    def option_pnl(self, S_dev: int, n: int, strike: float, is_call: bool, discount: float, premium: float, pnl_dev: int,
                   dtype_in=np.float64, dtype_out=np.float64, stride: int = 1):
        """pnl[i] = discount * payoff(S[i * stride]) - premium, device pointers in and out (asynchronous)."""
        ci = F64 if np.dtype(dtype_in) == np.float64 else F32
        co = F64 if np.dtype(dtype_out) == np.float64 else F32
        self._check(self.lib.b200mc_option_pnl(self.h, _vp(S_dev), int(n), int(stride), ci, float(strike), int(bool(is_call)),
                                                float(discount), float(premium), co, _vp(pnl_dev)))

    # multi-rank tail-metric primitives (host side: risk.compute_risk_metrics_sharded)
    def risk_begin(self, pnl, n: Optional[int] = None, dtype=None) -> np.ndarray:
        out = np.zeros(2, dtype=np.float64)
        if isinstance(pnl, int):
            code = F64 if np.dtype(dtype) == np.float64 else F32
            self._check(self.lib.b200mc_risk_begin(self.h, _vp(pnl), int(n), code, 1, out.ctypes.data_as(_dp)))
            return out
        a = np.asarray(pnl)
        if a.dtype != np.float32:
            a = a.astype(np.float64, copy=False)
        a = np.ascontiguousarray(a).ravel()
        self._risk_keepalive = a
        self._check(self.lib.b200mc_risk_begin(self.h, a.ctypes.data if a.size else None, a.size,
                                               F64 if a.dtype == np.float64 else F32, 0, out.ctypes.data_as(_dp)))
        return out

    def risk_hist(self, radix_pass: int, nsel: int, prefix) -> np.ndarray:
        pre = np.ascontiguousarray(prefix, dtype=np.uint64)
        hist = np.zeros((2, 256), dtype=np.uint64)
        self._check(self.lib.b200mc_risk_hist(self.h, int(radix_pass), int(nsel), pre.ctypes.data_as(C.POINTER(_u64)),
                                              hist.ctypes.data_as(C.POINTER(_u64))))
        return hist

    def risk_finish(self, mean: float, nsel: int, thr) -> np.ndarray:
        t = np.ascontiguousarray(thr, dtype=np.float64)
        out = np.zeros(6, dtype=np.float64)
        self._check(self.lib.b200mc_risk_finish(self.h, float(mean), int(nsel), t.ctypes.data_as(_dp),
                                                out.ctypes.data_as(_dp)))
        return out

    def dump_normals(self, seed, n_paths, n_steps, stream, which, path_offset=0, jump_prob=0.0) -> np.ndarray:
        """jump_prob = lambda_j * T / n_steps of the run being reproduced (only Z_jump_size of the SVJ stream uses it)."""
        out = np.empty((n_paths, n_steps), dtype=np.float64)
        self._check(self.lib.b200mc_dump_normals(self.h, int(seed) & (2 ** 64 - 1), int(path_offset), int(n_paths),
                                                  int(n_steps), int(stream), int(which), float(jump_prob),
                                                  out.ctypes.data))
        return out

    def normal_moments(self, seed, n_paths, n_blocks, path_offset=0) -> np.ndarray:
        out = np.empty(6, dtype=np.float64)
        self._check(self.lib.b200mc_normal_moments(self.h, int(seed) & (2 ** 64 - 1), int(path_offset), int(n_paths),
                                                    int(n_blocks), out.ctypes.data_as(_dp)))
        return out

    def normal_hist2d(self, seed, n_paths, n_blocks, lag=0, path_offset=0) -> np.ndarray:
        """uint64[64, 64] counts of consecutive normals of the GBM stream on a grid of equiprobable cells (see the header)."""
        out = np.zeros(4096, dtype=np.uint64)
        self._check(self.lib.b200mc_normal_hist2d(self.h, int(seed) & (2 ** 64 - 1), int(path_offset), int(n_paths), int(n_blocks),
                                                   int(lag), out.ctypes.data))
        return out.reshape(64, 64)

    def dump_philox(self, seed, n_paths, n_blocks, stream, path_offset=0) -> np.ndarray:
        out = np.empty((n_paths, n_blocks, 4), dtype=np.uint32)
        self._check(self.lib.b200mc_dump_philox(self.h, int(seed) & (2 ** 64 - 1), int(path_offset), int(n_paths),
                                                 int(n_blocks), int(stream), out.ctypes.data))
        return out


_default = {}
_default_lock = threading.Lock()


class ReferenceDraws:
    """The four arrays of the reference's pseudo-random front end, generated ON THE DEVICE and kept there:
        g = default_rng(seed);  Z1, Z2, Z_jump_size = g.standard_normal((n, steps)) x 3
        Z_jump = default_rng(seed + 1).random((n, steps))          engine/monte_carlo.py:301-308, engine/greeks.py:33-41
    (uniform_seed=None: Z_jump continues the SAME generator after the normals, the order of get_sample_paths, :458-462).
    NumPy's doubles bit for bit (csrc/np_normal.cu, csrc/pcg64.cu); nothing crosses PCIe.  simulate() runs the reference
    recurrence over them (b200mc_simulate_given_normals_dev) and returns host arrays."""

    def __init__(self, handle, seed, n_paths: int, n_steps: int, uniform_seed="seed+1"):
        self.h, self.n, self.steps = handle, int(n_paths), int(n_steps)
        N = self.n * self.steps
        self.nbytes = 4 * N * 8 + 16 * self.n * 8
        self.capacity, self.buf = self._take(handle, self.nbytes)
        self.Z1, self.Z2, self.Zjs, self.Zj = (self.buf + i * N * 8 for i in range(4))
        self.S = self.buf + 4 * N * 8
        self.v = self.S + self.n * 8
        _, used = handle.numpy_fill(seed, 3 * N, NUMPY_STANDARD_NORMAL, 0, out_dev=self.Z1)
        if uniform_seed is None:
            handle.numpy_fill(seed, N, NUMPY_RANDOM, used, out_dev=self.Zj)
        else:
            handle.numpy_fill(seed + 1 if isinstance(uniform_seed, str) else uniform_seed, N, NUMPY_RANDOM, 0, out_dev=self.Zj)
        self.raws_consumed = used

    def simulate(self, params, S0, T, record_paths=False, negate=False):
        paths_dev = None
        if record_paths:
            paths_dev = self.h.malloc(self.n * (self.steps + 1) * 8)
        try:
            self.h.simulate_given_normals_dev(params, S0, T, self.n, self.steps, self.Z1, self.Z2, self.Zj, self.Zjs, self.S,
                                              self.v, paths_dev, negate)
            S, v = np.empty(self.n), np.empty(self.n)
            self.h.d2h(S, self.S)
            self.h.d2h(v, self.v)
            paths = None
            if record_paths:
                paths = np.empty((self.n, self.steps + 1))
                self.h.d2h(paths, paths_dev)
        finally:
            if paths_dev is not None:
                self.h.free(paths_dev)
        return S, v, paths

    def host_arrays(self):
        """(Z1, Z2, Z_jump, Z_jump_size) copied to the host (tests)."""
        out = []
        for ptr in (self.Z1, self.Z2, self.Zj, self.Zjs):
            a = np.empty((self.n, self.steps))
            self.h.d2h(a, ptr)
            out.append(a)
        return out

    # cudaMalloc / cudaFree of a 400 MB buffer cost milliseconds each: a closed ReferenceDraws hands its buffer back to a
    # one-slot pool on the handle, and the next one of the same or a smaller size takes it (calibration-style loops build
    # a new engine per candidate)
    @staticmethod
    def _take(handle, nbytes):
        slot = getattr(handle, "_draws_pool", None)
        if slot is not None and slot[0] >= nbytes:
            handle._draws_pool = None
            return slot
        if slot is not None:
            handle._draws_pool = None
            handle.free(slot[1])
        return nbytes, handle.malloc(nbytes)

    def close(self):
        if self.buf:
            if getattr(self.h, "h", None) is None:          # handle already destroyed: nothing to give back
                self.buf = 0
                return
            old = getattr(self.h, "_draws_pool", None)
            self.h._draws_pool = (self.capacity, self.buf)
            self.buf = 0
            if old is not None:
                self.h.free(old[1])

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


_default_device: Optional[int] = None


def default_handle(device: Optional[int] = None) -> Handle:
    """Process-wide handle per device (LOCAL_RANK picks the device under torchrun)."""
    global _default_device
    if device is None:
        if _default_device is None:      # resolved once per process (an engine per calibration candidate asks every time)
            _default_device = int(os.environ.get("B200MC_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        device = _default_device
        h = _default.get(device)         # lock-free hit: dict reads are atomic, handles are only ever added or replaced
        if h is not None and h.h is not None:
            return h
    with _default_lock:
        h = _default.get(device)
        if h is None or h.h is None:
            h = Handle(device)
            _default[device] = h
        return h

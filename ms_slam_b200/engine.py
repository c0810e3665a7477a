"""ctypes binding of libmss.so (C-ABI: include/mss.h).  Python host side for tests, benchmarks and tools.

The product path is the CUDA library; there is NO CPU fallback here: importing works anywhere, but creating an
``Engine`` raises ``MssError`` when libmss.so is missing or no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
import numpy as np

from .window import WindowView, PackedView

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmss.so")

MSS_OK, MSS_E_BADARG, MSS_E_CUDA, MSS_E_NCCL, MSS_E_NOMEM, MSS_E_NOCONVERGE, MSS_E_INTERNAL = 0, -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_DEVICE = 0, 1
RESULT_SAME, RESULT_HOST, RESULT_DEVICE = 0, 1, 2
LAYOUT_SOA, LAYOUT_PACKED, LAYOUT_PACKED16 = 0, 1, 2
UNIQUE_ID_BYTES = 128


class MssError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libmss status {status}: {msg}")
        self.status = status


class mss_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("min_points", C.c_int32), ("lambda_", C.c_float), ("grid_lambda", C.c_float),
                ("max_rounds", C.c_int32), ("all_rule_steps", C.c_int32), ("max_drop_rounds", C.c_int32),
                ("stall_den", C.c_int32)]


class mss_window_view(C.Structure):
    _fields_ = [("K", C.c_int32), ("H", C.c_int32), ("M", C.c_int32), ("F", C.c_int32), ("O", C.c_int32),
                ("memory", C.c_int32),
                ("feat_ptr", C.c_void_p), ("feat_mp", C.c_void_p), ("feat_cell", C.c_void_p), ("mp_nobs", C.c_void_p),
                ("mp_obs_ptr", C.c_void_p), ("mp_obs_kf", C.c_void_p), ("okf_total", C.c_void_p),
                ("layout", C.c_int32), ("n_max_floor", C.c_int32),
                ("slots", C.c_void_p), ("mp_nobs16", C.c_void_p), ("slots16", C.c_void_p), ("obs_pairs", C.c_void_p),
                ("result_memory", C.c_int32), ("nobs8", C.c_int32), ("mp_tie", C.c_void_p)]


def packed_c_view(K, H, M, F, O, memory, feat_ptr, slots, mp_nobs16, obs_pairs, okf_total, n_max_floor=0, tokens16=False,
                  nobs8=False) -> "mss_window_view":
    """MSS_LAYOUT_PACKED (u32 slots) or MSS_LAYOUT_PACKED16 (u16 tokens) view from raw addresses; nobs8: the nObs table holds bytes"""
    if tokens16:
        return mss_window_view(K, H, M, F, O, memory, feat_ptr, None, None, None, None, None, okf_total,
                               LAYOUT_PACKED16, n_max_floor, None, mp_nobs16, slots, obs_pairs, 0, 1 if nobs8 else 0)
    return mss_window_view(K, H, M, F, O, memory, feat_ptr, None, None, None, None, None, okf_total,
                           LAYOUT_PACKED, n_max_floor, slots, mp_nobs16, None, obs_pairs, 0, 1 if nobs8 else 0)


class mss_result(C.Structure):
    _fields_ = [("keep_bits", C.c_void_p), ("kf_cov", C.c_void_p), ("kf_slack", C.c_void_p),
                ("objective", C.c_double), ("dual_bound", C.c_double), ("sum_cost", C.c_int64),
                ("uncovered_cells", C.c_int32), ("total_slack", C.c_int32), ("n_max", C.c_int32),
                ("n_vars", C.c_int32), ("n_cells", C.c_int32), ("nnz", C.c_int32), ("n_kept", C.c_int32),
                ("rounds", C.c_int32), ("status", C.c_int32), ("time_build_us", C.c_float),
                ("time_solve_us", C.c_float)]


class mss_stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("solves", C.c_int64), ("last_device_ms", C.c_double),
                ("last_total_ms", C.c_double), ("last_h2d_bytes", C.c_int64), ("last_d2h_bytes", C.c_int64),
                ("device_bytes", C.c_int64), ("grid_ctas", C.c_int32), ("sm_count", C.c_int32),
                ("last_row_entries", C.c_int64), ("last_var_visits", C.c_int64)]


# every symbol include/mss.h declares (tests check the library exports all of them)
SYMBOLS = ["mss_version", "mss_create", "mss_destroy", "mss_last_error", "mss_set_params", "mss_solve",
           "mss_solve_batch", "mss_comm_unique_id", "mss_comm_init", "mss_comm_destroy", "mss_host_alloc",
           "mss_host_free", "mss_device_alloc", "mss_device_free", "mss_memcpy_h2d", "mss_memcpy_d2h",
           "mss_get_stats", "mss_stream", "mss_debug_trace", "mss_debug_get_trace", "mss_components",
           # persistent device mirror (bound in ms_slam_b200/mirror.py)
           "mss_mirror_create", "mss_mirror_destroy", "mss_mirror_add_keyframe", "mss_mirror_add_keyframes",
           "mss_mirror_set_map_points", "mss_mirror_apply", "mss_mirror_solve", "mss_mirror_build_view", "mss_mirror_get_stats", "mss_mirror_components",
           "mss_compact_keyframes", "mss_mirror_compact_keyframes", "mss_set_dual_bound",
           # BoW re-transform + keyframe database (bound in ms_slam_b200/bow.py)
           "mss_voc_create", "mss_voc_destroy", "mss_voc_words", "mss_bow_transform", "mss_kfdb_common_words", "mss_kfdb_postings",
           # several GPUs from one process
           "mss_multi_create", "mss_multi_destroy", "mss_multi_device_count", "mss_multi_last_error", "mss_multi_set_params",
           "mss_multi_solve_batch", "mss_multi_get_stats"]

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen libmss.so and declare prototypes. Raises MssError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise MssError(MSS_E_CUDA, f"{path} not found: build it with `python __graft_entry__.py` "
                                   f"(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.mss_version.restype = C.c_int
    lib.mss_create.argtypes = [C.POINTER(mss_config), C.POINTER(C.c_void_p)]
    lib.mss_create.restype = C.c_int
    lib.mss_destroy.argtypes = [C.c_void_p]
    lib.mss_destroy.restype = None
    lib.mss_last_error.argtypes = [C.c_void_p]
    lib.mss_last_error.restype = C.c_char_p
    lib.mss_set_params.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float]
    lib.mss_solve.argtypes = [C.c_void_p, C.POINTER(mss_window_view), C.POINTER(mss_result)]
    lib.mss_solve_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(mss_window_view), C.POINTER(mss_result)]
    lib.mss_components.argtypes = [C.c_void_p, C.POINTER(mss_window_view), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32)]
    lib.mss_comm_unique_id.argtypes = [C.c_void_p]
    lib.mss_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.mss_comm_destroy.argtypes = [C.c_void_p]
    lib.mss_host_alloc.argtypes = [C.c_size_t]
    lib.mss_host_alloc.restype = C.c_void_p
    lib.mss_host_free.argtypes = [C.c_void_p]
    lib.mss_host_free.restype = None
    lib.mss_device_alloc.argtypes = [C.c_void_p, C.c_size_t]
    lib.mss_device_alloc.restype = C.c_void_p
    lib.mss_device_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.mss_device_free.restype = None
    lib.mss_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.mss_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.mss_get_stats.argtypes = [C.c_void_p, C.POINTER(mss_stats)]
    lib.mss_stream.argtypes = [C.c_void_p]
    lib.mss_stream.restype = C.c_void_p
    lib.mss_set_dual_bound.argtypes = [C.c_void_p, C.c_int32]
    lib.mss_debug_trace.argtypes = [C.c_void_p, C.c_int32]
    lib.mss_debug_get_trace.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    _lib = lib
    return lib


@dataclass
class Result:
    keep: np.ndarray          # bool [M]
    keep_bits: np.ndarray     # uint32 [(M+31)//32]
    kf_cov: np.ndarray        # int32 [K+H]
    kf_slack: np.ndarray      # int32 [K+H]
    objective: float
    dual_bound: float
    sum_cost: int
    uncovered_cells: int
    total_slack: int
    n_max: int
    n_vars: int
    n_cells: int
    nnz: int
    n_kept: int
    rounds: int
    status: int
    time_build_us: float
    time_solve_us: float


def unpack_bits(words: np.ndarray, M: int) -> np.ndarray:
    bits = (words[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & np.uint32(1)
    return bits.reshape(-1)[:M].astype(bool)


class _Pinned:
    """numpy array over cudaHostAlloc memory (freed with the engine library)."""

    def __init__(self, lib, shape, dtype):
        self.lib = lib
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib.mss_host_alloc(max(n, 1))
        if not self.ptr:
            raise MssError(MSS_E_NOMEM, "cudaHostAlloc failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.mss_host_free(self.ptr)
            self.ptr = None


class DeviceView:
    """A window view resident in device memory (inputs already in HBM: bench `value`, MSS_MEM_DEVICE)."""

    _ARR = ("feat_ptr", "feat_mp", "feat_cell", "mp_nobs", "mp_obs_ptr", "mp_obs_kf", "okf_total")
    _ARR_PACKED = ("feat_ptr", "slots", "mp_nobs16", "obs_pairs", "okf_total")

    def __init__(self, engine: "Engine", view):
        self.engine = engine
        self.packed = isinstance(view, PackedView)
        if self.packed:
            self._ARR = self._ARR_PACKED
        self.K, self.H, self.M, self.F, self.O = view.K, view.H, view.M, view.F, view.O
        self.n_max_floor = int(getattr(view, "n_max_floor", 0))
        self.tokens16 = bool(self.packed and view.meta.get("tokens16"))
        self.nobs8 = bool(self.packed and view.meta.get("nobs8"))
        self.ptrs = {}
        lib, h = engine.lib, engine.handle
        for name in self._ARR:
            a = getattr(view, name)
            p = lib.mss_device_alloc(h, max(a.nbytes, 4))
            if not p:
                raise MssError(MSS_E_NOMEM, "device allocation failed")
            if a.nbytes:
                engine._check(lib.mss_memcpy_h2d(h, p, a.ctypes.data, a.nbytes))
            self.ptrs[name] = p
        self.d_tie = None
        tie = self._tie_of(view)
        if tie is not None:
            self.d_tie = lib.mss_device_alloc(h, max(tie.nbytes, 4))
            engine._check(lib.mss_memcpy_h2d(h, self.d_tie, tie.ctypes.data, tie.nbytes))
        words, rows = (self.M + 31) // 32, self.K + self.H
        self.d_keep = lib.mss_device_alloc(h, max(words * 4, 4))
        self.d_cov = lib.mss_device_alloc(h, max(rows * 4, 4))
        self.d_slack = lib.mss_device_alloc(h, max(rows * 4, 4))

    @staticmethod
    def _tie_of(view):
        """uint32 tie-break ranks the view carries, or None"""
        if isinstance(view, PackedView):
            return None if view.mp_tie is None else np.ascontiguousarray(view.mp_tie, np.uint32)
        if view.meta.get("tie_by_gid"):
            from .window import tie_ranks
            return tie_ranks(view)
        return None

    def c_view(self) -> mss_window_view:
        if self.packed:
            cv = packed_c_view(self.K, self.H, self.M, self.F, self.O, MEM_DEVICE, *[self.ptrs[n] for n in self._ARR],
                               n_max_floor=self.n_max_floor, tokens16=self.tokens16, nobs8=self.nobs8)
        else:
            cv = mss_window_view(self.K, self.H, self.M, self.F, self.O, MEM_DEVICE, *[self.ptrs[n] for n in self._ARR],
                                 LAYOUT_SOA, self.n_max_floor)
        cv.mp_tie = self.d_tie
        return cv

    def fetch(self):
        """copy the device-resident result arrays back to numpy (not part of any timed region)"""
        lib, h = self.engine.lib, self.engine.handle
        words, rows = (self.M + 31) // 32, self.K + self.H
        kb = np.zeros(words, np.uint32)
        cov = np.zeros(rows, np.int32)
        sl = np.zeros(rows, np.int32)
        if words:
            self.engine._check(lib.mss_memcpy_d2h(h, kb.ctypes.data, self.d_keep, words * 4))
        if rows:
            self.engine._check(lib.mss_memcpy_d2h(h, cov.ctypes.data, self.d_cov, rows * 4))
            self.engine._check(lib.mss_memcpy_d2h(h, sl.ctypes.data, self.d_slack, rows * 4))
        return kb, cov, sl

    def free(self):
        lib, h = self.engine.lib, self.engine.handle
        for p in list(self.ptrs.values()) + [self.d_keep, self.d_cov, self.d_slack] + ([self.d_tie] if self.d_tie else []):
            lib.mss_device_free(h, p)
        self.ptrs = {}


class Engine:
    """One engine handle = one CUDA device + stream (the reference's GRBEnv, MapSparsification.h:59)."""

    def __init__(self, N=100, lam=500.0, grid_lam=10.0, device=0, max_rounds=0, all_rule_steps=0,
                 max_drop_rounds=0, stall_den=0):
        self.lib = load_library()
        self.handle = C.c_void_p()
        cfg = mss_config(device, N, lam, grid_lam, max_rounds, all_rule_steps, max_drop_rounds, stall_den)
        rc = self.lib.mss_create(C.byref(cfg), C.byref(self.handle))
        if rc != MSS_OK:
            self.handle = None
            raise MssError(rc, "mss_create failed (no usable CUDA device? libmss has no CPU fallback)")
        self.N, self.lam, self.grid_lam = N, lam, grid_lam
        self.rank, self.nranks = 0, 1

    # -- plumbing -----------------------------------------------------------------------------------------
    def _check(self, rc, allow=()):
        if rc != MSS_OK and rc not in allow:
            raise MssError(rc, self.lib.mss_last_error(self.handle).decode())
        return rc

    def close(self):
        if self.handle:
            self.lib.mss_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, N, lam, grid_lam):
        self._check(self.lib.mss_set_params(self.handle, N, lam, grid_lam))
        self.N, self.lam, self.grid_lam = N, lam, grid_lam

    def set_dual_bound(self, enable=True):
        """per-window lower bound of the reference ILP proven on the device (Result.dual_bound; NaN while off)"""
        self._check(self.lib.mss_set_dual_bound(self.handle, 1 if enable else 0))

    def stats(self) -> dict:
        s = mss_stats()
        self._check(self.lib.mss_get_stats(self.handle, C.byref(s)))
        return {f: getattr(s, f) for f, _ in mss_stats._fields_}

    def pinned(self, shape, dtype) -> _Pinned:
        return _Pinned(self.lib, shape, dtype)

    def trace(self, enable=True):
        self._check(self.lib.mss_debug_trace(self.handle, 1 if enable else 0))

    def get_trace(self, local_window=0):
        """[(phase, free_left, ns)] of one window of the last call (phase: 10..14 build steps, 1 PROP, 2 GREEDY, 3 FORCE,
        4 D1, 5 D2, 6/7 EVAL)"""
        buf = np.zeros(2 * 255, np.uint32)
        n = self.lib.mss_debug_get_trace(self.handle, local_window, buf.ctypes.data, 255)
        if n < 0:
            raise MssError(n, "no trace (call trace(True) before solving)")
        return [(int(buf[2 * i] >> 24), int(buf[2 * i] & 0xFFFFFF), int(buf[2 * i + 1])) for i in range(n)]

    # -- multi-GPU ------------------------------------------------------------------------------------------
    def unique_id(self) -> bytes:
        buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
        self._check(self.lib.mss_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self.lib.mss_comm_init(self.handle, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    # -- solve ------------------------------------------------------------------------------------------------
    @staticmethod
    def _host_view(v) -> mss_window_view:
        if isinstance(v, PackedView):
            cv = packed_c_view(v.K, v.H, v.M, v.F, v.O, MEM_HOST, v.feat_ptr.ctypes.data, v.slots.ctypes.data,
                               v.mp_nobs16.ctypes.data, v.obs_pairs.ctypes.data, v.okf_total.ctypes.data,
                               n_max_floor=v.n_max_floor, tokens16=bool(v.meta.get("tokens16")), nobs8=bool(v.meta.get("nobs8")))
        else:
            cv = mss_window_view(v.K, v.H, v.M, v.F, v.O, MEM_HOST, v.feat_ptr.ctypes.data, v.feat_mp.ctypes.data,
                                 v.feat_cell.ctypes.data, v.mp_nobs.ctypes.data, v.mp_obs_ptr.ctypes.data,
                                 v.mp_obs_kf.ctypes.data, v.okf_total.ctypes.data, LAYOUT_SOA, v.n_max_floor)
        tie = DeviceView._tie_of(v)
        if tie is not None:
            v.meta["_tie_keepalive"] = tie                 # (the array must outlive the call)
            cv.mp_tie = tie.ctypes.data
        return cv

    def components(self, view):
        """Connected components of a host window view (WindowView or PackedView): (row_label[K+H], mp_label[M], ncomp, n_max)"""
        cv = self._host_view(view)
        rows = np.zeros(view.K + view.H, np.int32)
        mps = np.zeros(view.M, np.int32)
        nc, nm = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.mss_components(self.handle, C.byref(cv), rows.ctypes.data, mps.ctypes.data, C.byref(nc), C.byref(nm)))
        return rows, mps, int(nc.value), int(nm.value)

    def solve_batch(self, views, raise_on_status=True):
        """views: list of WindowView (host) or DeviceView. Returns list of Result (all windows, on every rank)."""
        n = len(views)
        cv = (mss_window_view * n)()
        cr = (mss_result * n)()
        host_bufs = []
        for i, v in enumerate(views):
            if isinstance(v, DeviceView):
                cv[i] = v.c_view()
                cr[i].keep_bits, cr[i].kf_cov, cr[i].kf_slack = v.d_keep, v.d_cov, v.d_slack
                host_bufs.append(None)
            else:
                owned = (i % self.nranks) == self.rank
                cv[i] = self._host_view(v) if owned else mss_window_view(v.K, v.H, v.M, 0, 0, MEM_HOST)
                kb = np.zeros((v.M + 31) // 32, np.uint32)
                cov = np.zeros(v.K + v.H, np.int32)
                sl = np.zeros(v.K + v.H, np.int32)
                cr[i].keep_bits, cr[i].kf_cov, cr[i].kf_slack = kb.ctypes.data, cov.ctypes.data, sl.ctypes.data
                host_bufs.append((kb, cov, sl))
        rc = self.lib.mss_solve_batch(self.handle, n, cv, cr)
        if raise_on_status:
            self._check(rc)
        elif rc not in (MSS_OK, MSS_E_BADARG, MSS_E_NOCONVERGE):
            self._check(rc)
        out = []
        for i, v in enumerate(views):
            kb, cov, sl = host_bufs[i] if host_bufs[i] is not None else v.fetch()
            r = cr[i]
            out.append(Result(keep=unpack_bits(kb, v.M), keep_bits=kb, kf_cov=cov, kf_slack=sl, objective=r.objective,
                              dual_bound=r.dual_bound, sum_cost=r.sum_cost, uncovered_cells=r.uncovered_cells,
                              total_slack=r.total_slack, n_max=r.n_max, n_vars=r.n_vars, n_cells=r.n_cells, nnz=r.nnz,
                              n_kept=r.n_kept, rounds=r.rounds, status=r.status, time_build_us=r.time_build_us,
                              time_solve_us=r.time_solve_us))
        return out

    def solve(self, view, raise_on_status=True) -> Result:
        return self.solve_batch([view], raise_on_status=raise_on_status)[0]

    def solve_batch_raw(self, cviews, cresults, n):
        """Timed-loop entry: prebuilt ctypes arrays, no numpy work. Returns the status code."""
        return self.lib.mss_solve_batch(self.handle, n, cviews, cresults)


class MultiEngine:
    """Several GPUs driven from one process (include/mss.h mss_multi_*): window w of a batch -> device w % n."""

    def __init__(self, devices, N=100, lam=500.0, grid_lam=10.0):
        self.lib = load_library()
        self.lib.mss_multi_create.argtypes = [C.POINTER(mss_config), C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        self.lib.mss_multi_destroy.argtypes = [C.c_void_p]
        self.lib.mss_multi_destroy.restype = None
        self.lib.mss_multi_last_error.argtypes = [C.c_void_p]
        self.lib.mss_multi_last_error.restype = C.c_char_p
        self.lib.mss_multi_solve_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(mss_window_view), C.POINTER(mss_result)]
        self.lib.mss_multi_get_stats.argtypes = [C.c_void_p, C.c_int32, C.POINTER(mss_stats)]
        self.lib.mss_multi_set_params.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float]
        self.lib.mss_multi_device_count.argtypes = [C.c_void_p]
        dev = np.ascontiguousarray(devices, np.int32)
        cfg = mss_config(0, N, lam, grid_lam, 0, 0, 0, 0)
        self.handle = C.c_void_p()
        rc = self.lib.mss_multi_create(C.byref(cfg), dev.ctypes.data, dev.size, C.byref(self.handle))
        if rc != MSS_OK:
            self.handle = None
            raise MssError(rc, "mss_multi_create failed")
        self.n = dev.size

    def close(self):
        if self.handle:
            self.lib.mss_multi_destroy(self.handle)
            self.handle = None

    def stats(self, i):
        s = mss_stats()
        self.lib.mss_multi_get_stats(self.handle, i, C.byref(s))
        return {f: getattr(s, f) for f, _ in mss_stats._fields_}

    def solve_batch(self, views):
        n = len(views)
        cv = (mss_window_view * n)()
        cr = (mss_result * n)()
        bufs = []
        for i, v in enumerate(views):
            cv[i] = Engine._host_view(v)
            kb, cov, sl = np.zeros((v.M + 31) // 32, np.uint32), np.zeros(v.K + v.H, np.int32), np.zeros(v.K + v.H, np.int32)
            cr[i].keep_bits, cr[i].kf_cov, cr[i].kf_slack = kb.ctypes.data, cov.ctypes.data, sl.ctypes.data
            bufs.append((kb, cov, sl))
        rc = self.lib.mss_multi_solve_batch(self.handle, n, cv, cr)
        if rc != MSS_OK:
            raise MssError(rc, self.lib.mss_multi_last_error(self.handle).decode())
        out = []
        for (kb, cov, sl), v, r in zip(bufs, views, cr):
            out.append(Result(keep=unpack_bits(kb, v.M), keep_bits=kb, kf_cov=cov, kf_slack=sl, objective=r.objective,
                              dual_bound=r.dual_bound, sum_cost=r.sum_cost, uncovered_cells=r.uncovered_cells,
                              total_slack=r.total_slack, n_max=r.n_max, n_vars=r.n_vars, n_cells=r.n_cells, nnz=r.nnz,
                              n_kept=r.n_kept, rounds=r.rounds, status=r.status, time_build_us=r.time_build_us,
                              time_solve_us=r.time_solve_us))
        return out

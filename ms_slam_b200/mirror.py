"""ctypes binding of the persistent device mirror (include/mss.h ``mss_mirror_*``).

The mirror keeps the keyframe x map-point incidence of the whole map in HBM (keyframe-major slot arrays + per-map-point
attributes) and follows the map through small deltas; a window is then ``K`` keyframe handles instead of a flattened
view, and what comes back is a bitmask over map-point handles.  Python host side for tests and benchmarks; no compute here.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from .engine import Engine, MssError, Result, mss_result, unpack_bits, MSS_OK, MSS_E_BADARG, MSS_E_NOCONVERGE
from .window import PackedView, CELL_NONE

MOP_SLOT, MOP_OBS, MOP_MP, MOP_KF_COMPACT = 1, 2, 3, 4
OP_DTYPE = np.dtype([("kind", np.int32), ("a", np.int32), ("b", np.int32), ("c", np.int32)])


class mss_mirror_stats(C.Structure):
    _fields_ = [("n_keyframes", C.c_int32), ("n_map_points", C.c_int32), ("slots_per_kf", C.c_int32), ("reserved_", C.c_int32),
                ("device_bytes", C.c_int64), ("ops_applied", C.c_int64), ("windows_built", C.c_int64),
                ("last_build_ms", C.c_double), ("last_solve_ms", C.c_double), ("last_total_ms", C.c_double),
                ("last_h2d_bytes", C.c_int64), ("last_d2h_bytes", C.c_int64)]


class mss_kf_payload(C.Structure):
    _fields_ = [("n", C.c_int32), ("kf", C.c_int32), ("keep", C.c_void_p), ("descriptors", C.c_void_p), ("keypoints", C.c_void_p),
                ("uright", C.c_void_p), ("depth", C.c_void_p)]


class mss_mirror_window(C.Structure):
    _fields_ = [("K", C.c_int32), ("n_max_floor", C.c_int32), ("kf", C.c_void_p), ("del_bits", C.c_void_p),
                ("del_words", C.c_int32), ("h_lo", C.c_int32), ("h_hi", C.c_int32),
                ("M", C.c_int32), ("H", C.c_int32), ("F", C.c_int32), ("O", C.c_int32), ("n_deleted", C.c_int32),
                ("apply", C.c_int32), ("mp_handle", C.c_void_p), ("mp_cap", C.c_int32), ("reserved_", C.c_int32)]


SYMBOLS = ["mss_mirror_create", "mss_mirror_destroy", "mss_mirror_add_keyframe", "mss_mirror_add_keyframes",
           "mss_mirror_set_map_points", "mss_mirror_apply", "mss_mirror_solve", "mss_mirror_build_view", "mss_mirror_get_stats",
           "mss_mirror_components", "mss_compact_keyframes", "mss_mirror_compact_keyframes"]


def _declare(lib):
    if getattr(lib, "_mirror_declared", False):
        return
    lib.mss_mirror_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    lib.mss_mirror_destroy.argtypes = [C.c_void_p]
    lib.mss_mirror_destroy.restype = None
    lib.mss_mirror_add_keyframe.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mss_mirror_add_keyframes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mss_mirror_set_map_points.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.mss_mirror_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.mss_mirror_solve.argtypes = [C.c_void_p, C.c_int32, C.POINTER(mss_mirror_window), C.POINTER(mss_result)]
    lib.mss_mirror_build_view.argtypes = [C.c_void_p, C.POINTER(mss_mirror_window), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mss_mirror_get_stats.argtypes = [C.c_void_p, C.POINTER(mss_mirror_stats)]
    lib.mss_mirror_components.argtypes = [C.c_void_p, C.POINTER(mss_mirror_window), C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.mss_compact_keyframes.argtypes = [C.c_void_p, C.c_int32, C.POINTER(mss_kf_payload), C.c_void_p]
    lib.mss_mirror_compact_keyframes.argtypes = [C.c_void_p, C.c_int32, C.POINTER(mss_kf_payload), C.c_void_p]
    lib._mirror_declared = True


def _p(a):
    return None if a is None else a.ctypes.data


class KeyframePayload:
    """The per-keypoint arrays of one keyframe in device memory (what KeyFrame::EraseBadDescriptor compacts,
    /root/reference/src/KeyFrame.cc:311-361): descriptors u8 [n,32], keypoints 7 x 32-bit words, uRight, depth."""
    _ARR = ("keep", "descriptors", "keypoints", "uright", "depth")

    def __init__(self, engine: Engine, n, keep=None, descriptors=None, keypoints=None, uright=None, depth=None, kf=-1):
        self.engine, self.n, self.kf = engine, int(n), int(kf)
        self.ptr, self.meta = {}, {}
        lib, h = engine.lib, engine.handle
        given = dict(keep=None if keep is None else np.ascontiguousarray(keep, np.uint8),
                     descriptors=None if descriptors is None else np.ascontiguousarray(descriptors, np.uint8).reshape(self.n, 32),
                     keypoints=None if keypoints is None else np.ascontiguousarray(keypoints).view(np.uint32).reshape(self.n, 7),
                     uright=None if uright is None else np.ascontiguousarray(uright, np.float32),
                     depth=None if depth is None else np.ascontiguousarray(depth, np.float32))
        for name, a in given.items():
            if a is None:
                self.ptr[name] = None
                continue
            p = lib.mss_device_alloc(h, max(a.nbytes, 16))
            if not p:
                raise MssError(-4, "device allocation failed")
            if a.nbytes:
                engine._check(lib.mss_memcpy_h2d(h, p, a.ctypes.data, a.nbytes))
            self.ptr[name], self.meta[name] = p, (a.dtype, a.shape)

    def c_struct(self) -> "mss_kf_payload":
        return mss_kf_payload(self.n, self.kf, *[self.ptr[k] for k in self._ARR])

    def fetch(self, n_rows):
        """the first n_rows rows of every array (after a compaction: the survivors)"""
        out = {}
        for name in self._ARR[1:]:
            if self.ptr[name] is None:
                out[name] = None
                continue
            dt, shape = self.meta[name]
            a = np.zeros((n_rows,) + tuple(shape[1:]), dt)
            if a.nbytes:
                self.engine._check(self.engine.lib.mss_memcpy_d2h(self.engine.handle, a.ctypes.data, self.ptr[name], a.nbytes))
            out[name] = a
        return out

    def free(self):
        for p in self.ptr.values():
            if p:
                self.engine.lib.mss_device_free(self.engine.handle, p)
        self.ptr = {}


def compact_keyframes(engine: Engine, payloads, mirror: "Mirror" = None):
    """mss_compact_keyframes / mss_mirror_compact_keyframes over a list of KeyframePayload -> rows left per keyframe"""
    _declare(engine.lib)
    n = len(payloads)
    arr = (mss_kf_payload * max(n, 1))(*[p.c_struct() for p in payloads])
    out = np.zeros(max(n, 1), np.int32)
    if mirror is None:
        engine._check(engine.lib.mss_compact_keyframes(engine.handle, n, arr, out.ctypes.data))
    else:
        engine._check(engine.lib.mss_mirror_compact_keyframes(mirror.handle, n, arr, out.ctypes.data))
    return out[:n]


class MirrorResult:
    """one window of Mirror.solve"""
    def __init__(self, result: Result, deleted: np.ndarray, mp_handle, win: mss_mirror_window):
        self.result = result
        self.deleted = deleted                 # sorted map-point handles the selection dropped
        self.mp_handle = mp_handle             # table index -> handle (bit order of result.keep), when asked for
        self.M, self.H, self.F, self.O = win.M, win.H, win.F, win.O
        self.h_lo, self.h_hi, self.n_deleted = win.h_lo, win.h_hi, win.n_deleted


class Mirror:
    def __init__(self, engine: Engine, slots_per_kf: int):
        self.engine, self.lib = engine, engine.lib
        _declare(self.lib)
        self.handle = C.c_void_p()
        engine._check(self.lib.mss_mirror_create(engine.handle, slots_per_kf, C.byref(self.handle)))
        self.S = slots_per_kf

    def close(self):
        if self.handle:
            self.lib.mss_mirror_destroy(self.handle)
            self.handle = None

    def _check(self, rc, allow=()):
        return self.engine._check(rc, allow)

    def stats(self) -> dict:
        s = mss_mirror_stats()
        self._check(self.lib.mss_mirror_get_stats(self.handle, C.byref(s)))
        return {f: getattr(s, f) for f, _ in mss_mirror_stats._fields_}

    # ---- loading / deltas --------------------------------------------------------------------------------------
    def add_keyframes(self, kf0, n_slots, cells, slot_mp, obs_mp=None, sort_key=None):
        n = len(n_slots)
        n_slots = np.ascontiguousarray(n_slots, np.int32)
        cells = np.ascontiguousarray(cells, np.uint16).reshape(n, self.S)
        slot_mp = np.ascontiguousarray(slot_mp, np.int32).reshape(n, self.S)
        obs_mp = None if obs_mp is None else np.ascontiguousarray(obs_mp, np.int32).reshape(n, self.S)
        sort_key = None if sort_key is None else np.ascontiguousarray(sort_key, np.uint32)
        self._check(self.lib.mss_mirror_add_keyframes(self.handle, kf0, n, _p(sort_key), _p(n_slots), _p(cells), _p(slot_mp), _p(obs_mp)))

    def add_keyframe(self, kf, n_slots, cells, slot_mp, obs_mp=None, sort_key=None):
        cells = np.ascontiguousarray(cells, np.uint16)
        slot_mp = np.ascontiguousarray(slot_mp, np.int32)
        obs_mp = None if obs_mp is None else np.ascontiguousarray(obs_mp, np.int32)
        self._check(self.lib.mss_mirror_add_keyframe(self.handle, kf, kf if sort_key is None else sort_key, n_slots, _p(cells),
                                                     _p(slot_mp), _p(obs_mp)))

    def set_map_points(self, mp0, nobs, bad=None):
        nobs = np.ascontiguousarray(nobs, np.int32)
        bad = None if bad is None else np.ascontiguousarray(bad, np.uint8)
        self._check(self.lib.mss_mirror_set_map_points(self.handle, mp0, nobs.size, _p(nobs), _p(bad)))

    def apply(self, ops):
        """ops: structured array (OP_DTYPE) or iterable of (kind, a, b, c); applied in order"""
        arr = ops if isinstance(ops, np.ndarray) and ops.dtype == OP_DTYPE else np.array([tuple(o) for o in ops], OP_DTYPE)
        arr = np.ascontiguousarray(arr)
        self._check(self.lib.mss_mirror_apply(self.handle, _p(arr), arr.size))

    def load(self, packed: dict):
        """bulk load of a dict made by oracle.mirror_model.load_view (tests / bench)"""
        self.add_keyframes(packed["kf0"], packed["n_slots"], packed["cells"], packed["slot_mp"], packed["obs_mp"])
        self.set_map_points(packed["mp0"], packed["nobs"])

    # ---- windows ----------------------------------------------------------------------------------------------
    def solve(self, windows, apply=False, want_arrays=True, want_handles=True, n_max_floor=0, raise_on_status=True):
        """windows: list of int32 arrays of keyframe handles (independent windows).  -> list of MirrorResult"""
        n = len(windows)
        st = self.stats()
        words = (st["n_map_points"] + 31) // 32
        cw = (mss_mirror_window * n)()
        cr = (mss_result * n)()
        keep_alive = []
        for i, kfs in enumerate(windows):
            kfs = np.ascontiguousarray(kfs, np.int32)
            del_bits = np.zeros(max(words, 1), np.uint32)
            cap = kfs.size * self.S
            mp_handle = np.full(max(cap, 1), -1, np.int32) if want_handles else None
            cw[i].K, cw[i].n_max_floor, cw[i].kf = kfs.size, n_max_floor, _p(kfs)
            cw[i].del_bits, cw[i].del_words, cw[i].apply = _p(del_bits), del_bits.size, 1 if apply else 0
            cw[i].mp_handle, cw[i].mp_cap = _p(mp_handle), 0 if mp_handle is None else mp_handle.size
            bufs = None
            if want_arrays:
                # sizes are not known before the call: room for the largest table the window can have
                bufs = (np.zeros((cap + 31) // 32 + 1, np.uint32), np.zeros(kfs.size + 4096, np.int32), np.zeros(kfs.size + 4096, np.int32))
                cr[i].keep_bits, cr[i].kf_cov, cr[i].kf_slack = (_p(b) for b in bufs)
            keep_alive.append((kfs, del_bits, mp_handle, bufs))
        rc = self.lib.mss_mirror_solve(self.handle, n, cw, cr)
        if raise_on_status or rc not in (MSS_OK, MSS_E_BADARG, MSS_E_NOCONVERGE):
            self._check(rc)
        out = []
        for i in range(n):
            kfs, del_bits, mp_handle, bufs = keep_alive[i]
            w, r = cw[i], cr[i]
            M, R = w.M, w.K + w.H
            if bufs is not None:
                kb, cov, sl = bufs[0][:(M + 31) // 32].copy(), bufs[1][:R].copy(), bufs[2][:R].copy()
            else:
                kb, cov, sl = np.zeros(0, np.uint32), np.zeros(0, np.int32), np.zeros(0, np.int32)
            res = Result(keep=unpack_bits(kb, M) if bufs is not None else np.zeros(0, bool), keep_bits=kb, kf_cov=cov, kf_slack=sl,
                         objective=r.objective, dual_bound=r.dual_bound, sum_cost=r.sum_cost, uncovered_cells=r.uncovered_cells,
                         total_slack=r.total_slack, n_max=r.n_max, n_vars=r.n_vars, n_cells=r.n_cells, nnz=r.nnz, n_kept=r.n_kept,
                         rounds=r.rounds, status=r.status, time_build_us=r.time_build_us, time_solve_us=r.time_solve_us)
            lo, hi = w.h_lo >> 5, (w.h_hi + 31) >> 5
            bits = unpack_bits(del_bits[lo:hi], (hi - lo) * 32) if hi > lo else np.zeros(0, bool)
            deleted = (np.nonzero(bits)[0] + lo * 32).astype(np.int32)
            out.append(MirrorResult(res, deleted, None if mp_handle is None else mp_handle[:M].copy(), w))
        return out

    def components(self, kfs):
        """(kf_label[K], ncomp, n_max) of one window"""
        kfs = np.ascontiguousarray(kfs, np.int32)
        w = mss_mirror_window()
        w.K, w.kf = kfs.size, _p(kfs)
        lab = np.zeros(max(kfs.size, 1), np.int32)
        nc, nm = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.mss_mirror_components(self.handle, C.byref(w), _p(lab), C.byref(nc), C.byref(nm)))
        return lab[:kfs.size], int(nc.value), int(nm.value)

    def build_view(self, kfs, n_max_floor=0):
        """the view the mirror assembles for one window: (PackedView, mp_handle[M], okf_handle[H])"""
        kfs = np.ascontiguousarray(kfs, np.int32)
        w = mss_mirror_window()
        w.K, w.n_max_floor, w.kf = kfs.size, n_max_floor, _p(kfs)
        sizes = np.zeros(5, np.int32)
        cap = max(kfs.size * self.S, 1)
        feat_ptr = np.zeros(kfs.size + 1, np.int32)
        slots = np.zeros(cap, np.uint32)
        nobs16 = np.zeros(cap, np.uint16)
        mp_handle = np.zeros(cap, np.int32)
        okf_total = np.zeros(4096, np.int32)
        okf_handle = np.zeros(4096, np.int32)
        # the pair count has no small a-priori bound: ask for the sizes first
        self._check(self.lib.mss_mirror_build_view(self.handle, C.byref(w), _p(sizes), None, None, None, None, None, None, None))
        pairs = np.zeros(max(int(sizes[4]), 1), np.uint32)
        self._check(self.lib.mss_mirror_build_view(self.handle, C.byref(w), _p(sizes), _p(feat_ptr), _p(slots), _p(nobs16), _p(pairs),
                                                   _p(okf_total), _p(mp_handle), _p(okf_handle)))
        K, H, M, F, O = (int(x) for x in sizes)
        view = PackedView(K=K, H=H, M=M, feat_ptr=feat_ptr, slots=slots[:F].copy(), mp_nobs16=nobs16[:M].copy(),
                          obs_pairs=pairs[:O].copy(), okf_total=okf_total[:H].copy(), meta=dict(packed=True), n_max_floor=n_max_floor)
        return view, mp_handle[:M].copy(), okf_handle[:H].copy()


def arrays_from_view(view, S=None, seed=0, kf0=0, mp0=0, shuffle=True):
    """A map that flattens to `view`: window keyframes get handles kf0..kf0+K-1, outside keyframes kf0+K.., map points
    mp0 + a random permutation of their table index (so the discovery numbering is not the identity), filler points behind
    them bring every outside keyframe to its GetNumberMPs().  Returns dict(S, n_slots, cells, slot_mp, obs_mp, nobs, bad,
    mp_of_table, window) ready for Mirror.load (tests, bench; oracle/mirror_model.py uses the same arrays)."""
    K, H, M = view.K, view.H, view.M
    rng = np.random.default_rng(seed)
    perm = rng.permutation(M) if shuffle else np.arange(M)
    mp_of = (mp0 + perm).astype(np.int64)                        # table index -> handle
    obs_mp_tab = np.repeat(np.arange(M, dtype=np.int64), np.diff(view.mp_obs_ptr))
    out = view.mp_obs_kf >= K
    cnt_out = np.bincount(view.mp_obs_kf[out] - K, minlength=H) if H else np.zeros(0, np.int64)
    n_win = np.diff(view.feat_ptr).astype(np.int64)
    n_out = np.maximum(cnt_out, view.okf_total.astype(np.int64)) if H else np.zeros(0, np.int64)
    if S is None:
        S = int(max(n_win.max(initial=1), n_out.max(initial=1)))
    n_slots = np.concatenate([n_win, n_out]).astype(np.int32)
    slot_mp = np.full((K + H, S), -1, np.int32)
    obs = np.full((K + H, S), -1, np.int32)
    cells = np.full((K + H, S), CELL_NONE, np.uint16)
    kf_of_slot = np.repeat(np.arange(K), n_win)
    idx_in_kf = np.arange(view.F) - np.repeat(view.feat_ptr[:-1].astype(np.int64), n_win)
    has = view.feat_mp >= 0
    slot_mp[kf_of_slot[has], idx_in_kf[has]] = mp_of[view.feat_mp[has]]
    obs[kf_of_slot[has], idx_in_kf[has]] = mp_of[view.feat_mp[has]]
    cells[kf_of_slot, idx_in_kf] = view.feat_cell
    next_mp = mp0 + M
    fill_nobs = []
    for j in range(H):
        mps = obs_mp_tab[out][view.mp_obs_kf[out] - K == j]
        tot = int(view.okf_total[j])
        # GetNumberMPs() must come out as okf_total[j]: the first `tot` observers sit in a slot, further observers (a view may
        # list more observers than valid slots) only observe; filler points bring the slots up to `tot`
        obs[K + j, :mps.size] = mp_of[mps]
        slot_mp[K + j, :min(mps.size, tot)] = mp_of[mps[:tot]]
        extra = tot - mps.size
        if extra > 0:
            slot_mp[K + j, mps.size:tot] = np.arange(next_mp, next_mp + extra)
            obs[K + j, mps.size:tot] = np.arange(next_mp, next_mp + extra)
            next_mp += extra
            fill_nobs += [3] * extra
    nobs = np.zeros(next_mp - mp0, np.int32)
    nobs[perm] = view.mp_nobs
    nobs[M:] = fill_nobs
    return dict(S=S, n_slots=n_slots, cells=cells, slot_mp=slot_mp, obs_mp=obs, nobs=nobs, mp_of_table=mp_of.astype(np.int32),
                window=np.arange(kf0, kf0 + K, dtype=np.int32), kf0=kf0, mp0=mp0, n_mp=int(next_mp - mp0))

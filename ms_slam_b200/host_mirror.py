"""ctypes binding of libmss_host.so: the C++ ``ORB_SLAM3::MapSparsification`` mirror (ms_slam_b200/host/) driven through
its extern "C" harness.  Test / benchmark infrastructure for the drop-in boundary: a flat window
view is turned back into an ORB-SLAM3-shaped pointer graph, the sparsifier thread is run over it exactly the way
System / LocalMapping / LoopClosing drive it upstream, and the observable side effects are read back as arrays."""
from __future__ import annotations

import ctypes as C
import os
import tempfile
import numpy as np

from .window import WindowView

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libmss_host.so")
SYMBOLS = ["msh_create", "msh_destroy", "msh_engine_ready", "msh_build_world", "msh_flatten_only", "msh_snapshot_sizes",
           "msh_snapshot_copy", "msh_start", "msh_feed", "msh_nonlocal_after", "msh_forwarded_count", "msh_wait_forwarded",
           "msh_stop_handshake", "msh_consume", "msh_finish", "msh_bad_flags", "msh_forwarded_ids", "msh_keyframe_state",
           "msh_map_counts", "msh_reports", "msh_set_min_points", "msh_flatten_us", "msh_reports2", "msh_mirror_active"]
_lib = None


def load_library(path: str = LIB_PATH):
    global _lib
    if _lib is None:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found: build it with `python __graft_entry__.py`")
        lib = C.CDLL(path)
        lib.msh_create.argtypes = [C.c_char_p, C.c_int]
        lib.msh_create.restype = C.c_void_p
        for name in SYMBOLS[1:]:
            getattr(lib, name).restype = C.c_int
        lib.msh_destroy.restype = None
        lib.msh_snapshot_sizes.restype = None
        lib.msh_snapshot_copy.restype = None
        lib.msh_bad_flags.restype = None
        lib.msh_keyframe_state.restype = None
        lib.msh_snapshot_packed_sizes.restype = None
        lib.msh_snapshot_packed_copy.restype = None
        _lib = lib
    return _lib


def write_settings(path, N=100, lam=500.0, grid_lam=10.0, window_length=30, non_local=30):
    """The Sparsification.* block of an MS-SLAM settings yaml (Examples/Stereo/KITTI00-02.yaml:69-74)."""
    with open(path, "w") as f:
        f.write("%YAML:1.0\n# Parameters for sliding window map sparsification\n"
                f"Sparsification.N: {N}\nSparsification.Lambda: {lam}\nSparsification.GridLambda: {grid_lam}\n"
                f"Sparsification.WindowLength: {window_length}\n# Threshold to determine non-local keyframes\n"
                f"Sparsification.NonLocalKF: {non_local}\n")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class World:
    """One Atlas + LoopClosing + MapSparsification, populated from a WindowView."""

    def __init__(self, view: WindowView, N=100, lam=500.0, grid_lam=10.0, window_length=None, inertial=False, non_local=30,
                 mirror=None, batched_handback=None, devices=None, dual_bound=None):
        """mirror: None = the class default (device mirror on when a GPU is present), False = flatten every window from the
        pointer graph (MSS_MIRROR=0), True = default.  batched_handback=False: per-point SetBadFlag like the reference."""
        self.lib = load_library()
        self.view = view
        self.tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
        self.tmp.close()
        write_settings(self.tmp.name, N, lam, grid_lam, window_length if window_length is not None else max(view.K, 1), non_local)
        env = {"MSS_MIRROR": None if mirror is None else ("1" if mirror else "0"),
               "MSS_BATCHED_HANDBACK": None if batched_handback is None else ("1" if batched_handback else "0"),
               "MSS_DEVICES": None if devices is None else str(devices),
               "MSS_DUAL_BOUND": None if dual_bound is None else ("1" if dual_bound else "0")}
        saved = {k: os.environ.get(k) for k in env}
        for k, v in env.items():
            if v is not None:
                os.environ[k] = v
        try:
            self.h = C.c_void_p(self.lib.msh_create(self.tmp.name.encode(), 1 if inertial else 0))
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        self.lib.msh_build_world(self.h, view.K, view.H, view.M, _p(view.feat_ptr), _p(view.feat_mp), _p(view.feat_cell),
                                 _p(view.mp_nobs), _p(view.mp_obs_ptr), _p(view.mp_obs_kf), _p(view.okf_total))

    def close(self):
        if self.h:
            self.lib.msh_destroy(self.h)
            self.h = None
            os.unlink(self.tmp.name)

    def engine_ready(self):
        return bool(self.lib.msh_engine_ready(self.h))

    def flatten_only(self):
        self.lib.msh_flatten_only(self.h)
        return self.snapshot(0)

    def flatten_ms(self, reps=3):
        """host time of FlattenWindow (incl. packing into the pinned blob) on this world, best of reps, in ms"""
        best = None
        for _ in range(reps):
            self.lib.msh_flatten_only(self.h)
            t = self.lib.msh_flatten_us(self.h) / 1000.0
            best = t if best is None else min(best, t)
        return best

    def snapshot(self, which=1):
        """(WindowView, mp_ids, okf_ids, is_var) of the last flattened window"""
        sz = np.zeros(5, np.int32)
        self.lib.msh_snapshot_sizes(self.h, which, _p(sz))
        K, H, M, F, O = (int(x) for x in sz)
        a = dict(feat_ptr=np.zeros(K + 1, np.int32), feat_mp=np.zeros(F, np.int32), feat_cell=np.zeros(F, np.uint16),
                 mp_nobs=np.zeros(M, np.int32), mp_obs_ptr=np.zeros(M + 1, np.int32), mp_obs_kf=np.zeros(O, np.int32),
                 okf_total=np.zeros(H, np.int32))
        mp_ids, okf_ids, is_var = np.zeros(M, np.int64), np.zeros(H, np.int64), np.zeros(M, np.uint8)
        self.lib.msh_snapshot_copy(self.h, which, _p(a["feat_ptr"]), _p(a["feat_mp"]), _p(a["feat_cell"]), _p(a["mp_nobs"]),
                                   _p(a["mp_obs_ptr"]), _p(a["mp_obs_kf"]), _p(a["okf_total"]), _p(mp_ids), _p(okf_ids), _p(is_var))
        return WindowView(K=K, H=H, **a), mp_ids, okf_ids, is_var.astype(bool)

    def snapshot_packed(self, which=1):
        """The packed transport blob FlattenWindow built for the last window: dict(tok_ptr, tokens, nobs16, pairs, blob_bytes)
        or None when the window did not fit the packed ranges"""
        sz = np.zeros(4, np.int64)
        self.lib.msh_snapshot_packed_sizes(self.h, which, _p(sz))
        if not sz[2]:
            return None
        s5 = np.zeros(5, np.int32)
        self.lib.msh_snapshot_sizes(self.h, which, _p(s5))
        K, M = int(s5[0]), int(s5[2])
        out = dict(tok_ptr=np.zeros(K + 1, np.int32), tokens=np.zeros(int(sz[0]), np.uint16), nobs16=np.zeros(M, np.uint16),
                   pairs=np.zeros(int(sz[1]), np.uint32), blob_bytes=int(sz[3]))
        self.lib.msh_snapshot_packed_copy(self.h, which, _p(out["tok_ptr"]), _p(out["tokens"]), _p(out["nobs16"]), _p(out["pairs"]))
        out["nobs8"] = bool(self.lib.msh_snapshot_nobs8(self.h, which))       # the blob carries one byte per map point
        return out

    # the calls the rest of the SLAM system makes
    def start(self): return self.lib.msh_start(self.h)
    def feed(self, first, count): return self.lib.msh_feed(self.h, first, count)
    def wait_forwarded(self, n, timeout_ms=20000): return self.lib.msh_wait_forwarded(self.h, n, timeout_ms)
    def stop_handshake(self, timeout_ms=20000): return self.lib.msh_stop_handshake(self.h, timeout_ms)
    def consume(self): return self.lib.msh_consume(self.h)
    def finish(self, timeout_ms=60000): return self.lib.msh_finish(self.h, timeout_ms)
    def nonlocal_after(self, k, max_updates=1000): return self.lib.msh_nonlocal_after(self.h, k, max_updates)
    def set_min_points(self, n): return self.lib.msh_set_min_points(self.h, n)

    def bad_flags(self):
        out = np.zeros(self.view.M, np.uint8)
        self.lib.msh_bad_flags(self.h, _p(out))
        return out.astype(bool)

    def forwarded_ids(self):
        out = np.zeros(4 * (self.view.K + self.view.H) + 16, np.int64)
        n = self.lib.msh_forwarded_ids(self.h, _p(out), out.size)
        return out[:min(n, out.size)].tolist()

    def keyframe_state(self):
        out = np.zeros(3 * (self.view.K + self.view.H), np.int32)
        self.lib.msh_keyframe_state(self.h, _p(out))
        return out.reshape(-1, 3)

    def map_counts(self):
        out = np.zeros(3, np.int64)
        self.lib.msh_map_counts(self.h, _p(out))
        return dict(map_points=int(out[0]), sparsified_map_points=int(out[1]), sparsified_keyframes=int(out[2]))

    def reports(self):
        out = np.zeros(20 * 64, np.float64)
        n = min(self.lib.msh_reports2(self.h, _p(out), 64), 64)
        keys = ["status", "K", "H", "M", "n_vars", "n_kept", "n_deleted", "rounds", "objective", "flatten_ms", "solve_ms", "apply_ms",
                "components", "mirror", "delta_ops", "build_ms", "h2d_bytes", "d2h_bytes", "devices", "dual_bound"]
        return [dict(zip(keys, out[20 * i:20 * i + 20].tolist())) for i in range(n)]

    def mirror_active(self):
        return bool(self.lib.msh_mirror_active(self.h))

"""ctypes binding of the BoW re-transform + keyframe database (include/mss.h mss_voc_* / mss_bow_transform / mss_kfdb_*,
SURVEY.md section 8 f4).  Python host side for tests and benchmarks; no compute here."""
from __future__ import annotations

import ctypes as C
import numpy as np

from .engine import Engine, MssError


class mss_bow_keyframe(C.Structure):
    _fields_ = [("n", C.c_int32), ("kf_id", C.c_int32), ("descriptors", C.c_void_p), ("word", C.c_void_p), ("node", C.c_void_p),
                ("bow_word", C.c_void_p), ("bow_value", C.c_void_p), ("n_bow", C.c_void_p), ("fv_node", C.c_void_p),
                ("fv_feature", C.c_void_p), ("n_fv", C.c_void_p)]


def _declare(lib):
    if getattr(lib, "_bow_declared", False):
        return
    lib.mss_voc_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mss_voc_destroy.argtypes = [C.c_void_p]
    lib.mss_voc_destroy.restype = None
    lib.mss_voc_words.argtypes = [C.c_void_p]
    lib.mss_bow_transform.argtypes = [C.c_void_p, C.c_int32, C.POINTER(mss_bow_keyframe), C.c_int32, C.c_int32]
    lib.mss_kfdb_common_words.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.mss_kfdb_postings.argtypes = [C.c_void_p]
    lib._bow_declared = True


def synthetic_vocabulary(k=10, L=3, seed=0, stop_frac=0.05, ragged=True):
    """A vocabulary tree in DBoW2 text-file order (node 0 = root; parent[i] < i): complete k-ary tree of L levels, except
    that with `ragged` a few inner nodes have fewer children or are leaves early.  Random 256-bit descriptors, positive
    idf-like weights, a fraction of stopped words (weight 0).  -> dict(parent, is_leaf, desc, weight, k, L)"""
    rng = np.random.default_rng(seed)
    if not ragged:                                           # complete tree in breadth-first order (fast path for large trees)
        n = (k ** (L + 1) - 1) // (k - 1)
        parent = np.maximum((np.arange(n, dtype=np.int64) - 1) // k, 0).astype(np.int32)
        is_leaf = (np.arange(n) >= (k ** L - 1) // (k - 1)).astype(np.uint8)
        desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        weight = np.where(is_leaf == 1, rng.random(n) * 8.0 + 0.01, 0.0)
        weight[(is_leaf == 1) & (rng.random(n) < stop_frac)] = 0.0
        return dict(parent=parent, is_leaf=is_leaf, desc=desc, weight=weight.astype(np.float64), k=k, L=L)
    parent, level = [0], [0]
    frontier = [0]
    for lev in range(1, L + 1):
        nxt = []
        for p in frontier:
            nc = k
            if ragged and lev > 1:
                r = rng.random()
                nc = 0 if r < 0.03 else (int(rng.integers(1, k + 1)) if r < 0.15 else k)
            for _ in range(nc):
                parent.append(p); level.append(lev); nxt.append(len(parent) - 1)
        frontier = nxt
    n = len(parent)
    parent = np.array(parent, np.int32)
    has_child = np.zeros(n, bool)
    has_child[parent[1:]] = True
    is_leaf = (~has_child).astype(np.uint8)
    is_leaf[0] = 0
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    weight = np.where(is_leaf == 1, rng.random(n) * 8.0 + 0.01, 0.0)
    weight[(is_leaf == 1) & (rng.random(n) < stop_frac)] = 0.0
    # DBoW2 writes the nodes parent-first but not level by level: shuffle the file order among nodes whose parent precedes them
    return dict(parent=parent, is_leaf=is_leaf, desc=desc, weight=weight.astype(np.float64), k=k, L=L)


class Vocabulary:
    def __init__(self, engine: Engine, voc: dict):
        self.engine, self.lib = engine, engine.lib
        _declare(self.lib)
        self.handle = C.c_void_p()
        parent = np.ascontiguousarray(voc["parent"], np.int32)
        leaf = np.ascontiguousarray(voc["is_leaf"], np.uint8)
        desc = np.ascontiguousarray(voc["desc"], np.uint8)
        weight = np.ascontiguousarray(voc["weight"], np.float64)
        engine._check(self.lib.mss_voc_create(engine.handle, parent.size, int(voc["L"]), parent.ctypes.data, leaf.ctypes.data,
                                               desc.ctypes.data, weight.ctypes.data, C.byref(self.handle)))
        self.n_words = int(self.lib.mss_voc_words(self.handle))

    def close(self):
        if self.handle:
            self.lib.mss_voc_destroy(self.handle)
            self.handle = None

    def transform(self, device_descriptors, counts, kf_ids=None, levelsup=4, add_to_database=False):
        """device_descriptors: list of device pointers ([n][32] bytes each), counts: rows per keyframe.
        -> list of dict(word, node, bow_word, bow_value, fv_node, fv_feature)"""
        nkf = len(counts)
        arr = (mss_bow_keyframe * max(nkf, 1))()
        outs = []
        for q, (p, n) in enumerate(zip(device_descriptors, counts)):
            n = int(n)
            o = dict(word=np.zeros(max(n, 1), np.int32), node=np.zeros(max(n, 1), np.int32), bow_word=np.zeros(max(n, 1), np.int32),
                     bow_value=np.zeros(max(n, 1), np.float64), n_bow=np.zeros(1, np.int32), fv_node=np.zeros(max(n, 1), np.int32),
                     fv_feature=np.zeros(max(n, 1), np.int32), n_fv=np.zeros(1, np.int32))
            arr[q] = mss_bow_keyframe(n, q if kf_ids is None else int(kf_ids[q]), p, *[o[k].ctypes.data for k in
                                      ("word", "node", "bow_word", "bow_value", "n_bow", "fv_node", "fv_feature", "n_fv")])
            outs.append(o)
        self.engine._check(self.lib.mss_bow_transform(self.handle, nkf, arr, levelsup, 1 if add_to_database else 0))
        res = []
        for o, n in zip(outs, counts):
            nb, nf = int(o["n_bow"][0]), int(o["n_fv"][0])
            res.append(dict(word=o["word"][:n], node=o["node"][:n], bow_word=o["bow_word"][:nb], bow_value=o["bow_value"][:nb],
                            fv_node=o["fv_node"][:nf], fv_feature=o["fv_feature"][:nf]))
        return res

    def common_words(self, query_words, kf_cap):
        q = np.ascontiguousarray(query_words, np.int32)
        out = np.zeros(max(kf_cap, 1), np.int32)
        self.engine._check(self.lib.mss_kfdb_common_words(self.handle, q.size, q.ctypes.data, kf_cap, out.ctypes.data))
        return out[:kf_cap]

    def postings(self):
        return int(self.lib.mss_kfdb_postings(self.handle))

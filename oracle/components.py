"""ORACLE (test infrastructure, not product code) -- connected components of a window, on the CPU.

The reference has no such step: its final flush hands every unsparsified keyframe to ONE GUROBI model
(/root/reference/src/MapSparsification.cc:38-47).  That model is block-diagonal along the connected components of the
graph whose nodes are the model's rows' keyframes (window keyframes :119-122 with their cell rows :111-116, outside
keyframes :146-150) and its variables (:91-99), with an edge wherever a variable has a coefficient in a keyframe's rows.
This file computes those components with scipy.sparse.csgraph and labels them the way libmss's mss_components does
(dense ids in order of the first keyframe row of each component; -1 for map points that are not variables), so the GPU
labels can be compared exactly.  PARITY UNPINNED upstream (no counterpart, no fixtures); pinned here by hand-built cases.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components

CELL_NONE = 0xFFFF


def components(view):
    """-> (row_label[K+H] int32, mp_label[M] int32, ncomp, n_max)"""
    K, H, M = view.K, view.H, view.M
    R = K + H
    feat_kf = np.repeat(np.arange(K, dtype=np.int64), np.diff(view.feat_ptr))
    valid = view.feat_mp >= 0
    n_max = int(view.mp_nobs[view.feat_mp[valid]].max()) if valid.any() else 0
    grid = valid & (view.feat_cell != CELL_NONE)
    e_row = feat_kf[grid]
    e_var = view.feat_mp[grid].astype(np.int64)
    isvar = np.zeros(M, bool)
    isvar[e_var] = True
    obs_mp = np.repeat(np.arange(M, dtype=np.int64), np.diff(view.mp_obs_ptr))
    om = (view.mp_obs_kf >= K) & isvar[obs_mp]
    rows = np.concatenate([e_row, view.mp_obs_kf[om].astype(np.int64)])
    cols = np.concatenate([e_var, obs_mp[om]]) + R
    n = R + M
    g = sp.coo_matrix((np.ones(rows.size, np.int8), (rows, cols)), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    # canonical ids: rank of the smallest ROW index of the component among all components that contain a row
    first_row = np.full(lab.max() + 1 if n else 0, n, np.int64)
    np.minimum.at(first_row, lab[:R], np.arange(R))
    has_row = first_row < n
    order = np.argsort(first_row, kind="stable")
    dense = np.full(first_row.size, -1, np.int64)
    dense[order[:int(has_row.sum())]] = np.arange(int(has_row.sum()))
    row_label = dense[lab[:R]].astype(np.int32)
    mp_label = np.where(isvar, dense[lab[R:]], -1).astype(np.int32)
    return row_label, mp_label, int(has_row.sum()), n_max

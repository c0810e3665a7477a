"""ORACLE (test infrastructure, not product code) -- numpy model of the persistent device mirror (include/mss.h
"Persistent device mirror", ms_slam_b200/csrc/mss_mirror.{cu,cuh}).

It keeps the same keyframe-major arrays on the CPU, applies the same ops with sequential semantics, and assembles a
window view by restating what the reference's passes read (/root/reference/src/MapSparsification.cc):
  :67-76, :82-107   valid slots of the window keyframes in window / slot order (non-null, not bad); a map point is
                    numbered at its first appearance (mnIndexForSparsification, :91-99); it is a variable when one of
                    its slots lies in a grid cell
  :125-142          observations of the variables by keyframes that are not in the window -> outside keyframes
  :146              GetNumberMPs() of every outside keyframe (src/KeyFrame.cc:286-297)
SetBadFlag (src/MapPoint.cc:227-255) and EraseBadDescriptor (src/KeyFrame.cc:311-361) are restated for the apply / compact
paths.  PARITY UNPINNED upstream (the reference has no such structure); the model is pinned against the host FlattenWindow
(tests/test_mirror.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

from ms_slam_b200.window import WindowView, CELL_NONE

MOP_SLOT, MOP_OBS, MOP_MP, MOP_KF_COMPACT = 1, 2, 3, 4


class MirrorModel:
    def __init__(self, slots_per_kf: int):
        self.S = int(slots_per_kf)
        self.slot_mp = np.zeros((0, self.S), np.int32)
        self.obs_mp = np.zeros((0, self.S), np.int32)
        self.slot_cell = np.zeros((0, self.S), np.uint16)
        self.kf_n = np.zeros(0, np.int32)
        self.kf_key = np.zeros(0, np.uint32)
        self.mp_nobs = np.zeros(0, np.int32)
        self.mp_bad = np.zeros(0, bool)

    # ---- storage ------------------------------------------------------------------------------------------
    def _kfs(self, n):
        if n > self.kf_n.size:
            add = n - self.kf_n.size
            self.slot_mp = np.vstack([self.slot_mp, np.full((add, self.S), -1, np.int32)])
            self.obs_mp = np.vstack([self.obs_mp, np.full((add, self.S), -1, np.int32)])
            self.slot_cell = np.vstack([self.slot_cell, np.full((add, self.S), CELL_NONE, np.uint16)])
            self.kf_n = np.concatenate([self.kf_n, np.zeros(add, np.int32)])
            self.kf_key = np.concatenate([self.kf_key, np.zeros(add, np.uint32)])

    def _mps(self, n):
        if n > self.mp_nobs.size:
            add = n - self.mp_nobs.size
            self.mp_nobs = np.concatenate([self.mp_nobs, np.zeros(add, np.int32)])
            self.mp_bad = np.concatenate([self.mp_bad, np.zeros(add, bool)])

    def add_keyframes(self, kf0, sort_key, n_slots, cells, slot_mp, obs_mp=None):
        n = len(n_slots)
        self._kfs(kf0 + n)
        slot_mp = np.asarray(slot_mp, np.int32).reshape(n, self.S)
        obs = np.full((n, self.S), -1, np.int32) if obs_mp is None else np.asarray(obs_mp, np.int32).reshape(n, self.S)
        self._mps(int(max(slot_mp.max(initial=-1), obs.max(initial=-1))) + 1)
        self.slot_mp[kf0:kf0 + n] = slot_mp
        self.obs_mp[kf0:kf0 + n] = obs
        self.slot_cell[kf0:kf0 + n] = np.asarray(cells, np.uint16).reshape(n, self.S)
        self.kf_n[kf0:kf0 + n] = n_slots
        self.kf_key[kf0:kf0 + n] = np.arange(kf0, kf0 + n) if sort_key is None else sort_key

    def set_map_points(self, mp0, nobs, bad=None):
        n = len(nobs)
        self._mps(mp0 + n)
        self.mp_nobs[mp0:mp0 + n] = nobs
        self.mp_bad[mp0:mp0 + n] = False if bad is None else np.asarray(bad, bool)

    def apply(self, ops):
        """ops: iterable of (kind, a, b, c), applied in order"""
        for kind, a, b, c in ops:
            if kind == MOP_SLOT:
                self._kfs(a + 1); self._mps(c + 1)
                self.slot_mp[a, b] = c
                self.kf_n[a] = max(self.kf_n[a], b + 1)
            elif kind == MOP_OBS:
                self._kfs(a + 1); self._mps(c + 1)
                self.obs_mp[a, b] = c
            elif kind == MOP_MP:
                self._mps(a + 1)
                self.mp_nobs[a] = b
                self.mp_bad[a] = bool(c)
            elif kind == MOP_KF_COMPACT:
                self.compact(a)
            else:
                raise ValueError(kind)

    def compact(self, kf):
        """KeyFrame::EraseBadDescriptor: non-empty slots kept in order, every kept point observes kf at its new index"""
        n = int(self.kf_n[kf])
        kept = self.slot_mp[kf, :n][self.slot_mp[kf, :n] >= 0]
        self.slot_mp[kf] = -1
        self.obs_mp[kf] = -1
        self.slot_mp[kf, :kept.size] = kept
        self.obs_mp[kf, :kept.size] = kept
        self.slot_cell[kf] = CELL_NONE
        self.kf_n[kf] = kept.size

    def delete(self, handles):
        """MapPoint::SetBadFlag for every handle: bad, observations dropped, the observed slots emptied"""
        handles = np.asarray(handles, np.int64)
        self.mp_bad[handles] = True
        hit = np.isin(self.obs_mp, handles)
        self.slot_mp[hit] = -1
        self.obs_mp[hit] = -1

    # ---- window assembly -------------------------------------------------------------------------------------
    def build(self, kfs, n_max_floor=0):
        """-> (WindowView in the compact transport form with discovery-order numbering, mp_handle[M], okf_handle[H])"""
        kfs = np.asarray(kfs, np.int64)
        K = kfs.size
        S = self.S
        sl = self.slot_mp[kfs] if K else np.zeros((0, S), np.int32)
        cell = self.slot_cell[kfs] if K else np.zeros((0, S), np.uint16)
        inuse = np.arange(S)[None, :] < self.kf_n[kfs][:, None] if K else np.zeros((0, S), bool)
        valid = inuse & (sl >= 0)
        valid[valid] &= ~self.mp_bad[sl[valid]]
        flat_h = sl[valid].astype(np.int64)                      # window order, slot order
        flat_c = cell[valid]
        uniq, first_pos = np.unique(flat_h, return_index=True)
        order = np.argsort(first_pos, kind="stable")
        mp_handle = uniq[order]                                  # discovery order
        loc = np.full(self.mp_nobs.size, -1, np.int64)
        loc[mp_handle] = np.arange(mp_handle.size)
        feat_mp = loc[flat_h]
        feat_ptr = np.zeros(K + 1, np.int64)
        feat_ptr[1:] = np.cumsum(valid.sum(axis=1))
        M = mp_handle.size
        isvar = np.zeros(M, bool)
        isvar[feat_mp[flat_c != CELL_NONE]] = True
        # outside observations of the variables
        in_win = np.zeros(self.kf_n.size, bool)
        in_win[kfs] = True
        okf_rows, okf_cols = np.nonzero(self.obs_mp >= 0)
        h = self.obs_mp[okf_rows, okf_cols].astype(np.int64)
        sel = (loc[h] >= 0) & ~in_win[okf_rows]
        sel[sel] &= isvar[loc[h[sel]]]
        p_mp, p_kf = loc[h[sel]], okf_rows[sel]
        okf = np.unique(p_kf)
        okf = okf[np.lexsort((okf, self.kf_key[okf]))]
        jidx = np.full(self.kf_n.size, -1, np.int64)
        jidx[okf] = np.arange(okf.size)
        o = np.lexsort((jidx[p_kf], p_mp))                       # grouped by map point
        p_mp, p_kf = p_mp[o], p_kf[o]
        obs_ptr = np.zeros(M + 1, np.int64)
        obs_ptr[1:] = np.cumsum(np.bincount(p_mp, minlength=M))
        tot = np.zeros(okf.size, np.int64)
        for j, kf in enumerate(okf):
            s = self.slot_mp[kf, :self.kf_n[kf]]
            s = s[s >= 0]
            tot[j] = int(np.count_nonzero(~self.mp_bad[s]))
        view = WindowView(K=K, H=int(okf.size), feat_ptr=feat_ptr, feat_mp=feat_mp, feat_cell=flat_c,
                          mp_nobs=self.mp_nobs[mp_handle], mp_obs_ptr=obs_ptr, mp_obs_kf=K + jidx[p_kf], okf_total=tot,
                          n_max_floor=n_max_floor)
        return view, mp_handle.astype(np.int32), okf.astype(np.int32)


from ms_slam_b200.mirror import arrays_from_view as load_view      # noqa: E402,F401  (test helper: a map that flattens to a view)


def erase_bad_descriptor_rows(keep, descriptors=None, keypoints=None, uright=None, depth=None):
    """KeyFrame::EraseBadDescriptor (/root/reference/src/KeyFrame.cc:311-361) restated for the per-keypoint arrays: the rows
    whose slot still holds a map point survive, in order (DescriptorsNew.push_back(mDescriptors.row(i)), vKeysUnNew,
    vuRightNew, vDepthNew at :331-342).  keep: bool [n]; descriptors u8 [n,32]; keypoints [n,7] 32-bit words
    (cv::KeyPoint = pt.x, pt.y, size, angle, response, octave, class_id); uright / depth f32 [n].  Returns the new arrays."""
    keep = np.asarray(keep, bool)
    return tuple(None if a is None else np.asarray(a)[keep].copy() for a in (descriptors, keypoints, uright, depth))

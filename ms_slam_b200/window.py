"""Flattened (SoA) view of one sparsification window.

This is the host-side mirror of ``mss_window_view`` in ``include/mss.h``: the pointer graph the
reference walks in ``MapSparsification::Sparsifying`` (/root/reference/src/MapSparsification.cc:58-151)
-- ``KeyFrame::mvpMapPoints`` + ``KeyFrame::mGrid`` on the keyframe side, ``MapPoint::mObservations`` /
``MapPoint::nObs`` on the map-point side -- snapshotted once into plain arrays.

Conventions:
  * KF table = K window keyframes followed by H "outside" keyframes (keyframes that are not in the window
    but observe at least one window map point, MapSparsification.cc:125-151).
  * ``feat_ptr[K+1]``  slot ranges of the window keyframes (a slot is one entry of mvpMapPoints).
  * ``feat_mp[F]``     map-point table index held by the slot, -1 = empty or bad (MapSparsification.cc:70,90).
  * ``feat_cell[F]``   image-grid cell ``col*48+row`` of the slot's keypoint (64x48 grid, include/Frame.h:44-45),
                       0xFFFF = keypoint not in the grid (Frame::PosInGrid false, src/Frame.cc:657-668).
  * ``mp_nobs[M]``     MapPoint::Observations() (stereo observation counts 2, src/MapPoint.cc:155-158).
  * ``mp_obs_ptr[M+1]``/``mp_obs_kf[O]``  MapPoint::mObservations as KF-table indices (>= K means outside KF).
  * ``okf_total[H]``   KeyFrame::GetNumberMPs() of each outside keyframe (src/KeyFrame.cc:286-297).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

GRID_COLS = 64          # include/Frame.h:44
GRID_ROWS = 48          # include/Frame.h:45
N_CELLS = GRID_COLS * GRID_ROWS
CELL_NONE = 0xFFFF


@dataclass
class WindowView:
    K: int
    H: int
    feat_ptr: np.ndarray      # int32 [K+1]
    feat_mp: np.ndarray       # int32 [F]
    feat_cell: np.ndarray     # uint16 [F]
    mp_nobs: np.ndarray       # int32 [M]
    mp_obs_ptr: np.ndarray    # int32 [M+1]
    mp_obs_kf: np.ndarray     # int32 [O]
    okf_total: np.ndarray     # int32 [H]
    kf_gid: np.ndarray = None  # uint64 [K+H]
    mp_gid: np.ndarray = None  # uint64 [M]
    meta: dict = field(default_factory=dict)
    n_max_floor: int = 0       # nMax is at least this (component of a larger window, include/mss.h)

    def __post_init__(self):
        c = np.ascontiguousarray
        self.feat_ptr = c(self.feat_ptr, dtype=np.int32)
        self.feat_mp = c(self.feat_mp, dtype=np.int32)
        self.feat_cell = c(self.feat_cell, dtype=np.uint16)
        self.mp_nobs = c(self.mp_nobs, dtype=np.int32)
        self.mp_obs_ptr = c(self.mp_obs_ptr, dtype=np.int32)
        self.mp_obs_kf = c(self.mp_obs_kf, dtype=np.int32)
        self.okf_total = c(self.okf_total, dtype=np.int32)
        if self.kf_gid is None:
            self.kf_gid = np.arange(self.K + self.H, dtype=np.uint64)
        if self.mp_gid is None:
            self.mp_gid = np.arange(self.M, dtype=np.uint64)
        self.kf_gid = c(self.kf_gid, dtype=np.uint64)
        self.mp_gid = c(self.mp_gid, dtype=np.uint64)

    # sizes ---------------------------------------------------------------------------------------------
    @property
    def F(self) -> int:
        return int(self.feat_mp.shape[0])

    @property
    def M(self) -> int:
        return int(self.mp_nobs.shape[0])

    @property
    def O(self) -> int:
        return int(self.mp_obs_kf.shape[0])

    def input_bytes(self) -> int:
        """Bytes the engine has to read from the view (the H2D payload when the view lives on the host)."""
        return int(self.feat_ptr.nbytes + self.feat_mp.nbytes + self.feat_cell.nbytes + self.mp_nobs.nbytes
                   + self.mp_obs_ptr.nbytes + self.mp_obs_kf.nbytes + self.okf_total.nbytes)

    def compact(self) -> "WindowView":
        """The same window in the compact transport form the C++ FlattenWindow emits: empty / bad slots (feat_mp == -1) and
        observations by window keyframes (mp_obs_kf < K, which the engine never reads) are omitted.  Map-point and keyframe
        numbering is unchanged, so results are directly comparable (and bit-identical)."""
        keep = self.feat_mp >= 0
        kf = np.repeat(np.arange(self.K, dtype=np.int64), np.diff(self.feat_ptr))
        feat_ptr = np.zeros(self.K + 1, np.int64)
        feat_ptr[1:] = np.cumsum(np.bincount(kf[keep], minlength=self.K))
        out = self.mp_obs_kf >= self.K
        mp = np.repeat(np.arange(self.M, dtype=np.int64), np.diff(self.mp_obs_ptr))
        obs_ptr = np.zeros(self.M + 1, np.int64)
        obs_ptr[1:] = np.cumsum(np.bincount(mp[out], minlength=self.M))
        v = WindowView(K=self.K, H=self.H, feat_ptr=feat_ptr, feat_mp=self.feat_mp[keep], feat_cell=self.feat_cell[keep],
                       mp_nobs=self.mp_nobs, mp_obs_ptr=obs_ptr, mp_obs_kf=self.mp_obs_kf[out], okf_total=self.okf_total,
                       kf_gid=self.kf_gid, mp_gid=self.mp_gid, n_max_floor=self.n_max_floor)
        v.meta = dict(self.meta, compact=True)
        return v

    def discovery_order(self) -> "WindowView":
        """The same window with the map-point table renumbered the way the C++ FlattenWindow (and the reference, through
        mnIndexForSparsification, /root/reference/src/MapSparsification.cc:91-99) numbers it: in order of first appearance
        when the keyframes are walked in window order and their slots in slot order.  Map points that sit in no slot of a
        window keyframe (they cannot be reached by that walk) keep their relative order behind all others.  Returns a new
        view; `meta["mp_perm"][new] = old` maps results back."""
        valid = self.feat_mp >= 0
        first = np.full(self.M, np.iinfo(np.int64).max, np.int64)
        pos = np.nonzero(valid)[0]
        np.minimum.at(first, self.feat_mp[pos], pos)
        perm = np.argsort(first, kind="stable")             # new -> old
        inv = np.empty(self.M, np.int64)
        inv[perm] = np.arange(self.M)
        feat_mp = np.where(valid, inv[np.maximum(self.feat_mp, 0)], -1).astype(self.feat_mp.dtype)
        cnt = np.diff(self.mp_obs_ptr)[perm]
        obs_ptr = np.zeros(self.M + 1, self.mp_obs_ptr.dtype)
        obs_ptr[1:] = np.cumsum(cnt)
        src = np.repeat(self.mp_obs_ptr[:-1][perm] - obs_ptr[:-1], cnt) + np.arange(int(obs_ptr[-1]))
        v = WindowView(K=self.K, H=self.H, feat_ptr=self.feat_ptr, feat_mp=feat_mp, feat_cell=self.feat_cell,
                       mp_nobs=self.mp_nobs[perm], mp_obs_ptr=obs_ptr, mp_obs_kf=self.mp_obs_kf[src], okf_total=self.okf_total,
                       kf_gid=self.kf_gid, mp_gid=self.mp_gid[perm] if self.mp_gid is not None and len(self.mp_gid) == self.M else self.mp_gid,
                       n_max_floor=self.n_max_floor)
        v.meta = dict(self.meta, discovery_order=True, mp_perm=perm)
        return v

    def validate(self) -> None:
        K, H, F, M, O = self.K, self.H, self.F, self.M, self.O
        if K < 0 or H < 0:
            raise ValueError("negative K/H")
        if self.feat_ptr.shape != (K + 1,) or self.feat_ptr[0] != 0 or int(self.feat_ptr[-1]) != F:
            raise ValueError("feat_ptr must be [K+1], start at 0 and end at F")
        if np.any(np.diff(self.feat_ptr) < 0):
            raise ValueError("feat_ptr must be non-decreasing")
        if self.feat_cell.shape != (F,):
            raise ValueError("feat_cell/feat_mp length mismatch")
        if F and (self.feat_mp.min() < -1 or self.feat_mp.max() >= M):
            raise ValueError("feat_mp out of range")
        ok = (self.feat_cell < N_CELLS) | (self.feat_cell == CELL_NONE)
        if not bool(np.all(ok)):
            raise ValueError("feat_cell must be < 3072 or 0xFFFF")
        if self.mp_obs_ptr.shape != (M + 1,) or self.mp_obs_ptr[0] != 0 or int(self.mp_obs_ptr[-1]) != O:
            raise ValueError("mp_obs_ptr must be [M+1], start at 0 and end at O")
        if np.any(np.diff(self.mp_obs_ptr) < 0):
            raise ValueError("mp_obs_ptr must be non-decreasing")
        if O and (self.mp_obs_kf.min() < 0 or self.mp_obs_kf.max() >= K + H):
            raise ValueError("mp_obs_kf out of range")
        if self.okf_total.shape != (H,):
            raise ValueError("okf_total must be [H]")

    # io ------------------------------------------------------------------------------------------------
    _ARRAYS = ("feat_ptr", "feat_mp", "feat_cell", "mp_nobs", "mp_obs_ptr", "mp_obs_kf", "okf_total",
               "kf_gid", "mp_gid")

    def save(self, path: str) -> None:
        np.savez_compressed(path, K=self.K, H=self.H, **{n: getattr(self, n) for n in self._ARRAYS})

    @classmethod
    def load(cls, path: str) -> "WindowView":
        z = np.load(path)
        return cls(K=int(z["K"]), H=int(z["H"]), **{n: z[n] for n in cls._ARRAYS})


def split_components(view: WindowView, row_label, mp_label, n_max: int):
    """Independent sub-windows of a window, from the labels mss_components returns (include/mss.h).  Only components that
    contain a window keyframe become windows (an outside keyframe that sees no variable constrains nothing).  Every part
    carries the window-wide nMax as n_max_floor.  Returns [(part, kf_idx, mp_idx)]: part.K keyframes = view keyframes kf_idx
    (in order), part map points = view map points mp_idx (in order; non-variables that only count for nMax are dropped,
    nMax being carried explicitly)."""
    row_label = np.asarray(row_label)
    mp_label = np.asarray(mp_label)
    K, H = view.K, view.H
    feat_kf = np.repeat(np.arange(K, dtype=np.int64), np.diff(view.feat_ptr))
    obs_mp = np.repeat(np.arange(view.M, dtype=np.int64), np.diff(view.mp_obs_ptr))
    parts = []
    for c in np.unique(row_label[:K]):
        kf_idx = np.nonzero(row_label[:K] == c)[0]
        okf_idx = np.nonzero(row_label[K:] == c)[0]
        mp_idx = np.nonzero(mp_label == c)[0]
        mp_new = np.full(view.M, -1, np.int64)
        mp_new[mp_idx] = np.arange(mp_idx.size)
        kf_new = np.full(K + H, -1, np.int64)
        kf_new[kf_idx] = np.arange(kf_idx.size)
        kf_new[K + okf_idx] = kf_idx.size + np.arange(okf_idx.size)
        sl = np.isin(feat_kf, kf_idx)
        f_mp = view.feat_mp[sl].astype(np.int64)
        f_mp = np.where(f_mp >= 0, mp_new[np.maximum(f_mp, 0)], -1)      # slots of non-variables become empty
        feat_ptr = np.zeros(kf_idx.size + 1, np.int64)
        feat_ptr[1:] = np.cumsum(np.diff(view.feat_ptr)[kf_idx])
        ob = (mp_new[obs_mp] >= 0) & (kf_new[view.mp_obs_kf] >= 0)
        cnt = np.bincount(mp_new[obs_mp[ob]], minlength=mp_idx.size)
        obs_ptr = np.zeros(mp_idx.size + 1, np.int64)
        obs_ptr[1:] = np.cumsum(cnt)
        part = WindowView(K=int(kf_idx.size), H=int(okf_idx.size), feat_ptr=feat_ptr, feat_mp=f_mp, feat_cell=view.feat_cell[sl],
                          mp_nobs=view.mp_nobs[mp_idx], mp_obs_ptr=obs_ptr, mp_obs_kf=kf_new[view.mp_obs_kf[ob]],
                          okf_total=view.okf_total[okf_idx], kf_gid=np.concatenate([view.kf_gid[kf_idx], view.kf_gid[K + okf_idx]]),
                          mp_gid=view.mp_gid[mp_idx], n_max_floor=int(n_max))
        parts.append((part, kf_idx, mp_idx))
    return parts


def merge_views(views, interleave: bool = True) -> WindowView:
    """One window made of several independent ones (disjoint map points and keyframes): what a final flush over an atlas
    with several covisibility components looks like.  With interleave the window keyframes of the inputs alternate, so
    components are not contiguous keyframe ranges.  Map points are concatenated in input order."""
    Ks = [v.K for v in views]
    order = []                                            # (view, local kf) in merged window order
    if interleave:
        for i in range(max(Ks) if Ks else 0):
            order += [(a, i) for a, v in enumerate(views) if i < v.K]
    else:
        for a, v in enumerate(views):
            order += [(a, i) for i in range(v.K)]
    K = len(order)
    mp_base = np.concatenate([[0], np.cumsum([v.M for v in views])]).astype(np.int64)
    okf_base = K + np.concatenate([[0], np.cumsum([v.H for v in views])]).astype(np.int64)
    kf_pos = [np.full(v.K, -1, np.int64) for v in views]
    for pos, (a, i) in enumerate(order):
        kf_pos[a][i] = pos
    feat_ptr, feat_mp, feat_cell = [0], [], []
    for a, i in order:
        v = views[a]
        s, e = int(v.feat_ptr[i]), int(v.feat_ptr[i + 1])
        m = v.feat_mp[s:e].astype(np.int64)
        feat_mp.append(np.where(m >= 0, m + mp_base[a], -1))
        feat_cell.append(v.feat_cell[s:e])
        feat_ptr.append(feat_ptr[-1] + (e - s))
    obs_ptr, obs_kf = [np.zeros(1, np.int64)], []
    off = 0
    for a, v in enumerate(views):
        kf = v.mp_obs_kf.astype(np.int64)
        obs_kf.append(np.where(kf < v.K, kf_pos[a][np.minimum(kf, max(v.K - 1, 0))] if v.K else kf, kf - v.K + okf_base[a]))
        obs_ptr.append(v.mp_obs_ptr[1:].astype(np.int64) + off)
        off += v.O
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return WindowView(K=K, H=int(sum(v.H for v in views)), feat_ptr=np.asarray(feat_ptr), feat_mp=cat(feat_mp, np.int32),
                      feat_cell=cat(feat_cell, np.uint16), mp_nobs=cat([v.mp_nobs for v in views], np.int32),
                      mp_obs_ptr=np.concatenate(obs_ptr), mp_obs_kf=cat(obs_kf, np.int32),
                      okf_total=cat([v.okf_total for v in views], np.int32))


SLOT_CELL_NONE = 0xFFF        # include/mss.h MSS_SLOT_CELL_NONE
SLOT_EMPTY = 0xFFFFFFFF       # include/mss.h MSS_SLOT_EMPTY


@dataclass
class PackedView:
    """MSS_LAYOUT_PACKED form of a window (include/mss.h): u32 slots = (map point << 12) | cell, u16 tables.  Same window,
    same map-point numbering, about 0.6x the bytes of the SoA form."""
    K: int
    H: int
    M: int
    feat_ptr: np.ndarray      # int32 [K+1]
    slots: np.ndarray         # uint32 [F] (MSS_LAYOUT_PACKED) or uint16 tokens [F] (MSS_LAYOUT_PACKED16, see pack_view)
    mp_nobs16: np.ndarray     # uint16 [M]; uint8 [M] when meta["nobs8"] (mss_window_view::nobs8: every Observations() <= 255)
    obs_pairs: np.ndarray     # uint32 [O] (map point << 12) | outside keyframe j (KF-table index K + j)
    okf_total: np.ndarray     # int32 [H]
    meta: dict = field(default_factory=dict)
    n_max_floor: int = 0
    mp_tie: np.ndarray = None  # uint32 [M] tie-break rank (mss_window_view::mp_tie), None = ties break on the table index

    @property
    def F(self) -> int:
        return int(self.slots.shape[0])

    @property
    def O(self) -> int:
        return int(self.obs_pairs.shape[0])

    def input_bytes(self) -> int:
        return int(self.feat_ptr.nbytes + self.slots.nbytes + self.mp_nobs16.nbytes + self.obs_pairs.nbytes + self.okf_total.nbytes)


def tie_ranks(v) -> np.ndarray:
    """uint32 [M]: rank of every map point's gid (mss_window_view::mp_tie: lower wins a tie) -- the same for every numbering
    of the same window"""
    r = np.empty(v.M, np.uint32)
    r[np.argsort(v.mp_gid, kind="stable")] = np.arange(v.M, dtype=np.uint32)
    return r


def _tokens16(slots_sorted, feat_ptr):
    """u32 slots, sorted inside every keyframe -> (u16 tokens, token feat_ptr) of MSS_LAYOUT_PACKED16 (include/mss.h)"""
    K = feat_ptr.size - 1
    n = slots_sorted.size
    kf = np.repeat(np.arange(K, dtype=np.int64), np.diff(feat_ptr))
    mp = (slots_sorted >> 12).astype(np.int64)
    cell = (slots_sorted & 0xFFF).astype(np.int64)
    first = np.zeros(n, bool)
    first[feat_ptr[:-1][np.diff(feat_ptr) > 0]] = True
    prev = np.concatenate([[0], mp[:-1]])
    gap = np.where(first, mp, mp - prev)                   # the running index starts at 0 in every keyframe
    units = gap // 15                                      # advanced by escape tokens, 15 * (low + 1) each, low <= 4095
    nesc = (units + 4095) // 4096
    ntok = nesc + 1
    start = np.concatenate([[0], np.cumsum(ntok)])
    tok = np.zeros(int(start[-1]), np.uint16)
    tok[start[1:] - 1] = (((gap % 15) << 12) | cell).astype(np.uint16)          # the slot token closes its group
    for j in range(int(nesc.max()) if n else 0):           # j-th escape of the slots that need more than j
        sel = nesc > j
        u = np.minimum(units[sel] - 4096 * j, 4096)
        tok[start[:-1][sel] + j] = ((15 << 12) | (u - 1)).astype(np.uint16)
    tptr = np.zeros(K + 1, np.int64)
    tptr[1:] = np.cumsum(np.bincount(kf, weights=ntok, minlength=K)).astype(np.int64)
    return tok, tptr


def pack_view(v: WindowView, sort_slots: bool = False, tokens16: bool = False, nobs8=None, tie: bool = False) -> PackedView:
    """WindowView -> PackedView.  Raises ValueError when the window exceeds the packed form's ranges.
    sort_slots: order the slots of every keyframe by map-point index (the order of the slots inside a keyframe carries no
    meaning for the model; sorted, the 32 entries a warp handles touch neighbouring map points, which turns the state
    gathers of the row phases into nearly coalesced accesses).  FlattenWindow emits this order.
    nobs8: Observations() as one byte per map point; None = what FlattenWindow does: with tokens16 whenever every value fits.
    tie: carry the tie-break ranks of the map points' gids (mp_tie)."""
    if v.M > (1 << 20) or v.H > 4095:
        raise ValueError("window too large for the packed layout")
    if v.M and int(v.mp_nobs.max()) > 65535:
        raise ValueError("Observations() above 65535")
    mp = v.feat_mp.astype(np.int64)
    cell = np.where(v.feat_cell == CELL_NONE, SLOT_CELL_NONE, v.feat_cell).astype(np.int64)
    slots = np.where(mp >= 0, (mp << 12) | cell, SLOT_EMPTY).astype(np.uint32)
    if (sort_slots or tokens16) and slots.size:
        kf = np.repeat(np.arange(v.K, dtype=np.int64), np.diff(v.feat_ptr))
        slots = slots[np.lexsort((slots, kf))]
    owner = np.repeat(np.arange(v.M, dtype=np.int64), np.diff(v.mp_obs_ptr))
    outside = v.mp_obs_kf >= v.K                     # observations by window keyframes are not part of the pair list
    pairs = ((owner[outside] << 12) | (v.mp_obs_kf[outside].astype(np.int64) - v.K)).astype(np.uint32)
    if nobs8 is None:
        nobs8 = bool(tokens16 and (v.M == 0 or int(v.mp_nobs.max()) <= 255))
    if nobs8 and v.M and int(v.mp_nobs.max()) > 255:
        raise ValueError("Observations() above 255: nobs8 does not apply")
    nobs = np.ascontiguousarray(v.mp_nobs.astype(np.uint8 if nobs8 else np.uint16))
    if tokens16:
        # MSS_LAYOUT_PACKED16: valid slots only, sorted by map point inside every keyframe, delta-coded as u16 tokens
        kf = np.repeat(np.arange(v.K, dtype=np.int64), np.diff(v.feat_ptr))
        ok = slots != SLOT_EMPTY
        ptr = np.zeros(v.K + 1, np.int64)
        ptr[1:] = np.cumsum(np.bincount(kf[ok], minlength=v.K))
        tok, tptr = _tokens16(slots[ok], ptr)
        return PackedView(K=v.K, H=v.H, M=v.M, feat_ptr=np.ascontiguousarray(tptr, dtype=np.int32), slots=np.ascontiguousarray(tok),
                          mp_nobs16=nobs, obs_pairs=np.ascontiguousarray(pairs),
                          okf_total=v.okf_total, meta=dict(v.meta, packed=True, tokens16=True, nobs8=bool(nobs8)), n_max_floor=v.n_max_floor,
                          mp_tie=tie_ranks(v) if tie else None)
    return PackedView(K=v.K, H=v.H, M=v.M, feat_ptr=v.feat_ptr, slots=np.ascontiguousarray(slots),
                      mp_nobs16=nobs, obs_pairs=np.ascontiguousarray(pairs),
                      okf_total=v.okf_total, meta=dict(v.meta, packed=True, nobs8=bool(nobs8)), n_max_floor=v.n_max_floor,
                      mp_tie=tie_ranks(v) if tie else None)


def make_view(K, kf_slots, mp_nobs, outside=None, okf_total=None) -> WindowView:
    """Small helper for hand-written windows (tests, fixtures).

    kf_slots : list (len K) of lists of (mp_index or -1, cell or None) per slot.
    outside  : list (len H) of lists of mp indices observed by each outside keyframe.
    """
    outside = outside or []
    H = len(outside)
    M = len(mp_nobs)
    feat_ptr = [0]
    feat_mp, feat_cell = [], []
    obs = [[] for _ in range(M)]
    for k, slots in enumerate(kf_slots):
        for (p, cell) in slots:
            feat_mp.append(-1 if p is None else p)
            feat_cell.append(CELL_NONE if cell is None else cell)
            if p is not None and p >= 0 and k not in obs[p]:
                obs[p].append(k)
        feat_ptr.append(len(feat_mp))
    for j, plist in enumerate(outside):
        for p in plist:
            obs[p].append(K + j)
    ptr = np.zeros(M + 1, np.int32)
    ptr[1:] = np.cumsum([len(o) for o in obs])
    flat = np.array([k for o in obs for k in o], np.int32)
    if okf_total is None:
        okf_total = [len(pl) for pl in outside]
    return WindowView(K=K, H=H, feat_ptr=np.array(feat_ptr), feat_mp=np.array(feat_mp, np.int32),
                      feat_cell=np.array(feat_cell, np.uint16), mp_nobs=np.array(mp_nobs, np.int32),
                      mp_obs_ptr=ptr, mp_obs_kf=flat, okf_total=np.array(okf_total, np.int32))

"""msgen-v1: seeded synthetic sparsification windows.

The reference ships no windows, fixtures or benchmark inputs for MapSparsification::Sparsifying
(/root/reference/src/MapSparsification.cc:58-171), so BASELINE.json's configs are realised by this generator.
It produces exactly the flattened view the C-ABI consumes (``window.WindowView``); all randomness comes from
``numpy.random.default_rng(seed)`` (PCG64), so a (config, seed) pair names one window bit-for-bit.

Shapes follow the reference's shipped settings: 64x48 feature grid (include/Frame.h:44-45), cell of a keypoint
= round((u-minX)*64/W), round((v-minY)*48/H) rejected outside the grid (src/Frame.cc:657-668), nFeatures and
image sizes from Examples/Stereo/KITTI00-02.yaml, Examples/Stereo/EuRoC.yaml, Examples/Stereo-Inertial/4season.yaml,
nObs counting a stereo observation twice (src/MapPoint.cc:155-158).
"""
from __future__ import annotations

import numpy as np

from .window import WindowView, GRID_COLS, GRID_ROWS, CELL_NONE

# BASELINE.json `configs`, in order
CONFIGS = {
    # 20 KF x 2000 MP random-visibility plumbing window (reference CPU path config)
    "c1": dict(K=20, M=2000, kind="random", p_vis=0.2, n_feat=1200, width=752, height=480, N=100, H=0),
    # KITTI-00-shaped north-star window
    "c2": dict(K=500, M=200_000, kind="banded", tau=4.0, rho=0.02, n_feat=2000, width=1241, height=376,
               N=100, H=40),
    # EuRoC-MH-shaped dense-overlap window
    "c3": dict(K=100, M=30_000, kind="banded", tau=4.0, rho=0.5, n_feat=1200, width=752, height=480,
               N=75, H=20),
    # one window of the 64-window multi-GPU batch (seeds 1000+w)
    "c4": dict(K=100, M=30_000, kind="banded", tau=4.0, rho=0.02, n_feat=2000, width=1241, height=376,
               N=100, H=0),
    # 4Seasons-shaped stress window
    "c5": dict(K=2000, M=1_000_000, kind="banded", tau=3.0, rho=0.02, n_feat=2000, width=800, height=400,
               N=100, H=40),
    # live-window-sized case (WindowLength 30, Examples/Stereo/KITTI00-02.yaml:72)
    "live": dict(K=30, M=6000, kind="banded", tau=5.0, rho=0.02, n_feat=2000, width=1241, height=376,
                 N=100, H=10),
}
LAMBDA = 500.0       # Sparsification.Lambda (Examples/Stereo/KITTI00-02.yaml:70)
GRID_LAMBDA = 10.0   # Sparsification.GridLambda (:71)


def _cells(u, v, width, height):
    """Frame::PosInGrid (src/Frame.cc:657-668) with minX=minY=0."""
    px = np.rint(u * (GRID_COLS / float(width))).astype(np.int64)
    py = np.rint(v * (GRID_ROWS / float(height))).astype(np.int64)
    inside = (px >= 0) & (px < GRID_COLS) & (py >= 0) & (py < GRID_ROWS)
    cell = np.where(inside, px * GRID_ROWS + py, CELL_NONE)
    return cell.astype(np.uint16)


def generate(K, M, kind="banded", tau=4.0, rho=0.02, p_vis=0.2, n_feat=2000, width=1241, height=376,
             N=100, H=0, seed=0, gid_base=0, lam=LAMBDA) -> WindowView:
    rng = np.random.default_rng(seed)
    Hb = H // 2                       # halo keyframes before / after the window on the timeline
    T = K + H                         # timeline: t in [0,T), window = [Hb, Hb+K)

    # ---- observations (p, t) ----------------------------------------------------------------------------
    if kind == "random":
        vis = rng.random((K, M)) < p_vis
        kk, pp = np.nonzero(vis)
        obs_p = pp.astype(np.int64)
        obs_t = kk.astype(np.int64) + Hb
        # independent pixel per observation
        u = rng.random(obs_p.size) * width
        v = rng.random(obs_p.size) * height
    else:
        def runs(n, start_lo, start_hi):
            length = 1 + rng.geometric(1.0 / max(tau - 1.0, 1.0), size=n)      # >= 2, mean tau
            start = rng.integers(start_lo, start_hi, size=n)
            return start, length

        start, length = runs(M, 0, T - (H - Hb))           # start anywhere up to the last window KF
        miss = start + length <= Hb                         # entirely inside the leading halo
        if miss.any():
            start[miss] = rng.integers(Hb, Hb + K, size=int(miss.sum()))
        p_ids = np.arange(M, dtype=np.int64)
        revisit = rng.random(M) < rho
        r_start, r_len = runs(int(revisit.sum()), Hb, Hb + K)
        all_p = np.concatenate([p_ids, p_ids[revisit]])
        all_s = np.concatenate([start, r_start])
        all_l = np.concatenate([length, r_len])
        all_l = np.minimum(all_l, T - all_s)
        obs_p = np.repeat(all_p, all_l)
        run_first = np.cumsum(all_l) - all_l
        within = np.arange(obs_p.size) - np.repeat(run_first, all_l)
        obs_t = np.repeat(all_s, all_l) + within
        # pixel: uniform first observation, sigma = 20 px random walk afterwards (per run)
        u0 = rng.random(all_p.size) * width
        v0 = rng.random(all_p.size) * height
        du = rng.normal(0.0, 20.0, obs_p.size)
        dv = rng.normal(0.0, 20.0, obs_p.size)
        du[run_first] = 0.0
        dv[run_first] = 0.0
        cu, cv = np.cumsum(du), np.cumsum(dv)
        u = np.repeat(u0, all_l) + cu - np.repeat(cu[run_first], all_l)
        v = np.repeat(v0, all_l) + cv - np.repeat(cv[run_first], all_l)
        # a revisit run may overlap the first run: keep one observation per (p, t)
        key = obs_p * T + obs_t
        _, first_idx = np.unique(key, return_index=True)
        first_idx.sort()
        obs_p, obs_t, u, v = obs_p[first_idx], obs_t[first_idx], u[first_idx], v[first_idx]

    # ---- slot cap: at most n_feat observations per keyframe, random victims --------------------------------
    order = np.lexsort((rng.random(obs_p.size), obs_t))
    obs_p, obs_t, u, v = obs_p[order], obs_t[order], u[order], v[order]
    t_first = np.searchsorted(obs_t, np.arange(T))
    rank = np.arange(obs_p.size) - t_first[obs_t]
    keep = rank < n_feat
    obs_p, obs_t, u, v, rank = obs_p[keep], obs_t[keep], u[keep], v[keep], rank[keep]

    in_win = (obs_t >= Hb) & (obs_t < Hb + K)
    # KF-table index: window keyframes 0..K-1, then the outside (halo) keyframes K..K+H-1
    # (before-halo t < Hb -> K + t;   after-halo t >= Hb+K -> K + Hb + (t - Hb - K) = t)
    kf_tab = np.where(in_win, obs_t - Hb, np.where(obs_t < Hb, K + obs_t, obs_t))

    # ---- window keyframe slots ----------------------------------------------------------------------------
    F = K * n_feat
    feat_mp = np.full(F, -1, np.int32)
    feat_cell = np.full(F, CELL_NONE, np.uint16)
    perm = np.argsort(rng.random((K, n_feat)), axis=1)          # slot permutation per keyframe
    wk = kf_tab[in_win]
    slot = perm[wk, rank[in_win]]
    pos = wk * n_feat + slot
    feat_mp[pos] = obs_p[in_win]
    feat_cell[pos] = _cells(u[in_win], v[in_win], width, height)
    feat_ptr = (np.arange(K + 1, dtype=np.int64) * n_feat).astype(np.int32)

    # ---- map-point table ------------------------------------------------------------------------------------
    o2 = np.lexsort((kf_tab, obs_p))
    mp_obs_kf = kf_tab[o2].astype(np.int32)
    cnt = np.bincount(obs_p, minlength=M)
    mp_obs_ptr = np.zeros(M + 1, np.int64)
    mp_obs_ptr[1:] = np.cumsum(cnt)
    stereo = rng.random(M) < 0.8
    mp_nobs = np.maximum(np.where(stereo, 2, 1) * cnt, 3).astype(np.int32)

    # ---- outside keyframes: total valid points >= the points they share with the window ------------------------
    okf_total = np.zeros(H, np.int32)
    if H:
        # shared count over *variables* (map points reachable through a grid cell of a window keyframe)
        is_var = np.zeros(M, bool)
        ok = (feat_mp >= 0) & (feat_cell != CELL_NONE)
        is_var[feat_mp[ok]] = True
        out = ~in_win
        shared = np.bincount(kf_tab[out][is_var[obs_p[out]]] - K, minlength=H)
        hi = np.maximum(shared, n_feat // 2)
        okf_total = (shared + (rng.random(H) * (hi - shared + 1)).astype(np.int64)).astype(np.int32)
        okf_total = np.maximum(okf_total, 1)

    # generator contract: no point may cost more than a unit of keyframe slack
    valid = feat_mp >= 0
    n_max = int(mp_nobs[feat_mp[valid]].max()) if valid.any() else 0
    assert n_max - 3 < lam, "msgen-v1 contract: max c_p < Lambda"

    view = WindowView(K=K, H=H, feat_ptr=feat_ptr, feat_mp=feat_mp, feat_cell=feat_cell, mp_nobs=mp_nobs,
                      mp_obs_ptr=mp_obs_ptr.astype(np.int32), mp_obs_kf=mp_obs_kf, okf_total=okf_total,
                      kf_gid=np.uint64(gid_base) + np.arange(K + H, dtype=np.uint64),
                      mp_gid=np.uint64(gid_base) + np.arange(M, dtype=np.uint64))
    view.meta = dict(N=N, seed=seed, kind=kind, n_feat=n_feat)
    return view


def make_config(name: str, seed: int = 0, **override):
    """Returns (view, N) for one of BASELINE.json's configs."""
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    view = generate(seed=seed, **cfg)
    return view, int(cfg["N"])

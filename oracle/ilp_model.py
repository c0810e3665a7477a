"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference's sparsification ILP.

PARITY UNPINNED BY UPSTREAM: the reference (fishmarch/MS-SLAM @ e4730ec) has no tests, fixtures or golden vectors
for this path, and the arithmetic of the solve lives in GUROBI 10.0.2 (closed source, not vendored; pinned only by
/root/reference/cmake_modules/FindGUROBI.cmake:25,39).  GUROBI, OpenCV and Eigen are absent here, so the reference
cannot be compiled (oracle/_ref does not exist).  This file restates the model the reference hands to GUROBI,
line by line from /root/reference/src/MapSparsification.cc:58-171 (formalised in SURVEY.md Appendix A), and solves it
with HiGHS (scipy.optimize.milp) at the reference's own MIPGap = 0.002 (:155-156).  The oracle itself is pinned by
(i) brute-force enumeration on micro windows and (ii) the hand-derived known-answer fixtures in tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Model:
    nMax = max nObs over valid slots of window keyframes                                  (MapSparsification.cc:66-76)
    c_p  = (float)(nMax - nObs_p)      one binary x_p per distinct valid MP met through a grid cell     (:91-99)
    cell row   sum_{p in cell} x_p + g >= 1,  g binary, cost GridLambda                                  (:111-116)
    KF row     sum_p a_kp x_p + t_k >= N,     t_k integer in [0,1000], cost Lambda                       (:119-122)
    outside    sum_{p in V observed by j} x_p + u_j >= (float)cnt_j/(float)total_j*N, u_j as t_k         (:125-151)
    minimise   sum c_p x_p + GridLambda sum g + Lambda (sum t + sum u)                                   (:153)
    delete p iff X_p <= 0                                                                               (:159-166)
"""
from __future__ import annotations

from dataclasses import dataclass
import time
import numpy as np
import scipy.sparse as sp
from scipy.optimize import milp, LinearConstraint, Bounds

N_CELLS = 64 * 48         # include/Frame.h:44-45
CELL_NONE = 0xFFFF
SLACK_UB = 1000           # MapSparsification.cc:119,148
MIP_GAP = float(np.float32(0.0020))   # MapSparsification.cc:155-156 ((double)(float)0.0020)


def outside_need(cnt, total, N):
    """Canonical integer rhs of the outside rows.

    Reference: ``float nMini = (float)cnt / nTotal * mnMinNum`` (MapSparsification.cc:146-147) used as the rhs of a
    row whose lhs is integral -> ceil, with a 1e-5 guard for fp32 round-up.  total == 0 cannot occur for a keyframe
    that observes a valid map point; it is mapped to need = 0.
    """
    cnt = np.asarray(cnt, np.float32)
    total = np.asarray(total, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (cnt / total).astype(np.float32) * np.float32(N)
    r = np.where(np.asarray(total) > 0, r.astype(np.float64), 0.0)
    return np.ceil(r - 1e-5).astype(np.int64)


@dataclass
class Model:
    """Incidence structure of one window (what GUROBI is given), in arrays."""
    n_max: int
    var_mp: np.ndarray      # [V] MP-table index of each variable, discovery order (MapSparsification.cc:91-99)
    mp_var: np.ndarray      # [M] variable index or -1
    cost: np.ndarray        # [V] float64, integral values
    ent_var: np.ndarray     # [Z] variable of each (kf, cell, mp) incidence (valid grid-listed slot)
    ent_kf: np.ndarray      # [Z] window keyframe
    ent_cell: np.ndarray    # [Z] cell-row id (0..G-1)
    G: int
    K: int
    H: int
    out_var: np.ndarray     # [E] variable of each outside incidence
    out_kf: np.ndarray      # [E] outside keyframe (0..H-1)
    out_cnt: np.ndarray     # [H] cnt_j
    out_need: np.ndarray    # [H] need_j (integer rhs)
    avail: np.ndarray       # [K] sum_p a_kp = number of valid grid-listed slots


def build_model(view, N) -> Model:
    K, H = view.K, view.H
    feat_kf = np.repeat(np.arange(K, dtype=np.int64), np.diff(view.feat_ptr))
    valid = view.feat_mp >= 0
    n_max = int(view.mp_nobs[view.feat_mp[valid]].max()) if valid.any() else 0
    n_max = max(n_max, int(getattr(view, "n_max_floor", 0)))      # component of a larger window: window-wide nMax (mss.h)
    grid = valid & (view.feat_cell != CELL_NONE)
    idx = np.nonzero(grid)[0]
    e_kf = feat_kf[idx]
    e_cell_raw = view.feat_cell[idx].astype(np.int64)
    e_mp = view.feat_mp[idx].astype(np.int64)
    # discovery order of the reference: keyframes in window order, cells col-major (= increasing col*48+row),
    # features in grid-list order (= increasing slot index, Frame::AssignFeaturesToGrid pushes in index order)
    order = np.lexsort((idx, e_cell_raw, e_kf))
    e_kf, e_cell_raw, e_mp = e_kf[order], e_cell_raw[order], e_mp[order]
    M = view.M
    mp_var = np.full(M, -1, np.int64)
    uniq, first = np.unique(e_mp, return_index=True)
    disc = np.argsort(first, kind="stable")
    var_mp = uniq[disc]
    mp_var[var_mp] = np.arange(var_mp.size)
    cost = (n_max - view.mp_nobs[var_mp]).astype(np.float32).astype(np.float64)
    cell_key = e_kf * N_CELLS + e_cell_raw
    ukeys, ent_cell = np.unique(cell_key, return_inverse=True)
    avail = np.bincount(e_kf, minlength=K).astype(np.int64)
    # outside rows
    obs_mp = np.repeat(np.arange(M, dtype=np.int64), np.diff(view.mp_obs_ptr))
    om = (view.mp_obs_kf >= K) & (mp_var[obs_mp] >= 0)
    out_var = mp_var[obs_mp[om]]
    out_kf = view.mp_obs_kf[om].astype(np.int64) - K
    out_cnt = np.bincount(out_kf, minlength=H).astype(np.int64)
    out_need = outside_need(out_cnt, view.okf_total, N)
    out_need = np.where(out_cnt > 0, out_need, 0)
    return Model(n_max=n_max, var_mp=var_mp, mp_var=mp_var, cost=cost, ent_var=mp_var[e_mp], ent_kf=e_kf,
                 ent_cell=ent_cell.astype(np.int64), G=int(ukeys.size), K=K, H=H, out_var=out_var, out_kf=out_kf,
                 out_cnt=out_cnt, out_need=out_need, avail=avail)


# --------------------------------------------------------------------------------------------------------------
# penalty form F(x) and row checks
# --------------------------------------------------------------------------------------------------------------
def keep_to_x(model: Model, keep_mask) -> np.ndarray:
    """keep_mask: bool/0-1 array over the MP table -> x over variables."""
    return np.asarray(keep_mask)[model.var_mp].astype(np.int64)


def coverage(model: Model, x):
    x = np.asarray(x).astype(np.int64)
    kf_cov = np.bincount(model.ent_kf, weights=x[model.ent_var], minlength=model.K).astype(np.int64)
    cell_cov = np.bincount(model.ent_cell, weights=x[model.ent_var], minlength=model.G).astype(np.int64)
    out_cov = np.bincount(model.out_kf, weights=x[model.out_var], minlength=model.H).astype(np.int64)
    return kf_cov, cell_cov, out_cov


def objective(model: Model, x, N, lam, grid_lam, parts=False):
    """F(x): slack variables eliminated at their optimum (valid while N - cov <= 1000)."""
    x = np.asarray(x).astype(np.int64)
    kf_cov, cell_cov, out_cov = coverage(model, x)
    pts = float(np.dot(model.cost, x))
    cells = int(np.count_nonzero(cell_cov == 0))
    t = np.maximum(0, N - kf_cov)
    u = np.maximum(0, model.out_need - out_cov)
    F = pts + grid_lam * cells + lam * (int(t.sum()) + int(u.sum()))
    if parts:
        return F, dict(points=pts, n_kept=int(x.sum()), uncovered_cells=cells, kf_slack=t, out_slack=u,
                       kf_cov=kf_cov, out_cov=out_cov)
    return F


def rows_satisfied(model: Model, x, N):
    """Every coverage row at its best attainable level (valid when max c_p < Lambda)."""
    kf_cov, _, out_cov = coverage(model, x)
    ok_kf = kf_cov >= np.minimum(N, model.avail)
    ok_out = out_cov >= np.minimum(model.out_need, model.out_cnt)
    return bool(ok_kf.all() and ok_out.all()), ok_kf, ok_out


# --------------------------------------------------------------------------------------------------------------
# the model as the reference states it (explicit slack variables), solved with HiGHS
# --------------------------------------------------------------------------------------------------------------
def assemble(model: Model, N, lam, grid_lam):
    """Columns: [x (V) | g (G) | t (K) | u (Hn)], rows: [cells (G) | KFs (K) | outside (Hn)], all '>='."""
    V, G, K = model.var_mp.size, model.G, model.K
    hrows = np.nonzero(model.out_cnt > 0)[0]
    hmap = np.full(model.H, -1, np.int64)
    hmap[hrows] = np.arange(hrows.size)
    Hn = hrows.size
    ncol = V + G + K + Hn
    nrow = G + K + Hn
    Z, E = model.ent_var.size, model.out_var.size
    r = np.concatenate([model.ent_cell, np.arange(G), G + model.ent_kf, G + np.arange(K),
                        G + K + hmap[model.out_kf], G + K + np.arange(Hn)])
    c = np.concatenate([model.ent_var, V + np.arange(G), model.ent_var, V + G + np.arange(K),
                        model.out_var, V + G + K + np.arange(Hn)])
    A = sp.csr_matrix((np.ones(r.size), (r, c)), shape=(nrow, ncol))   # duplicates sum -> multiplicity a_kp
    obj = np.concatenate([model.cost, np.full(G, grid_lam), np.full(K, lam), np.full(Hn, lam)])
    rhs = np.concatenate([np.ones(G), np.full(K, float(N)), model.out_need[hrows].astype(np.float64)])
    ub = np.concatenate([np.ones(V + G), np.full(K + Hn, float(SLACK_UB))])
    return A, obj, rhs, ub, V


@dataclass
class Solution:
    status: int
    message: str
    objective: float        # solver objective (with slacks)
    x: np.ndarray           # [V] 0/1 (MILP) or fractional (LP)
    seconds: float
    assemble_seconds: float
    mip_gap: float | None = None
    dual_bound: float | None = None


def solve_ilp(view, N, lam, grid_lam, mip_rel_gap=MIP_GAP, time_limit=None, model=None) -> Solution:
    """The reference's solve: MILP at MIPGap 0.002 (HiGHS stand-in for GUROBI)."""
    t0 = time.perf_counter()
    model = model or build_model(view, N)
    A, obj, rhs, ub, V = assemble(model, N, lam, grid_lam)
    t1 = time.perf_counter()
    opts = {"mip_rel_gap": mip_rel_gap, "disp": False}
    if time_limit:
        opts["time_limit"] = float(time_limit)
    res = milp(obj, constraints=LinearConstraint(A, rhs, np.inf), bounds=Bounds(0, ub),
               integrality=np.ones(obj.size), options=opts)
    t2 = time.perf_counter()
    x = np.rint(res.x[:V]).astype(np.int64) if res.x is not None else None
    return Solution(status=res.status, message=res.message, objective=float(res.fun) if res.x is not None else np.nan,
                    x=x, seconds=t2 - t1, assemble_seconds=t1 - t0,
                    mip_gap=getattr(res, "mip_gap", None), dual_bound=getattr(res, "mip_dual_bound", None))


def solve_lp(view, N, lam, grid_lam, model=None) -> Solution:
    """LP relaxation: a lower bound on the ILP optimum (LP* <= ILP* <= F(x_hat))."""
    t0 = time.perf_counter()
    model = model or build_model(view, N)
    A, obj, rhs, ub, V = assemble(model, N, lam, grid_lam)
    t1 = time.perf_counter()
    res = milp(obj, constraints=LinearConstraint(A, rhs, np.inf), bounds=Bounds(0, ub),
               integrality=np.zeros(obj.size), options={"disp": False})
    t2 = time.perf_counter()
    return Solution(status=res.status, message=res.message, objective=float(res.fun), x=res.x[:V],
                    seconds=t2 - t1, assemble_seconds=t1 - t0)


def brute_force(view, N, lam, grid_lam, max_vars=20):
    """Enumerate all 2^V selections (micro windows only). Returns (F*, list of optimal x)."""
    model = build_model(view, N)
    V = model.var_mp.size
    if V > max_vars:
        raise ValueError("brute force is for micro windows only")
    best, arg = None, []
    for m in range(1 << V):
        x = (m >> np.arange(V)) & 1
        f = objective(model, x, N, lam, grid_lam)
        if best is None or f < best - 1e-9:
            best, arg = f, [x]
        elif abs(f - best) <= 1e-9:
            arg.append(x)
    return best, arg


def x_to_keep(view, model: Model, x) -> np.ndarray:
    """Reference read-out (MapSparsification.cc:159-166): delete variable MPs with X <= 0; others untouched."""
    keep = np.ones(view.M, bool)
    keep[model.var_mp[np.asarray(x) <= 0]] = False
    return keep

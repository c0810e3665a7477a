"""ORACLE (test infrastructure, not product code) -- a certified lower bound of the reference ILP without an LP solver.

Plan for the device-side `mss_result.dual_bound` (NaN today), restated on the CPU so that the next
step has its checker: the PROP rules of the device algorithm are exact dominance rules -- also for the LP relaxation of
the reference model (/root/reference/src/MapSparsification.cc:58-157) -- so every optimum agrees with
the state S reached by propagation alone on the decided variables, and

    ILP* >= LP* >= F_fix(S) + D_cells + D_rows

  F_fix   cost of the points propagation takes + GridLambda * cells with neither a taken nor an undecided point
          + Lambda * the part of every row's deficit that exceeds its undecided points
  D_cells value of a feasible dual of the cell rows of the residual problem (undecided points only), found by a few rounds
          of parallel dual ascent: every active cell raises its dual by min(GridLambda - z, min over its points of slack /
          #active cells of the point); a point's duals never exceed its cost, so all reduced costs stay >= 0
  D_rows  for every still deficient row the best single multiplier against the slack the cells left, split evenly over
          the point's deficient rows: max_y d'*y - sum_p max(0, y - share_p)   (the Lagrangian term of that row alone)

PARITY UNPINNED upstream (no counterpart in the reference).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

from . import emulate as em


def dual_bound(view, N, lam, grid_lam, rounds=30):
    """-> dict(bound, f_fix, d_cells, d_rows, n_free): bound <= LP* of the reference model"""
    E = em.Emulator(view, N, lam, grid_lam, stall_den=0)
    while True:
        changed, _ = E._prop_round()
        if changed == 0:
            break
    st = E.st
    nin_c, nfree_c, cov, free_r = E._stats()
    occupied = np.bincount(E.e_cell, minlength=E.K * em.N_CELLS) > 0
    u0 = int(np.count_nonzero(occupied & (nin_c == 0) & (nfree_c == 0)))
    d = np.maximum(0, E.need - cov)
    s0 = int(np.maximum(0, d - free_r).sum())
    f_fix = float(E.cost[st == em.IN].sum()) + grid_lam * u0 + lam * s0
    free = st == em.FREE
    M = E.M
    fe = free[E.e_var] & (nin_c[E.e_cell] == 0)
    cell_ids, cinv = np.unique(E.e_cell[fe], return_inverse=True)
    cvar = E.e_var[fe]
    nC = cell_ids.size
    z = np.zeros(nC)
    slack = E.cost.astype(float).copy()
    active = np.ones(nC, bool)
    for _ in range(rounds):
        if not active.any():
            break
        nact = np.bincount(cvar, weights=active[cinv], minlength=M)
        share = np.where(nact > 0, slack / np.maximum(nact, 1), np.inf)
        m = np.full(nC, np.inf)
        np.minimum.at(m, cinv, share[cvar])
        delta = np.maximum(np.where(active, np.minimum(grid_lam - z, m), 0.0), 0.0)
        z += delta
        slack = np.maximum(slack - np.bincount(cvar, weights=delta[cinv], minlength=M), 0.0)
        dead = np.zeros(nC, bool)
        np.logical_or.at(dead, cinv, (slack <= 1e-12)[cvar])
        active &= ~dead & (z < grid_lam - 1e-12)
    d_cells = float(z.sum())
    fr = free[E.r_var] & (d[E.r_row] > 0)
    rrow, rvar = E.r_row[fr], E.r_var[fr]
    dprime = np.minimum(d, free_r)
    d_rows = 0.0
    if rrow.size:
        share = slack[rvar] / np.maximum(np.bincount(rvar, minlength=M)[rvar], 1)
        order = np.lexsort((share, rrow))
        rr, sh = rrow[order], share[order]
        starts = np.flatnonzero(np.concatenate([[True], rr[1:] != rr[:-1]]))
        ends = np.concatenate([starts[1:], [rr.size]])
        for a, b in zip(starts, ends):
            dp = int(dprime[rr[a]])
            if dp <= 0:
                continue
            vals = sh[a:b]
            y = min(lam, vals[min(dp, vals.size) - 1])
            d_rows += dp * y - float(np.maximum(0.0, y - vals).sum())
    return dict(bound=f_fix + d_cells + d_rows, f_fix=f_fix, d_cells=d_cells, d_rows=d_rows, n_free=int(free.sum()))


# ---------------------------------------------------------------------------------------------------------------------------
# Integer twin of the DEVICE dual bound (ms_slam_b200/csrc/mss_bound.cuh, mss_result.dual_bound): same snapshot, same
# fixed-point arithmetic, so the device's counters can be compared bit for bit.
#
#   snapshot S  the state the live lists of the last PROP row phase before the FIRST greedy step describe (the state before
#               that round's variable phase): every decision in S was taken by an exact dominance rule
#   one round of dual ascent on the cell rows of the residual (share = floor(cost * 2^SC / #uncovered cells of the point)),
#   then the best single multiplier per deficient row against the slack the cells left (split over the point's deficient rows)
#   u0 (cells whose points are all rejected in S) is not counted at the snapshot but recovered at the end:
#   u0 = uncovered cells of the final selection - residual cells the final selection leaves uncovered
#   (points taken by dominance are never dropped again and rejected points never come back)
#
#   bound = cost_in + GridLambda * u0 + Lambda * s0 + (zsum + drows) / 2^SC          <= LP* <= ILP* <= F(x)
# ---------------------------------------------------------------------------------------------------------------------------
SC_BITS = 10
COST_CAP = (1 << 20) - 1


def device_twin(view, N, lam, grid_lam, **kw):
    """-> dict(flag, cost_in, s0, zsum, drows, res_unc, u0, bound, objective, n_free): what the device reports.
    flag 0: no certificate (round cap hit), 1: snapshot at the first greedy step, 2: propagation alone decided everything
    (the selection is optimal, bound = objective)"""
    E = em.Emulator(view, N, lam, grid_lam, **kw)
    snap = None
    flag = 2
    # Emulator.run() with the snapshot hook
    while True:
        prev = E.st.copy()
        changed, nfree = E._prop_round()
        greedy = False
        if changed > 0 and E.rounds < E.max_rounds:
            if not (E.stall_den > 0 and E.rounds >= 2 and changed * E.stall_den < nfree):
                continue
            greedy = True
        elif nfree == 0:
            break
        elif E.rounds >= E.max_rounds:
            E.st[E.st == em.FREE] = em.IN
            flag = 0
            break
        else:
            greedy = True
        if greedy:
            if snap is None:
                snap = prev
                flag = 1
            E._greedy_round()
    for _ in range(E.max_drop_rounds):
        if E._drop_round() == 0:
            break
    res = E.result()
    out = dict(flag=flag, objective=res["objective"], result=res)
    if flag == 0:
        out.update(bound=float("nan"))
        return out
    if flag == 2:
        out.update(bound=res["objective"], n_free=0)
        return out
    st = snap
    M, K = E.M, E.K
    sc = 1 << SC_BITS
    glam_fx, lam_fx = int(np.floor(float(np.float32(grid_lam)) * sc)), int(np.floor(float(np.float32(lam)) * sc))
    free = st == em.FREE
    nin_c = np.bincount(E.e_cell, weights=(st[E.e_var] == em.IN), minlength=K * em.N_CELLS).astype(np.int64)
    cov = np.bincount(E.r_row, weights=(st[E.r_var] == em.IN), minlength=E.R).astype(np.int64)
    free_r = np.bincount(E.r_row, weights=free[E.r_var], minlength=E.R).astype(np.int64)
    d = np.maximum(0, E.need - cov)
    s0 = int(np.maximum(0, d - free_r).sum())
    cost_in = int(E.cost[st == em.IN].sum())
    fe = free[E.e_var] & (nin_c[E.e_cell] == 0)              # live entries with an uncovered cell
    cvar, ccell = E.e_var[fe], E.e_cell[fe]
    fr = free[E.r_var] & (d[E.r_row] > 0)                     # live entries of deficient rows
    rvar, rrow = E.r_var[fr], E.r_row[fr]
    nact = np.bincount(cvar, minlength=M).astype(np.int64)
    ndef = np.bincount(rvar, minlength=M).astype(np.int64)
    slack0 = np.minimum(E.cost.astype(np.int64), COST_CAP) << SC_BITS
    share0 = np.where(nact > 0, slack0 // np.maximum(nact, 1), np.int64(0xFFFFFFFF))
    cell_ids, cinv = np.unique(ccell, return_inverse=True)
    mn = np.full(cell_ids.size, np.int64(0xFFFFFFFF))
    np.minimum.at(mn, cinv, share0[cvar])
    delta = np.minimum(mn, glam_fx)
    zsum = int(delta.sum())
    red1 = np.bincount(cvar, weights=delta[cinv].astype(np.float64), minlength=M).astype(np.int64)
    slack1 = np.maximum(slack0 - red1, 0)
    share_r = np.where(ndef > 0, slack1 // np.maximum(ndef, 1), np.int64(0xFFFFFFFF))
    drows = 0
    if rrow.size:
        vals_all = share_r[rvar]
        order = np.lexsort((vals_all, rrow))
        rr, sh = rrow[order], vals_all[order]
        starts = np.flatnonzero(np.concatenate([[True], rr[1:] != rr[:-1]]))
        ends = np.concatenate([starts[1:], [rr.size]])
        for a, b in zip(starts, ends):
            dp = int(min(d[rr[a]], free_r[rr[a]]))
            if dp <= 0:
                continue
            vals = sh[a:b]
            y = min(lam_fx, int(vals[dp - 1]))
            drows += dp * y - int(np.maximum(0, y - vals).sum())
    # residual cells the final selection leaves uncovered
    fin_in = (E.st[cvar] == em.IN)
    cov_c = np.zeros(cell_ids.size, bool)
    np.logical_or.at(cov_c, cinv, fin_in)
    res_unc = int(np.count_nonzero(~cov_c))
    u0 = int(res["uncovered"]) - res_unc
    bound = float(cost_in) + float(np.float32(grid_lam)) * u0 + float(np.float32(lam)) * s0 + (zsum + drows) / sc
    out.update(cost_in=cost_in, s0=s0, zsum=zsum, drows=drows, res_unc=res_unc, u0=u0, bound=bound, n_free=int(free.sum()))
    return out

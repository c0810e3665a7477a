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

"""ORACLE (test infrastructure, not product code) -- numpy emulation of the device algorithm, phase by phase.

PARITY UNPINNED BY UPSTREAM (no tests / fixtures / golden vectors for this path; GUROBI is closed source and the reference
cannot be built here, see oracle/ilp_model.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

This is NOT the reference's algorithm (that is the ILP in oracle/ilp_model.py, solved by GUROBI upstream,
/root/reference/src/MapSparsification.cc:153-157).  It restates, on the CPU and with identical integer arithmetic,
tie-breaks and control flow, what ms_slam_b200/csrc/mss_kernels.cu does on the device, so that the GPU keep-bitmask can be
compared BIT-EXACTLY with a CPU computation, while oracle/ilp_model.py judges the quality of that bitmask
(F(x) within 1 % of the ILP optimum / LP bound, every coverage row satisfied).

Algorithm ("dominance propagation + conflict-free greedy + budgeted reverse delete"), all on F(x) of SURVEY A.3:
  state per variable: FREE / IN / OUT.
  PROP   exact dominance, iterated to a fixed point.  For a FREE point p with cost c_p
           ub_p = GridLambda * #{cells of p with no IN point} + Lambda * #{rows of p still deficient} - c_p
           lb_p = GridLambda * #{uncovered cells where p is the only FREE point}
                  + Lambda * #{deficient rows that need every FREE point they have} - c_p
         ub_p <= 0  ->  OUT (no completion can make p pay);  lb_p >= 0  ->  IN (p pays in every completion).
  GREEDY when PROP stalls: every FREE point has gain ub_p > 0; a point is taken iff it is the best
         (gain, then lower index) FREE candidate of every uncovered cell it lies in and within the top-deficit
         candidates of every deficient row it lies in (conflict-free = what sequential greedy would also take).
  DROP   once no FREE point is left: remove IN points whose removal lowers F; per cell at most nin-1 removals, per row
         at most (cov - need) removals per round, so the summed deltas are exact and F decreases monotonically.
"""
from __future__ import annotations

import numpy as np

N_CELLS = 64 * 48
CELL_NONE = 0xFFFF
FREE, IN, OUT, NOTVAR, CAND = 0, 1, 2, 3, 4
FLAG_BLOCKED, FLAG_NOMINATED = 1, 2


def f32_key(g):
    """order-preserving uint32 image of float32(g)"""
    b = np.asarray(g, np.float64).astype(np.float32).view(np.uint32).astype(np.uint64)
    mask = np.where((b >> np.uint64(31)) != 0, np.uint64(0xFFFFFFFF), np.uint64(0x80000000))
    return b ^ mask


def make_key(g, p):
    """larger key = better: higher gain first, then lower variable index"""
    return (f32_key(g) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(p, np.uint64))


def outside_need(cnt, total, N):
    cnt32 = np.asarray(cnt).astype(np.float32)
    tot32 = np.asarray(total).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = ((cnt32 / tot32).astype(np.float32) * np.float32(N)).astype(np.float32)
    need = np.ceil(r.astype(np.float64) - 1e-5).astype(np.int64)
    return np.where((np.asarray(total) > 0) & (np.asarray(cnt) > 0), need, 0)


def _admit_topk(rows, keys, k_row):
    """admitted[e] iff key[e] > (k+1)-th largest key of its row (all admitted when the row has <= k entries).
    rows: row id per entry, keys: uint64, k_row: dict-like array of budget per row id."""
    if rows.size == 0:
        return np.zeros(0, bool)
    order = np.lexsort((keys, rows))          # ascending key within row
    r_s, k_s = rows[order], keys[order]
    last = np.searchsorted(r_s, r_s, side="right")          # one past the last entry of the row
    first = np.searchsorted(r_s, r_s, side="left")
    n_row = last - first
    budget = k_row[r_s]
    # (k+1)-th largest key position = last - 1 - k
    thr_pos = last - 1 - budget
    has_thr = (n_row > budget)
    thr = k_s[np.clip(thr_pos, 0, r_s.size - 1)]
    adm_s = np.where(has_thr, k_s > thr, True) & (budget > 0)
    adm = np.zeros(rows.size, bool)
    adm[order] = adm_s
    return adm


class Emulator:
    def __init__(self, view, N, lam, grid_lam, max_rounds=256, all_rule_steps=0, max_drop_rounds=16, stall_den=8):
        self.N, self.lam, self.glam = int(N), float(lam), float(grid_lam)
        self.max_rounds, self.all_rule_steps, self.max_drop_rounds = max_rounds, all_rule_steps, max_drop_rounds
        self.stall_den = stall_den
        K, H, M = view.K, view.H, view.M
        self.K, self.H, self.M = K, H, M
        feat_kf = np.repeat(np.arange(K, dtype=np.int64), np.diff(view.feat_ptr))
        valid = view.feat_mp >= 0
        self.n_max = max(int(view.mp_nobs[view.feat_mp[valid]].max()) if valid.any() else 0, int(getattr(view, "n_max_floor", 0)))
        grid = valid & (view.feat_cell != CELL_NONE)
        self.e_row = feat_kf[grid]
        self.e_var = view.feat_mp[grid].astype(np.int64)
        self.e_cell = self.e_row * N_CELLS + view.feat_cell[grid].astype(np.int64)
        self.st = np.full(M, NOTVAR, np.int8)
        self.st[self.e_var] = FREE
        self.cost = (self.n_max - view.mp_nobs).astype(np.int64)
        obs_mp = np.repeat(np.arange(M, dtype=np.int64), np.diff(view.mp_obs_ptr))
        om = (view.mp_obs_kf >= K) & (self.st[obs_mp] == FREE)
        self.o_var = obs_mp[om]
        self.o_row = view.mp_obs_kf[om].astype(np.int64)          # K + j
        cnt = np.bincount(self.o_row - K, minlength=H)
        self.out_cnt = cnt
        self.need = np.concatenate([np.full(K, self.N, np.int64), outside_need(cnt, view.okf_total, self.N)])
        self.avail = np.concatenate([np.bincount(self.e_row, minlength=K), cnt]).astype(np.int64)
        # unified row entries
        self.r_row = np.concatenate([self.e_row, self.o_row])
        self.r_var = np.concatenate([self.e_var, self.o_var])
        self.R = K + H
        self.rounds = 0
        self.greedy_steps = 0
        self.log = []
        # tie-break key of every variable: the table index, or -- mss_window_view::mp_tie -- the rank of its gid (the selection
        # is then the same for every numbering of the window)
        tie = getattr(view, "mp_tie", None)
        if tie is None and getattr(view, "meta", {}).get("tie_by_gid"):
            tie = np.empty(M, np.int64)
            tie[np.argsort(view.mp_gid, kind="stable")] = np.arange(M)
        self.tie = np.arange(M) if tie is None else np.asarray(tie, np.int64)

    # ---- shared per-round statistics -------------------------------------------------------------------
    def _stats(self):
        st = self.st
        nin_c = np.bincount(self.e_cell, weights=(st[self.e_var] == IN), minlength=self.K * N_CELLS).astype(np.int64)
        nfree_c = np.bincount(self.e_cell, weights=(st[self.e_var] == FREE), minlength=self.K * N_CELLS).astype(np.int64)
        cov = np.bincount(self.r_row, weights=(st[self.r_var] == IN), minlength=self.R).astype(np.int64)
        free_r = np.bincount(self.r_row, weights=(st[self.r_var] == FREE), minlength=self.R).astype(np.int64)
        return nin_c, nfree_c, cov, free_r

    def _prop_round(self):
        st, M = self.st, self.M
        nin_c, nfree_c, cov, free_r = self._stats()
        self.row_view = (nin_c, cov)          # what the row phase of this round saw (the device's cell flags / row_cov)
        d = np.maximum(0, self.need - cov)
        fe = st[self.e_var] == FREE
        unc = fe & (nin_c[self.e_cell] == 0)
        crit_c = unc & (nfree_c[self.e_cell] == 1)
        fr = st[self.r_var] == FREE
        defi = fr & (d[self.r_row] > 0)
        crit_r = defi & (d[self.r_row] >= free_r[self.r_row])
        ubc = np.bincount(self.e_var, weights=unc, minlength=M)
        lbc = np.bincount(self.e_var, weights=crit_c, minlength=M)
        ubr = np.bincount(self.r_var, weights=defi, minlength=M)
        lbr = np.bincount(self.r_var, weights=crit_r, minlength=M)
        ub = self.glam * ubc + self.lam * ubr - self.cost
        lb = self.glam * lbc + self.lam * lbr - self.cost
        isfree = st == FREE
        to_out = isfree & (ub <= 0)
        to_in = isfree & ~to_out & (lb >= 0)
        st[to_out] = OUT
        st[to_in] = IN
        self.gain = ub
        self.rounds += 1
        changed = int(to_out.sum() + to_in.sum())
        nfree = int((st == FREE).sum())
        self.log.append(("prop", changed, nfree))
        return changed, nfree

    def _greedy_round(self):
        st, M = self.st, self.M
        any_rule = self.greedy_steps >= self.all_rule_steps
        # GREEDY reads the live lists of the last PROP row phase: cell-covered flags and row coverage are the ones that phase
        # saw (exact when that round changed nothing; one variable phase behind when GREEDY was triggered by a stall),
        # state and gain are current
        nin_c, cov = self.row_view
        d = np.maximum(0, self.need - cov)
        key = make_key(self.gain, self.tie)
        flags = np.zeros(M, np.int8)
        fe = (st[self.e_var] == FREE) & (nin_c[self.e_cell] == 0)
        best = np.zeros(self.K * N_CELLS, np.uint64)
        np.maximum.at(best, self.e_cell[fe], key[self.e_var[fe]])
        lose = fe & (key[self.e_var] != best[self.e_cell])
        flags[self.e_var[lose]] |= FLAG_BLOCKED
        fr = (st[self.r_var] == FREE) & (d[self.r_row] > 0)
        idx = np.nonzero(fr)[0]
        adm = _admit_topk(self.r_row[idx], key[self.r_var[idx]], d)
        flags[self.r_var[idx[~adm]]] |= FLAG_BLOCKED
        nom = np.zeros(M, bool)
        nom[self.r_var[idx[adm]]] = True
        isfree = st == FREE
        sel = isfree & (self.gain > 0) & ((flags & FLAG_BLOCKED) == 0)
        if any_rule:
            sel |= isfree & nom
        st[sel] = IN
        self.greedy_steps += 1
        self.rounds += 1
        self.log.append(("greedy", int(sel.sum()), int((st == FREE).sum())))

    def _drop_round(self):
        st, M = self.st, self.M
        nin_c, _, cov, _ = self._stats()
        ie = st[self.e_var] == IN
        ir = st[self.r_var] == IN
        crit_c = np.bincount(self.e_var, weights=ie & (nin_c[self.e_cell] == 1), minlength=M)
        crit_r = np.bincount(self.r_var, weights=ir & (cov[self.r_row] <= self.need[self.r_row]), minlength=M)
        dF = -self.cost + self.glam * crit_c + self.lam * crit_r
        cand = (st == IN) & (dF < 0)
        self.rounds += 1
        ncand = int(cand.sum())
        if ncand == 0:
            self.log.append(("drop", 0, 0))
            return 0
        key = make_key(-dF, self.tie)
        blocked = np.zeros(M, bool)
        ce = cand[self.e_var] & (nin_c[self.e_cell] >= 2)
        best = np.zeros(self.K * N_CELLS, np.uint64)
        np.maximum.at(best, self.e_cell[ce], key[self.e_var[ce]])
        lose = ce & (key[self.e_var] != best[self.e_cell])
        blocked[self.e_var[lose]] = True
        re_ = cand[self.r_var] & (cov[self.r_row] > self.need[self.r_row])
        idx = np.nonzero(re_)[0]
        adm = _admit_topk(self.r_row[idx], key[self.r_var[idx]], np.maximum(cov - self.need, 0))
        blocked[self.r_var[idx[~adm]]] = True
        drop = cand & ~blocked
        st[drop] = OUT
        self.log.append(("drop", ncand, int(drop.sum())))
        return ncand

    # ---- control flow (mirrors the persistent kernel's phase machine) -----------------------------------------
    def run(self):
        while True:
            changed, nfree = self._prop_round()
            if changed > 0 and self.rounds < self.max_rounds:
                # stall: propagation still moves, but slowly (deficient rows with many candidates creep through the window
                # keyframe by keyframe) -> take a greedy step now instead of waiting for the fixed point
                if not (self.stall_den > 0 and self.rounds >= 2 and changed * self.stall_den < nfree):
                    continue
                self._greedy_round()
                continue
            if nfree == 0:
                break
            if self.rounds >= self.max_rounds:
                self.st[self.st == FREE] = IN          # safe: only adds coverage
                break
            self._greedy_round()
        for _ in range(self.max_drop_rounds):
            if self._drop_round() == 0:
                break
        return self.result()

    def result(self):
        st = self.st
        keep = st != OUT
        nin_c, _, cov, _ = self._stats()
        occupied = np.bincount(self.e_cell, minlength=self.K * N_CELLS) > 0
        unc = int(np.count_nonzero(occupied & (nin_c == 0)))
        slack = np.maximum(0, self.need - cov)
        sum_cost = int(self.cost[st == IN].sum())
        F = float(sum_cost) + self.glam * unc + self.lam * int(slack.sum())
        return dict(keep=keep, cov=cov, slack=slack, objective=F, sum_cost=sum_cost, uncovered=unc,
                    n_vars=int((st != NOTVAR).sum()), n_cells=int(occupied.sum()), nnz=int(self.e_var.size),
                    rounds=self.rounds, n_max=self.n_max, need=self.need, n_kept=int((st == IN).sum()))


def solve(view, N, lam, grid_lam, **kw):
    return Emulator(view, N, lam, grid_lam, **kw).run()


def pack_bits(keep):
    """keep mask -> uint32 words, bit i of word w = MP 32*w+i (the C-ABI's keep_bits layout)"""
    keep = np.asarray(keep, bool)
    pad = (-keep.size) % 32
    b = np.concatenate([keep, np.zeros(pad, bool)]).reshape(-1, 32)
    return (b.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)

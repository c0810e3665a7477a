// mss_bound.cuh -- device-side lower bound of the reference ILP (mss_result.dual_bound).
// Included by mss_kernels.cuh (inside namespace mss, after the row-phase helpers).
//
// The reference gets its optimality certificate from GUROBI's branch and bound (MIPGap 0.002,
// /root/reference/src/MapSparsification.cc:153-157); the device algorithm is a heuristic on the same model, so a window
// carries a certificate only if the device produces a bound itself.  It does so from the SNAPSHOT S = the state described
// by the live lists of the last PROP row phase before the first GREEDY step: every decision in S was taken by an exact
// dominance rule (valid for the LP relaxation too), so
//
//     LP* >= cost(IN in S) + GridLambda * u0 + Lambda * s0 + D_cells + D_rows
//
//   u0       cells whose points are all rejected in S
//   s0       sum over rows of the part of the deficit that exceeds the row's undecided points
//   D_cells  one round of dual ascent on the uncovered cells of the residual: cell c gets z_c = min(GridLambda,
//            min over its FREE points of cost_p / #uncovered cells of p) -- every point pays at most its cost
//   D_rows   per still deficient row the best single multiplier y against the slack the cells left, a point's slack split
//            evenly over its deficient rows: d' * y - sum_p max(0, y - share_p), maximal at the d'-th smallest share
//
// All of it in integer fixed point (2^-10), so the sums are order independent and oracle/dual_bound.py device_twin()
// reproduces every counter bit for bit.  u0 is not counted here: rejected points never come back and points taken by
// dominance are never dropped, so u0 = (uncovered cells of the final selection) - (residual cells of S the final selection
// leaves uncovered); the latter is counted at the end from the saved copy of the live lists (bound_final).
// Cost: four short phases over the residual (c2: ~1 200 points, ~5 000 entries) + one at the end.
#pragma once

constexpr int kBndScBits = 10;
constexpr unsigned kBndInf = 0xFFFFFFFFu;
constexpr unsigned kBndCostCap = (1u << 20) - 1u;     // costs enter the duals capped (a smaller cost only weakens the bound): shares < 2^30

struct BoundBufs {
    uint32_t* snap;          // [Ftot + Otot] copy of the live lists at the snapshot (same segments as ent / live)
    int* snap_n;             // [Rtot] entries of the copy
    int* snap_d;             // [Rtot] deficit of the row in S
    unsigned* share;         // [Mpad] share of the point's left-over slack per deficient row
    unsigned* red;           // [Mpad] what the cell duals take from the point
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// B0: copy the lists, count per point the uncovered cells (low word of acc) and deficient rows (high word) it lies in;
//     s0; cost of the points that are IN now (corrected to S by bound_b1); clear the reduction accumulators
__device__ void bound_b0(const Params& P, const BoundBufs& B, const WinDesc& D, WinState& ws, int cta, int ncta, const uint32_t* lists,
                         const int* listn, const int* vprev, int nv) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows = D.K + D.H;
    unsigned long long* acc_w = P.acc + D.var_base;
    unsigned s0 = 0;
    for (int r = cta + wid * ncta; r < rows; r += ncta * kWarps) {
        const int R = D.row_base + r;
        const int n = listn[R], off = P.row_off[R];
        const int d = max(0, P.row_need[R] - P.row_cov[R]);
        if (lane == 0) { B.snap_n[R] = n; B.snap_d[R] = d; s0 += (unsigned)max(0, d - n); }
        for (int j = lane; j < n; j += 32) {
            const uint32_t e = lists[off + j];
            B.snap[off + j] = e;
            const unsigned long long add = ((e & kCellCov) != kCellCov ? 1ull : 0ull) | (d > 0 ? 1ull << 32 : 0ull);
            if (add) atomicAdd(&acc_w[e >> kCellBits], add);
        }
    }
    const int gt = cta * kThreads + (int)threadIdx.x, gsz = ncta * kThreads;
    unsigned long long cin = 0;
    for (int i = gt; i < D.M; i += gsz)
        if (P.st[D.var_base + i] == ST_IN) cin += (unsigned long long)(ws.n_max - ld_nobs(D, i));
    for (int i = gt; i < nv; i += gsz) B.red[D.var_base + vprev[i]] = 0u;
    cin = warp_sum_u64(cin);
    s0 = __reduce_add_sync(0xFFFFFFFFu, s0);
    if (lane == 0) {
        if (cin) atomicAdd(&ws.b_cost_in, cin);
        if (s0) atomicAdd(&ws.b_s0, s0);
    }
}

// share of a point's cost per uncovered cell it lies in (low word of acc = that count, written by bound_b0)
__device__ __forceinline__ unsigned bound_share0(const Params& P, const WinDesc& D, const WinState& ws, unsigned mp) {
    const unsigned nact = (unsigned)P.acc[D.var_base + mp];
    const unsigned cost = (unsigned)(ws.n_max - ld_nobs(D, (int)mp));
    return nact ? (min(cost, kBndCostCap) << kBndScBits) / nact : kBndInf;
}

// B2: per uncovered cell of the residual z = min(GridLambda, min share); every point of the cell pays z
__device__ void bound_b2(const Params& P, const BoundBufs& B, const WinDesc& D, WinState& ws, int cta, int ncta, unsigned* tab) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows = D.K + D.H;
    const unsigned glam_fx = (unsigned)floor(P.glam * (double)(1 << kBndScBits));
    unsigned long long z = 0;
    // long lists: the CTA, cell table in shared memory
    for (int r = cta; r < rows; r += ncta) {
        const int R = D.row_base + r;
        const int n = B.snap_n[R];
        if (n <= 32) continue;
        const uint32_t* src = B.snap + P.row_off[R];
        for (int j = threadIdx.x; j < n; j += kThreads) { const unsigned c = src[j] & kCellCov; if (c != kCellCov) tab[c] = kBndInf; }
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kThreads) {
            const uint32_t e = src[j];
            const unsigned c = e & kCellCov;
            if (c != kCellCov) atomicMin(&tab[c], bound_share0(P, D, ws, e >> kCellBits));
        }
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kThreads) {
            const uint32_t e = src[j];
            const unsigned c = e & kCellCov;
            if (c == kCellCov) continue;
            const unsigned old = atomicOr(&tab[c], 0x80000000u);             // shares are < 2^27: bit 31 = "cell counted"
            const unsigned delta = min(glam_fx, old & 0x7FFFFFFFu);
            atomicAdd(&B.red[D.var_base + (e >> kCellBits)], delta);
            if (!(old & 0x80000000u)) z += delta;
        }
        __syncthreads();
    }
    // short lists: one warp, cells matched by comparison
    for (int r = cta + wid * ncta; r < rows; r += ncta * kWarps) {
        const int R = D.row_base + r;
        const int n = B.snap_n[R];
        if (n <= 0 || n > 32) continue;
        const uint32_t e = lane < n ? B.snap[P.row_off[R] + lane] : kEntInvalid;
        const unsigned c = (e == kEntInvalid || (e & kCellCov) == kCellCov) ? 0x10000u + (unsigned)lane : (e & kCellCov);
        const unsigned sh = c < 0x10000u ? bound_share0(P, D, ws, e >> kCellBits) : kBndInf;
        unsigned m = sh;
        int leader = lane;
        for (int j = 0; j < n; ++j) {
            const unsigned cj = __shfl_sync(0xFFFFFFFFu, c, j), sj = __shfl_sync(0xFFFFFFFFu, sh, j);
            if (cj == c) { m = min(m, sj); leader = min(leader, j); }
        }
        if (c < 0x10000u) {
            const unsigned delta = min(glam_fx, m);
            atomicAdd(&B.red[D.var_base + (e >> kCellBits)], delta);
            if (leader == lane) z += delta;
        }
    }
    z = warp_sum_u64(z);
    if (lane == 0 && z) atomicAdd(&ws.b_zsum, z);
}

// B3: per FREE point of S: slack the cells left, split over the point's deficient rows (high word of acc); the counters go
//     back to zero; the points the last variable phase took are not IN in S
__device__ void bound_b3(const Params& P, const BoundBufs& B, const WinDesc& D, WinState& ws, int cta, int ncta, const int* vprev, int nv) {
    const int gt = cta * kThreads + (int)threadIdx.x, gsz = ncta * kThreads;
    unsigned long long sub = 0;
    for (int i = gt; i < nv; i += gsz) {
        const int mp = vprev[i], g = D.var_base + mp;
        const unsigned long long a = P.acc[g];
        if (a) P.acc[g] = 0ull;
        const unsigned nd = (unsigned)(a >> 32);
        const unsigned cost = (unsigned)(ws.n_max - ld_nobs(D, mp));
        const unsigned slack0 = min(cost, kBndCostCap) << kBndScBits;
        const unsigned red = B.red[g];
        const unsigned slack1 = slack0 > red ? slack0 - red : 0u;
        B.share[g] = nd ? slack1 / nd : kBndInf;
        if (P.st[g] == ST_IN) sub += cost;
    }
    sub = warp_sum_u64(sub);
    if ((threadIdx.x & 31) == 0 && sub) atomicAdd(&ws.b_cost_in, 0ull - sub);
}

// B4: deficient rows: y = min(Lambda, d'-th smallest share), D = d' * y - sum max(0, y - share)
__device__ void bound_b4(const Params& P, const BoundBufs& B, const WinDesc& D, WinState& ws, int cta, int ncta, BlockScratch& S) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows = D.K + D.H;
    const unsigned lam_fx = (unsigned)floor(P.lam * (double)(1 << kBndScBits));
    unsigned long long tot = 0;
    for (int r = cta; r < rows; r += ncta) {
        const int R = D.row_base + r;
        const int n = B.snap_n[R], d = B.snap_d[R];
        if (n <= 32 || d <= 0) continue;
        const uint32_t* src = B.snap + P.row_off[R];
        const int dp = min(d, n);
        const unsigned long long kth = block_kth_largest(S, n - dp, [&](auto f) {
            for (int j = threadIdx.x; j < n; j += kThreads) f((unsigned long long)B.share[D.var_base + (src[j] >> kCellBits)] << 32);
        }, 32);
        const unsigned y = min(lam_fx, (unsigned)(kth >> 32));
        unsigned long long under = 0;
        for (int j = threadIdx.x; j < n; j += kThreads) {
            const unsigned v = B.share[D.var_base + (src[j] >> kCellBits)];
            if (v < y) under += (unsigned long long)(y - v);
        }
        under = warp_sum_u64(under);
        __syncthreads();
        if (threadIdx.x == 0) S.work = 0ull;
        __syncthreads();
        if (lane == 0 && under) atomicAdd(&S.work, under);
        __syncthreads();
        if (threadIdx.x == 0) tot += (unsigned long long)dp * y - S.work;
    }
    for (int r = cta + wid * ncta; r < rows; r += ncta * kWarps) {
        const int R = D.row_base + r;
        const int n = B.snap_n[R], d = B.snap_d[R];
        if (n <= 0 || n > 32 || d <= 0) continue;
        const int dp = min(d, n);
        const unsigned v = lane < n ? B.share[D.var_base + (B.snap[P.row_off[R] + lane] >> kCellBits)] : kBndInf;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const unsigned vj = __shfl_sync(0xFFFFFFFFu, v, j);
            rank += (vj < v || (vj == v && j < lane)) ? 1 : 0;
        }
        const unsigned pick = __ballot_sync(0xFFFFFFFFu, lane < n && rank == dp - 1);
        const unsigned y = min(lam_fx, __shfl_sync(0xFFFFFFFFu, v, __ffs(pick) - 1));
        unsigned long long under = (lane < n && v < y) ? (unsigned long long)(y - v) : 0ull;
        under = warp_sum_u64(under);
        if (lane == 0) tot += (unsigned long long)dp * y - under;
    }
    tot = warp_sum_u64(tot);
    if (lane == 0 && tot) atomicAdd(&ws.b_drows, tot);
}

// At the end: residual cells of S the final selection leaves uncovered
__device__ void bound_final(const Params& P, const BoundBufs& B, const WinDesc& D, WinState& ws, int cta, int ncta, unsigned* tab) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows = D.K + D.H;
    const uint8_t* st_w = P.st + D.var_base;
    unsigned cnt = 0;
    for (int r = cta; r < rows; r += ncta) {
        const int R = D.row_base + r;
        const int n = B.snap_n[R];
        if (n <= 32) continue;
        const uint32_t* src = B.snap + P.row_off[R];
        for (int j = threadIdx.x; j < n; j += kThreads) { const unsigned c = src[j] & kCellCov; if (c != kCellCov) tab[c] = 0u; }
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kThreads) {
            const uint32_t e = src[j];
            const unsigned c = e & kCellCov;
            if (c != kCellCov && st_w[e >> kCellBits] == ST_IN) tab[c] = 1u;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kThreads) {
            const unsigned c = src[j] & kCellCov;
            if (c != kCellCov && atomicOr(&tab[c], 2u) == 0u) ++cnt;
        }
        __syncthreads();
    }
    for (int r = cta + wid * ncta; r < rows; r += ncta * kWarps) {
        const int R = D.row_base + r;
        const int n = B.snap_n[R];
        if (n <= 0 || n > 32) continue;
        const uint32_t e = lane < n ? B.snap[P.row_off[R] + lane] : kEntInvalid;
        const unsigned c = (e == kEntInvalid || (e & kCellCov) == kCellCov) ? 0x10000u + (unsigned)lane : (e & kCellCov);
        const bool in = c < 0x10000u && st_w[e >> kCellBits] == ST_IN;
        bool any = in;
        int leader = lane;
        for (int j = 0; j < n; ++j) {
            const unsigned cj = __shfl_sync(0xFFFFFFFFu, c, j);
            const bool ij = __shfl_sync(0xFFFFFFFFu, in ? 1 : 0, j) != 0;
            if (cj == c) { any = any || ij; leader = min(leader, j); }
        }
        if (c < 0x10000u && leader == lane && !any) ++cnt;
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0 && cnt) atomicAdd(&ws.b_res_unc, cnt);
}

// mss_kernels.cuh -- device side of the sparsification engine (sm_100a).
//
// One persistent cooperative kernel solves a whole batch of independent windows.  What it replaces in the reference
// (/root/reference/src/MapSparsification.cc):
//   :66-76   nMaxObservation scan                         -> phase P1 (per keyframe row, block max -> atomicMax)
//   :78-123  variable discovery + cell rows + KF rows     -> P1 marks variables; cell rows are never materialised:
//                                                            a CTA owns one keyframe row and keeps its 64x48 cell table
//                                                            in shared memory (cell rows are private to a keyframe)
//   :125-151 outside-keyframe rows                        -> P2 count / P3 scan+rhs / P4 fill (CSR of outside rows)
//   :153-157 GUROBI optimize()                            -> per-window phase machine PROP / GREEDY / DROP (below)
//   :159-166 read-out of GRB_DoubleAttr_X                 -> EVAL: ballot-packed keep bits + row coverage + F(x)
//
// Selection algorithm on the penalty form F(x), per variable state FREE / IN / OUT:
//   PROP   exact dominance to a fixed point: ub_p <= 0 -> OUT, lb_p >= 0 -> IN  (bounds on p's marginal gain over
//          every completion of the FREE points; see oracle/emulate.py for the formulas)
//   GREEDY conflict-free step: a FREE point is taken iff it is the best candidate of every uncovered cell it lies in
//          and within the top-deficit candidates of every deficient row it lies in
//   DROP   budgeted reverse delete; per-cell / per-row budgets make the summed deltas exact
// Every decision is made from integer counters accumulated with integer atomics and from keys with a unique
// tie-break (gain, then lower map-point index), so the result is deterministic and independent of scheduling;
// oracle/emulate.py reproduces it bit for bit on the CPU.
#pragma once

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace mss {

namespace cg = cooperative_groups;

constexpr int kThreads = 256;
constexpr int kCells = 64 * 48;
constexpr int kVarTile = 256;          // var_base alignment; one var-pass tile belongs to exactly one window
constexpr int kHdrWords = 16;          // result-slot header
constexpr unsigned kCellNone = 0xFFFFu;
constexpr int kCellFieldMax = 1023;    // 10-bit per-cell counters

enum : uint8_t { ST_FREE = 0, ST_IN = 1, ST_OUT = 2, ST_NOTVAR = 3, ST_CAND = 4 };
enum : int { MODE_DONE = 0, MODE_PROP = 1, MODE_GREEDY = 2, MODE_FORCE = 3, MODE_D1 = 4, MODE_D2 = 5, MODE_EVAL = 6 };
enum : unsigned long long { FLAG_BLOCKED = 1ull, FLAG_NOMINATED = 2ull };
enum : unsigned { ERR_INDEX = 1u, ERR_CELL_OVERFLOW = 2u, ERR_PTR = 4u };

struct WinDesc {
    const int* feat_ptr;
    const int* feat_mp;
    const uint16_t* feat_cell;
    const int* mp_nobs;
    const int* mp_obs_ptr;
    const int* mp_obs_kf;
    const int* okf_total;
    int K, H, M, F, O;
    int row_base;    // global id of the first keyframe row
    int orow_base;   // index of the first outside row among all outside rows (global row id = Ktot + orow_base + j)
    int var_base;    // global variable index of map point 0 (multiple of kVarTile)
    int out_off;     // u32 word offset of this window's result slot
    int owned;       // solved on this rank
    int pad_;
};

struct WinState {
    int mode, rounds, greedy_steps, drop_rounds;
    unsigned changed, nfree, ncand, error;
    int n_max, n_vars, n_cells, nnz;
    int n_kept, uncovered, total_slack, status;
    unsigned long long sum_cost;
    unsigned long long pad_;
};

struct Ctrl {
    unsigned ticket;
    int n_active;
    unsigned long long t_start, t_build, t_end;
    int iters;
    int pad_;
};

struct Params {
    const WinDesc* win;
    WinState* ws;
    const int* row_win;     // [Rtot] window of every row (keyframe rows first, then outside rows)
    const int* tile_win;    // [ntiles]
    uint8_t* st;            // [Mpad]
    unsigned long long* acc;// [Mpad] packed counters / flags
    float* gain;            // [Mpad]
    int* row_need;          // [Rtot]
    int* ocnt;              // [Htot]
    int* orow_ptr;          // [Htot+1]
    int* ocursor;           // [Htot]
    int* orow_var;          // [Ocap]
    uint32_t* out;          // result slots
    Ctrl* ctrl;
    int nwin, Ktot, Htot, Rtot, Mpad, ntiles;
    int N;
    int max_rounds, all_rule_steps, max_drop_rounds;
    double lam, glam;
};

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned f32_orderable(float g) {
    unsigned b = __float_as_uint(g);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
// larger key = better: higher gain first, then lower (window-local) map-point index
__device__ __forceinline__ unsigned long long make_key(float g, unsigned local_idx) {
    return ((unsigned long long)f32_orderable(g) << 32) | (unsigned long long)(0xFFFFFFFFu - local_idx);
}

struct BlockScratch {
    int red[3][kThreads / 32];
    int bcast[4];
    unsigned hist[256];
    unsigned long long sel_prefix;
};

// block-wide sum of up to three ints, result broadcast to every thread
__device__ __forceinline__ void block_sum3(BlockScratch& S, int& a, int& b, int& c) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    }
    __syncthreads();                       // protect S.red from the previous use
    if (lane == 0) { S.red[0][wid] = a; S.red[1][wid] = b; S.red[2][wid] = c; }
    __syncthreads();
    a = b = c = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) { a += S.red[0][w]; b += S.red[1][w]; c += S.red[2][w]; }
}

__device__ __forceinline__ int block_max(BlockScratch& S, int a) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = max(a, __shfl_xor_sync(0xFFFFFFFFu, a, o));
    __syncthreads();
    if (lane == 0) S.red[0][wid] = a;
    __syncthreads();
    int m = S.red[0][0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) m = max(m, S.red[0][w]);
    return m;
}

// A row of the coverage problem: a window keyframe (entries = valid grid-listed slots, each in a cell) or an
// outside keyframe (entries = variables it observes, no cells).
struct Row {
    const int* mp;            // keyframe row: feat_mp ; outside row: orow_var (global variable indices)
    const uint16_t* cell;     // keyframe row only
    int beg, end;
    int var_base;
    int need;
    bool is_kf;
};

template <class Fn>
__device__ __forceinline__ void for_each_entry(const Row& R, Fn fn) {
    if (R.is_kf) {
        for (int i = R.beg + (int)threadIdx.x; i < R.end; i += kThreads) {
            const int mp = __ldg(R.mp + i);
            if (mp < 0) continue;
            const unsigned c = __ldg(R.cell + i);
            if (c == kCellNone) continue;
            fn(R.var_base + mp, (int)c);
        }
    } else {
        for (int i = R.beg + (int)threadIdx.x; i < R.end; i += kThreads) fn(R.mp[i], -1);
    }
}

__device__ __forceinline__ Row make_row(const Params& P, int r, const WinDesc& D) {
    Row R;
    R.need = P.row_need[r];
    R.var_base = D.var_base;
    if (r < P.Ktot) {
        const int k = r - D.row_base;
        R.is_kf = true;
        R.mp = D.feat_mp;
        R.cell = D.feat_cell;
        R.beg = __ldg(D.feat_ptr + k);
        R.end = __ldg(D.feat_ptr + k + 1);
    } else {
        const int jj = r - P.Ktot;
        R.is_kf = false;
        R.mp = P.orow_var;
        R.cell = nullptr;
        R.beg = P.orow_ptr[jj];
        R.end = P.orow_ptr[jj + 1];
    }
    return R;
}

// (rank+1)-th largest key (rank elements are larger, counting multiplicity) among the keys produced by `emit`.
// The caller guarantees that more than `rank` keys exist.  8 radix passes of 8 bits, most significant first.
template <class Emit>
__device__ unsigned long long block_kth_largest(BlockScratch& S, int rank, Emit emit) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        __syncthreads();
        if (threadIdx.x < 256) S.hist[threadIdx.x] = 0;
        __syncthreads();
        emit([&](unsigned long long key) {
            if ((key & mask) == prefix) atomicAdd(&S.hist[(unsigned)(key >> shift) & 255u], 1u);
        });
        __syncthreads();
        if (threadIdx.x == 0) {
            int cum = 0, b = 255;
            for (; b > 0; --b) {
                const int h = (int)S.hist[b];
                if (cum + h > rank) break;
                cum += h;
            }
            S.bcast[0] = b;
            S.bcast[1] = rank - cum;
        }
        __syncthreads();
        prefix |= (unsigned long long)S.bcast[0] << shift;
        mask |= 255ull << shift;
        rank = S.bcast[1];
    }
    return prefix;
}

// cell table word: [total:12 | nin:10 | nlow:10]   (nlow = FREE count in PROP/GREEDY, CAND count in D2)
__device__ __forceinline__ int tab_low(unsigned t) { return (int)(t & 0x3FFu); }
__device__ __forceinline__ int tab_in(unsigned t) { return (int)((t >> 10) & 0x3FFu); }
__device__ __forceinline__ int tab_total(unsigned t) { return (int)(t >> 20); }

__device__ __forceinline__ void zero_tab(unsigned* tab) {
    for (int c = threadIdx.x; c < kCells; c += kThreads) tab[c] = 0u;
}
__device__ __forceinline__ void zero_keytab(unsigned long long* kt) {
    for (int c = threadIdx.x; c < kCells; c += kThreads) kt[c] = 0ull;
}

// ---------------------------------------------------------------------------------------------------------------
// build phases
// ---------------------------------------------------------------------------------------------------------------
__device__ void p1_scan_row(const Params& P, int r, unsigned* tab, BlockScratch& S) {
    const int w = P.row_win[r];
    const WinDesc D = P.win[w];
    const int k = r - D.row_base;
    const int beg = __ldg(D.feat_ptr + k), end = __ldg(D.feat_ptr + k + 1);
    unsigned err = 0;
    if (beg < 0 || end < beg || end > D.F) {
        if (threadIdx.x == 0) atomicOr(&P.ws[w].error, ERR_PTR);
        if (threadIdx.x == 0) P.row_need[r] = P.N;
        return;
    }
    zero_tab(tab);
    __syncthreads();
    int nmax = 0, nz = 0;
    for (int i = beg + (int)threadIdx.x; i < end; i += kThreads) {
        const int mp = __ldg(D.feat_mp + i);
        if (mp < 0) { if (mp < -1) err |= ERR_INDEX; continue; }
        if (mp >= D.M) { err |= ERR_INDEX; continue; }
        nmax = max(nmax, __ldg(D.mp_nobs + mp));                 // MapSparsification.cc:69-75: every valid slot
        const unsigned c = __ldg(D.feat_cell + i);
        if (c == kCellNone) continue;                           // not in mGrid: not a variable through this slot
        if (c >= (unsigned)kCells) { err |= ERR_INDEX; continue; }
        P.st[D.var_base + mp] = ST_FREE;                        // MapSparsification.cc:91-99 (same value from every writer)
        atomicAdd(&tab[c], 1u);
        ++nz;
    }
    __syncthreads();
    int ncell = 0, z0 = 0;
    for (int c = threadIdx.x; c < kCells; c += kThreads) {
        const unsigned t = tab[c];
        if (t) ++ncell;
        if (t > (unsigned)kCellFieldMax) err |= ERR_CELL_OVERFLOW;
    }
    nmax = block_max(S, nmax);
    block_sum3(S, nz, ncell, z0);
    if (threadIdx.x == 0) {
        atomicMax(&P.ws[w].n_max, nmax);
        if (nz) atomicAdd(&P.ws[w].nnz, nz);
        if (ncell) atomicAdd(&P.ws[w].n_cells, ncell);
        P.row_need[r] = P.N;
    }
    if (err) atomicOr(&P.ws[w].error, err);
}

// per variable: count it, and count/fill its observations by outside keyframes (MapSparsification.cc:127-142)
template <bool FILL>
__device__ void p24_outside(const Params& P, int tile, BlockScratch& S) {
    const int w = P.tile_win[tile];
    const WinDesc D = P.win[w];
    if (!D.owned) return;
    const int g = tile * kVarTile + (int)threadIdx.x;
    const int mp = g - D.var_base;
    int nv = 0, z0 = 0, z1 = 0;
    if (mp < D.M && P.st[g] == ST_FREE) {
        nv = 1;
        if (D.H > 0) {
            const int ob = __ldg(D.mp_obs_ptr + mp), oe = __ldg(D.mp_obs_ptr + mp + 1);
            if (ob < 0 || oe < ob || oe > D.O) {
                atomicOr(&P.ws[w].error, ERR_PTR);
            } else {
                for (int o = ob; o < oe; ++o) {
                    const int kf = __ldg(D.mp_obs_kf + o);
                    if (kf < D.K) { if (kf < 0) atomicOr(&P.ws[w].error, ERR_INDEX); continue; }
                    if (kf >= D.K + D.H) { atomicOr(&P.ws[w].error, ERR_INDEX); continue; }
                    const int jj = D.orow_base + (kf - D.K);
                    if (FILL) {
                        const int pos = atomicAdd(&P.ocursor[jj], 1);
                        P.orow_var[pos] = g;
                    } else {
                        atomicAdd(&P.ocnt[jj], 1);
                    }
                }
            }
        }
    }
    if (!FILL) {
        block_sum3(S, nv, z0, z1);
        if (threadIdx.x == 0 && nv) atomicAdd(&P.ws[w].n_vars, nv);
    }
}

// canonical integer rhs of an outside row (reference: MapSparsification.cc:146-147)
__device__ __forceinline__ int outside_need(int cnt, int total, int N) {
    if (cnt <= 0 || total <= 0) return 0;
    const float r = __fmul_rn(__fdiv_rn((float)cnt, (float)total), (float)N);
    return (int)ceil((double)r - 1e-5);
}

// block 0: exclusive scan of the outside-row counts, rhs of the outside rows, window error gate
__device__ void p3_scan(const Params& P, BlockScratch& S) {
    __shared__ int carry;
    __shared__ int warp_tot[kThreads / 32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < P.Htot; base += kThreads) {
        const int jj = base + (int)threadIdx.x;
        const int v = (jj < P.Htot) ? P.ocnt[jj] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        int woff = 0;
        for (int q = 0; q < wid; ++q) woff += warp_tot[q];
        const int excl = carry + woff + x - v;
        if (jj < P.Htot) {
            P.orow_ptr[jj] = excl;
            P.ocursor[jj] = excl;
            const int w = P.row_win[P.Ktot + jj];
            const WinDesc& D = P.win[w];
            const int total = D.owned ? __ldg(D.okf_total + (jj - D.orow_base)) : 0;
            P.row_need[P.Ktot + jj] = outside_need(v, total, P.N);
        }
        __syncthreads();
        if (threadIdx.x == kThreads - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) P.orow_ptr[P.Htot] = carry;
    // windows whose view failed validation never enter the phase machine
    int nact = 0, z0 = 0, z1 = 0;
    for (int w = threadIdx.x; w < P.nwin; w += kThreads) {
        WinState& s = P.ws[w];
        if (s.mode != MODE_DONE) {
            if (s.error) { s.mode = MODE_DONE; s.status = -1; }
            else ++nact;
        }
    }
    block_sum3(S, nact, z0, z1);
    if (threadIdx.x == 0) P.ctrl->n_active = nact;
}

// ---------------------------------------------------------------------------------------------------------------
// row phases (one CTA per row at a time)
// ---------------------------------------------------------------------------------------------------------------
// sweep 1 shared by all modes: per-cell and per-row counts of (IN-like, low-like) entries
template <class Classify>
__device__ __forceinline__ void count_sweep(const Row& R, const Params& P, unsigned* tab, BlockScratch& S,
                                            int& n_in, int& n_low, Classify cls) {
    int a = 0, b = 0, c = 0;
    for_each_entry(R, [&](int g, int cell) {
        const unsigned add = cls(P.st[g]);          // bit10 = counts as IN, bit0 = counts as low, bit20 = total
        if (add & (1u << 10)) ++a;
        if (add & 1u) ++b;
        if (cell >= 0 && add) atomicAdd(&tab[cell], add);
    });
    block_sum3(S, a, b, c);        // contains the barriers that publish tab
    n_in = a;
    n_low = b;
}

__device__ void row_prop(const Params& P, const Row& R, unsigned* tab, BlockScratch& S) {
    if (R.is_kf) { zero_tab(tab); __syncthreads(); }
    int cov, nfree;
    count_sweep(R, P, tab, S, cov, nfree,
                [](uint8_t s) -> unsigned { return s == ST_IN ? (1u << 10) : (s == ST_FREE ? 1u : 0u); });
    if (nfree == 0) return;
    const int d = max(0, R.need - cov);
    const bool defi = d > 0;
    const bool critr = defi && d >= nfree;
    for_each_entry(R, [&](int g, int cell) {
        if (P.st[g] != ST_FREE) return;
        unsigned long long add = 0;
        if (cell >= 0) {
            const unsigned t = tab[cell];
            if (tab_in(t) == 0) {
                add |= 1ull;
                if (tab_low(t) == 1) add |= 1ull << 16;
            }
        }
        if (defi) add |= 1ull << 32;
        if (critr) add |= 1ull << 48;
        if (add) atomicAdd(&P.acc[g], add);
    });
}

__device__ void row_greedy(const Params& P, const Row& R, unsigned* tab, unsigned long long* keytab, BlockScratch& S) {
    if (R.is_kf) { zero_tab(tab); zero_keytab(keytab); __syncthreads(); }
    int cov, nfree;
    count_sweep(R, P, tab, S, cov, nfree,
                [](uint8_t s) -> unsigned { return s == ST_IN ? (1u << 10) : (s == ST_FREE ? 1u : 0u); });
    if (nfree == 0) return;
    const int d = max(0, R.need - cov);
    const unsigned vb = (unsigned)R.var_base;
    if (R.is_kf) {
        for_each_entry(R, [&](int g, int cell) {
            if (P.st[g] != ST_FREE) return;
            if (tab_in(tab[cell]) != 0) return;
            atomicMax(&keytab[cell], make_key(P.gain[g], (unsigned)g - vb));
        });
        __syncthreads();
        for_each_entry(R, [&](int g, int cell) {
            if (P.st[g] != ST_FREE) return;
            if (tab_in(tab[cell]) != 0) return;
            if (make_key(P.gain[g], (unsigned)g - vb) != keytab[cell]) atomicOr(&P.acc[g], FLAG_BLOCKED);
        });
    }
    if (d > 0) {
        if (nfree > d) {
            const unsigned long long thr = block_kth_largest(S, d, [&](auto sink) {
                for_each_entry(R, [&](int g, int) {
                    if (P.st[g] == ST_FREE) sink(make_key(P.gain[g], (unsigned)g - vb));
                });
            });
            for_each_entry(R, [&](int g, int) {
                if (P.st[g] != ST_FREE) return;
                const bool adm = make_key(P.gain[g], (unsigned)g - vb) > thr;
                atomicOr(&P.acc[g], adm ? FLAG_NOMINATED : FLAG_BLOCKED);
            });
        } else {
            for_each_entry(R, [&](int g, int) {
                if (P.st[g] == ST_FREE) atomicOr(&P.acc[g], FLAG_NOMINATED);
            });
        }
    }
}

__device__ void row_d1(const Params& P, const Row& R, unsigned* tab, BlockScratch& S) {
    if (R.is_kf) { zero_tab(tab); __syncthreads(); }
    int cov, unused;
    count_sweep(R, P, tab, S, cov, unused, [](uint8_t s) -> unsigned { return s == ST_IN ? (1u << 10) : 0u; });
    if (cov == 0) return;
    const bool critr = cov <= R.need;
    for_each_entry(R, [&](int g, int cell) {
        if (P.st[g] != ST_IN) return;
        unsigned long long add = 0;
        if (cell >= 0 && tab_in(tab[cell]) == 1) add |= 1ull;
        if (critr) add |= 1ull << 32;
        if (add) atomicAdd(&P.acc[g], add);
    });
}

__device__ void row_d2(const Params& P, const Row& R, unsigned* tab, unsigned long long* keytab, BlockScratch& S) {
    if (R.is_kf) { zero_tab(tab); zero_keytab(keytab); __syncthreads(); }
    int cov, ncand;
    count_sweep(R, P, tab, S, cov, ncand, [](uint8_t s) -> unsigned {
        return s == ST_IN ? (1u << 10) : (s == ST_CAND ? ((1u << 10) | 1u) : 0u);
    });
    if (ncand == 0) return;
    const unsigned vb = (unsigned)R.var_base;
    if (R.is_kf) {
        for_each_entry(R, [&](int g, int cell) {
            if (P.st[g] != ST_CAND) return;
            if (tab_in(tab[cell]) < 2) return;
            atomicMax(&keytab[cell], make_key(P.gain[g], (unsigned)g - vb));
        });
        __syncthreads();
        for_each_entry(R, [&](int g, int cell) {
            if (P.st[g] != ST_CAND) return;
            if (tab_in(tab[cell]) < 2) return;
            if (make_key(P.gain[g], (unsigned)g - vb) != keytab[cell]) atomicOr(&P.acc[g], FLAG_BLOCKED);
        });
    }
    const int u = cov - R.need;
    if (u > 0 && ncand > u) {
        const unsigned long long thr = block_kth_largest(S, u, [&](auto sink) {
            for_each_entry(R, [&](int g, int) {
                if (P.st[g] == ST_CAND) sink(make_key(P.gain[g], (unsigned)g - vb));
            });
        });
        for_each_entry(R, [&](int g, int) {
            if (P.st[g] != ST_CAND) return;
            if (!(make_key(P.gain[g], (unsigned)g - vb) > thr)) atomicOr(&P.acc[g], FLAG_BLOCKED);
        });
    }
}

__device__ void row_eval(const Params& P, const Row& R, int r, int w, const WinDesc& D, unsigned* tab, BlockScratch& S) {
    if (R.is_kf) { zero_tab(tab); __syncthreads(); }
    int cov, unused;
    count_sweep(R, P, tab, S, cov, unused,
                [](uint8_t s) -> unsigned { return (1u << 20) | (s == ST_IN ? (1u << 10) : 0u); });
    int unc = 0, z0 = 0, z1 = 0;
    if (R.is_kf) {
        for (int c = threadIdx.x; c < kCells; c += kThreads) {
            const unsigned t = tab[c];
            if (tab_total(t) > 0 && tab_in(t) == 0) ++unc;
        }
        block_sum3(S, unc, z0, z1);
    }
    if (threadIdx.x == 0) {
        const int slack = max(0, R.need - cov);
        const int local = (r < P.Ktot) ? (r - D.row_base) : (D.K + (r - P.Ktot - D.orow_base));
        const int words = (D.M + 31) >> 5;
        uint32_t* slot = P.out + D.out_off + kHdrWords + words;
        slot[local] = (uint32_t)cov;
        slot[D.K + D.H + local] = (uint32_t)slack;
        if (unc) atomicAdd(&P.ws[w].uncovered, unc);
        if (slack) atomicAdd(&P.ws[w].total_slack, slack);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// variable phases (one thread per map point of a 256-wide tile)
// ---------------------------------------------------------------------------------------------------------------
__device__ void var_phase(const Params& P, int tile, BlockScratch& S) {
    const int w = P.tile_win[tile];
    WinState& ws = P.ws[w];
    const int mode = ws.mode;
    if (mode == MODE_DONE) return;
    const WinDesc D = P.win[w];
    const int g = tile * kVarTile + (int)threadIdx.x;
    const int mp = g - D.var_base;
    const bool inb = mp < D.M;
    const uint8_t s = inb ? P.st[g] : (uint8_t)ST_NOTVAR;
    int c0 = 0, c1 = 0, c2 = 0;
    switch (mode) {
    case MODE_PROP: {
        if (s == ST_FREE) {
            const unsigned long long a = P.acc[g];
            if (a) P.acc[g] = 0;
            const double ubc = (double)(a & 0xFFFFu), lbc = (double)((a >> 16) & 0xFFFFu);
            const double ubr = (double)((a >> 32) & 0xFFFFu), lbr = (double)(a >> 48);
            const double cost = (double)(ws.n_max - __ldg(D.mp_nobs + mp));
            const double ub = __dsub_rn(__dadd_rn(__dmul_rn(P.glam, ubc), __dmul_rn(P.lam, ubr)), cost);
            const double lb = __dsub_rn(__dadd_rn(__dmul_rn(P.glam, lbc), __dmul_rn(P.lam, lbr)), cost);
            if (ub <= 0.0) { P.st[g] = ST_OUT; c0 = 1; }
            else if (lb >= 0.0) { P.st[g] = ST_IN; c0 = 1; }
            else { P.gain[g] = (float)ub; c1 = 1; }
        }
        block_sum3(S, c0, c1, c2);
        if (threadIdx.x == 0) {
            if (c0) atomicAdd(&ws.changed, (unsigned)c0);
            if (c1) atomicAdd(&ws.nfree, (unsigned)c1);
        }
    } break;
    case MODE_GREEDY: {
        if (s == ST_FREE) {
            const unsigned long long a = P.acc[g];
            if (a) P.acc[g] = 0;
            const bool any_rule = ws.greedy_steps >= P.all_rule_steps;
            const bool sel = (P.gain[g] > 0.0f && !(a & FLAG_BLOCKED)) || (any_rule && (a & FLAG_NOMINATED));
            if (sel) P.st[g] = ST_IN;
        }
    } break;
    case MODE_FORCE: {
        if (s == ST_FREE) P.st[g] = ST_IN;
    } break;
    case MODE_D1: {
        if (s == ST_IN) {
            const unsigned long long a = P.acc[g];
            if (a) P.acc[g] = 0;
            const double critc = (double)(a & 0xFFFFu), critr = (double)((a >> 32) & 0xFFFFu);
            const double cost = (double)(ws.n_max - __ldg(D.mp_nobs + mp));
            const double dF = __dadd_rn(__dadd_rn(-cost, __dmul_rn(P.glam, critc)), __dmul_rn(P.lam, critr));
            if (dF < 0.0) { P.st[g] = ST_CAND; P.gain[g] = (float)(-dF); c0 = 1; }
        }
        block_sum3(S, c0, c1, c2);
        if (threadIdx.x == 0 && c0) atomicAdd(&ws.ncand, (unsigned)c0);
    } break;
    case MODE_D2: {
        if (s == ST_CAND) {
            const unsigned long long a = P.acc[g];
            if (a) P.acc[g] = 0;
            P.st[g] = (a & FLAG_BLOCKED) ? ST_IN : ST_OUT;
        }
    } break;
    case MODE_EVAL: {
        // read-out (MapSparsification.cc:159-166): bit = 0 only for variables the solve rejected
        const bool keep = inb && (s != ST_OUT);
        const unsigned word = __ballot_sync(0xFFFFFFFFu, keep);
        const int words = (D.M + 31) >> 5;
        const int widx = mp >> 5;
        if ((threadIdx.x & 31) == 0 && widx < words) P.out[D.out_off + kHdrWords + widx] = word;
        int kept = (s == ST_IN) ? 1 : 0;
        int cost = kept ? (ws.n_max - __ldg(D.mp_nobs + mp)) : 0;    // < 2^31 per block: 256 * nMax
        block_sum3(S, kept, cost, c2);
        if (threadIdx.x == 0 && kept) {
            atomicAdd(&ws.n_kept, kept);
            atomicAdd(&ws.sum_cost, (unsigned long long)cost);
        }
    } break;
    default: break;
    }
}

// per-window phase machine; run by the last CTA to finish the variable phase of an iteration
__device__ void transition(const Params& P, int w) {
    WinState& s = P.ws[w];
    const WinDesc& D = P.win[w];
    const int drop_mode = (P.max_drop_rounds > 0) ? MODE_D1 : MODE_EVAL;
    switch (s.mode) {
    case MODE_PROP: {
        s.rounds++;
        const unsigned changed = s.changed, nfree = s.nfree;
        s.changed = 0; s.nfree = 0;
        if (changed > 0 && s.rounds < P.max_rounds) s.mode = MODE_PROP;
        else if (nfree == 0) s.mode = drop_mode;
        else if (s.rounds >= P.max_rounds) { s.mode = MODE_FORCE; s.status = -5; }
        else s.mode = MODE_GREEDY;
    } break;
    case MODE_GREEDY: s.greedy_steps++; s.rounds++; s.mode = MODE_PROP; break;
    case MODE_FORCE: s.mode = drop_mode; break;
    case MODE_D1: {
        s.rounds++;
        const unsigned nc = s.ncand;
        s.ncand = 0;
        s.mode = (nc == 0) ? MODE_EVAL : MODE_D2;
    } break;
    case MODE_D2: s.drop_rounds++; s.mode = (s.drop_rounds >= P.max_drop_rounds) ? MODE_EVAL : MODE_D1; break;
    case MODE_EVAL: {
        uint32_t* hdr = P.out + D.out_off;
        hdr[0] = (uint32_t)s.status;
        hdr[1] = (uint32_t)s.rounds;
        hdr[2] = (uint32_t)s.n_max;
        hdr[3] = (uint32_t)s.n_vars;
        hdr[4] = (uint32_t)s.n_cells;
        hdr[5] = (uint32_t)s.nnz;
        hdr[6] = (uint32_t)s.n_kept;
        hdr[7] = (uint32_t)s.uncovered;
        hdr[8] = (uint32_t)s.total_slack;
        hdr[9] = (uint32_t)(s.sum_cost & 0xFFFFFFFFull);
        hdr[10] = (uint32_t)(s.sum_cost >> 32);
        hdr[11] = s.error;
        hdr[12] = (uint32_t)s.greedy_steps;
        hdr[13] = (uint32_t)s.drop_rounds;
        hdr[14] = 0x4D535331u;      // "MSS1": slot written
        hdr[15] = (uint32_t)w;
        s.mode = MODE_DONE;
    } break;
    default: break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the persistent cooperative kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) mss_persistent_kernel(const Params P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned tab[kCells];
    __shared__ unsigned long long keytab[kCells];
    __shared__ BlockScratch S;
    __shared__ int s_last;

    const int gtid = blockIdx.x * kThreads + threadIdx.x;
    const int gsize = gridDim.x * kThreads;

    // ---- P0: state init --------------------------------------------------------------------------------------
    if (gtid == 0) { P.ctrl->t_start = globaltimer_ns(); P.ctrl->ticket = 0u; P.ctrl->n_active = 0; P.ctrl->iters = 0; }
    {
        uint32_t* st32 = reinterpret_cast<uint32_t*>(P.st);
        for (int i = gtid; i < P.Mpad / 4; i += gsize) st32[i] = 0x03030303u;       // ST_NOTVAR
        for (int i = gtid; i < P.Mpad; i += gsize) P.acc[i] = 0ull;
        for (int i = gtid; i < P.Htot; i += gsize) P.ocnt[i] = 0;
        for (int w = gtid; w < P.nwin; w += gsize) {
            WinState z;
            memset(&z, 0, sizeof(z));
            z.mode = P.win[w].owned ? MODE_PROP : MODE_DONE;
            P.ws[w] = z;
            if (P.win[w].owned) {
                // header "not written" marker until EVAL completes
                P.out[P.win[w].out_off + 14] = 0u;
            }
        }
    }
    grid.sync();
    // ---- P1: keyframe rows: nMax, variable marking, cell statistics -------------------------------------------------
    for (int r = blockIdx.x; r < P.Ktot; r += gridDim.x) {
        if (!P.win[P.row_win[r]].owned) continue;
        p1_scan_row(P, r, tab, S);
        __syncthreads();
    }
    grid.sync();
    // ---- P2..P4: outside rows ----------------------------------------------------------------------------------
    for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x) p24_outside<false>(P, t, S);
    grid.sync();
    if (blockIdx.x == 0) p3_scan(P, S);
    grid.sync();
    if (P.Htot > 0) {
        for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x) p24_outside<true>(P, t, S);
        grid.sync();
    }
    if (gtid == 0) P.ctrl->t_build = globaltimer_ns();

    // ---- phase machine -----------------------------------------------------------------------------------------
    unsigned iter = 0;
    while (true) {
        const int n_active = *((volatile int*)&P.ctrl->n_active);
        if (n_active == 0) break;
        for (int r = blockIdx.x; r < P.Rtot; r += gridDim.x) {
            const int w = P.row_win[r];
            const int mode = P.ws[w].mode;
            if (mode == MODE_DONE || mode == MODE_FORCE) continue;
            const WinDesc D = P.win[w];
            const Row R = make_row(P, r, D);
            switch (mode) {
            case MODE_PROP: row_prop(P, R, tab, S); break;
            case MODE_GREEDY: row_greedy(P, R, tab, keytab, S); break;
            case MODE_D1: row_d1(P, R, tab, S); break;
            case MODE_D2: row_d2(P, R, tab, keytab, S); break;
            case MODE_EVAL: row_eval(P, R, r, w, D, tab, S); break;
            default: break;
            }
            __syncthreads();
        }
        grid.sync();
        for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x) var_phase(P, t, S);
        // last CTA to arrive advances every window's phase machine
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(&P.ctrl->ticket, 1u);
            s_last = (ticket == (iter + 1u) * gridDim.x - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            int nact = 0, z0 = 0, z1 = 0;
            for (int w = threadIdx.x; w < P.nwin; w += kThreads) {
                if (P.ws[w].mode != MODE_DONE) {
                    transition(P, w);
                    if (P.ws[w].mode != MODE_DONE) ++nact;
                }
            }
            block_sum3(S, nact, z0, z1);
            if (threadIdx.x == 0) { P.ctrl->n_active = nact; P.ctrl->iters = (int)iter + 1; }
            __threadfence();
        }
        ++iter;
        grid.sync();
    }
    if (gtid == 0) P.ctrl->t_end = globaltimer_ns();
}

}  // namespace mss

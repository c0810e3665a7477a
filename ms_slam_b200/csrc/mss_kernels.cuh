// mss_kernels.cuh -- device side of the sparsification engine (sm_100a), kernel generation 2.
//
// One persistent cooperative kernel solves a whole batch of independent windows.  CTAs are partitioned into GROUPS, one
// group per window (or a queue of windows per group when there are more windows than CTAs); a group synchronises with
// its own release/acquire barrier in global memory, so windows never wait for each other (no grid-wide sync anywhere).
//
// What it replaces in the reference (/root/reference/src/MapSparsification.cc):
//   :66-76   nMaxObservation scan            -> W1 (per keyframe row: block max -> atomicMax)
//   :78-123  variables + cell rows + KF rows -> W1 marks variables and writes the keyframe rows as a CSR of packed
//                                               (map point, cell) entries, counting-sorted by cell in shared memory, so the
//                                               entries of one cell row are a contiguous run of its keyframe row
//   :125-151 outside-keyframe rows           -> W2 count / W3 scan + rhs / W4 fill (CSR of the outside rows)
//   :153-157 GUROBI optimize()               -> per-window phase machine PROP / GREEDY / DROP on F(x)
//   :159-166 read-out of GRB_DoubleAttr_X    -> EVAL: ballot-packed keep bits + row coverage + F(x)
//
// Selection algorithm, per variable state FREE / IN / OUT (oracle/emulate.py restates it on the CPU, bit for bit):
//   PROP   exact dominance to a fixed point: ub_p <= 0 -> OUT, lb_p >= 0 -> IN
//   GREEDY conflict-free step: a FREE point is taken iff it is the best candidate of every uncovered cell it lies in
//          and within the top-deficit candidates of every deficient row it lies in
//   DROP   budgeted reverse delete; per-cell / per-row budgets make the summed deltas exact
// Every decision is made from integer counters (integer atomics) and keys with a unique tie-break, so the result does not
// depend on scheduling, on the group partition or on the order of entries inside a list.
//
// Work per round is proportional to the UNDECIDED part of the window: round 1 is fused into the build (W1/W4), round 2
// streams the CSR once, and from then on every row keeps a compacted "live list" of its FREE entries (cell field
// rewritten to kCellCov once the cell is covered) plus a running row coverage, so later rounds touch only those.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mss {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCells = 64 * 48;
constexpr int kCellsPerThread = kCells / kThreads;     // 12
constexpr int kVarTile = 256;          // var_base alignment; one var-pass tile belongs to exactly one window
constexpr int kHdrWords = 16;          // result-slot header
constexpr unsigned kCellNone = 0xFFFFu;     // view: slot whose keypoint is not in the grid
constexpr unsigned kCellCov = 0xFFFu;       // packed entry: "no cell / cell already covered"
constexpr int kCellBits = 12;
constexpr int kMaxWindowMps = 1 << 20;      // packed entry = (map point << 12) | cell
constexpr int kMaxWindowRows = 65535;       // 16-bit per-variable row counters
constexpr unsigned kEntInvalid = 0xFFFFFFFFu;
constexpr unsigned kTabCov = 0x80000000u;   // cell table: bit 31 = cell has an IN point, low 16 bits = FREE count
constexpr int kEpt = 8;                     // entries per thread held in registers (rows up to 2048 entries)
constexpr int kRegRow = kThreads * kEpt;

enum : uint8_t { ST_FREE = 0, ST_IN = 1, ST_OUT = 2, ST_NOTVAR = 3, ST_CAND = 4 };
enum : int { MODE_DONE = 0, MODE_PROP = 1, MODE_GREEDY = 2, MODE_FORCE = 3, MODE_D1 = 4, MODE_D2 = 5, MODE_EVAL = 6,
             MODE_EVALV = 7 };
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
    int row_base;    // first row of this window in the per-row arrays: K keyframe rows, then H outside rows
    int slot_base;   // first entry of this window's keyframe rows in ent[] / live[]
    int obs_base;    // first entry of this window's outside rows in ent[] / live[] (after all keyframe segments)
    int var_base;    // global variable index of map point 0 (multiple of kVarTile)
    int out_off;     // u32 word offset of this window's result slot
    int pad_[3];
};

// per-phase counters; three copies rotate so that a copy is zeroed two phases before it is used again
struct RoundCnt {
    unsigned changed, nfree, sumdeg, ncand;
    unsigned uncovered, slack, nkept, rows_live;
    unsigned long long sumcost;
    unsigned long long pad_;
};

struct WinState {
    int n_max, n_vars, n_cells, nnz;
    unsigned error;
    int pad_[3];
    RoundCnt rc[3];
};

struct GroupDesc {
    int cta0, ncta;      // CTAs [cta0, cta0 + ncta) work on this group's windows
    int wbeg, wend;      // windows gwin[wbeg .. wend), solved one after the other
};

struct Ctrl {
    unsigned long long t_start, t_build, t_end;
    int abort;           // set by the watchdog: a barrier waited longer than watchdog_ns
    int pad_;
};

struct Params {
    const WinDesc* win;
    WinState* ws;
    const GroupDesc* grp;
    const int* cta_grp;      // [grid] group of every CTA
    const int* gwin;         // window lists of the groups
    unsigned* gbar;          // one barrier counter per group, 128 B apart, zeroed before the launch
    uint8_t* st;             // [Mpad] variable state
    unsigned long long* acc; // [Mpad] packed 4 x 16-bit counters / flags
    float* gain;             // [Mpad]
    unsigned* deg;           // [Mpad] number of row entries of the variable
    uint32_t* ent;           // [Ftot + Otot] CSR entries: keyframe rows at slot_base + feat_ptr[k], then outside rows
    uint32_t* live;          // [Ftot + Otot] live lists (same segments)
    int* row_off;            // [Rtot] first entry of the row's segment
    int* ent_n;              // [Rtot] CSR entries of the row
    int* live_n;             // [Rtot] live entries of the row
    int* row_need;           // [Rtot]
    int* row_cov;            // [Rtot] IN entries of the row (running)
    int* row_ncell;          // [Rtot] occupied cells of the row
    int* ocursor;            // [Rtot] fill cursor of the outside rows
    uint32_t* out;           // result slots
    Ctrl* ctrl;
    int nwin, ngroups;
    int Ftot;                // outside-row segments start at ent + Ftot
    int N;
    int max_rounds, all_rule_steps, max_drop_rounds;
    double lam, glam;
    unsigned long long watchdog_ns;
};

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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
    int red[3][kWarps];
    int scan[kWarps];
    int bcast[4];
    unsigned hist[256];
};

struct GroupCtx {
    unsigned* bar;
    unsigned gen;
    int ncta;
    int cta;        // index of this CTA inside the group
    int* s_abort;   // shared flag
    unsigned long long t0;   // this CTA's start time (watchdog reference)
};

// Barrier of the CTAs of one group.  Thread 0 publishes the CTA's writes with a release add and waits with acquire
// loads; bar.sync on both sides extends the ordering to the whole CTA (same construction as a cooperative grid sync,
// but scoped to the CTAs that actually share the window).  Returns false when the launch is being aborted.
__device__ __forceinline__ bool group_sync(const Params& P, GroupCtx& G) {
    __syncthreads();
    G.gen += 1u;
    if (threadIdx.x == 0) {
        int ab = 0;
        if (G.ncta > 1) {
            red_release_add_u32(G.bar, 1u);
            const unsigned target = G.gen * (unsigned)G.ncta;
            unsigned spins = 0;
            while (ld_acquire_u32(G.bar) < target) {
                if ((++spins & 0xFFu) == 0u) {
                    if (*(volatile int*)&P.ctrl->abort) { ab = 1; break; }
                    if (globaltimer_ns() - G.t0 > P.watchdog_ns) {
                        atomicExch(&P.ctrl->abort, 1);
                        ab = 1;
                        break;
                    }
                }
            }
        }
        if (!ab) ab = *(volatile int*)&P.ctrl->abort;
        *G.s_abort = ab;
    }
    __syncthreads();
    return *G.s_abort == 0;
}

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
    for (int w = 0; w < kWarps; ++w) { a += S.red[0][w]; b += S.red[1][w]; c += S.red[2][w]; }
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
    for (int w = 1; w < kWarps; ++w) m = max(m, S.red[0][w]);
    return m;
}

// exclusive prefix of one int per thread (thread order); total returned to every thread
__device__ __forceinline__ int block_excl_scan(BlockScratch& S, int v, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) S.scan[wid] = x;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        const int t = S.scan[q];
        if (q < wid) woff += t;
        tot += t;
    }
    total = tot;
    return woff + x - v;
}

// exclusive prefix of one int per WARP (value taken from lane 0 of each warp); returns this warp's offset
__device__ __forceinline__ int warp_excl_scan(BlockScratch& S, int warp_val, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) S.scan[wid] = warp_val;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        const int t = S.scan[q];
        if (q < wid) woff += t;
        tot += t;
    }
    total = tot;
    return woff;
}

// (rank+1)-th largest key (rank elements are larger, counting multiplicity) among the keys produced by `emit`.
// The caller guarantees that more than `rank` keys exist.  8 radix passes of 8 bits, most significant first.
template <class Emit>
__device__ unsigned long long block_kth_largest(BlockScratch& S, int rank, Emit emit) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        __syncthreads();
        S.hist[threadIdx.x] = 0;            // kThreads == 256 bins
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

__device__ __forceinline__ void zero_tab(unsigned* tab) {
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) tab[threadIdx.x + j * kThreads] = 0u;
}

// canonical integer rhs of an outside row (reference: MapSparsification.cc:146-147)
__device__ __forceinline__ int outside_need(int cnt, int total, int N) {
    if (cnt <= 0 || total <= 0) return 0;
    const float r = __fmul_rn(__fdiv_rn((float)cnt, (float)total), (float)N);
    return (int)ceil((double)r - 1e-5);
}

// A row's entries held in registers: warp w owns a contiguous chunk of the list, lane-strided inside it, so global
// accesses are coalesced and (warp, b, lane) order is list order (needed for the stable compaction).
struct RowRegs {
    uint32_t e[kEpt];
    uint8_t s[kEpt];
    int per, nb;
};

__device__ __forceinline__ void load_row(RowRegs& X, const uint32_t* __restrict__ src, int n, const uint8_t* __restrict__ st_w) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    X.per = (((n + kWarps - 1) / kWarps) + 31) & ~31;
    X.nb = X.per >> 5;
#pragma unroll
    for (int b = 0; b < kEpt; ++b) {
        const int idx = wid * X.per + b * 32 + lane;
        X.e[b] = (b < X.nb && idx < n) ? src[idx] : kEntInvalid;
    }
#pragma unroll
    for (int b = 0; b < kEpt; ++b) X.s[b] = (X.e[b] != kEntInvalid) ? st_w[X.e[b] >> kCellBits] : (uint8_t)ST_NOTVAR;
}

// ---------------------------------------------------------------------------------------------------------------
// W1: keyframe row -> nMax, variable marking, cell-sorted CSR segment, round-1 contributions
// ---------------------------------------------------------------------------------------------------------------
__device__ void w1_build_row(const Params& P, const WinDesc& D, WinState& ws, int k, unsigned* tab, unsigned* cursor,
                             BlockScratch& S) {
    const int R = D.row_base + k;
    const int beg = __ldg(D.feat_ptr + k), end = __ldg(D.feat_ptr + k + 1);
    if (beg < 0 || end < beg || end > D.F) {
        if (threadIdx.x == 0) {
            atomicOr(&ws.error, ERR_PTR);
            P.ent_n[R] = 0; P.live_n[R] = 0; P.row_need[R] = P.N; P.row_cov[R] = 0; P.row_ncell[R] = 0; P.row_off[R] = 0;
        }
        return;
    }
    zero_tab(tab);
    __syncthreads();
    unsigned err = 0;
    int nmax = 0, nz = 0;
    uint8_t* st_w = P.st + D.var_base;
    for (int i = beg + (int)threadIdx.x; i < end; i += kThreads) {
        const int mp = __ldg(D.feat_mp + i);
        if (mp < 0) { if (mp < -1) err |= ERR_INDEX; continue; }
        if (mp >= D.M) { err |= ERR_INDEX; continue; }
        nmax = max(nmax, __ldg(D.mp_nobs + mp));                 // MapSparsification.cc:69-75: every valid slot
        const unsigned c = __ldg(D.feat_cell + i);
        if (c == kCellNone) continue;                           // not in mGrid: not a variable through this slot
        if (c >= (unsigned)kCells) { err |= ERR_INDEX; continue; }
        st_w[mp] = ST_FREE;                                     // MapSparsification.cc:91-99 (same value from every writer)
        atomicAdd(&tab[c], 1u);
        ++nz;
    }
    nmax = block_max(S, nmax);
    int z0 = 0, z1 = 0;
    block_sum3(S, nz, z0, z1);                                  // nz = entries of the row; barriers publish tab
    // exclusive scan of the cell histogram -> scatter cursors (counting sort by cell)
    int local = 0, ncell = 0;
    const int c0 = (int)threadIdx.x * kCellsPerThread;
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) {
        const unsigned t = tab[c0 + j];
        local += (int)t;
        ncell += t ? 1 : 0;
        if (t > 0xFFFFu) err |= ERR_CELL_OVERFLOW;
    }
    int total;
    int base = block_excl_scan(S, local, total);
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) {
        cursor[c0 + j] = (unsigned)base;
        base += (int)tab[c0 + j];
    }
    block_sum3(S, ncell, z0, z1);                               // barriers publish cursor
    const int seg = D.slot_base + beg;
    // round 1 (everything FREE, nothing IN): every cell is uncovered, every row with N > 0 is deficient
    const bool defi = P.N > 0;
    const bool critr = defi && P.N >= nz;
    uint32_t* ent = P.ent + seg;
    unsigned long long* acc_w = P.acc + D.var_base;
    for (int i = beg + (int)threadIdx.x; i < end; i += kThreads) {
        const int mp = __ldg(D.feat_mp + i);
        if (mp < 0 || mp >= D.M) continue;
        const unsigned c = __ldg(D.feat_cell + i);
        if (c >= (unsigned)kCells) continue;
        const unsigned pos = atomicAdd(&cursor[c], 1u);
        ent[pos] = ((uint32_t)mp << kCellBits) | c;
        unsigned long long add = 1ull;
        if (tab[c] == 1u) add |= 1ull << 16;
        if (defi) add |= 1ull << 32;
        if (critr) add |= 1ull << 48;
        atomicAdd(&acc_w[mp], add);
    }
    if (threadIdx.x == 0) {
        P.ent_n[R] = nz; P.live_n[R] = 0; P.row_need[R] = P.N; P.row_cov[R] = 0; P.row_ncell[R] = ncell; P.row_off[R] = seg;
        atomicMax(&ws.n_max, nmax);
        if (nz) atomicAdd(&ws.nnz, nz);
        if (ncell) atomicAdd(&ws.n_cells, ncell);
    }
    if (err) atomicOr(&ws.error, err);
}

// W2: per variable, count it and count its observations by outside keyframes (MapSparsification.cc:127-142)
__device__ void w2_count_outside(const Params& P, const WinDesc& D, WinState& ws, int tile, BlockScratch& S) {
    const int mp = tile * kVarTile + (int)threadIdx.x;
    int nv = 0, z0 = 0, z1 = 0;
    if (mp < D.M && P.st[D.var_base + mp] == ST_FREE) {
        nv = 1;
        if (D.H > 0) {
            const int ob = __ldg(D.mp_obs_ptr + mp), oe = __ldg(D.mp_obs_ptr + mp + 1);
            if (ob < 0 || oe < ob || oe > D.O) {
                atomicOr(&ws.error, ERR_PTR);
            } else {
                for (int o = ob; o < oe; ++o) {
                    const int kf = __ldg(D.mp_obs_kf + o);
                    if (kf < D.K) { if (kf < 0) atomicOr(&ws.error, ERR_INDEX); continue; }
                    if (kf >= D.K + D.H) { atomicOr(&ws.error, ERR_INDEX); continue; }
                    atomicAdd(&P.ent_n[D.row_base + kf], 1);
                }
            }
        }
    }
    block_sum3(S, nv, z0, z1);
    if (threadIdx.x == 0 && nv) atomicAdd(&ws.n_vars, nv);
}

// W3 (one CTA per window): exclusive scan of the outside-row counts -> segments, rhs of the outside rows
__device__ void w3_scan_outside(const Params& P, const WinDesc& D, BlockScratch& S) {
    int carry = 0;
    for (int base = 0; base < D.H; base += kThreads) {
        const int j = base + (int)threadIdx.x;
        const int R = D.row_base + D.K + j;
        const int v = (j < D.H) ? P.ent_n[R] : 0;
        int total;
        const int excl = carry + block_excl_scan(S, v, total);
        if (j < D.H) {
            const int off = P.Ftot + D.obs_base + excl;
            P.row_off[R] = off;
            P.ocursor[R] = off;
            P.row_need[R] = outside_need(v, __ldg(D.okf_total + j), P.N);
            P.row_cov[R] = 0;
            P.row_ncell[R] = 0;
            P.live_n[R] = 0;
        }
        carry += total;
    }
}

// Decision of a FREE variable from its packed counters (shared by W4 and the PROP variable phase)
__device__ __forceinline__ void prop_decide(const Params& P, unsigned long long a, int cost_i, uint8_t* st, float* gain, unsigned deg,
                                            int& changed, int& nfree, int& sumdeg) {
    const double ubc = (double)(a & 0xFFFFu), lbc = (double)((a >> 16) & 0xFFFFu);
    const double ubr = (double)((a >> 32) & 0xFFFFu), lbr = (double)(a >> 48);
    const double cost = (double)cost_i;
    const double ub = __dsub_rn(__dadd_rn(__dmul_rn(P.glam, ubc), __dmul_rn(P.lam, ubr)), cost);
    const double lb = __dsub_rn(__dadd_rn(__dmul_rn(P.glam, lbc), __dmul_rn(P.lam, lbr)), cost);
    if (ub <= 0.0) { *st = ST_OUT; changed += 1; }
    else if (lb >= 0.0) { *st = ST_IN; changed += 1; }
    else { *gain = (float)ub; nfree += 1; sumdeg += (int)deg; }
}

// W4: per variable, fill the outside rows, add their round-1 contributions and take the round-1 decision
__device__ void w4_fill_and_round1(const Params& P, const WinDesc& D, WinState& ws, int tile, BlockScratch& S) {
    const int mp = tile * kVarTile + (int)threadIdx.x;
    const int g = D.var_base + mp;
    int changed = 0, nfree = 0, sumdeg = 0;
    if (mp < D.M && P.st[g] == ST_FREE) {
        unsigned long long a = P.acc[g];
        P.acc[g] = 0ull;
        unsigned nout = 0;
        if (D.H > 0) {
            const int ob = __ldg(D.mp_obs_ptr + mp), oe = __ldg(D.mp_obs_ptr + mp + 1);
            for (int o = ob; o < oe; ++o) {
                const int kf = __ldg(D.mp_obs_kf + o);
                if (kf < D.K) continue;
                const int R = D.row_base + kf;
                const int pos = atomicAdd(&P.ocursor[R], 1);
                P.ent[pos] = ((uint32_t)mp << kCellBits) | kCellCov;
                ++nout;
                const int need = P.row_need[R];
                if (need > 0) {
                    a += 1ull << 32;
                    if (need >= P.ent_n[R]) a += 1ull << 48;
                }
            }
        }
        const unsigned deg = (unsigned)(a & 0xFFFFu) + nout;
        P.deg[g] = deg;
        prop_decide(P, a, ws.n_max - __ldg(D.mp_nobs + mp), &P.st[g], &P.gain[g], deg, changed, nfree, sumdeg);
    }
    block_sum3(S, changed, nfree, sumdeg);
    if (threadIdx.x == 0) {
        RoundCnt& rc = ws.rc[0];
        if (changed) atomicAdd(&rc.changed, (unsigned)changed);
        if (nfree) atomicAdd(&rc.nfree, (unsigned)nfree);
        if (sumdeg) atomicAdd(&rc.sumdeg, (unsigned)sumdeg);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// row phases (one CTA per row at a time)
// ---------------------------------------------------------------------------------------------------------------
// PROP: counts on the current state, contributions to the FREE variables, and the row's new live list.
// Source list: the CSR (first PROP row phase of the window) or the previous live list.  Entries found IN are added to the
// running coverage exactly once (they are not copied to the new list); entries found OUT are dropped.
__device__ void row_prop(const Params& P, const WinDesc& D, RoundCnt& rc, int R, bool from_csr, unsigned* tab, BlockScratch& S) {
    const int n = from_csr ? P.ent_n[R] : P.live_n[R];
    if (n == 0) return;                                     // live_n[R] is already 0 (W1 / W3 / previous round)
    const int off = P.row_off[R];
    const uint32_t* src = (from_csr ? P.ent : P.live) + off;
    uint32_t* dst = P.live + off;
    const uint8_t* st_w = P.st + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R];
    const int cov0 = P.row_cov[R];
    const int lane = threadIdx.x & 31;
    if (n <= kRegRow) {
        RowRegs X;
        load_row(X, src, n, st_w);
#pragma unroll
        for (int b = 0; b < kEpt; ++b)
            if (X.e[b] != kEntInvalid && (X.e[b] & kCellCov) != kCellCov) tab[X.e[b] & kCellCov] = 0u;
        __syncthreads();
        int cin = 0, cfree = 0, z = 0;
#pragma unroll
        for (int b = 0; b < kEpt; ++b) {
            if (X.e[b] == kEntInvalid) continue;
            const unsigned cell = X.e[b] & kCellCov;
            if (X.s[b] == ST_IN) { ++cin; if (cell != kCellCov) atomicOr(&tab[cell], kTabCov); }
            else if (X.s[b] == ST_FREE) { ++cfree; if (cell != kCellCov) atomicAdd(&tab[cell], 1u); }
        }
        block_sum3(S, cin, cfree, z);                       // barriers publish tab
        const int cov = cov0 + cin;
        const int d = max(0, need - cov);
        const bool defi = d > 0, critr = defi && d >= cfree;
        int wcnt = 0;
        unsigned m[kEpt];
#pragma unroll
        for (int b = 0; b < kEpt; ++b) {
            const bool fr = (X.e[b] != kEntInvalid) && X.s[b] == ST_FREE;
            if (fr) {
                const unsigned cell = X.e[b] & kCellCov;
                unsigned long long add = 0;
                bool covered = true;
                if (cell != kCellCov) {
                    const unsigned t = tab[cell];
                    covered = (t & kTabCov) != 0u;
                    if (!covered) { add |= 1ull; if ((t & 0xFFFFu) == 1u) add |= 1ull << 16; }
                }
                if (defi) add |= 1ull << 32;
                if (critr) add |= 1ull << 48;
                if (add) atomicAdd(&acc_w[X.e[b] >> kCellBits], add);
                if (covered) X.e[b] |= kCellCov;
            }
            m[b] = __ballot_sync(0xFFFFFFFFu, fr);
            wcnt += __popc(m[b]);
        }
        int total;
        int pos = warp_excl_scan(S, wcnt, total);           // barrier: every read of src precedes the in-place writes
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int b = 0; b < kEpt; ++b) {
            if ((m[b] >> lane) & 1u) dst[pos + __popc(m[b] & lt)] = X.e[b];
            pos += __popc(m[b]);
        }
        if (threadIdx.x == 0) {
            P.row_cov[R] = cov;
            P.live_n[R] = cfree;
            if (cfree) atomicAdd(&rc.rows_live, 1u);
        }
    } else {
        // long row: two passes over the list in global memory
        zero_tab(tab);
        __syncthreads();
        int cin = 0, cfree = 0, z = 0;
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const uint32_t e = src[i];
            const uint8_t s = st_w[e >> kCellBits];
            const unsigned cell = e & kCellCov;
            if (s == ST_IN) { ++cin; if (cell != kCellCov) atomicOr(&tab[cell], kTabCov); }
            else if (s == ST_FREE) { ++cfree; if (cell != kCellCov) atomicAdd(&tab[cell], 1u); }
        }
        block_sum3(S, cin, cfree, z);
        const int cov = cov0 + cin;
        const int d = max(0, need - cov);
        const bool defi = d > 0, critr = defi && d >= cfree;
        int out_base = 0;
        for (int base = 0; base < n; base += kThreads) {
            const int i = base + (int)threadIdx.x;
            uint32_t e = kEntInvalid;
            bool fr = false;
            if (i < n) {
                e = src[i];
                fr = st_w[e >> kCellBits] == ST_FREE;
            }
            if (fr) {
                const unsigned cell = e & kCellCov;
                unsigned long long add = 0;
                bool covered = true;
                if (cell != kCellCov) {
                    const unsigned t = tab[cell];
                    covered = (t & kTabCov) != 0u;
                    if (!covered) { add |= 1ull; if ((t & 0xFFFFu) == 1u) add |= 1ull << 16; }
                }
                if (defi) add |= 1ull << 32;
                if (critr) add |= 1ull << 48;
                if (add) atomicAdd(&acc_w[e >> kCellBits], add);
                if (covered) e |= kCellCov;
            }
            int total;
            const int p = block_excl_scan(S, fr ? 1 : 0, total);    // barriers: chunk read before it is overwritten
            if (fr) dst[out_base + p] = e;
            out_base += total;
        }
        if (threadIdx.x == 0) {
            P.row_cov[R] = cov;
            P.live_n[R] = cfree;
            if (cfree) atomicAdd(&rc.rows_live, 1u);
        }
    }
}

// GREEDY: runs right after a PROP round that changed nothing, so the live list is exact (all FREE, cell field = covered
// flag, row_cov current).
__device__ void row_greedy(const Params& P, const WinDesc& D, int R, unsigned long long* keytab, BlockScratch& S) {
    const int n = P.live_n[R];
    if (n == 0) return;
    const uint32_t* src = P.live + P.row_off[R];
    const uint8_t* st_w = P.st + D.var_base;
    const float* gain_w = P.gain + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int d = max(0, P.row_need[R] - P.row_cov[R]);
    if (n <= kRegRow) {
        RowRegs X;
        load_row(X, src, n, st_w);
        unsigned long long key[kEpt];
        int nfree = 0;
#pragma unroll
        for (int b = 0; b < kEpt; ++b) {
            const bool fr = X.e[b] != kEntInvalid && X.s[b] == ST_FREE;
            key[b] = fr ? make_key(gain_w[X.e[b] >> kCellBits], X.e[b] >> kCellBits) : 0ull;
            if (fr) { ++nfree; if ((X.e[b] & kCellCov) != kCellCov) keytab[X.e[b] & kCellCov] = 0ull; }
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < kEpt; ++b)
            if (key[b] && (X.e[b] & kCellCov) != kCellCov) atomicMax(&keytab[X.e[b] & kCellCov], key[b]);
        __syncthreads();
#pragma unroll
        for (int b = 0; b < kEpt; ++b)
            if (key[b] && (X.e[b] & kCellCov) != kCellCov && keytab[X.e[b] & kCellCov] != key[b])
                atomicOr(&acc_w[X.e[b] >> kCellBits], FLAG_BLOCKED);
        if (d > 0) {
            int z0 = 0, z1 = 0;
            block_sum3(S, nfree, z0, z1);
            if (nfree > d) {
                const unsigned long long thr = block_kth_largest(S, d, [&](auto sink) {
#pragma unroll
                    for (int b = 0; b < kEpt; ++b) if (key[b]) sink(key[b]);
                });
#pragma unroll
                for (int b = 0; b < kEpt; ++b)
                    if (key[b]) atomicOr(&acc_w[X.e[b] >> kCellBits], key[b] > thr ? FLAG_NOMINATED : FLAG_BLOCKED);
            } else {
#pragma unroll
                for (int b = 0; b < kEpt; ++b) if (key[b]) atomicOr(&acc_w[X.e[b] >> kCellBits], FLAG_NOMINATED);
            }
        }
    } else {
        for (int c = threadIdx.x; c < kCells; c += kThreads) keytab[c] = 0ull;
        __syncthreads();
        int nfree = 0;
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const uint32_t e = src[i];
            const unsigned v = e >> kCellBits;
            if (st_w[v] != ST_FREE) continue;
            ++nfree;
            if ((e & kCellCov) != kCellCov) atomicMax(&keytab[e & kCellCov], make_key(gain_w[v], v));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const uint32_t e = src[i];
            const unsigned v = e >> kCellBits;
            if (st_w[v] != ST_FREE || (e & kCellCov) == kCellCov) continue;
            if (keytab[e & kCellCov] != make_key(gain_w[v], v)) atomicOr(&acc_w[v], FLAG_BLOCKED);
        }
        if (d > 0) {
            int z0 = 0, z1 = 0;
            block_sum3(S, nfree, z0, z1);
            if (nfree > d) {
                const unsigned long long thr = block_kth_largest(S, d, [&](auto sink) {
                    for (int i = threadIdx.x; i < n; i += kThreads) {
                        const unsigned v = src[i] >> kCellBits;
                        if (st_w[v] == ST_FREE) sink(make_key(gain_w[v], v));
                    }
                });
                for (int i = threadIdx.x; i < n; i += kThreads) {
                    const unsigned v = src[i] >> kCellBits;
                    if (st_w[v] != ST_FREE) continue;
                    atomicOr(&acc_w[v], make_key(gain_w[v], v) > thr ? FLAG_NOMINATED : FLAG_BLOCKED);
                }
            } else {
                for (int i = threadIdx.x; i < n; i += kThreads) {
                    const unsigned v = src[i] >> kCellBits;
                    if (st_w[v] == ST_FREE) atomicOr(&acc_w[v], FLAG_NOMINATED);
                }
            }
        }
    }
}

// D1 (and EVAL): one sweep of the row's CSR segment: IN counts per cell and per row; D1 adds the criticality counters
// of the IN points; both write the row's coverage / slack and the uncovered-cell count (the read-out uses the values of
// the last sweep, which is the one that found nothing left to drop).
__device__ void row_d1_eval(const Params& P, const WinDesc& D, RoundCnt& rc, int R, bool accumulate, unsigned* tab, BlockScratch& S) {
    const int n = P.ent_n[R];
    const uint32_t* src = P.ent + P.row_off[R];
    const uint8_t* st_w = P.st + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R];
    int cin = 0, ccells = 0, z = 0;
    if (n > 0 && n <= kRegRow) {
        RowRegs X;
        load_row(X, src, n, st_w);
#pragma unroll
        for (int b = 0; b < kEpt; ++b)
            if (X.e[b] != kEntInvalid && (X.e[b] & kCellCov) != kCellCov) tab[X.e[b] & kCellCov] = 0u;
        __syncthreads();
#pragma unroll
        for (int b = 0; b < kEpt; ++b) {
            if (X.e[b] == kEntInvalid || X.s[b] != ST_IN) continue;
            ++cin;
            const unsigned cell = X.e[b] & kCellCov;
            if (cell != kCellCov && atomicAdd(&tab[cell], 1u) == 0u) ++ccells;
        }
        block_sum3(S, cin, ccells, z);
        if (accumulate && cin > 0) {
            const bool critr = cin <= need;
#pragma unroll
            for (int b = 0; b < kEpt; ++b) {
                if (X.e[b] == kEntInvalid || X.s[b] != ST_IN) continue;
                const unsigned cell = X.e[b] & kCellCov;
                unsigned long long add = 0;
                if (cell != kCellCov && tab[cell] == 1u) add |= 1ull;
                if (critr) add |= 1ull << 32;
                if (add) atomicAdd(&acc_w[X.e[b] >> kCellBits], add);
            }
        }
    } else if (n > 0) {
        zero_tab(tab);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const uint32_t e = src[i];
            if (st_w[e >> kCellBits] != ST_IN) continue;
            ++cin;
            const unsigned cell = e & kCellCov;
            if (cell != kCellCov && atomicAdd(&tab[cell], 1u) == 0u) ++ccells;
        }
        block_sum3(S, cin, ccells, z);
        if (accumulate && cin > 0) {
            const bool critr = cin <= need;
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const uint32_t e = src[i];
                if (st_w[e >> kCellBits] != ST_IN) continue;
                const unsigned cell = e & kCellCov;
                unsigned long long add = 0;
                if (cell != kCellCov && tab[cell] == 1u) add |= 1ull;
                if (critr) add |= 1ull << 32;
                if (add) atomicAdd(&acc_w[e >> kCellBits], add);
            }
        }
    }
    if (threadIdx.x == 0) {
        const int slack = max(0, need - cin);
        const int local = R - D.row_base;
        const int words = (D.M + 31) >> 5;
        uint32_t* slot = P.out + D.out_off + kHdrWords + words;
        slot[local] = (uint32_t)cin;
        slot[D.K + D.H + local] = (uint32_t)slack;
        const int unc = P.row_ncell[R] - ccells;
        if (unc) atomicAdd(&rc.uncovered, (unsigned)unc);
        if (slack) atomicAdd(&rc.slack, (unsigned)slack);
    }
}

// D2: budgets of the reverse delete (per cell: keep at least one IN point; per row: at most cov - need removals)
__device__ void row_d2(const Params& P, const WinDesc& D, int R, unsigned* tab, unsigned long long* keytab, BlockScratch& S) {
    const int n = P.ent_n[R];
    if (n == 0) return;
    const uint32_t* src = P.ent + P.row_off[R];
    const uint8_t* st_w = P.st + D.var_base;
    const float* gain_w = P.gain + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R];
    zero_tab(tab);
    for (int c = threadIdx.x; c < kCells; c += kThreads) keytab[c] = 0ull;
    __syncthreads();
    int cov = 0, ncand = 0, z = 0;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const uint32_t e = src[i];
        const uint8_t s = st_w[e >> kCellBits];
        if (s != ST_IN && s != ST_CAND) continue;
        ++cov;
        if (s == ST_CAND) ++ncand;
        if ((e & kCellCov) != kCellCov) atomicAdd(&tab[e & kCellCov], 1u);
    }
    block_sum3(S, cov, ncand, z);
    if (ncand == 0) return;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const uint32_t e = src[i];
        const unsigned v = e >> kCellBits, cell = e & kCellCov;
        if (st_w[v] != ST_CAND || cell == kCellCov || tab[cell] < 2u) continue;
        atomicMax(&keytab[cell], make_key(gain_w[v], v));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const uint32_t e = src[i];
        const unsigned v = e >> kCellBits, cell = e & kCellCov;
        if (st_w[v] != ST_CAND || cell == kCellCov || tab[cell] < 2u) continue;
        if (make_key(gain_w[v], v) != keytab[cell]) atomicOr(&acc_w[v], FLAG_BLOCKED);
    }
    const int u = cov - need;
    if (u > 0 && ncand > u) {
        const unsigned long long thr = block_kth_largest(S, u, [&](auto sink) {
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const unsigned v = src[i] >> kCellBits;
                if (st_w[v] == ST_CAND) sink(make_key(gain_w[v], v));
            }
        });
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const unsigned v = src[i] >> kCellBits;
            if (st_w[v] != ST_CAND) continue;
            if (!(make_key(gain_w[v], v) > thr)) atomicOr(&acc_w[v], FLAG_BLOCKED);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// variable phases (one thread per map point of a 256-wide tile)
// ---------------------------------------------------------------------------------------------------------------
__device__ void var_phase(const Params& P, const WinDesc& D, WinState& ws, RoundCnt& rc, int mode, int greedy_steps, int tile,
                          BlockScratch& S) {
    const int mp = tile * kVarTile + (int)threadIdx.x;
    const int g = D.var_base + mp;
    const bool inb = mp < D.M;
    const uint8_t s = inb ? P.st[g] : (uint8_t)ST_NOTVAR;
    int c0 = 0, c1 = 0, c2 = 0;
    switch (mode) {
    case MODE_PROP: {
        if (s == ST_FREE) {
            const unsigned long long a = P.acc[g];
            if (a) P.acc[g] = 0ull;
            prop_decide(P, a, ws.n_max - __ldg(D.mp_nobs + mp), &P.st[g], &P.gain[g], P.deg[g], c0, c1, c2);
        }
        if (__syncthreads_or(s == ST_FREE)) {
            block_sum3(S, c0, c1, c2);
            if (threadIdx.x == 0) {
                if (c0) atomicAdd(&rc.changed, (unsigned)c0);
                if (c1) atomicAdd(&rc.nfree, (unsigned)c1);
                if (c2) atomicAdd(&rc.sumdeg, (unsigned)c2);
            }
        }
    } break;
    case MODE_GREEDY: {
        if (s == ST_FREE) {
            const unsigned long long a = P.acc[g];
            if (a) P.acc[g] = 0ull;
            const bool any_rule = greedy_steps >= P.all_rule_steps;
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
            if (a) P.acc[g] = 0ull;
            const double critc = (double)(a & 0xFFFFu), critr = (double)((a >> 32) & 0xFFFFu);
            const double cost = (double)(ws.n_max - __ldg(D.mp_nobs + mp));
            const double dF = __dadd_rn(__dadd_rn(-cost, __dmul_rn(P.glam, critc)), __dmul_rn(P.lam, critr));
            if (dF < 0.0) { P.st[g] = ST_CAND; P.gain[g] = (float)(-dF); c0 = 1; }
        }
        if (__syncthreads_or(c0)) {
            block_sum3(S, c0, c1, c2);
            if (threadIdx.x == 0 && c0) atomicAdd(&rc.ncand, (unsigned)c0);
        }
    } break;
    case MODE_D2: {
        if (s == ST_CAND) {
            const unsigned long long a = P.acc[g];
            if (a) P.acc[g] = 0ull;
            P.st[g] = (a & FLAG_BLOCKED) ? ST_IN : ST_OUT;
        }
    } break;
    case MODE_EVAL:
    case MODE_EVALV: {
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
            atomicAdd(&rc.nkept, (unsigned)kept);
            atomicAdd(&rc.sumcost, (unsigned long long)cost);
        }
    } break;
    default: break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// one window, solved by the CTAs of one group
// ---------------------------------------------------------------------------------------------------------------
__device__ bool solve_window(const Params& P, GroupCtx& G, int w, unsigned* tab, unsigned long long* keytab, BlockScratch& S) {
    const WinDesc D = P.win[w];
    WinState& ws = P.ws[w];
    const int rows = D.K + D.H;
    const int tiles = (D.M + kVarTile - 1) / kVarTile;
    const int gt = G.cta * kThreads + (int)threadIdx.x, gsz = G.ncta * kThreads;

    // ---- W0: state init ------------------------------------------------------------------------------------------
    {
        const int mpad = ((max(D.M, 1) + kVarTile - 1) / kVarTile) * kVarTile;
        uint32_t* st32 = reinterpret_cast<uint32_t*>(P.st + D.var_base);
        for (int i = gt; i < mpad / 4; i += gsz) st32[i] = 0x03030303u;            // ST_NOTVAR
        for (int i = gt; i < mpad; i += gsz) P.acc[D.var_base + i] = 0ull;
        for (int j = gt; j < D.H; j += gsz) P.ent_n[D.row_base + D.K + j] = 0;
        if (G.cta == 0) {
            uint32_t* z = reinterpret_cast<uint32_t*>(&ws);
            for (int i = threadIdx.x; i < (int)(sizeof(WinState) / 4); i += kThreads) z[i] = 0u;
            if (threadIdx.x == 0) P.out[D.out_off + 14] = 0u;                          // "slot not written"
        }
    }
    if (!group_sync(P, G)) return false;
    // ---- W1: keyframe rows ---------------------------------------------------------------------------------------
    for (int k = G.cta; k < D.K; k += G.ncta) {
        w1_build_row(P, D, ws, k, tab, reinterpret_cast<unsigned*>(keytab), S);
        __syncthreads();
    }
    if (!group_sync(P, G)) return false;
    // ---- W2..W4: outside rows + round 1 ----------------------------------------------------------------------------
    for (int t = G.cta; t < tiles; t += G.ncta) w2_count_outside(P, D, ws, t, S);
    if (!group_sync(P, G)) return false;
    if (G.cta == 0) w3_scan_outside(P, D, S);
    if (!group_sync(P, G)) return false;
    if (ws.error) return true;                // view failed validation: the slot stays unwritten (host keeps every point)
    for (int t = G.cta; t < tiles; t += G.ncta) w4_fill_and_round1(P, D, ws, t, S);
    if (!group_sync(P, G)) return false;
    if (w == P.gwin[P.grp[0].wbeg] && G.cta == 0 && threadIdx.x == 0) P.ctrl->t_build = globaltimer_ns();

    // ---- phase machine (every CTA of the group takes the same decisions from the same counters) ----------------------
    const int drop_mode = (P.max_drop_rounds > 0) ? MODE_D1 : MODE_EVAL;
    int rounds = 1, greedy_steps = 0, drop_rounds = 0, status = 0;
    int seq = 0, mode;
    bool from_csr = true;
    unsigned unc_final = 0, slack_final = 0;
    auto after_prop = [&](unsigned changed, unsigned nfree) {
        if (changed > 0 && rounds < P.max_rounds) return (int)MODE_PROP;
        if (nfree == 0) return drop_mode;
        if (rounds >= P.max_rounds) { status = -5; return (int)MODE_FORCE; }
        return (int)MODE_GREEDY;
    };
    mode = after_prop(ws.rc[0].changed, ws.rc[0].nfree);
    while (mode != MODE_DONE) {
        ++seq;
        RoundCnt& rc = ws.rc[seq % 3];
        if (G.cta == 0 && threadIdx.x < (int)(sizeof(RoundCnt) / 4))
            reinterpret_cast<uint32_t*>(&ws.rc[(seq + 1) % 3])[threadIdx.x] = 0u;     // used by phase seq + 1
        // row phase
        if (mode != MODE_FORCE && mode != MODE_EVALV) {
            for (int r = G.cta; r < rows; r += G.ncta) {
                const int R = D.row_base + r;
                switch (mode) {
                case MODE_PROP: row_prop(P, D, rc, R, from_csr, tab, S); break;
                case MODE_GREEDY: row_greedy(P, D, R, keytab, S); break;
                case MODE_D1: row_d1_eval(P, D, rc, R, true, tab, S); break;
                case MODE_D2: row_d2(P, D, R, tab, keytab, S); break;
                case MODE_EVAL: row_d1_eval(P, D, rc, R, false, tab, S); break;
                default: break;
                }
                __syncthreads();
            }
            if (!group_sync(P, G)) return false;
        }
        if (mode == MODE_PROP) from_csr = false;
        // variable phase
        for (int t = G.cta; t < tiles; t += G.ncta) var_phase(P, D, ws, rc, mode, greedy_steps, t, S);
        if (!group_sync(P, G)) return false;
        // transition
        switch (mode) {
        case MODE_PROP: ++rounds; mode = after_prop(rc.changed, rc.nfree); break;
        case MODE_GREEDY: ++greedy_steps; ++rounds; mode = MODE_PROP; break;
        case MODE_FORCE: mode = drop_mode; break;
        case MODE_D1:
            ++rounds;
            if (rc.ncand == 0) { unc_final = rc.uncovered; slack_final = rc.slack; mode = MODE_EVALV; }
            else mode = MODE_D2;
            break;
        case MODE_D2: ++drop_rounds; mode = (drop_rounds >= P.max_drop_rounds) ? MODE_EVAL : MODE_D1; break;
        case MODE_EVAL: unc_final = rc.uncovered; slack_final = rc.slack;      // fallthrough
        case MODE_EVALV: {
            if (G.cta == 0 && threadIdx.x == 0) {
                uint32_t* hdr = P.out + D.out_off;
                hdr[0] = (uint32_t)status;
                hdr[1] = (uint32_t)rounds;
                hdr[2] = (uint32_t)ws.n_max;
                hdr[3] = (uint32_t)ws.n_vars;
                hdr[4] = (uint32_t)ws.n_cells;
                hdr[5] = (uint32_t)ws.nnz;
                hdr[6] = rc.nkept;
                hdr[7] = unc_final;
                hdr[8] = slack_final;
                hdr[9] = (uint32_t)(rc.sumcost & 0xFFFFFFFFull);
                hdr[10] = (uint32_t)(rc.sumcost >> 32);
                hdr[11] = ws.error;
                hdr[12] = (uint32_t)greedy_steps;
                hdr[13] = (uint32_t)drop_rounds;
                hdr[14] = 0x4D535331u;      // "MSS1": slot written
                hdr[15] = (uint32_t)w;
            }
            mode = MODE_DONE;
        } break;
        default: mode = MODE_DONE; break;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// the persistent cooperative kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) mss_persistent_kernel(const Params P) {
    __shared__ unsigned tab[kCells];
    __shared__ unsigned long long keytab[kCells];
    __shared__ BlockScratch S;
    __shared__ int s_abort;

    if (blockIdx.x == 0 && threadIdx.x == 0) P.ctrl->t_start = globaltimer_ns();
    const int gi = P.cta_grp[blockIdx.x];
    const GroupDesc gd = P.grp[gi];
    GroupCtx G;
    G.bar = P.gbar + (size_t)gi * 32;
    G.gen = 0u;
    G.ncta = gd.ncta;
    G.cta = (int)blockIdx.x - gd.cta0;
    G.s_abort = &s_abort;
    G.t0 = globaltimer_ns();
    if (threadIdx.x == 0) s_abort = 0;
    __syncthreads();
    for (int wi = gd.wbeg; wi < gd.wend; ++wi) {
        if (!solve_window(P, G, P.gwin[wi], tab, keytab, S)) break;
    }
    if (G.cta == 0 && threadIdx.x == 0) atomicMax(&P.ctrl->t_end, globaltimer_ns());
}

}  // namespace mss

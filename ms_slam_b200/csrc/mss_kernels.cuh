// mss_kernels.cuh -- device side of the sparsification engine (sm_100a), kernel generation 2.
//
// One persistent cooperative kernel solves a whole batch of independent windows.  CTAs are partitioned into GROUPS, one
// group per window (or a queue of windows per group when there are more windows than CTAs); a group synchronises with
// its own release/acquire barrier in global memory, so windows never wait for each other (no grid-wide sync anywhere).
//
// What it replaces in the reference (/root/reference/src/MapSparsification.cc):
//   :66-76   nMaxObservation scan            -> W1 (per keyframe row: block max -> atomicMax)
//   :78-123  variables + cell rows + KF rows -> W1 writes the keyframe rows as a CSR of packed (map point, cell) entries
//                                               (slot order, coalesced); a cell row is the set of entries of one keyframe
//                                               row with the same cell id and is never materialised
//   :125-151 outside-keyframe rows           -> W2 count / W3 scan + rhs / W4 fill (CSR of the outside rows)
//   :153-157 GUROBI optimize()               -> per-window phase machine PROP / GREEDY / DROP on F(x)
//   :159-166 read-out of GRB_DoubleAttr_X    -> EVAL: ballot-packed keep bits + row coverage + F(x)
//
// Selection algorithm, per variable state FREE / IN / OUT (oracle/emulate.py restates it on the CPU, bit for bit):
//   PROP   exact dominance to a fixed point: ub_p <= 0 -> OUT, lb_p >= 0 -> IN
//   GREEDY when PROP stops, or stalls (a round decides fewer than 1/stall_den of the undecided points): a FREE point is
//          taken iff it has positive gain and is the best candidate of every uncovered cell it lies in and within the
//          top-deficit candidates of every deficient row it lies in, or a deficient row nominates it
//   DROP   budgeted reverse delete; per-cell / per-row budgets make the summed deltas exact
// Every decision is made from integer counters (integer atomics) and keys with a unique tie-break, so the result does not
// depend on scheduling, on the group partition or on the order of entries inside a list.
//
// Work per round is proportional to the UNDECIDED part of the window: round 1 is fused into the build (W1/W4), round 2
// streams the CSR once, and from then on every row keeps a compacted "live list" of its FREE entries (cell field
// rewritten to kCellCov once the cell is covered) plus a running row coverage, so later rounds touch only those.  The
// reverse delete sweeps the CSR once, leaves each row's IN list in the (then dead) live-list segment and reads only
// those afterwards.
//
// Views come in two layouts (include/mss.h mss_layout): SoA arrays, or the packed transport form (u32 slots, u16 nObs,
// outside observations as a flat pair list) for which W2 / W4 are flat passes without shared staging.  Host views may
// still be in flight when the kernel starts: a group waits for the ready flag of the window it draws (Params::ready).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mss {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCells = 64 * 48;
constexpr int kCellsPerThread = kCells / kThreads;     // 12
constexpr int kVarTile = 256;          // var_base alignment; one var-pass tile belongs to exactly one window
constexpr int kHdrWords = 32;          // result-slot header (16 words of counters + 16 of the dual bound)
constexpr unsigned kCellNone = 0xFFFFu;     // view: slot whose keypoint is not in the grid
constexpr unsigned kCellCov = 0xFFFu;       // packed entry: "no cell / cell already covered"
constexpr int kCellBits = 12;
constexpr int kMaxWindowMps = 1 << 20;      // packed entry = (map point << 12) | cell
constexpr int kMaxWindowRows = 65535;       // 16-bit per-variable row counters
constexpr unsigned kEntInvalid = 0xFFFFFFFFu;
constexpr unsigned kTabCov = 0x80000000u;   // cell table: bit 31 = cell has an IN point, low 16 bits = FREE count
constexpr int kEpt = 8;                     // entries per thread held in registers (rows up to 2048 entries)
constexpr int kRegRow = kThreads * kEpt;
constexpr int kWarpRow = 32;                // lists up to this long are handled by one warp, cells matched with match.any
constexpr int kWarpTabRow = 256;            // lists up to this long are handled by one warp with its own nibble cell table
constexpr int kWarpTabWords = kCells / 8;   // 4 bits per cell: 1536 B per warp, the eight tables fill the CTA's cell table
constexpr int kWarpEpt = kWarpTabRow / 32;  // entries per lane
constexpr int kTraceCap = 256;
// shared-memory tail solver: capacity of the residual problem one CTA finishes on its own
constexpr int kTailEnts = 2048;
constexpr int kTailVars = 512;
constexpr int kTailRows = 512;
constexpr int kTailRowMax = 256;            // longest keyframe-row list the tail accepts (cell matching is O(n^2 / 32))

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
    const unsigned* mp_tie;   // optional tie-break rank per map point (mss_window_view::mp_tie); nullptr = the table index
    int K, H, M, F, O;
    int row_base;    // first row of this window in the per-row arrays: K keyframe rows, then H outside rows
    int slot_base;   // first entry of this window's keyframe rows in ent[] / live[]
    int obs_base;    // first entry of this window's outside rows in ent[] / live[] (after all keyframe segments)
    int var_base;    // global variable index of map point 0 (multiple of kVarTile)
    int out_off;     // u32 word offset of this window's result slot
    int packed;      // 1 = MSS_LAYOUT_PACKED: feat_mp holds u32 (map point << 12 | cell) slots, mp_nobs is a u16 array, mp_obs_kf holds
                     // the outside observations as u32 pairs (map point << 12 | outside keyframe), mp_obs_ptr is unused
                     // 2 = MSS_LAYOUT_PACKED16: as 1, but feat_mp holds u16 tokens (delta of the map-point index << 12 | cell)
    int n_max_floor; // nMax is at least this (a component of a larger window keeps the window-wide nMax)
    int nobs8;       // packed layouts: mp_nobs is a u8 array (mss_window_view::nobs8)
};

// per-phase counters; three copies rotate so that a copy is zeroed two phases before it is used again
struct RoundCnt {
    unsigned changed, nfree, sumdeg, ncand;
    unsigned uncovered, slack, nkept, rows_live;
    unsigned long long sumcost;
    unsigned maxlive;        // longest live list with cells written by this PROP row phase
    unsigned pad_;
};

struct WinState {
    int n_max, n_vars, n_cells, nnz;
    unsigned error;
    int t_rounds, t_greedy, t_status;      // written by the tail CTA for the rest of the group
    RoundCnt rc[3];
    // dual bound (mss_bound.cuh): cost of the points taken in the snapshot, cell duals, row multipliers (fixed point 2^-10),
    // deficit beyond the undecided points, residual cells left uncovered by the final selection
    unsigned long long b_cost_in, b_zsum, b_drows;
    unsigned b_s0, b_res_unc;
};

struct GroupDesc {
    int cta0, ncta;      // CTAs [cta0, cta0 + ncta) form the group; groups pull windows from one queue (gwin order)
    int pad_[2];
};

struct Ctrl {
    unsigned long long t_start, t_build, t_end;
    int abort;           // set by the watchdog (one barrier / ready-flag wait longer than watchdog_ns) or by the host (failed copy)
    unsigned queue;      // next position of gwin[] to hand out
    unsigned long long row_entries;   // list / CSR entries read by the row phases of this launch (work accounting)
    unsigned long long var_visits;    // map points visited by the variable phases of this launch
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
    uint8_t* seen;           // [Mpad] map point sits in a valid slot outside the grid (counts for nMax only)
    int* vlist;              // [2][Mpad] compact lists of the FREE map points of every window (ping-pong)
    uint2* trace;            // optional [nwin][kTraceCap] (phase, ns since the window started); nullptr = off
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
    int nwin, ngroups, Mpad;
    int Ftot;                // outside-row segments start at ent + Ftot
    int N;
    int max_rounds, all_rule_steps, max_drop_rounds;
    int stall_den;               // a PROP round that decides fewer than nfree / stall_den points is followed by GREEDY (0 = off)
    double lam, glam;
    unsigned long long watchdog_ns;
    int tail_vars, tail_ents;    // residual size handed to the shared-memory tail (<= kTailVars / kTailEnts; 0 = never)
    // dual bound (optional, all nullptr = off; mss_bound.cuh)
    uint32_t* b_snap;            // [Ftot + Otot]
    int* b_snap_n;               // [Rtot]
    int* b_snap_d;               // [Rtot]
    unsigned* b_share;           // [Mpad]
    unsigned* b_red;             // [Mpad]
    int w1_tma;                  // PACKED16: stage the token rows with bulk copies (1 = default, MSS_W1_TMA=0 reads them through L1)
    int pad3_;
    const unsigned* ready;       // optional [nwin] (queue order): set to 1 by a stream-ordered host->device copy once the views
                                 // of that window have landed in the staging buffer; nullptr = everything is resident
};

#ifndef MSS_KERNELS_TYPES_ONLY      // (translation units that only need the descriptors define this: mss_mirror.cu)
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

// Loads of VIEW data (the caller's arrays, possibly staged from the host while this kernel already runs: Params::ready).
// Plain ld.global, not ld.global.nc: the non-coherent path requires data that is read-only for the whole kernel, which a
// window still in flight is not; the ready flag's acquire + the group barrier order these loads after the copy.
template <class T>
__device__ __forceinline__ T ldv(const T* p) { return *p; }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned f32_orderable(float g) {
    unsigned b = __float_as_uint(g);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
// larger key = better: higher gain first, then lower (window-local) map-point index
__device__ __forceinline__ unsigned long long make_key(float g, unsigned local_idx) {
    return ((unsigned long long)f32_orderable(g) << 32) | (unsigned long long)(0xFFFFFFFFu - local_idx);
}

// ---- bulk copies by the TMA unit + mbarrier (sm_90+; sm_100a here).  One elected thread arms the barrier with the byte count
//      and issues the copy; every consumer thread waits on the barrier's phase parity.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// tie-break key of map point v of window D: the caller's rank, or the table index
__device__ __forceinline__ unsigned tie_of(const WinDesc& D, unsigned v) { return D.mp_tie ? ldv(D.mp_tie + v) : v; }

struct BlockScratch {
    int red[3][kWarps];
    int scan[kWarps];
    int bcast[4];
    unsigned hist[256];
    // row queue of the current chunk of this CTA's rows: long rows from the front, short rows from the back
    unsigned short rowq[kThreads];
    int rown[kThreads];
    int rowoff[kThreads];
    int pred[2][3][kWarps];  // parity-buffered partial sums / warp counts of the row phases (saves the protective barriers)
    int pscan[2][kWarps];
    int qn[2];
    unsigned rows_live, maxlive;
    unsigned long long work;   // entries read by this CTA in the current row phase
    int tcnt[4];             // tail: changed, nfree
};

struct GroupCtx {
    unsigned* bar;
    unsigned gen;
    int ncta;
    int cta;        // index of this CTA inside the group
    int* s_abort;   // shared flag
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
            unsigned long long t_wait = 0ull;          // taken at the first check: the limit is per wait, not per launch
            while (ld_acquire_u32(G.bar) < target) {
                if (spins > 64u) __nanosleep(64);
                if ((++spins & 0xFFu) == 0u) {
                    if (*(volatile int*)&P.ctrl->abort) { ab = 1; break; }
                    if (t_wait == 0ull) t_wait = globaltimer_ns();
                    else if (globaltimer_ns() - t_wait > P.watchdog_ns) {
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
__device__ unsigned long long block_kth_largest(BlockScratch& S, int rank, Emit emit, int low_shift = 0) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= low_shift; shift -= 8) {          // (keys whose bits below low_shift are all zero)
        __syncthreads();
        S.hist[threadIdx.x] = 0;            // kThreads == 256 bins
        __syncthreads();
        emit([&](unsigned long long key) {
            if ((key & mask) == prefix) atomicAdd(&S.hist[(unsigned)(key >> shift) & 255u], 1u);
        });
        __syncthreads();
        {
            // thread t owns bin 255 - t: the exclusive prefix over threads is the number of keys in larger bins
            const int h = (int)S.hist[255 - (int)threadIdx.x];
            int total;
            const int above = block_excl_scan(S, h, total);
            if (above <= rank && rank < above + h) { S.bcast[0] = 255 - (int)threadIdx.x; S.bcast[1] = rank - above; }
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

// View accessors: the SoA layout carries i32 / u16 arrays, the packed transport layout u32 slots and u16 tables
// (include/mss.h, mss_layout); the branch is uniform per window.
__device__ __forceinline__ int ld_nobs(const WinDesc& D, int mp) {
    if (!D.packed) return ldv(D.mp_nobs + mp);
    return D.nobs8 ? (int)ldv(reinterpret_cast<const uint8_t*>(D.mp_nobs) + mp) : (int)ldv(reinterpret_cast<const uint16_t*>(D.mp_nobs) + mp);
}
__device__ __forceinline__ int ld_obs_kf(const WinDesc& D, int o) { return ldv(D.mp_obs_kf + o); }      // SoA layout only
// MSS_LAYOUT_PACKED16: the slots of a keyframe, sorted by map-point index, as 16-bit tokens.  d = t >> 12, low = t & 0xFFF:
//   d < 15  a slot: map point = previous map point + d (0 at the start of the keyframe), cell = low (0xFFF = not in mGrid)
//   d == 15 no slot: the running map-point index advances by 15 * (low + 1)
// tok_decode returns the advance of the running index; slot / cell describe the token
__device__ __forceinline__ int tok_decode(unsigned t, bool& slot, unsigned& cell) {
    const unsigned d = t >> 12, low = t & 0xFFFu;
    slot = d < 15u;
    cell = low == 0xFFFu ? kCellNone : low;
    return slot ? (int)d : 15 * (int)(low + 1u);
}

// slot i of the view -> (map point or -1, cell or kCellNone)  [SoA and MSS_LAYOUT_PACKED; tokens are decoded by their readers]
__device__ __forceinline__ void ld_slot(const WinDesc& D, int i, int& mp, unsigned& c) {
    if (D.packed) {
        const uint32_t s = ldv(reinterpret_cast<const uint32_t*>(D.feat_mp) + i);
        if (s == kEntInvalid) { mp = -1; c = kCellNone; return; }
        mp = (int)(s >> kCellBits);
        c = s & kCellCov;
        if (c == kCellCov) c = kCellNone;
    } else {
        mp = ldv(D.feat_mp + i);
        c = ldv(D.feat_cell + i);
    }
}

// A row's entries held in registers: warp w owns a contiguous chunk of the list, lane-strided inside it, so global
// accesses are coalesced and (warp, b, lane) order is list order (needed for the stable compaction).
template <int EPT>
struct RowRegs {
    uint32_t e[EPT];
    uint8_t s[EPT];
    int per, nb;
};

// EPT = entries per thread: 1, 2, 4 or 8, the smallest that covers the list (short lists must not pay for eight slots)
template <int EPT>
__device__ __forceinline__ void load_row(RowRegs<EPT>& X, const uint32_t* __restrict__ src, int n, const uint8_t* __restrict__ st_w) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    X.per = (((n + kWarps - 1) / kWarps) + 31) & ~31;
    X.nb = X.per >> 5;
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        const int idx = wid * X.per + b * 32 + lane;
        X.e[b] = (b < X.nb && idx < n) ? src[idx] : kEntInvalid;
    }
#pragma unroll
    for (int b = 0; b < EPT; ++b) X.s[b] = (X.e[b] != kEntInvalid) ? st_w[X.e[b] >> kCellBits] : (uint8_t)ST_NOTVAR;
}

// ---------------------------------------------------------------------------------------------------------------
// W1: keyframe row -> CSR segment (coalesced, slot order), cell statistics, round-1 contributions.
// One spread-address memory operation per entry (the 64-bit reduction on the map point's counters); everything else is
// coalesced or shared memory.  A map point becomes a variable exactly when its counters are non-zero after W1, so W2 can
// derive the state array with coalesced stores; valid slots outside the grid only leave a "seen" mark for nMax.
// ---------------------------------------------------------------------------------------------------------------
// stok: this row's tokens staged in shared memory by a bulk copy (token 0 of the row), or nullptr: read them from global memory
__device__ void w1_build_row(const Params& P, const WinDesc& D, WinState& ws, int k, int par, unsigned* tab, BlockScratch& S,
                             const uint16_t* stok = nullptr) {
    const int R = D.row_base + k;
    const int beg = ldv(D.feat_ptr + k), end = ldv(D.feat_ptr + k + 1);
    if (beg < 0 || end < beg || end > D.F) {
        if (threadIdx.x == 0) {
            atomicOr(&ws.error, ERR_PTR);
            P.ent_n[R] = 0; P.live_n[R] = 0; P.row_need[R] = P.N; P.row_cov[R] = 0; P.row_ncell[R] = 0; P.row_off[R] = 0;
        }
        return;
    }
    const int nslots = end - beg;
    const int seg = D.slot_base + beg;
    uint32_t* ent = P.ent + seg;
    unsigned long long* acc_w = P.acc + D.var_base;
    uint8_t* seen_w = P.seen + D.var_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool defi = P.N > 0;
    const bool tok = D.packed == 2;
    unsigned err = 0;
    if (nslots <= kRegRow) {
        // slots in registers: warp w owns a contiguous chunk, lane-strided (coalesced loads, list order = slot order)
        const int per = (((nslots + kWarps - 1) / kWarps) + 31) & ~31;
        const int nb = per >> 5;
        uint32_t e[kEpt];
        int pre[kEpt];                  // tokens: inclusive prefix of the index advances inside this warp's chunk
        unsigned offgrid = 0u;          // tokens: bit b = my token of chunk b is a valid slot outside the grid
        int chunk_total = 0;
        if (tok) {
            const uint16_t* tk = reinterpret_cast<const uint16_t*>(D.feat_mp) + beg;
            int run = 0;
#pragma unroll
            for (int b = 0; b < kEpt; ++b) {
                const int idx = wid * per + b * 32 + lane;
                int adv = 0;
                bool slot = false;
                unsigned c = kCellNone;
                if (b < nb && idx < nslots) adv = tok_decode(stok ? (unsigned)stok[idx] : (unsigned)ldv(tk + idx), slot, c);
                int x = adv;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                    if (lane >= o) x += y;
                }
                pre[b] = run + x;
                run += __shfl_sync(0xFFFFFFFFu, x, 31);
                e[b] = kEntInvalid;                                             // provisionally: the cell alone (the map point
                if (slot) {                                                     // index needs the other warps' totals)
                    if (c == kCellNone) offgrid |= 1u << b;
                    else if (c >= (unsigned)kCells) err |= ERR_INDEX;
                    else e[b] = c;
                }
            }
            chunk_total = run;
        } else {
#pragma unroll
            for (int b = 0; b < kEpt; ++b) {
                const int idx = wid * per + b * 32 + lane;
                int mp = -1;
                unsigned c = kCellNone;
                if (b < nb && idx < nslots) ld_slot(D, beg + idx, mp, c);
                e[b] = kEntInvalid;
                pre[b] = 0;
                if (mp < -1 || mp >= D.M) err |= ERR_INDEX;
                else if (mp >= 0) {
                    if (c == kCellNone) seen_w[mp] = 1;                         // valid slot, not in mGrid (MapSparsification.cc:69-75)
                    else if (c >= (unsigned)kCells) err |= ERR_INDEX;
                    else e[b] = ((uint32_t)mp << kCellBits) | c;
                }
            }
        }
#pragma unroll
        for (int b = 0; b < kEpt; ++b) if (e[b] != kEntInvalid) tab[e[b] & kCellCov] = 0u;
        __syncthreads();
        int nz = 0, ncell = 0;
        unsigned m[kEpt];
#pragma unroll
        for (int b = 0; b < kEpt; ++b) {
            const bool v = e[b] != kEntInvalid;
            if (v && atomicAdd(&tab[e[b] & kCellCov], 1u) == 0u) ++ncell;
            m[b] = __ballot_sync(0xFFFFFFFFu, v);
        }
        int wcnt = 0;
#pragma unroll
        for (int b = 0; b < kEpt; ++b) wcnt += __popc(m[b]);
        ncell = __reduce_add_sync(0xFFFFFFFFu, ncell);
        // one exchange for the row totals and the warps' write offsets; the scratch is double-buffered by row parity, so
        // the only other barriers of a row are the one after the lazy zeroing and the caller's
        if (lane == 0) { S.pred[par][0][wid] = wcnt; S.pred[par][1][wid] = ncell; S.pred[par][2][wid] = chunk_total; }
        __syncthreads();                                                // publishes the cell table and the partial sums
        int pos = 0, mp_base = 0;
        ncell = 0;
#pragma unroll
        for (int q = 0; q < kWarps; ++q) {
            if (q < wid) { pos += S.pred[par][0][q]; mp_base += S.pred[par][2][q]; }
            nz += S.pred[par][0][q]; ncell += S.pred[par][1][q];
        }
        if (tok) {                                                      // the map-point indices are known now
#pragma unroll
            for (int b = 0; b < kEpt; ++b) {
                const unsigned mp = (unsigned)(mp_base + pre[b]);
                if (e[b] != kEntInvalid) {
                    if (mp >= (unsigned)D.M) err |= ERR_INDEX;          // (the window is rejected; index 0 keeps W1 in bounds)
                    e[b] = ((mp < (unsigned)D.M ? mp : 0u) << kCellBits) | e[b];
                } else if ((offgrid >> b) & 1u) {
                    if (mp < (unsigned)D.M) seen_w[mp] = 1; else err |= ERR_INDEX;
                }
            }
        }
        const bool critr = defi && P.N >= nz;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int b = 0; b < kEpt; ++b) {
            if (e[b] != kEntInvalid) {
                const unsigned t = tab[e[b] & kCellCov];
                if (t > 0xFFFFu) err |= ERR_CELL_OVERFLOW;
                // round 1 (everything FREE, nothing IN): every cell is uncovered, every row with N > 0 is deficient
                unsigned long long add = 1ull;
                if (t == 1u) add |= 1ull << 16;
                if (defi) add |= 1ull << 32;
                if (critr) add |= 1ull << 48;
                atomicAdd(&acc_w[e[b] >> kCellBits], add);
                ent[pos + __popc(m[b] & lt)] = e[b];
            }
            pos += __popc(m[b]);
        }
        if (threadIdx.x == 0) {
            P.ent_n[R] = nz; P.live_n[R] = 0; P.row_need[R] = P.N; P.row_cov[R] = 0; P.row_ncell[R] = ncell; P.row_off[R] = seg;
            if (nz) atomicAdd(&ws.nnz, nz);
            if (ncell) atomicAdd(&ws.n_cells, ncell);
        }
    } else {
        // long keyframe: same thing in two passes over global memory
        zero_tab(tab);
        __syncthreads();
        int nz = 0, ncell = 0, z = 0;
        const uint16_t* tk = reinterpret_cast<const uint16_t*>(D.feat_mp) + beg;
        for (int i = threadIdx.x; i < nslots; i += kThreads) {
            int mp = 0;
            unsigned c;
            if (tok) {                                                  // pass 1 needs the cells only
                bool slot;
                tok_decode(ldv(tk + i), slot, c);
                if (!slot || c == kCellNone) continue;
            } else {
                ld_slot(D, beg + i, mp, c);
                if (mp < 0) { if (mp < -1) err |= ERR_INDEX; continue; }
                if (mp >= D.M) { err |= ERR_INDEX; continue; }
                if (c == kCellNone) { seen_w[mp] = 1; continue; }
            }
            if (c >= (unsigned)kCells) { err |= ERR_INDEX; continue; }
            ++nz;
            if (atomicAdd(&tab[c], 1u) == 0u) ++ncell;
        }
        block_sum3(S, nz, ncell, z);
        const bool critr = defi && P.N >= nz;
        int out_base = 0, mp_carry = 0;
        for (int base = 0; base < nslots; base += kThreads) {
            const int i = base + (int)threadIdx.x;
            uint32_t e = kEntInvalid;
            if (tok) {
                int adv = 0;
                bool slot = false;
                unsigned c = kCellNone;
                if (i < nslots) adv = tok_decode(ldv(tk + i), slot, c);
                int tot;
                const int mp = mp_carry + block_excl_scan(S, adv, tot) + adv;       // running map-point index, token order
                mp_carry += tot;
                if (slot) {
                    if ((unsigned)mp >= (unsigned)D.M) err |= ERR_INDEX;       // (a wrapped running index is negative)
                    else if (c == kCellNone) seen_w[mp] = 1;
                    else if (c < (unsigned)kCells) e = ((uint32_t)mp << kCellBits) | c;
                }
            } else if (i < nslots) {
                int mp;
                unsigned c;
                ld_slot(D, beg + i, mp, c);
                if (mp >= 0 && mp < D.M && c < (unsigned)kCells) e = ((uint32_t)mp << kCellBits) | c;
            }
            int total;
            const int p = block_excl_scan(S, e != kEntInvalid ? 1 : 0, total);
            if (e != kEntInvalid) {
                const unsigned t = tab[e & kCellCov];
                if (t > 0xFFFFu) err |= ERR_CELL_OVERFLOW;
                unsigned long long add = 1ull;
                if (t == 1u) add |= 1ull << 16;
                if (defi) add |= 1ull << 32;
                if (critr) add |= 1ull << 48;
                atomicAdd(&acc_w[e >> kCellBits], add);
                ent[out_base + p] = e;
            }
            out_base += total;
        }
        if (threadIdx.x == 0) {
            P.ent_n[R] = nz; P.live_n[R] = 0; P.row_need[R] = P.N; P.row_cov[R] = 0; P.row_ncell[R] = ncell; P.row_off[R] = seg;
            if (nz) atomicAdd(&ws.nnz, nz);
            if (ncell) atomicAdd(&ws.n_cells, ncell);
        }
    }
    if (err) atomicOr(&ws.error, err);
}

// Observations of the map points of one super-tile (kSuper consecutive map points, kVpt per thread), flattened: they
// are one contiguous range of mp_obs_kf, read coalesced; the owner of observation o is found by binary search in the
// super-tile's pointers (shared memory).  Several map points per thread keep that many independent loads in flight.
constexpr int kVpt = 4;
constexpr int kSuper = kVarTile * kVpt;
struct ObsTile {
    int ptr[kSuper + 1];
    unsigned add[kSuper];        // W4: [outside rows that are deficient : 16 | ... that need every candidate : 16]
    unsigned nout[kSuper];       // W4: outside observations
    uint8_t isvar[kSuper];
};

__device__ __forceinline__ int obs_owner(const ObsTile& O, int o) {
    int lo = 0, hi = kSuper;                // ptr[lo] <= o < ptr[hi]
#pragma unroll
    for (int step = 0; step < 10; ++step) {
        const int mid = (lo + hi) >> 1;
        if (O.ptr[mid] <= o) lo = mid; else hi = mid;
    }
    return lo;
}

// loads the super-tile's pointers; returns false (and flags the window) when they are not a valid CSR
__device__ __forceinline__ bool obs_tile_load(const WinDesc& D, WinState& ws, int base, ObsTile& O) {
    bool bad = false;
#pragma unroll
    for (int j = 0; j < kVpt; ++j) O.ptr[j * kVarTile + threadIdx.x] = ldv(D.mp_obs_ptr + min(base + j * kVarTile + (int)threadIdx.x, D.M));
    if (threadIdx.x == 0) O.ptr[kSuper] = ldv(D.mp_obs_ptr + min(base + kSuper, D.M));
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kVpt; ++j) {
        const int p = O.ptr[j * kVarTile + threadIdx.x], q = O.ptr[j * kVarTile + threadIdx.x + 1];
        bad |= p < 0 || p > D.O || q < p || q > D.O;
    }
    if (__syncthreads_or(bad)) {
        if (threadIdx.x == 0) atomicOr(&ws.error, ERR_PTR);
        return false;
    }
    return true;
}

// W2: per super-tile of map points: state array from the W1 counters, nMax (MapSparsification.cc:66-76), variable count,
// and the number of variables every outside keyframe observes (MapSparsification.cc:127-142)
__device__ void w2_vars_and_outside_counts(const Params& P, const WinDesc& D, WinState& ws, int stile, ObsTile& O, BlockScratch& S) {
    const int base = stile * kSuper;
    bool isvar[kVpt];
    int nmax = 0, nv = 0;
    unsigned long long a[kVpt];
    uint8_t sn[kVpt];
#pragma unroll
    for (int j = 0; j < kVpt; ++j) {
        const int mp = base + j * kVarTile + (int)threadIdx.x;
        a[j] = mp < D.M ? P.acc[D.var_base + mp] : 0ull;
        sn[j] = mp < D.M ? P.seen[D.var_base + mp] : (uint8_t)0;
    }
#pragma unroll
    for (int j = 0; j < kVpt; ++j) {
        const int mp = base + j * kVarTile + (int)threadIdx.x;
        isvar[j] = a[j] != 0ull;
        if (mp < D.M || mp < ((D.M + kVarTile - 1) / kVarTile) * kVarTile) P.st[D.var_base + mp] = isvar[j] ? (uint8_t)ST_FREE : (uint8_t)ST_NOTVAR;
        if (isvar[j] || sn[j]) nmax = max(nmax, ld_nobs(D, mp));
        nv += isvar[j] ? 1 : 0;
    }
    nmax = block_max(S, nmax);
    int z0 = 0, z1 = 0;
    block_sum3(S, nv, z0, z1);
    if (threadIdx.x == 0) {
        if (nmax > 0) atomicMax(&ws.n_max, nmax);
        if (nv) atomicAdd(&ws.n_vars, nv);
    }
    if (D.H == 0 || nv == 0 || D.packed) return;           // packed layout: the pair list is walked by w2_pairs
#pragma unroll
    for (int j = 0; j < kVpt; ++j) O.isvar[j * kVarTile + threadIdx.x] = isvar[j] ? 1 : 0;
    if (!obs_tile_load(D, ws, base, O)) return;
    const int p0 = O.ptr[0], p1 = O.ptr[kSuper];
    unsigned err = 0;
    for (int o0 = p0; o0 < p1; o0 += kThreads * kVpt) {
        int kf[kVpt];
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int o = o0 + j * kThreads + (int)threadIdx.x;
            kf[j] = o < p1 ? ld_obs_kf(D, o) : 0;
        }
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int o = o0 + j * kThreads + (int)threadIdx.x;
            if (o >= p1 || kf[j] < D.K) { if (kf[j] < 0) err |= ERR_INDEX; continue; }
            if (kf[j] >= D.K + D.H) { err |= ERR_INDEX; continue; }
            if (O.isvar[obs_owner(O, o)]) atomicAdd(&P.ent_n[D.row_base + kf[j]], 1);
        }
    }
    if (err) atomicOr(&ws.error, err);
    __syncthreads();
}

// Packed layout: the outside observations come as one flat list of (map point, outside keyframe) pairs, so no owner
// search is needed.  W2 part: how many variables every outside keyframe observes (MapSparsification.cc:127-142).  A map
// point is a variable exactly when its W1 counters are non-zero.
__device__ void w2_pairs(const Params& P, const WinDesc& D, WinState& ws, int gt, int gsz) {
    const uint32_t* pairs = reinterpret_cast<const uint32_t*>(D.mp_obs_kf);
    unsigned err = 0;
    for (int o = gt; o < D.O; o += gsz) {
        const uint32_t pr = ldv(pairs + o);
        const int mp = (int)(pr >> kCellBits), j = (int)(pr & kCellCov);
        if (mp >= D.M || j >= D.H) { err |= ERR_INDEX; continue; }
        if (P.acc[D.var_base + mp] != 0ull) atomicAdd(&P.ent_n[D.row_base + D.K + j], 1);
    }
    if (err) atomicOr(&ws.error, err);
}

// W4 part: fill the outside rows and add their round-1 contributions straight to the map points' counters; deg[] (zeroed
// in W0) collects the number of outside rows of every variable
__device__ void w4_pairs(const Params& P, const WinDesc& D, int gt, int gsz) {
    const uint32_t* pairs = reinterpret_cast<const uint32_t*>(D.mp_obs_kf);
    for (int o = gt; o < D.O; o += gsz) {
        const uint32_t pr = ldv(pairs + o);
        const int mp = (int)(pr >> kCellBits), j = (int)(pr & kCellCov);
        const int g = D.var_base + mp;
        if (P.st[g] != ST_FREE) continue;                   // (indices were validated by w2_pairs)
        const int R = D.row_base + D.K + j;
        const int pos = atomicAdd(&P.ocursor[R], 1);
        P.ent[pos] = ((uint32_t)mp << kCellBits) | kCellCov;
        const int need = P.row_need[R];
        if (need > 0) atomicAdd(&P.acc[g], (1ull << 32) | (need >= P.ent_n[R] ? (1ull << 48) : 0ull));
        atomicAdd(&P.deg[g], 1u);
    }
}

__device__ __forceinline__ void prop_decide(const Params& P, unsigned long long a, int cost_i, uint8_t* st, float* gain, unsigned deg,
                                            int& changed, int& nfree, int& sumdeg);
__device__ __forceinline__ void prop_commit(RoundCnt& rc, int* vdst, int mp, int changed, int nfree, int sumdeg);
__device__ __forceinline__ void prop_commit4(RoundCnt& rc, int* vdst, const int* mp, const int* nfree);
__device__ __forceinline__ void prop_flush(RoundCnt& rc, int changed, int sumdeg);

// Packed layout: W2 and the round-1 decision of W4 as flat passes over the map points (no owner search, hence no shared
// staging and no per-tile barriers; four independent map points in flight per thread, consecutive threads on consecutive
// map points).  Same arithmetic as the tile versions.
__device__ void w2_flat(const Params& P, const WinDesc& D, WinState& ws, int gt, int gsz, BlockScratch& S) {
    const int mpad = ((max(D.M, 1) + kVarTile - 1) / kVarTile) * kVarTile;
    int nmax = 0, nv = 0;
    for (int base = gt; base < mpad; base += gsz * kVpt) {
        unsigned long long a[kVpt];
        uint8_t sn[kVpt];
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int mp = base + j * gsz;
            a[j] = mp < D.M ? P.acc[D.var_base + mp] : 0ull;
            sn[j] = mp < D.M ? P.seen[D.var_base + mp] : (uint8_t)0;
        }
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int mp = base + j * gsz;
            if (mp < mpad) P.st[D.var_base + mp] = a[j] != 0ull ? (uint8_t)ST_FREE : (uint8_t)ST_NOTVAR;
            if (a[j] != 0ull || sn[j]) nmax = max(nmax, ld_nobs(D, mp));
            nv += a[j] != 0ull ? 1 : 0;
        }
    }
    nmax = block_max(S, nmax);
    int z0 = 0, z1 = 0;
    block_sum3(S, nv, z0, z1);
    if (threadIdx.x == 0) {
        if (nmax > 0) atomicMax(&ws.n_max, nmax);
        if (nv) atomicAdd(&ws.n_vars, nv);
    }
}

__device__ void w4_flat(const Params& P, const WinDesc& D, WinState& ws, int gt, int gsz) {
    const int lim = (D.M + 31) & ~31;                       // warp-uniform bound: the commit uses full-warp ballots
    int changed = 0, sumdeg = 0;
    for (int base = gt; base < lim; base += gsz * kVpt) {
        bool isvar[kVpt];
        unsigned long long a[kVpt];
        int nobs[kVpt];
        unsigned nout[kVpt];
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int mp = base + j * gsz;
            isvar[j] = mp < D.M && P.st[D.var_base + mp] == ST_FREE;
        }
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int g = D.var_base + base + j * gsz;
            a[j] = 0ull; nobs[j] = 0; nout[j] = 0u;
            if (isvar[j]) { a[j] = P.acc[g]; nobs[j] = ld_nobs(D, base + j * gsz); if (D.H > 0) nout[j] = P.deg[g]; }
        }
        int mps[kVpt], nfr[kVpt];
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int mp = base + j * gsz;
            const int g = D.var_base + mp;
            mps[j] = mp;
            nfr[j] = 0;
            if (isvar[j]) {
                const unsigned deg = (unsigned)(a[j] & 0xFFFFu) + nout[j];
                P.deg[g] = deg;
                P.acc[g] = 0ull;
                prop_decide(P, a[j], ws.n_max - nobs[j], &P.st[g], &P.gain[g], deg, changed, nfr[j], sumdeg);
            }
        }
        prop_commit4(ws.rc[0], P.vlist + D.var_base, mps, nfr);
    }
    prop_flush(ws.rc[0], changed, sumdeg);
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
            P.row_need[R] = outside_need(v, ldv(D.okf_total + j), P.N);
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

// warp-aggregated bookkeeping of a PROP decision: counters + append of the still-FREE map points to the new list
__device__ __forceinline__ void prop_commit(RoundCnt& rc, int* vdst, int mp, int changed, int nfree, int sumdeg) {
    const int lane = threadIdx.x & 31;
    const unsigned mk = __ballot_sync(0xFFFFFFFFu, nfree != 0);
    const unsigned mc = __ballot_sync(0xFFFFFFFFu, changed != 0);
    if (mk == 0u && mc == 0u) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sumdeg += __shfl_xor_sync(0xFFFFFFFFu, sumdeg, o);
    int pos = 0;
    if (lane == 0) {
        if (mk) pos = (int)atomicAdd(&rc.nfree, (unsigned)__popc(mk));
        if (mc) atomicAdd(&rc.changed, (unsigned)__popc(mc));
        if (sumdeg) atomicAdd(&rc.sumdeg, (unsigned)sumdeg);
    }
    pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
    if (nfree) vdst[pos + __popc(mk & ((1u << lane) - 1u))] = mp;
}

// The same for kVpt decisions per thread at once: ONE reservation per warp for all of them (the counters of a window live
// in one cache line, so per-decision atomics from every warp of the group serialise in L2); changed / sumdeg are left to
// the caller, which adds them up over its whole loop and flushes them once (prop_flush).
__device__ __forceinline__ void prop_commit4(RoundCnt& rc, int* vdst, const int* mp, const int* nfree) {
    const int lane = threadIdx.x & 31;
    unsigned mk[4];
    int tot = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { mk[j] = __ballot_sync(0xFFFFFFFFu, nfree[j] != 0); tot += __popc(mk[j]); }
    if (tot == 0) return;
    int pos = 0;
    if (lane == 0) pos = (int)atomicAdd(&rc.nfree, (unsigned)tot);
    pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (nfree[j]) vdst[pos + __popc(mk[j] & lt)] = mp[j];
        pos += __popc(mk[j]);
    }
}
__device__ __forceinline__ void prop_flush(RoundCnt& rc, int changed, int sumdeg) {
    changed = __reduce_add_sync(0xFFFFFFFFu, changed);
    sumdeg = __reduce_add_sync(0xFFFFFFFFu, sumdeg);
    if ((threadIdx.x & 31) == 0) {
        if (changed) atomicAdd(&rc.changed, (unsigned)changed);
        if (sumdeg) atomicAdd(&rc.sumdeg, (unsigned)sumdeg);
    }
}

// W4: per super-tile of map points: fill the outside rows, add their round-1 contributions, take the round-1 decision
__device__ void w4_fill_and_round1(const Params& P, const WinDesc& D, WinState& ws, int stile, ObsTile& O) {
    const int base = stile * kSuper;
    bool isvar[kVpt];
    unsigned long long a[kVpt];
    int nobs[kVpt];
    bool any = false;
#pragma unroll
    for (int j = 0; j < kVpt; ++j) {
        const int mp = base + j * kVarTile + (int)threadIdx.x;
        isvar[j] = mp < D.M && P.st[D.var_base + mp] == ST_FREE;
    }
#pragma unroll
    for (int j = 0; j < kVpt; ++j) {
        const int mp = base + j * kVarTile + (int)threadIdx.x;
        a[j] = 0ull;
        nobs[j] = 0;
        if (isvar[j]) { a[j] = P.acc[D.var_base + mp]; nobs[j] = ld_nobs(D, mp); any = true; }
    }
    unsigned nout[kVpt] = {0u, 0u, 0u, 0u};
    if (D.packed) {
        if (D.H > 0) {
#pragma unroll
            for (int j = 0; j < kVpt; ++j)
                if (isvar[j]) nout[j] = P.deg[D.var_base + base + j * kVarTile + (int)threadIdx.x];     // written by w4_pairs
        }
    } else if (D.H > 0 && __syncthreads_or(any)) {
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            O.isvar[j * kVarTile + threadIdx.x] = isvar[j] ? 1 : 0;
            O.add[j * kVarTile + threadIdx.x] = 0u;
            O.nout[j * kVarTile + threadIdx.x] = 0u;
        }
        if (obs_tile_load(D, ws, base, O)) {            // (validated in W2 already; W4 only runs on windows without errors)
            const int p0 = O.ptr[0], p1 = O.ptr[kSuper];
            for (int o0 = p0; o0 < p1; o0 += kThreads * kVpt) {
                int kf[kVpt];
#pragma unroll
                for (int j = 0; j < kVpt; ++j) {
                    const int o = o0 + j * kThreads + (int)threadIdx.x;
                    kf[j] = o < p1 ? ld_obs_kf(D, o) : 0;
                }
#pragma unroll
                for (int j = 0; j < kVpt; ++j) {
                    const int o = o0 + j * kThreads + (int)threadIdx.x;
                    if (o >= p1 || kf[j] < D.K) continue;
                    const int v = obs_owner(O, o);
                    if (!O.isvar[v]) continue;
                    const int R = D.row_base + kf[j];
                    const int pos = atomicAdd(&P.ocursor[R], 1);
                    P.ent[pos] = ((uint32_t)(base + v) << kCellBits) | kCellCov;
                    const int need = P.row_need[R];
                    unsigned add = 0u;
                    if (need > 0) { add = 1u; if (need >= P.ent_n[R]) add |= 1u << 16; }
                    if (add) atomicAdd(&O.add[v], add);
                    atomicAdd(&O.nout[v], 1u);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const unsigned ad = O.add[j * kVarTile + threadIdx.x];
            a[j] += ((unsigned long long)(ad & 0xFFFFu) << 32) + ((unsigned long long)(ad >> 16) << 48);
            nout[j] = O.nout[j * kVarTile + threadIdx.x];
        }
        __syncthreads();                                 // O is reused by the next super-tile
    }
#pragma unroll
    for (int j = 0; j < kVpt; ++j) {
        const int mp = base + j * kVarTile + (int)threadIdx.x;
        const int g = D.var_base + mp;
        int changed = 0, nfree = 0, sumdeg = 0;
        if (isvar[j]) {
            const unsigned deg = (unsigned)(a[j] & 0xFFFFu) + nout[j];
            P.deg[g] = deg;
            P.acc[g] = 0ull;
            prop_decide(P, a[j], ws.n_max - nobs[j], &P.st[g], &P.gain[g], deg, changed, nfree, sumdeg);
        }
        prop_commit(ws.rc[0], P.vlist + D.var_base, mp, changed, nfree, sumdeg);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// row phases (one CTA per row at a time)
// ---------------------------------------------------------------------------------------------------------------
// PROP: counts on the current state, contributions to the FREE variables, and the row's new live list.
// Source list: the CSR (first PROP row phase of the window) or the previous live list.  Entries found IN are added to the
// running coverage exactly once (they are not copied to the new list); entries found OUT are dropped.
// PROP on a list held in registers (EPT entries per thread)
template <int EPT>
__device__ __forceinline__ void row_prop_regs(const Params& P, const WinDesc& D, int R, int n, const uint32_t* src, uint32_t* dst,
                                              int need, int cov0, int par, unsigned* tab, BlockScratch& S) {
    const uint8_t* st_w = P.st + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    // two block barriers per row: the cell table and the scratch are double-buffered by row parity (the caller alternates
    // `tab` between two tables), so neither the lazy zeroing of the next row nor its partial sums can overtake a slow warp
    RowRegs<EPT> X;
    load_row(X, src, n, st_w);
#pragma unroll
    for (int b = 0; b < EPT; ++b)
        if (X.e[b] != kEntInvalid && (X.e[b] & kCellCov) != kCellCov) tab[X.e[b] & kCellCov] = 0u;
    __syncthreads();
    int cin = 0, cfree = 0;
    unsigned m[EPT];
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        const bool fr = X.e[b] != kEntInvalid && X.s[b] == ST_FREE;
        m[b] = __ballot_sync(0xFFFFFFFFu, fr);
        if (X.e[b] == kEntInvalid) continue;
        const unsigned cell = X.e[b] & kCellCov;
        if (X.s[b] == ST_IN) { ++cin; if (cell != kCellCov) atomicOr(&tab[cell], kTabCov); }
        else if (fr && cell != kCellCov) atomicAdd(&tab[cell], 1u);
    }
    cin = __reduce_add_sync(0xFFFFFFFFu, cin);
#pragma unroll
    for (int b = 0; b < EPT; ++b) cfree += __popc(m[b]);                 // this warp's FREE entries = its share of the new list
    if (lane == 0) { S.pred[par][0][wid] = cin; S.pred[par][1][wid] = cfree; }
    __syncthreads();                                    // publishes the cell table and the partial sums
    int pos = 0;
    cin = 0; cfree = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) { if (q < wid) pos += S.pred[par][1][q]; cin += S.pred[par][0][q]; cfree += S.pred[par][1][q]; }
    const int cov = cov0 + cin;
    const int d = max(0, need - cov);
    const bool defi = d > 0, critr = defi && d >= cfree;
    const unsigned lt = (1u << lane) - 1u;
    // every entry of src is in registers (load_row precedes the first barrier), so the in-place compaction can start at once
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        if ((m[b] >> lane) & 1u) {
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
            dst[pos + __popc(m[b] & lt)] = covered ? (X.e[b] | kCellCov) : X.e[b];
        }
        pos += __popc(m[b]);
    }
    if (threadIdx.x == 0) {
        P.row_cov[R] = cov;
        P.live_n[R] = cfree;
        if (cfree) { S.rows_live += 1u; if (R - D.row_base < D.K) S.maxlive = max(S.maxlive, (unsigned)cfree); }   // thread 0 only
    }
}

__device__ void row_prop(const Params& P, const WinDesc& D, int R, int n, int off, int par, bool from_csr, unsigned* tab,
                         BlockScratch& S) {
    if (n == 0) return;                                     // live_n[R] is already 0 (W1 / W3 / previous round)
    const uint32_t* src = (from_csr ? P.ent : P.live) + off;
    uint32_t* dst = P.live + off;
    const uint8_t* st_w = P.st + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R];
    const int cov0 = P.row_cov[R];
    // entries per thread = exactly the 32-entry blocks a warp's share of the list needs (a c2 keyframe row of 1 250 .. 1 600
    // entries takes 5 .. 7, not 8: the unrolled row code has no idle iterations)
    if (n <= kThreads) row_prop_regs<1>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else if (n <= 2 * kThreads) row_prop_regs<2>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else if (n <= 3 * kThreads) row_prop_regs<3>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else if (n <= 4 * kThreads) row_prop_regs<4>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else if (n <= 5 * kThreads) row_prop_regs<5>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else if (n <= 6 * kThreads) row_prop_regs<6>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else if (n <= 7 * kThreads) row_prop_regs<7>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else if (n <= kRegRow) row_prop_regs<8>(P, D, R, n, src, dst, need, cov0, par, tab, S);
    else {
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
            if (cfree) { S.rows_live += 1u; if (R - D.row_base < D.K) S.maxlive = max(S.maxlive, (unsigned)cfree); }   // thread 0 only
        }
    }
}

// GREEDY: runs right after a PROP round that changed nothing, so the live list is exact (all FREE, cell field = covered
// flag, row_cov current).
template <int EPT>
__device__ __forceinline__ void row_greedy_regs(const Params& P, const WinDesc& D, int n, const uint32_t* src, int d,
                                                unsigned long long* keytab, BlockScratch& S) {
    const uint8_t* st_w = P.st + D.var_base;
    const float* gain_w = P.gain + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;

    RowRegs<EPT> X;
    load_row(X, src, n, st_w);
    unsigned long long key[EPT];
    int nfree = 0;
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        const bool fr = X.e[b] != kEntInvalid && X.s[b] == ST_FREE;
        key[b] = fr ? make_key(gain_w[X.e[b] >> kCellBits], tie_of(D, X.e[b] >> kCellBits)) : 0ull;
        if (fr) { ++nfree; if ((X.e[b] & kCellCov) != kCellCov) keytab[X.e[b] & kCellCov] = 0ull; }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < EPT; ++b)
        if (key[b] && (X.e[b] & kCellCov) != kCellCov) atomicMax(&keytab[X.e[b] & kCellCov], key[b]);
    __syncthreads();
#pragma unroll
    for (int b = 0; b < EPT; ++b)
        if (key[b] && (X.e[b] & kCellCov) != kCellCov && keytab[X.e[b] & kCellCov] != key[b])
            atomicOr(&acc_w[X.e[b] >> kCellBits], FLAG_BLOCKED);
    if (d > 0) {
        int z0 = 0, z1 = 0;
        block_sum3(S, nfree, z0, z1);
        if (nfree > d) {
            const unsigned long long thr = block_kth_largest(S, d, [&](auto sink) {
#pragma unroll
                for (int b = 0; b < EPT; ++b) if (key[b]) sink(key[b]);
            });
#pragma unroll
            for (int b = 0; b < EPT; ++b)
                if (key[b]) atomicOr(&acc_w[X.e[b] >> kCellBits], key[b] > thr ? FLAG_NOMINATED : FLAG_BLOCKED);
        } else {
#pragma unroll
            for (int b = 0; b < EPT; ++b) if (key[b]) atomicOr(&acc_w[X.e[b] >> kCellBits], FLAG_NOMINATED);
        }
    }
}

__device__ void row_greedy(const Params& P, const WinDesc& D, int R, int n, unsigned long long* keytab, BlockScratch& S) {
    if (n == 0) return;
    const uint32_t* src = P.live + P.row_off[R];
    const uint8_t* st_w = P.st + D.var_base;
    const float* gain_w = P.gain + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int d = max(0, P.row_need[R] - P.row_cov[R]);
    if (n <= kThreads) row_greedy_regs<1>(P, D, n, src, d, keytab, S);
    else if (n <= 4 * kThreads) row_greedy_regs<4>(P, D, n, src, d, keytab, S);
    else if (n <= kRegRow) row_greedy_regs<8>(P, D, n, src, d, keytab, S);
    else {
        for (int c = threadIdx.x; c < kCells; c += kThreads) keytab[c] = 0ull;
        __syncthreads();
        int nfree = 0;
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const uint32_t e = src[i];
            const unsigned v = e >> kCellBits;
            if (st_w[v] != ST_FREE) continue;
            ++nfree;
            if ((e & kCellCov) != kCellCov) atomicMax(&keytab[e & kCellCov], make_key(gain_w[v], tie_of(D, v)));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const uint32_t e = src[i];
            const unsigned v = e >> kCellBits;
            if (st_w[v] != ST_FREE || (e & kCellCov) == kCellCov) continue;
            if (keytab[e & kCellCov] != make_key(gain_w[v], tie_of(D, v))) atomicOr(&acc_w[v], FLAG_BLOCKED);
        }
        if (d > 0) {
            int z0 = 0, z1 = 0;
            block_sum3(S, nfree, z0, z1);
            if (nfree > d) {
                const unsigned long long thr = block_kth_largest(S, d, [&](auto sink) {
                    for (int i = threadIdx.x; i < n; i += kThreads) {
                        const unsigned v = src[i] >> kCellBits;
                        if (st_w[v] == ST_FREE) sink(make_key(gain_w[v], tie_of(D, v)));
                    }
                });
                for (int i = threadIdx.x; i < n; i += kThreads) {
                    const unsigned v = src[i] >> kCellBits;
                    if (st_w[v] != ST_FREE) continue;
                    atomicOr(&acc_w[v], make_key(gain_w[v], tie_of(D, v)) > thr ? FLAG_NOMINATED : FLAG_BLOCKED);
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

// Warp-per-row variants for lists of at most 32 entries (one entry per lane).  Entries of one cell are found with
// match.any instead of the shared-memory cell table, so eight short rows are in flight per CTA and none of them needs a
// block barrier.  Same arithmetic as the block versions.
__device__ __forceinline__ void warp_row_prop(const Params& P, const WinDesc& D, int R, int n, bool from_csr, unsigned& rows_live) {
    const int lane = threadIdx.x & 31;
    const int off = P.row_off[R];
    const uint32_t* src = (from_csr ? P.ent : P.live) + off;
    uint32_t* dst = P.live + off;
    const uint8_t* st_w = P.st + D.var_base;
    const bool valid = lane < n;
    uint32_t e = valid ? src[lane] : kEntInvalid;
    const int need = P.row_need[R], cov0 = P.row_cov[R];
    const uint8_t s = valid ? st_w[e >> kCellBits] : (uint8_t)ST_NOTVAR;
    const unsigned cell = e & kCellCov;
    const bool hascell = valid && cell != kCellCov;
    const unsigned mIN = __ballot_sync(0xFFFFFFFFu, s == ST_IN);
    const unsigned mFR = __ballot_sync(0xFFFFFFFFu, s == ST_FREE);
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, hascell ? cell : 0x10000u);
    const int cin = __popc(mIN), cfree = __popc(mFR);
    const int cov = cov0 + cin;
    const int d = max(0, need - cov);
    const bool defi = d > 0, critr = defi && d >= cfree;
    if (s == ST_FREE) {
        unsigned long long add = 0;
        bool covered = true;
        if (hascell) {
            covered = (grp & mIN) != 0u;
            if (!covered) { add |= 1ull; if (__popc(grp & mFR) == 1) add |= 1ull << 16; }
        }
        if (defi) add |= 1ull << 32;
        if (critr) add |= 1ull << 48;
        if (add) atomicAdd(&P.acc[D.var_base + (e >> kCellBits)], add);
        if (covered) e |= kCellCov;
    }
    __syncwarp();                                           // every lane has read src before the in-place writes
    if (s == ST_FREE) dst[__popc(mFR & ((1u << lane) - 1u))] = e;
    if (lane == 0) {
        if (cin) P.row_cov[R] = cov;
        P.live_n[R] = cfree;
        if (cfree) rows_live += 1u;
    }
}

__device__ __forceinline__ void warp_row_greedy(const Params& P, const WinDesc& D, int R, int n) {
    const int lane = threadIdx.x & 31;
    const uint32_t* src = P.live + P.row_off[R];
    const bool valid = lane < n;
    const uint32_t e = valid ? src[lane] : kEntInvalid;
    const unsigned v = e >> kCellBits;
    const int d = max(0, P.row_need[R] - P.row_cov[R]);
    const bool fr = valid && P.st[D.var_base + v] == ST_FREE;
    const unsigned long long key = fr ? make_key(P.gain[D.var_base + v], tie_of(D, v)) : 0ull;
    const unsigned cell = e & kCellCov;
    const bool unc = fr && cell != kCellCov;
    const unsigned mFR = __ballot_sync(0xFFFFFFFFu, fr);
    const unsigned mU = __ballot_sync(0xFFFFFFFFu, unc);
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, unc ? cell : 0x10000u);
    unsigned long long best = 0ull;
    for (unsigned rem = mU; rem; rem &= rem - 1u) {
        const int j = __ffs(rem) - 1;
        const unsigned long long kj = __shfl_sync(0xFFFFFFFFu, key, j);
        if ((grp >> j) & 1u) best = max(best, kj);
    }
    unsigned long long flags = 0ull;
    if (unc && key != best) flags |= FLAG_BLOCKED;
    if (d > 0) {
        const int nfree = __popc(mFR);
        if (nfree > d) {
            int cge = 0;
            for (unsigned rem = mFR; rem; rem &= rem - 1u) {
                const int j = __ffs(rem) - 1;
                const unsigned long long kj = __shfl_sync(0xFFFFFFFFu, key, j);
                cge += (kj >= key) ? 1 : 0;
            }
            if (fr) flags |= (cge <= d) ? FLAG_NOMINATED : FLAG_BLOCKED;     // key > (d+1)-th largest  <=>  #{keys >= key} <= d
        } else if (fr) {
            flags |= FLAG_NOMINATED;
        }
    }
    if (flags) atomicOr(&P.acc[D.var_base + v], flags);
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-per-row variants for lists of 33..256 entries (eight entries per lane, lane-strided: coalesced, list order =
// (chunk, lane) order).  Each warp owns a 4-bit-per-cell table in shared memory (the eight tables alias the CTA's cell
// table): bit 0 = "at least one", bit 1 = "at least two" (set by the second atomicOr that finds bit 0), bit 2 = "has an
// IN entry" -- exactly what the dominance rules read (a count matters only as 0 / 1 / more).  No block barrier: eight
// such rows are in flight per CTA.  Same arithmetic as the block versions; list order is preserved.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned nib_of(const unsigned* wt, unsigned cell) { return (wt[cell >> 3] >> ((cell & 7u) * 4u)) & 0xFu; }
// returns true when this call was the first to mark the cell
__device__ __forceinline__ bool nib_count(unsigned* wt, unsigned cell) {
    const unsigned sh = (cell & 7u) * 4u;
    const unsigned old = atomicOr(&wt[cell >> 3], 1u << sh);
    if ((old >> sh) & 1u) { atomicOr(&wt[cell >> 3], 2u << sh); return false; }
    return true;
}

__device__ __forceinline__ void warp_tab_load(uint32_t (&e)[kWarpEpt], uint8_t (&st)[kWarpEpt], const uint32_t* src, int n,
                                              const uint8_t* st_w) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int b = 0; b < kWarpEpt; ++b) {
        const int idx = b * 32 + lane;
        e[b] = idx < n ? src[idx] : kEntInvalid;
    }
#pragma unroll
    for (int b = 0; b < kWarpEpt; ++b) st[b] = e[b] != kEntInvalid ? st_w[e[b] >> kCellBits] : (uint8_t)ST_NOTVAR;
}

__device__ __forceinline__ void warp_tab_prop(const Params& P, const WinDesc& D, int R, int n, bool from_csr, unsigned* wt,
                                              unsigned& rows_live) {
    const int lane = threadIdx.x & 31;
    const int off = P.row_off[R];
    const uint32_t* src = (from_csr ? P.ent : P.live) + off;
    uint32_t* dst = P.live + off;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R], cov0 = P.row_cov[R];
    uint32_t e[kWarpEpt];
    uint8_t st[kWarpEpt];
    warp_tab_load(e, st, src, n, P.st + D.var_base);
#pragma unroll
    for (int b = 0; b < kWarpEpt; ++b)
        if (e[b] != kEntInvalid && (e[b] & kCellCov) != kCellCov) wt[(e[b] & kCellCov) >> 3] = 0u;
    __syncwarp();
    int cin = 0, cfree = 0;
    unsigned mfr[kWarpEpt];
#pragma unroll
    for (int b = 0; b < kWarpEpt; ++b) {
        const unsigned cell = e[b] & kCellCov;
        const bool in = st[b] == ST_IN, fr = st[b] == ST_FREE;          // (invalid lanes carry ST_NOTVAR)
        if (in && cell != kCellCov) atomicOr(&wt[cell >> 3], 4u << ((cell & 7u) * 4u));
        if (fr && cell != kCellCov) nib_count(wt, cell);
        cin += __popc(__ballot_sync(0xFFFFFFFFu, in));
        mfr[b] = __ballot_sync(0xFFFFFFFFu, fr);
        cfree += __popc(mfr[b]);
    }
    __syncwarp();
    const int cov = cov0 + cin;
    const int d = max(0, need - cov);
    const bool defi = d > 0, critr = defi && d >= cfree;
    const unsigned lt = (1u << lane) - 1u;
    int pos = 0;
#pragma unroll
    for (int b = 0; b < kWarpEpt; ++b) {
        if ((mfr[b] >> lane) & 1u) {
            const unsigned cell = e[b] & kCellCov;
            unsigned long long add = 0;
            bool covered = true;
            if (cell != kCellCov) {
                const unsigned t = nib_of(wt, cell);
                covered = (t & 4u) != 0u;
                if (!covered) { add |= 1ull; if (!(t & 2u)) add |= 1ull << 16; }
            }
            if (defi) add |= 1ull << 32;
            if (critr) add |= 1ull << 48;
            if (add) atomicAdd(&acc_w[e[b] >> kCellBits], add);
            dst[pos + __popc(mfr[b] & lt)] = covered ? (e[b] | kCellCov) : e[b];   // every lane holds its entries in registers
        }
        pos += __popc(mfr[b]);
    }
    if (lane == 0) {
        if (cin) P.row_cov[R] = cov;
        P.live_n[R] = cfree;
        if (cfree) rows_live += 1u;
    }
}

// sweep (D1 / EVAL) of a list of at most 256 entries: see row_d1_regs / row_d1_eval
__device__ __forceinline__ void warp_tab_d1(const Params& P, const WinDesc& D, RoundCnt& rc, int R, int n, bool accumulate, bool from_in,
                                            unsigned* wt) {
    const int lane = threadIdx.x & 31;
    const int off = P.row_off[R];
    const uint32_t* src = (from_in ? P.live : P.ent) + off;
    uint32_t* dst = P.live + off;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R];
    int cin = 0, ccells = 0;
    if (n > 0) {
        uint32_t e[kWarpEpt];
        uint8_t st[kWarpEpt];
        warp_tab_load(e, st, src, n, P.st + D.var_base);
#pragma unroll
        for (int b = 0; b < kWarpEpt; ++b)
            if (e[b] != kEntInvalid && (e[b] & kCellCov) != kCellCov) wt[(e[b] & kCellCov) >> 3] = 0u;
        __syncwarp();
        unsigned m[kWarpEpt];
#pragma unroll
        for (int b = 0; b < kWarpEpt; ++b) {
            const bool in = st[b] == ST_IN;
            const unsigned cell = e[b] & kCellCov;
            if (in && cell != kCellCov && nib_count(wt, cell)) ++ccells;
            m[b] = __ballot_sync(0xFFFFFFFFu, in);
            cin += __popc(m[b]);
        }
        ccells = __reduce_add_sync(0xFFFFFFFFu, ccells);
        __syncwarp();
        const bool critr = cin <= need;
        const unsigned lt = (1u << lane) - 1u;
        int pos = 0;
#pragma unroll
        for (int b = 0; b < kWarpEpt; ++b) {
            if ((m[b] >> lane) & 1u) {
                dst[pos + __popc(m[b] & lt)] = e[b];                      // the row's IN list
                if (accumulate) {
                    const unsigned cell = e[b] & kCellCov;
                    unsigned long long add = 0;
                    if (cell != kCellCov && nib_of(wt, cell) == 1u) add |= 1ull;
                    if (critr) add |= 1ull << 32;
                    if (add) atomicAdd(&acc_w[e[b] >> kCellBits], add);
                }
            }
            pos += __popc(m[b]);
        }
    }
    if (lane == 0) {
        const int slack = max(0, need - cin);
        const int local = R - D.row_base;
        const int words = (D.M + 31) >> 5;
        uint32_t* slot = P.out + D.out_off + kHdrWords + words;
        slot[local] = (uint32_t)cin;
        slot[D.K + D.H + local] = (uint32_t)slack;
        P.live_n[R] = cin;
        const int unc = P.row_ncell[R] - ccells;
        if (unc) atomicAdd(&rc.uncovered, (unsigned)unc);
        if (slack) atomicAdd(&rc.slack, (unsigned)slack);
    }
}

// D1 (and EVAL): one sweep of the row's CSR segment: IN counts per cell and per row; D1 adds the criticality counters
// of the IN points; both write the row's coverage / slack and the uncovered-cell count (the read-out uses the values of
// the last sweep, which is the one that found nothing left to drop).
template <int EPT>
__device__ __forceinline__ void row_d1_regs(const Params& P, const WinDesc& D, int n, const uint32_t* src, uint32_t* dst, int need,
                                            int par, bool accumulate, unsigned* tab, BlockScratch& S, int& cin, int& ccells) {
    const uint8_t* st_w = P.st + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    RowRegs<EPT> X;
    load_row(X, src, n, st_w);
#pragma unroll
    for (int b = 0; b < EPT; ++b)           // (only the cells of IN entries are counted and read below: the others need no zeroing)
        if (X.e[b] != kEntInvalid && X.s[b] == ST_IN && (X.e[b] & kCellCov) != kCellCov) tab[X.e[b] & kCellCov] = 0u;
    __syncthreads();
    unsigned m[EPT];
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        const bool in = X.e[b] != kEntInvalid && X.s[b] == ST_IN;
        m[b] = __ballot_sync(0xFFFFFFFFu, in);
        if (!in) continue;
        const unsigned cell = X.e[b] & kCellCov;
        if (cell != kCellCov && atomicAdd(&tab[cell], 1u) == 0u) ++ccells;
    }
#pragma unroll
    for (int b = 0; b < EPT; ++b) cin += __popc(m[b]);          // warp count
    ccells = __reduce_add_sync(0xFFFFFFFFu, ccells);
    if (lane == 0) { S.pred[par][0][wid] = cin; S.pred[par][1][wid] = ccells; }
    __syncthreads();                                    // every entry of src is in registers; counts and cell table published
    int pos = 0;
    cin = 0; ccells = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) { if (q < wid) pos += S.pred[par][0][q]; cin += S.pred[par][0][q]; ccells += S.pred[par][1][q]; }
    {
        // the row's IN list (the later sweeps of the reverse delete read it instead of the whole CSR segment)
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int b = 0; b < EPT; ++b) {
            if ((m[b] >> lane) & 1u) dst[pos + __popc(m[b] & lt)] = X.e[b];
            pos += __popc(m[b]);
        }
    }
    if (accumulate && cin > 0) {
        const bool critr = cin <= need;
#pragma unroll
        for (int b = 0; b < EPT; ++b) {
            if (X.e[b] == kEntInvalid || X.s[b] != ST_IN) continue;
            const unsigned cell = X.e[b] & kCellCov;
            unsigned long long add = 0;
            if (cell != kCellCov && tab[cell] == 1u) add |= 1ull;
            if (critr) add |= 1ull << 32;
            if (add) atomicAdd(&acc_w[X.e[b] >> kCellBits], add);
        }
    }
}

__device__ void row_d1_eval(const Params& P, const WinDesc& D, RoundCnt& rc, int R, int n, int off, int par, bool accumulate,
                            bool from_in, unsigned* tab, BlockScratch& S) {
    const uint32_t* src = (from_in ? P.live : P.ent) + off;
    uint32_t* dst = P.live + off;
    const uint8_t* st_w = P.st + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R];
    int cin = 0, ccells = 0, z = 0;
    if (n > 0 && n <= kThreads) row_d1_regs<1>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0 && n <= 2 * kThreads) row_d1_regs<2>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0 && n <= 3 * kThreads) row_d1_regs<3>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0 && n <= 4 * kThreads) row_d1_regs<4>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0 && n <= 5 * kThreads) row_d1_regs<5>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0 && n <= 6 * kThreads) row_d1_regs<6>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0 && n <= 7 * kThreads) row_d1_regs<7>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0 && n <= kRegRow) row_d1_regs<8>(P, D, n, src, dst, need, par, accumulate, tab, S, cin, ccells);
    else if (n > 0) {
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
        // IN list of a long row (chunk by chunk; the barriers of the scan order the reads of a chunk before its writes)
        int out_base = 0;
        for (int base = 0; base < n; base += kThreads) {
            const int i = base + (int)threadIdx.x;
            uint32_t e = kEntInvalid;
            bool in = false;
            if (i < n) { e = src[i]; in = st_w[e >> kCellBits] == ST_IN; }
            int total;
            const int p = block_excl_scan(S, in ? 1 : 0, total);
            if (in) dst[out_base + p] = e;
            out_base += total;
        }
    }
    if (threadIdx.x == 0) {
        const int slack = max(0, need - cin);
        const int local = R - D.row_base;
        const int words = (D.M + 31) >> 5;
        uint32_t* slot = P.out + D.out_off + kHdrWords + words;
        slot[local] = (uint32_t)cin;
        slot[D.K + D.H + local] = (uint32_t)slack;
        P.live_n[R] = cin;
        const int unc = P.row_ncell[R] - ccells;
        if (unc) atomicAdd(&rc.uncovered, (unsigned)unc);
        if (slack) atomicAdd(&rc.slack, (unsigned)slack);
    }
}

// D2 on an IN list held in registers (one load of the entries, one gather of the states)
template <int EPT>
__device__ __forceinline__ void row_d2_regs(const Params& P, const WinDesc& D, int n, const uint32_t* src, int need, unsigned* tab,
                                            unsigned long long* keytab, BlockScratch& S) {
    const float* gain_w = P.gain + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    RowRegs<EPT> X;
    load_row(X, src, n, P.st + D.var_base);
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        const unsigned cell = X.e[b] & kCellCov;
        if (X.e[b] != kEntInvalid && cell != kCellCov) { tab[cell] = 0u; keytab[cell] = 0ull; }
    }
    __syncthreads();
    int cov = 0, ncand = 0, z = 0;
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        if (X.e[b] == kEntInvalid || (X.s[b] != ST_IN && X.s[b] != ST_CAND)) continue;
        ++cov;
        if (X.s[b] == ST_CAND) ++ncand;
        if ((X.e[b] & kCellCov) != kCellCov) atomicAdd(&tab[X.e[b] & kCellCov], 1u);
    }
    block_sum3(S, cov, ncand, z);
    if (ncand == 0) return;
    unsigned long long key[EPT];
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        key[b] = 0ull;
        if (X.e[b] != kEntInvalid && X.s[b] == ST_CAND) {
            const unsigned v = X.e[b] >> kCellBits, cell = X.e[b] & kCellCov;
            key[b] = make_key(gain_w[v], tie_of(D, v));
            if (cell != kCellCov && tab[cell] >= 2u) atomicMax(&keytab[cell], key[b]);
        }
    }
    __syncthreads();
    unsigned blocked = 0u;
#pragma unroll
    for (int b = 0; b < EPT; ++b) {
        if (!key[b]) continue;
        const unsigned cell = X.e[b] & kCellCov;
        if (cell != kCellCov && tab[cell] >= 2u && keytab[cell] != key[b]) blocked |= 1u << b;
    }
    const int u = cov - need;
    if (u > 0 && ncand > u) {
        const unsigned long long thr = block_kth_largest(S, u, [&](auto sink) {
#pragma unroll
            for (int b = 0; b < EPT; ++b) if (key[b]) sink(key[b]);
        });
#pragma unroll
        for (int b = 0; b < EPT; ++b) if (key[b] && !(key[b] > thr)) blocked |= 1u << b;
    }
#pragma unroll
    for (int b = 0; b < EPT; ++b) if ((blocked >> b) & 1u) atomicOr(&acc_w[X.e[b] >> kCellBits], FLAG_BLOCKED);
}

// D2: budgets of the reverse delete (per cell: keep at least one IN point; per row: at most cov - need removals).
// Reads the row's IN list written by the D1 sweep just before (every entry is IN or CAND now).
__device__ void row_d2(const Params& P, const WinDesc& D, int R, unsigned* tab, unsigned long long* keytab, BlockScratch& S) {
    const int n = P.live_n[R];
    if (n == 0) return;
    const uint32_t* src = P.live + P.row_off[R];
    const uint8_t* st_w = P.st + D.var_base;
    const float* gain_w = P.gain + D.var_base;
    unsigned long long* acc_w = P.acc + D.var_base;
    const int need = P.row_need[R];
    if (n <= kThreads) { row_d2_regs<1>(P, D, n, src, need, tab, keytab, S); return; }
    if (n <= 2 * kThreads) { row_d2_regs<2>(P, D, n, src, need, tab, keytab, S); return; }
    if (n <= 4 * kThreads) { row_d2_regs<4>(P, D, n, src, need, tab, keytab, S); return; }
    for (int i = threadIdx.x; i < n; i += kThreads) {           // lazy zeroing: only the cells this list touches
        const unsigned cell = src[i] & kCellCov;
        if (cell != kCellCov) { tab[cell] = 0u; keytab[cell] = 0ull; }
    }
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
        atomicMax(&keytab[cell], make_key(gain_w[v], tie_of(D, v)));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const uint32_t e = src[i];
        const unsigned v = e >> kCellBits, cell = e & kCellCov;
        if (st_w[v] != ST_CAND || cell == kCellCov || tab[cell] < 2u) continue;
        if (make_key(gain_w[v], tie_of(D, v)) != keytab[cell]) atomicOr(&acc_w[v], FLAG_BLOCKED);
    }
    const int u = cov - need;
    if (u > 0 && ncand > u) {
        const unsigned long long thr = block_kth_largest(S, u, [&](auto sink) {
            for (int i = threadIdx.x; i < n; i += kThreads) {
                const unsigned v = src[i] >> kCellBits;
                if (st_w[v] == ST_CAND) sink(make_key(gain_w[v], tie_of(D, v)));
            }
        });
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const unsigned v = src[i] >> kCellBits;
            if (st_w[v] != ST_CAND) continue;
            if (!(make_key(gain_w[v], tie_of(D, v)) > thr)) atomicOr(&acc_w[v], FLAG_BLOCKED);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// variable phases
// ---------------------------------------------------------------------------------------------------------------
// PROP / GREEDY / FORCE touch only FREE map points: they walk the compact list written by the last PROP phase
// (one thread per listed map point; PROP writes the next list).  No block barrier inside.
__device__ void var_list_phase(const Params& P, const WinDesc& D, WinState& ws, RoundCnt& rc, int mode, int greedy_steps,
                               const int* vsrc, int nsrc, int* vdst, const GroupCtx& G) {
    const bool any_rule = greedy_steps >= P.all_rule_steps;
    if (mode == MODE_PROP) {
        // four list positions per thread and iteration: independent loads in flight, one list reservation per warp
        const int gsz = G.ncta * kThreads;
        int changed = 0, sumdeg = 0;
        for (int base = G.cta * kThreads + (int)threadIdx.x; base - (int)(threadIdx.x & 31) < nsrc; base += gsz * 4) {
            int mps[4], nfr[4];
            uint8_t s4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = base + j * gsz;
                mps[j] = i < nsrc ? vsrc[i] : -1;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = mps[j] >= 0 ? P.st[D.var_base + mps[j]] : (uint8_t)ST_NOTVAR;
            unsigned long long a4[4];
            int nobs4[4];
            unsigned deg4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a4[j] = 0ull; nobs4[j] = 0; deg4[j] = 0u;
                if (s4[j] == ST_FREE) {
                    const int g = D.var_base + mps[j];
                    a4[j] = P.acc[g]; nobs4[j] = ld_nobs(D, mps[j]); deg4[j] = P.deg[g];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                nfr[j] = 0;
                if (s4[j] == ST_FREE) {
                    const int g = D.var_base + mps[j];
                    if (a4[j]) P.acc[g] = 0ull;
                    prop_decide(P, a4[j], ws.n_max - nobs4[j], &P.st[g], &P.gain[g], deg4[j], changed, nfr[j], sumdeg);
                }
            }
            prop_commit4(rc, vdst, mps, nfr);
        }
        prop_flush(rc, changed, sumdeg);
        return;
    }
    for (int base = G.cta * kThreads; base < nsrc; base += G.ncta * kThreads) {
        const int i = base + (int)threadIdx.x;
        int mp = -1, g = 0;
        uint8_t s = ST_NOTVAR;
        if (i < nsrc) { mp = vsrc[i]; g = D.var_base + mp; s = P.st[g]; }
        if (mode == MODE_PROP) {
            int c0 = 0, c1 = 0, c2 = 0;
            if (s == ST_FREE) {
                const unsigned long long a = P.acc[g];
                if (a) P.acc[g] = 0ull;
                prop_decide(P, a, ws.n_max - ld_nobs(D, mp), &P.st[g], &P.gain[g], P.deg[g], c0, c1, c2);
            }
            prop_commit(rc, vdst, mp, c0, c1, c2);
        } else if (mode == MODE_GREEDY) {
            if (s == ST_FREE) {
                const unsigned long long a = P.acc[g];
                if (a) P.acc[g] = 0ull;
                const bool sel = (P.gain[g] > 0.0f && !(a & FLAG_BLOCKED)) || (any_rule && (a & FLAG_NOMINATED));
                if (sel) P.st[g] = ST_IN;
            }
        } else {    // MODE_FORCE
            if (s == ST_FREE) P.st[g] = ST_IN;
        }
    }
}

// D1 / D2 / EVAL look at every map point: one flat pass, four map points in flight per thread (a warp always holds 32
// consecutive map points, which is what the ballot-packed keep words need), counters flushed once per warp
__device__ void var_flat_phase(const Params& P, const WinDesc& D, WinState& ws, RoundCnt& rc, int mode, int gt, int gsz) {
    const int lim = (D.M + 31) & ~31;
    const int words = (D.M + 31) >> 5;
    int ncand = 0, kept = 0;
    long long cost = 0;
    for (int base = gt; base < lim; base += gsz * kVpt) {
        uint8_t s4[kVpt];
#pragma unroll
        for (int j = 0; j < kVpt; ++j) {
            const int mp = base + j * gsz;
            s4[j] = mp < D.M ? P.st[D.var_base + mp] : (uint8_t)ST_NOTVAR;
        }
        if (mode == MODE_D1) {
            unsigned long long a4[kVpt];
            int nobs4[kVpt];
#pragma unroll
            for (int j = 0; j < kVpt; ++j) {
                a4[j] = 0ull; nobs4[j] = 0;
                if (s4[j] == ST_IN) { a4[j] = P.acc[D.var_base + base + j * gsz]; nobs4[j] = ld_nobs(D, base + j * gsz); }
            }
#pragma unroll
            for (int j = 0; j < kVpt; ++j) {
                if (s4[j] != ST_IN) continue;
                const int g = D.var_base + base + j * gsz;
                if (a4[j]) P.acc[g] = 0ull;
                const double critc = (double)(a4[j] & 0xFFFFu), critr = (double)((a4[j] >> 32) & 0xFFFFu);
                const double c = (double)(ws.n_max - nobs4[j]);
                const double dF = __dadd_rn(__dadd_rn(-c, __dmul_rn(P.glam, critc)), __dmul_rn(P.lam, critr));
                if (dF < 0.0) { P.st[g] = ST_CAND; P.gain[g] = (float)(-dF); ++ncand; }
            }
        } else if (mode == MODE_D2) {
#pragma unroll
            for (int j = 0; j < kVpt; ++j) {
                if (s4[j] != ST_CAND) continue;
                const int g = D.var_base + base + j * gsz;
                const unsigned long long a = P.acc[g];
                if (a) P.acc[g] = 0ull;
                P.st[g] = (a & FLAG_BLOCKED) ? ST_IN : ST_OUT;
            }
        } else {    // MODE_EVAL / MODE_EVALV: read-out (MapSparsification.cc:159-166): bit = 0 only for variables the solve rejected
#pragma unroll
            for (int j = 0; j < kVpt; ++j) {
                const int mp = base + j * gsz;
                if (mp >= lim) break;                                   // (warp-uniform)
                const bool keep = mp < D.M && s4[j] != ST_OUT;
                const unsigned word = __ballot_sync(0xFFFFFFFFu, keep);
                if ((threadIdx.x & 31) == 0 && (mp >> 5) < words) P.out[D.out_off + kHdrWords + (mp >> 5)] = word;
                if (s4[j] == ST_IN) { ++kept; cost += (long long)(ws.n_max - ld_nobs(D, mp)); }
            }
        }
    }
    if (mode == MODE_D1) {
        ncand = __reduce_add_sync(0xFFFFFFFFu, ncand);
        if ((threadIdx.x & 31) == 0 && ncand) atomicAdd(&rc.ncand, (unsigned)ncand);
    } else if (mode == MODE_EVAL || mode == MODE_EVALV) {
        kept = __reduce_add_sync(0xFFFFFFFFu, kept);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xFFFFFFFFu, cost, o);
        if ((threadIdx.x & 31) == 0 && kept) {
            atomicAdd(&rc.nkept, (unsigned)kept);
            atomicAdd(&rc.sumcost, (unsigned long long)cost);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory tail: once the undecided part of a window is small, CTA 0 of the group copies it into shared memory
// (FREE map points renumbered 0..nv-1, their row entries, the rows' running deficits) and runs the remaining PROP /
// GREEDY rounds there with block barriers only; the other CTAs of the group wait at one group barrier.
// Same arithmetic and tie-breaks (keys use the original map-point index) as the group-wide phases.
// ---------------------------------------------------------------------------------------------------------------
struct TailSmem {
    uint32_t ent[kTailEnts];              // (local id << 12) | cell, row segments back to back
    uint32_t acc_lo[kTailVars];            // [ubc:16 | lbc:16] or the GREEDY flags (32-bit: native shared-memory atomics)
    uint32_t acc_hi[kTailVars];            // [ubr:16 | lbr:16]
    float gain[kTailVars];
    int cost[kTailVars];
    uint32_t mp[kTailVars];               // original map-point index (write-back)
    uint32_t tie[kTailVars];              // tie-break key (mp_tie rank or the index)
    int rdef[kTailRows];                  // need - coverage of the row (may be negative)
    unsigned short rptr[kTailRows];
    unsigned short rn[kTailRows];
    unsigned short rsrc[kTailRows];       // window-local row index
    uint8_t st[kTailVars];
};

// one row of the tail, any length: PROP row step (see row_prop)
__device__ void tail_row_prop(TailSmem& T, int lid) {
    const int lane = threadIdx.x & 31;
    uint32_t* lst = T.ent + T.rptr[lid];
    const int n = T.rn[lid];
    if (n == 0) return;
    const unsigned lt = (1u << lane) - 1u;
    if (n <= 32) {
        const bool valid = lane < n;
        uint32_t e = valid ? lst[lane] : kEntInvalid;
        const uint8_t s = valid ? T.st[e >> kCellBits] : (uint8_t)ST_NOTVAR;
        const unsigned cell = e & kCellCov;
        const bool hascell = valid && cell != kCellCov;
        const unsigned mIN = __ballot_sync(0xFFFFFFFFu, s == ST_IN);
        const unsigned mFR = __ballot_sync(0xFFFFFFFFu, s == ST_FREE);
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, hascell ? cell : 0x10000u);
        const int cin = __popc(mIN), cfree = __popc(mFR);
        const int def = T.rdef[lid] - cin;
        const int d = max(0, def);
        const bool defi = d > 0, critr = defi && d >= cfree;
        if (s == ST_FREE) {
            uint32_t lo = 0, hi = 0;
            bool covered = true;
            if (hascell) {
                covered = (grp & mIN) != 0u;
                if (!covered) { lo |= 1u; if (__popc(grp & mFR) == 1) lo |= 1u << 16; }
            }
            if (defi) hi |= 1u;
            if (critr) hi |= 1u << 16;
            if (lo) atomicAdd(&T.acc_lo[e >> kCellBits], lo);
            if (hi) atomicAdd(&T.acc_hi[e >> kCellBits], hi);
            if (covered) e |= kCellCov;
        }
        __syncwarp();
        if (s == ST_FREE) lst[__popc(mFR & lt)] = e;
        if (lane == 0) { T.rdef[lid] = def; T.rn[lid] = (unsigned short)cfree; }
        return;
    }
    // long list: pass A counts, pass B matches every FREE entry against the whole list, pass C compacts
    int cin = 0, cfree = 0;
    for (int c = 0; c < n; c += 32) {
        const int i = c + lane;
        const uint8_t s = (i < n) ? T.st[lst[i] >> kCellBits] : (uint8_t)ST_NOTVAR;
        cin += __popc(__ballot_sync(0xFFFFFFFFu, s == ST_IN));
        cfree += __popc(__ballot_sync(0xFFFFFFFFu, s == ST_FREE));
    }
    const int def = T.rdef[lid] - cin;
    const int d = max(0, def);
    const bool defi = d > 0, critr = defi && d >= cfree;
    for (int ca = 0; ca < n; ca += 32) {
        const int ia = ca + lane;
        const uint32_t ea = (ia < n) ? lst[ia] : kEntInvalid;
        const bool fra = (ia < n) && T.st[ea >> kCellBits] == ST_FREE;
        const unsigned cella = ea & kCellCov;
        const bool hasa = fra && cella != kCellCov;
        bool covered = false;
        int nf = 0;
        if (__any_sync(0xFFFFFFFFu, hasa)) {
            // only chunks whose cell range overlaps this chunk's range can hold a match (a cheap filter; it prunes most
            // pairs when neighbouring slots fall into neighbouring cells and is harmless otherwise)
            const unsigned amin = __reduce_min_sync(0xFFFFFFFFu, hasa ? cella : 0xFFFFu);
            const unsigned amax = __reduce_max_sync(0xFFFFFFFFu, hasa ? cella : 0u);
            for (int cb = 0; cb < n; cb += 32) {
                const int ib = cb + lane;
                const uint32_t eb = (ib < n) ? lst[ib] : kEntInvalid;
                const uint8_t sb = (ib < n) ? T.st[eb >> kCellBits] : (uint8_t)ST_NOTVAR;
                // covered FREE entries may already carry kCellCov (rewritten below): they no longer match, which is fine
                // because an IN entry of the same cell still does
                const unsigned cellb = (sb == ST_IN || sb == ST_FREE) ? (eb & kCellCov) : kCellCov;
                const bool hasb = cellb != kCellCov;
                const unsigned bmin = __reduce_min_sync(0xFFFFFFFFu, hasb ? cellb : 0xFFFFu);
                const unsigned bmax = __reduce_max_sync(0xFFFFFFFFu, hasb ? cellb : 0u);
                if (bmin > amax || bmax < amin) continue;
                for (unsigned rem = __ballot_sync(0xFFFFFFFFu, hasb && cellb >= amin && cellb <= amax); rem; rem &= rem - 1u) {
                    const int j = __ffs(rem) - 1;
                    const unsigned cj = __shfl_sync(0xFFFFFFFFu, cellb, j);
                    const unsigned sj = __shfl_sync(0xFFFFFFFFu, (unsigned)sb, j);
                    if (hasa && cj == cella) { if (sj == ST_IN) covered = true; else ++nf; }
                }
            }
        }
        __syncwarp();
        if (fra) {
            uint32_t lo = 0, hi = 0;
            const bool cov2 = !hasa || covered;
            if (!cov2) { lo |= 1u; if (nf == 1) lo |= 1u << 16; }
            if (defi) hi |= 1u;
            if (critr) hi |= 1u << 16;
            if (lo) atomicAdd(&T.acc_lo[ea >> kCellBits], lo);
            if (hi) atomicAdd(&T.acc_hi[ea >> kCellBits], hi);
            if (hasa && covered) lst[ia] = ea | kCellCov;
        }
        __syncwarp();
    }
    int out = 0;
    for (int c = 0; c < n; c += 32) {
        const int i = c + lane;
        const uint32_t e = (i < n) ? lst[i] : kEntInvalid;
        const bool fr = (i < n) && T.st[e >> kCellBits] == ST_FREE;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, fr);
        __syncwarp();
        if (fr) lst[out + __popc(m & lt)] = e;
        out += __popc(m);
        __syncwarp();
    }
    if (lane == 0) { T.rdef[lid] = def; T.rn[lid] = (unsigned short)cfree; }
}

// one row of the tail, any length: GREEDY row step (see row_greedy); the list is exact (all FREE)
__device__ void tail_row_greedy(TailSmem& T, int lid) {
    const int lane = threadIdx.x & 31;
    const uint32_t* lst = T.ent + T.rptr[lid];
    const int n = T.rn[lid];
    if (n == 0) return;
    const int d = max(0, T.rdef[lid]);
    // what does this row have to decide?  cells with FREE candidates, and / or a deficit smaller than its FREE count
    int nfr = 0;
    bool any_unc = false;
    for (int c = 0; c < n; c += 32) {
        const int i = c + lane;
        const uint32_t e = (i < n) ? lst[i] : kEntInvalid;
        const bool fr = (i < n) && T.st[e >> kCellBits] == ST_FREE;
        nfr += __popc(__ballot_sync(0xFFFFFFFFu, fr));
        any_unc |= __any_sync(0xFFFFFFFFu, fr && (e & kCellCov) != kCellCov) != 0;
    }
    const bool rank = d > 0 && nfr > d;
    if (!any_unc && !rank) {
        if (d > 0) {
            for (int i = lane; i < n; i += 32) {
                const unsigned v = lst[i] >> kCellBits;
                if (T.st[v] == ST_FREE) atomicOr(&T.acc_lo[v], (uint32_t)FLAG_NOMINATED);
            }
        }
        return;
    }
    // (d+1)-th largest key of the row by bisection on the key bits: thr = max x with #{keys >= x} >= d + 1
    unsigned long long thr = 0ull;
    if (rank) {
        if (n <= 32 * kEpt) {
            unsigned long long k[kEpt];
#pragma unroll
            for (int b = 0; b < kEpt; ++b) {
                const int i = b * 32 + lane;
                k[b] = 0ull;
                if (i < n) {
                    const unsigned v = lst[i] >> kCellBits;
                    if (T.st[v] == ST_FREE) k[b] = make_key(T.gain[v], T.tie[v]);
                }
            }
            for (int bit = 63; bit >= 0; --bit) {
                const unsigned long long cand = thr | (1ull << bit);
                int c = 0;
#pragma unroll
                for (int b = 0; b < kEpt; ++b) c += (k[b] >= cand) ? 1 : 0;
                c = __reduce_add_sync(0xFFFFFFFFu, c);
                if (c > d) thr = cand;
            }
        } else {
            for (int bit = 63; bit >= 0; --bit) {
                const unsigned long long cand = thr | (1ull << bit);
                int c = 0;
                for (int i = lane; i < n; i += 32) {
                    const unsigned v = lst[i] >> kCellBits;
                    c += (T.st[v] == ST_FREE && make_key(T.gain[v], T.tie[v]) >= cand) ? 1 : 0;
                }
                c = __reduce_add_sync(0xFFFFFFFFu, c);
                if (c > d) thr = cand;
            }
        }
    }
    for (int ca = 0; ca < n; ca += 32) {
        const int ia = ca + lane;
        const uint32_t ea = (ia < n) ? lst[ia] : kEntInvalid;
        const unsigned va = ea >> kCellBits;
        const bool fra = (ia < n) && T.st[va] == ST_FREE;
        const unsigned long long keya = fra ? make_key(T.gain[va], T.tie[va]) : 0ull;
        const unsigned cella = ea & kCellCov;
        const bool unca = fra && cella != kCellCov;
        unsigned long long best = 0ull;
        if (__any_sync(0xFFFFFFFFu, unca)) {
            const unsigned amin = __reduce_min_sync(0xFFFFFFFFu, unca ? cella : 0xFFFFu);
            const unsigned amax = __reduce_max_sync(0xFFFFFFFFu, unca ? cella : 0u);
            for (int cb = 0; cb < n; cb += 32) {
                const int ib = cb + lane;
                const uint32_t eb = (ib < n) ? lst[ib] : kEntInvalid;
                const unsigned vb = eb >> kCellBits;
                const bool ub = (ib < n) && T.st[vb] == ST_FREE && (eb & kCellCov) != kCellCov;
                const unsigned cellb = ub ? (eb & kCellCov) : kCellCov;
                const unsigned bmin = __reduce_min_sync(0xFFFFFFFFu, ub ? cellb : 0xFFFFu);
                const unsigned bmax = __reduce_max_sync(0xFFFFFFFFu, ub ? cellb : 0u);
                if (bmin > amax || bmax < amin) continue;
                const unsigned long long keyb = ub ? make_key(T.gain[vb], T.tie[vb]) : 0ull;
                for (unsigned rem = __ballot_sync(0xFFFFFFFFu, ub && cellb >= amin && cellb <= amax); rem; rem &= rem - 1u) {
                    const int j = __ffs(rem) - 1;
                    const unsigned long long kj = __shfl_sync(0xFFFFFFFFu, keyb, j);
                    const unsigned cj = __shfl_sync(0xFFFFFFFFu, cellb, j);
                    if (unca && cj == cella) best = max(best, kj);
                }
            }
        }
        uint32_t flags = 0u;
        if (unca && keya != best) flags |= (uint32_t)FLAG_BLOCKED;
        if (d > 0 && fra) flags |= (uint32_t)((!rank || keya > thr) ? FLAG_NOMINATED : FLAG_BLOCKED);
        if (flags) atomicOr(&T.acc_lo[va], flags);
    }
}

// Runs on CTA 0 of the group.  vprev/nv: FREE list of the state the live lists were written for.  On return the states of
// those map points are final (IN / OUT) in P.st, and rounds / greedy_steps / status are updated.
__device__ void tail_solve(const Params& P, const WinDesc& D, WinState& ws, TailSmem& T, BlockScratch& S, const int* vprev, int nv,
                           int mode, int drop_mode, int& rounds, int& greedy_steps, int& status, int w, int& tn,
                           unsigned long long t_win) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows = D.K + D.H;
    // ---- gather ----------------------------------------------------------------------------------------------------
    for (int i = threadIdx.x; i < nv; i += kThreads) {
        const int mp = vprev[i];
        const int g = D.var_base + mp;
        T.mp[i] = (uint32_t)mp;
        T.tie[i] = tie_of(D, (unsigned)mp);
        T.st[i] = P.st[g];
        T.cost[i] = ws.n_max - ld_nobs(D, mp);
        T.acc_lo[i] = 0u;
        T.acc_hi[i] = 0u;
        T.gain[i] = P.gain[g];
        P.deg[g] = (unsigned)i;                 // deg is not needed any more: reuse it as the local-id map
    }
    int nr = 0, ne = 0;
    for (int base = 0; base < rows; base += kThreads) {
        const int r = base + (int)threadIdx.x;
        const int n = (r < rows) ? P.live_n[D.row_base + r] : 0;
        int tr, te;
        const int lid = nr + block_excl_scan(S, n > 0 ? 1 : 0, tr);
        const int ptr = ne + block_excl_scan(S, n, te);
        if (n > 0) {
            T.rptr[lid] = (unsigned short)ptr;
            T.rn[lid] = (unsigned short)n;
            T.rsrc[lid] = (unsigned short)r;
            T.rdef[lid] = P.row_need[D.row_base + r] - P.row_cov[D.row_base + r];
        }
        nr += tr;
        ne += te;
    }
    __syncthreads();
    for (int lid = wid; lid < nr; lid += kWarps) {
        const uint32_t* src = P.live + P.row_off[D.row_base + T.rsrc[lid]];
        uint32_t* dst = T.ent + T.rptr[lid];
        const int n = T.rn[lid];
        for (int i = lane; i < n; i += 32) {
            const uint32_t e = src[i];
            dst[i] = (P.deg[D.var_base + (e >> kCellBits)] << kCellBits) | (e & kCellCov);
        }
    }
    __syncthreads();
    if (P.trace && threadIdx.x == 0 && tn < kTraceCap - 1) {
        P.trace[(size_t)w * kTraceCap + tn] = make_uint2((20u << 24) | (unsigned)ne, (unsigned)(globaltimer_ns() - t_win));
        ++tn;
    }
    // ---- rounds ----------------------------------------------------------------------------------------------------
    while (mode == MODE_PROP || mode == MODE_GREEDY) {
        const int mode0 = mode;
        if (mode == MODE_PROP) {
            if (threadIdx.x < 2) S.tcnt[threadIdx.x] = 0;
            for (int lid = wid; lid < nr; lid += kWarps) tail_row_prop(T, lid);
            __syncthreads();
            int c0 = 0, c1 = 0, c2 = 0;
            for (int i = threadIdx.x; i < nv; i += kThreads) {
                if (T.st[i] != ST_FREE) continue;
                const unsigned long long a = (unsigned long long)T.acc_lo[i] | ((unsigned long long)T.acc_hi[i] << 32);
                T.acc_lo[i] = 0u;
                T.acc_hi[i] = 0u;
                prop_decide(P, a, T.cost[i], &T.st[i], &T.gain[i], 0u, c0, c1, c2);
            }
            if (c0) atomicAdd(&S.tcnt[0], c0);
            if (c1) atomicAdd(&S.tcnt[1], c1);
            __syncthreads();
            ++rounds;
            const unsigned changed = (unsigned)S.tcnt[0], nfree = (unsigned)S.tcnt[1];
            __syncthreads();                         // everyone has read the counters before they are zeroed again
            if (changed > 0 && rounds < P.max_rounds)
                mode = (P.stall_den > 0 && rounds >= 2 && (unsigned long long)changed * (unsigned)P.stall_den < nfree) ? MODE_GREEDY : MODE_PROP;
            else if (nfree == 0) mode = drop_mode;
            else if (rounds >= P.max_rounds) { status = -5; mode = MODE_FORCE; }
            else mode = MODE_GREEDY;
        } else {
            for (int lid = wid; lid < nr; lid += kWarps) tail_row_greedy(T, lid);
            __syncthreads();
            const bool any_rule = greedy_steps >= P.all_rule_steps;
            for (int i = threadIdx.x; i < nv; i += kThreads) {
                if (T.st[i] != ST_FREE) continue;
                const uint32_t a = T.acc_lo[i];
                T.acc_lo[i] = 0u;
                const bool sel = (T.gain[i] > 0.0f && !(a & FLAG_BLOCKED)) || (any_rule && (a & FLAG_NOMINATED));
                if (sel) T.st[i] = ST_IN;
            }
            __syncthreads();
            ++greedy_steps;
            ++rounds;
            mode = MODE_PROP;
        }
        if (P.trace && threadIdx.x == 0 && tn < kTraceCap - 1) {
            P.trace[(size_t)w * kTraceCap + tn] = make_uint2(((20u + (unsigned)mode0) << 24) | (unsigned)S.tcnt[1], (unsigned)(globaltimer_ns() - t_win));
            ++tn;
        }
    }
    // ---- write-back (FORCE: every point still FREE is taken) --------------------------------------------------------------
    for (int i = threadIdx.x; i < nv; i += kThreads) {
        uint8_t s = T.st[i];
        if (s == ST_FREE) s = ST_IN;
        P.st[D.var_base + T.mp[i]] = s;
    }
    if (threadIdx.x == 0) { ws.t_rounds = rounds; ws.t_greedy = greedy_steps; ws.t_status = status; }
}

// Row phase of PROP / GREEDY / D1 / EVAL over the rows of this CTA: the rows are classified by list length in chunks of
// 256; long lists are processed by the whole CTA one after the other (the next row's entries are prefetched into L1 while
// the current one is processed), short PROP / GREEDY lists by one warp each.
__device__ void row_phase_lists(const Params& P, const WinDesc& D, RoundCnt& rc, const GroupCtx& G, int mode, bool from_csr,
                                bool from_in, unsigned* tab, unsigned long long* keytab, BlockScratch& S) {
    const int rows = D.K + D.H;
    const int mine = (rows - G.cta + G.ncta - 1) / G.ncta;          // rows G.cta, G.cta + ncta, ...
    const bool sweep = mode == MODE_D1 || mode == MODE_EVAL;        // every row (also empty ones) reports; the first sweep reads
    const bool csr = from_csr || (sweep && !from_in);               // the CSR and leaves IN lists, later sweeps read those
    const int* listn = csr ? P.ent_n : P.live_n;
    const uint32_t* lists = csr ? P.ent : P.live;
    const int wid = threadIdx.x >> 5;
    unsigned rows_live = 0;
    if (threadIdx.x == 0) { S.rows_live = 0u; S.maxlive = 0u; S.work = 0ull; }
    for (int base = 0; base < mine; base += kThreads) {
        if (threadIdx.x < 2) S.qn[threadIdx.x] = 0;
        __syncthreads();
        const int i = base + (int)threadIdx.x;
        if (i < mine) {
            const int R = D.row_base + G.cta + i * G.ncta;
            const int n = listn[R];
            S.rown[threadIdx.x] = n;
            S.rowoff[threadIdx.x] = P.row_off[R];
            // whole CTA: long lists, and GREEDY lists above the match.any size (it needs a 64-bit key per cell); one warp:
            // lists of up to 256 entries (PROP and the sweeps; a sweep also visits empty rows: they report their slack)
            const bool cta_row = n > kWarpTabRow || (mode == MODE_GREEDY && n > kWarpRow);
            if (cta_row) S.rowq[atomicAdd(&S.qn[0], 1)] = (unsigned short)threadIdx.x;
            else if (n > 0 || sweep) S.rowq[kThreads - 1 - atomicAdd(&S.qn[1], 1)] = (unsigned short)threadIdx.x;
        }
        __syncthreads();
        const int nlong = S.qn[0], nshort = S.qn[1];
        if (threadIdx.x == 0) {
            unsigned long long wsum = 0;
            for (int q = 0; q < nlong; ++q) wsum += (unsigned long long)S.rown[S.rowq[q]];
            for (int q = 0; q < nshort; ++q) wsum += (unsigned long long)S.rown[S.rowq[kThreads - 1 - q]];
            S.work += wsum;
        }
        for (int q = 0; q < nlong; ++q) {
            const int slot = S.rowq[q];
            const int R = D.row_base + G.cta + (base + slot) * G.ncta;
            if (q + 1 < nlong) {                                    // one 128-byte line per thread covers 8192 entries
                const int ns = S.rowq[q + 1];
                if ((int)threadIdx.x * 32 < S.rown[ns]) prefetch_l1(lists + S.rowoff[ns] + threadIdx.x * 32);
            }
            // PROP and the sweeps alternate between two cell tables (the second one lies in the key table, which only GREEDY
            // and D2 use): a row needs no barrier to protect its table from the next row's lazy zeroing
            unsigned* t = (q & 1) ? reinterpret_cast<unsigned*>(keytab) : tab;
            if (mode == MODE_PROP) row_prop(P, D, R, S.rown[slot], S.rowoff[slot], q & 1, from_csr, t, S);
            else if (sweep) row_d1_eval(P, D, rc, R, S.rown[slot], S.rowoff[slot], q & 1, mode == MODE_D1, from_in, t, S);
            else { row_greedy(P, D, R, S.rown[slot], keytab, S); __syncthreads(); }
        }
        if (nlong > 0 && nshort > 0 && mode != MODE_GREEDY) __syncthreads();     // the warps' tables alias the CTA's
        unsigned* wt = tab + wid * kWarpTabWords;          // this warp's nibble table (the CTA rows are done with `tab`)
        for (int q = wid; q < nshort; q += kWarps) {
            const int slot = S.rowq[kThreads - 1 - q];
            const int R = D.row_base + G.cta + (base + slot) * G.ncta;
            const int n = S.rown[slot];
            if (mode == MODE_PROP) {
                if (n <= kWarpRow) warp_row_prop(P, D, R, n, from_csr, rows_live);
                else warp_tab_prop(P, D, R, n, from_csr, wt, rows_live);
            } else if (sweep) {
                warp_tab_d1(P, D, rc, R, n, mode == MODE_D1, from_in, wt);
            } else {
                warp_row_greedy(P, D, R, n);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && S.work) atomicAdd(&P.ctrl->row_entries, S.work);
    if (mode == MODE_PROP) {
        if ((threadIdx.x & 31) == 0 && rows_live) atomicAdd(&S.rows_live, rows_live);
        __syncthreads();
        if (threadIdx.x == 0 && S.rows_live) atomicAdd(&rc.rows_live, S.rows_live);
        if (threadIdx.x == 0 && S.maxlive > (unsigned)kWarpRow) atomicMax(&rc.maxlive, S.maxlive);
    }
}

__device__ __forceinline__ void trace_mark(const Params& P, const GroupCtx& G, int w, int& tn, int mode, unsigned info,
                                           unsigned long long t_win) {
    if (P.trace && G.cta == 0 && threadIdx.x == 0 && tn < kTraceCap - 1) {
        P.trace[(size_t)w * kTraceCap + tn] = make_uint2(((unsigned)mode << 24) | min(info, 0xFFFFFFu),
                                                        (unsigned)(globaltimer_ns() - t_win));
        ++tn;
    }
}

#include "mss_bound.cuh"

// ---------------------------------------------------------------------------------------------------------------
// one window, solved by the CTAs of one group
// ---------------------------------------------------------------------------------------------------------------
__device__ bool solve_window(const Params& P, GroupCtx& G, int w, unsigned* tab, unsigned long long* keytab, TailSmem& T,
                             BlockScratch& S) {
    const WinDesc D = P.win[w];
    WinState& ws = P.ws[w];
    const int rows = D.K + D.H;
    const int gt = G.cta * kThreads + (int)threadIdx.x, gsz = G.ncta * kThreads;
    const unsigned long long t_win = globaltimer_ns();
    int tn = 0;

    // ---- W0: state init ------------------------------------------------------------------------------------------
    {
        const int mpad = ((max(D.M, 1) + kVarTile - 1) / kVarTile) * kVarTile;
        uint32_t* seen32 = reinterpret_cast<uint32_t*>(P.seen + D.var_base);
        for (int i = gt; i < mpad / 4; i += gsz) seen32[i] = 0u;
        for (int i = gt; i < mpad; i += gsz) P.acc[D.var_base + i] = 0ull;
        if (D.packed && D.H > 0) for (int i = gt; i < mpad; i += gsz) P.deg[D.var_base + i] = 0u;
        for (int j = gt; j < D.H; j += gsz) P.ent_n[D.row_base + D.K + j] = 0;
        if (G.cta == 0) {
            uint32_t* z = reinterpret_cast<uint32_t*>(&ws);
            for (int i = threadIdx.x; i < (int)(sizeof(WinState) / 4); i += kThreads) z[i] = 0u;
            __syncthreads();
            if (threadIdx.x == 0) ws.n_max = max(D.n_max_floor, 0);
            if (threadIdx.x == 0) P.out[D.out_off + 14] = 0u;                          // "slot not written"
        }
    }
    if (!group_sync(P, G)) return false;
    trace_mark(P, G, w, tn, 10, 0, t_win);
    // ---- W1: keyframe rows ---------------------------------------------------------------------------------------
    // PACKED16: the tokens of this CTA's NEXT row are staged in shared memory by the TMA unit (cp.async.bulk, completion on an
    // mbarrier) while the current row is processed -- they are there when the row starts, whatever the L1 held on to.  The two
    // 4 KB buffers and the barriers live in the upper third of the key-table area, which the build phases do not use (the
    // second cell table of the row parity scheme occupies its lower half).  A bulk copy needs 16-byte aligned addresses and
    // sizes, a row starts at any token: the copy covers the aligned superset of the row, rows that touch the last partial
    // 16 bytes of the token array (or are longer than 2048 tokens) are read from global memory as before.
    {
        unsigned char* free_area = reinterpret_cast<unsigned char*>(keytab) + (size_t)kCells * 4;      // 12 KB not used by W1
        uint16_t* tbuf[2] = {reinterpret_cast<uint16_t*>(free_area), reinterpret_cast<uint16_t*>(free_area + 4608)};
        unsigned long long* tbar = reinterpret_cast<unsigned long long*>(free_area + 2 * 4608);
        const uintptr_t tk_base = reinterpret_cast<uintptr_t>(D.feat_mp);
        const bool w1_tma = D.packed == 2 && (tk_base & 15u) == 0 && P.w1_tma;
        if (w1_tma && threadIdx.x == 0) {
            mbar_init(&tbar[0], 1);
            mbar_init(&tbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const uintptr_t lim = tk_base + (((size_t)D.F * 2) & ~(size_t)15);
        // arms buffer `b` with the tokens of row k; returns the offset (in tokens) of the row's first token inside it, -1 = no
        auto stage_row = [&](int k, int b) -> int {
            if (!w1_tma || k >= D.K) return -1;
            const int beg = ldv(D.feat_ptr + k), end = ldv(D.feat_ptr + k + 1);
            if (beg < 0 || end <= beg || end > D.F || end - beg > kRegRow) return -1;
            const uintptr_t a = tk_base + (size_t)beg * 2, a0 = a & ~(uintptr_t)15, a1 = (tk_base + (size_t)end * 2 + 15) & ~(uintptr_t)15;
            if (a1 > lim) return -1;
            if (threadIdx.x == 0) {
                mbar_expect_tx(&tbar[b], (uint32_t)(a1 - a0));
                bulk_g2s(tbuf[b], reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0), &tbar[b]);
            }
            return (int)((a - a0) >> 1);
        };
        unsigned uses[2] = {0u, 0u};
        int it = 0;
        int skip_cur = stage_row(G.cta, 0);
        for (int k = G.cta; k < D.K; k += G.ncta, ++it) {
            // (buffer (it + 1) & 1 was read by the row before this one; every thread has passed that row's barriers)
            const int skip_next = stage_row(k + G.ncta, (it + 1) & 1);
            if (!w1_tma && k + G.ncta < D.K) {           // other layouts: pull the next row's slots towards L1 while this one is processed
                const int nb = ldv(D.feat_ptr + k + G.ncta), ne = ldv(D.feat_ptr + k + G.ncta + 1);
                if (nb >= 0 && ne <= D.F) {
                    if (D.packed == 2) {
                        const uint16_t* tk = reinterpret_cast<const uint16_t*>(D.feat_mp);
                        for (int i = nb + (int)threadIdx.x * 64; i < ne; i += kThreads * 64) prefetch_l1(tk + i);
                    } else {
                        for (int i = nb + (int)threadIdx.x * 32; i < ne; i += kThreads * 32) prefetch_l1(D.feat_mp + i);
                    }
                    if (!D.packed) for (int i = nb + (int)threadIdx.x * 64; i < ne; i += kThreads * 64) prefetch_l1(D.feat_cell + i);
                }
            }
            const uint16_t* stok = nullptr;
            if (skip_cur >= 0) {
                mbar_wait(&tbar[it & 1], uses[it & 1] & 1u);
                ++uses[it & 1];
                stok = tbuf[it & 1] + skip_cur;
            }
            const int par = (k / G.ncta) & 1;              // cell table and scratch alternate: no barrier between rows
            w1_build_row(P, D, ws, k, par, par ? reinterpret_cast<unsigned*>(keytab) : tab, S, stok);
            skip_cur = skip_next;
        }
        __syncthreads();                                   // every staged row has been waited for and read
        if (w1_tma && threadIdx.x == 0) { mbar_inval(&tbar[0]); mbar_inval(&tbar[1]); }
    }
    if (!group_sync(P, G)) return false;
    trace_mark(P, G, w, tn, 11, 0, t_win);
    // ---- W2..W4: outside rows + round 1 ----------------------------------------------------------------------------
    ObsTile& OT = *reinterpret_cast<ObsTile*>(keytab);             // shared scratch of the variable passes
    const int stiles = (D.M + kSuper - 1) / kSuper;
    if (D.packed) {
        w2_flat(P, D, ws, gt, gsz, S);
        if (D.H > 0) w2_pairs(P, D, ws, gt, gsz);
    } else {
        for (int t = G.cta; t < stiles; t += G.ncta) w2_vars_and_outside_counts(P, D, ws, t, OT, S);
    }
    if (!group_sync(P, G)) return false;
    trace_mark(P, G, w, tn, 12, 0, t_win);
    if (D.H > 0) {                            // (no outside rows: nothing to scan, one barrier less)
        if (G.cta == 0) w3_scan_outside(P, D, S);
        if (!group_sync(P, G)) return false;
    }
    trace_mark(P, G, w, tn, 13, 0, t_win);
    if (ws.error) return true;                // view failed validation: the slot stays unwritten (host keeps every point)
    if (D.packed) {
        if (D.H > 0) {
            w4_pairs(P, D, gt, gsz);
            if (!group_sync(P, G)) return false;
            trace_mark(P, G, w, tn, 15, 0, t_win);
        }
        w4_flat(P, D, ws, gt, gsz);
    } else {
        for (int t = G.cta; t < stiles; t += G.ncta) w4_fill_and_round1(P, D, ws, t, OT);
    }
    if (!group_sync(P, G)) return false;
    if (w == P.gwin[0] && G.cta == 0 && threadIdx.x == 0) P.ctrl->t_build = globaltimer_ns();

    // ---- phase machine (every CTA of the group takes the same decisions from the same counters) ----------------------
    const int drop_mode = (P.max_drop_rounds > 0) ? MODE_D1 : MODE_EVAL;
    int rounds = 1, greedy_steps = 0, drop_rounds = 0, status = 0;
    int seq = 0, mode;
    bool from_csr = true;
    bool in_lists = false;          // the rows' IN lists exist (written by the first sweep of the reverse delete / read-out)
    int vbuf = 0, vcnt = (int)ws.rc[0].nfree;           // FREE list written by the last PROP phase
    unsigned prev_sumdeg = ws.rc[0].sumdeg;             // its total number of row entries
    unsigned unc_final = 0, slack_final = 0;
    auto after_prop = [&](unsigned changed, unsigned nfree) {
        if (changed > 0 && rounds < P.max_rounds) {
            // stall: propagation still moves, but slowly (deficient rows with many candidates creep through the window
            // keyframe by keyframe) -> greedy step now.  Its row phase reads the live lists of the PROP row phase just done:
            // cell flags and row coverage are one variable phase behind, state and gain are current (oracle/emulate.py
            // restates exactly that); any selection is feasible, DROP removes what turns out redundant.
            if (P.stall_den > 0 && rounds >= 2 && (unsigned long long)changed * (unsigned)P.stall_den < nfree) return (int)MODE_GREEDY;
            return (int)MODE_PROP;
        }
        if (nfree == 0) return drop_mode;
        if (rounds >= P.max_rounds) { status = -5; return (int)MODE_FORCE; }
        return (int)MODE_GREEDY;
    };
    // dual bound: the snapshot is taken right before the first greedy step (mss_bound.cuh); 0 = not taken
    const BoundBufs BB{P.b_snap, P.b_snap_n, P.b_snap_d, P.b_share, P.b_red};
    const bool want_bound = P.b_snap != nullptr;
    int snap_flag = 0;
    auto bound_snapshot = [&](const int* vprev, int nv, bool csr) -> bool {
        bound_b0(P, BB, D, ws, G.cta, G.ncta, csr ? P.ent : P.live, csr ? P.ent_n : P.live_n, vprev, nv);
        if (!group_sync(P, G)) return false;
        bound_b2(P, BB, D, ws, G.cta, G.ncta, tab);
        if (!group_sync(P, G)) return false;
        bound_b3(P, BB, D, ws, G.cta, G.ncta, vprev, nv);
        if (!group_sync(P, G)) return false;
        bound_b4(P, BB, D, ws, G.cta, G.ncta, S);        // reads only the snapshot copies: no barrier needed behind it
        __syncthreads();
        snap_flag = 1;
        trace_mark(P, G, w, tn, 9, (unsigned)nv, t_win);
        return true;
    };
    mode = after_prop(ws.rc[0].changed, ws.rc[0].nfree);
    trace_mark(P, G, w, tn, 14, ws.rc[0].nfree, t_win);
    if (want_bound && mode == MODE_GREEDY && !bound_snapshot(P.vlist + D.var_base, vcnt, true)) return false;
    while (mode != MODE_DONE) {
        ++seq;
        RoundCnt& rc = ws.rc[seq % 3];
        if (G.cta == 0 && threadIdx.x < (int)(sizeof(RoundCnt) / 4))
            reinterpret_cast<uint32_t*>(&ws.rc[(seq + 1) % 3])[threadIdx.x] = 0u;     // used by phase seq + 1
        // row phase
        if (mode == MODE_PROP || mode == MODE_GREEDY || mode == MODE_D1 || mode == MODE_EVAL) {
            row_phase_lists(P, D, rc, G, mode, from_csr, in_lists, tab, keytab, S);
            if (!group_sync(P, G)) return false;
            if (mode == MODE_D1 || mode == MODE_EVAL) in_lists = true;
        } else if (mode == MODE_D2) {
            unsigned long long wsum = 0;
            for (int r = G.cta; r < rows; r += G.ncta) {
                const int R = D.row_base + r;
                wsum += (unsigned long long)P.live_n[R];
                row_d2(P, D, R, tab, keytab, S);
                __syncthreads();
            }
            if (threadIdx.x == 0 && wsum) atomicAdd(&P.ctrl->row_entries, wsum);
            if (!group_sync(P, G)) return false;
        }
        if (mode == MODE_PROP) from_csr = false;
        // variable phase
        if (G.cta == 0 && threadIdx.x == 0)
            atomicAdd(&P.ctrl->var_visits, (unsigned long long)((mode == MODE_PROP || mode == MODE_GREEDY || mode == MODE_FORCE) ? vcnt : D.M));
        if (mode == MODE_PROP || mode == MODE_GREEDY || mode == MODE_FORCE) {
            var_list_phase(P, D, ws, rc, mode, greedy_steps, P.vlist + (size_t)vbuf * P.Mpad + D.var_base, vcnt,
                           P.vlist + (size_t)(vbuf ^ 1) * P.Mpad + D.var_base, G);
        } else {
            var_flat_phase(P, D, ws, rc, mode, gt, gsz);
        }
        if (!group_sync(P, G)) return false;
        trace_mark(P, G, w, tn, mode, rc.nfree, t_win);
        // transition
        switch (mode) {
        case MODE_PROP: {
            // the live lists written by this round's row phase are exact for the state BEFORE its variable phase, whose
            // FREE list is the source list of that variable phase
            const int src_buf = vbuf, src_cnt = vcnt;
            const unsigned src_deg = prev_sumdeg;
            ++rounds;
            vbuf ^= 1;
            vcnt = (int)rc.nfree;
            prev_sumdeg = rc.sumdeg;
            mode = after_prop(rc.changed, rc.nfree);
            if (want_bound && snap_flag == 0 && greedy_steps == 0 && mode == MODE_GREEDY &&
                !bound_snapshot(P.vlist + (size_t)src_buf * P.Mpad + D.var_base, src_cnt, false)) return false;
            // (with the dual bound on, the shared-memory tail -- which may take greedy steps of its own -- starts only after
            // the snapshot; it is result-neutral, so the selection is the same either way)
            if ((mode == MODE_PROP || mode == MODE_GREEDY) && (!want_bound || snap_flag != 0 || greedy_steps > 0) &&
                src_cnt <= P.tail_vars && src_deg <= (unsigned)P.tail_ents &&
                rc.rows_live <= (unsigned)kTailRows && rc.maxlive <= (unsigned)kTailRowMax) {
                if (G.cta == 0)
                    tail_solve(P, D, ws, T, S, P.vlist + (size_t)src_buf * P.Mpad + D.var_base, src_cnt, mode, drop_mode, rounds,
                               greedy_steps, status, w, tn, t_win);
                if (!group_sync(P, G)) return false;
                rounds = ws.t_rounds; greedy_steps = ws.t_greedy; status = ws.t_status;
                mode = drop_mode;
                trace_mark(P, G, w, tn, 8, (unsigned)rounds, t_win);
            }
        } break;
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
            if (snap_flag == 1) {
                bound_final(P, BB, D, ws, G.cta, G.ncta, tab);
                if (!group_sync(P, G)) return false;
            }
            if (G.cta == 0 && threadIdx.x == 0) {
                uint32_t* hdr = P.out + D.out_off;
                // dual bound: 1 = snapshot counters below, 2 = propagation alone decided everything (the selection is optimal),
                // 0 = none (not asked for, or the round cap forced the rest in)
                hdr[16] = !want_bound || status != 0 ? 0u : (snap_flag == 1 ? 1u : (greedy_steps == 0 ? 2u : 0u));
                hdr[17] = ws.b_s0;
                hdr[18] = ws.b_res_unc;
                hdr[19] = 0u;
                hdr[20] = (uint32_t)(ws.b_cost_in & 0xFFFFFFFFull); hdr[21] = (uint32_t)(ws.b_cost_in >> 32);
                hdr[22] = (uint32_t)(ws.b_zsum & 0xFFFFFFFFull);    hdr[23] = (uint32_t)(ws.b_zsum >> 32);
                hdr[24] = (uint32_t)(ws.b_drows & 0xFFFFFFFFull);   hdr[25] = (uint32_t)(ws.b_drows >> 32);
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
                if (P.trace) P.trace[(size_t)w * kTraceCap + kTraceCap - 1] = make_uint2((unsigned)tn, (unsigned)(globaltimer_ns() - t_win));
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
constexpr size_t kRowSmem = (size_t)kCells * 12;                                   // tab (u32) + keytab (u64)
constexpr size_t kUnionSmem = ((kRowSmem > sizeof(TailSmem) ? kRowSmem : sizeof(TailSmem)) + 15) & ~(size_t)15;
constexpr size_t kSmemBytes = kUnionSmem + ((sizeof(BlockScratch) + 15) & ~(size_t)15) + 16;

__global__ void __launch_bounds__(kThreads, 4) mss_persistent_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* tab = reinterpret_cast<unsigned*>(smem_raw);
    unsigned long long* keytab = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)kCells * 4);
    TailSmem& T = *reinterpret_cast<TailSmem*>(smem_raw);                        // aliases tab / keytab
    BlockScratch& S = *reinterpret_cast<BlockScratch*>(smem_raw + kUnionSmem);
    int& s_abort = *reinterpret_cast<int*>(smem_raw + kSmemBytes - 16);

    if (blockIdx.x == 0 && threadIdx.x == 0) P.ctrl->t_start = globaltimer_ns();
    const int gi = P.cta_grp[blockIdx.x];
    const GroupDesc gd = P.grp[gi];
    GroupCtx G;
    G.bar = P.gbar + (size_t)gi * 32;
    G.gen = 0u;
    G.ncta = gd.ncta;
    G.cta = (int)blockIdx.x - gd.cta0;
    G.s_abort = &s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    __syncthreads();
    // dynamic window queue: a group that finishes early takes the next window (largest windows are queued first)
    while (true) {
        if (G.cta == 0 && threadIdx.x == 0) {
            const unsigned q = atomicAdd(&P.ctrl->queue, 1u);
            if (P.ready && q < (unsigned)P.nwin) {
                // host views are copied while the kernel runs: wait for this window's flag (written by a copy that is
                // stream-ordered after the copies of its arrays); the group barrier publishes it to the other CTAs
                unsigned spins = 0;
                unsigned long long t_wait = 0ull;
                while (ld_acquire_sys_u32(P.ready + q) == 0u) {
                    __nanosleep(200);
                    if ((++spins & 0x3FFu) == 0u) {
                        if (*(volatile int*)&P.ctrl->abort) break;
                        if (t_wait == 0ull) t_wait = globaltimer_ns();
                        else if (globaltimer_ns() - t_wait > P.watchdog_ns) { atomicExch(&P.ctrl->abort, 1); break; }
                    }
                }
            }
            G.bar[1] = q;
        }
        if (!group_sync(P, G)) break;
        const unsigned wi = *(volatile unsigned*)&G.bar[1];
        if (wi >= (unsigned)P.nwin) break;
        if (!solve_window(P, G, P.gwin[wi], tab, keytab, T, S)) break;
    }
    if (G.cta == 0 && threadIdx.x == 0) atomicMax(&P.ctrl->t_end, globaltimer_ns());
}

#endif  // MSS_KERNELS_TYPES_ONLY

}  // namespace mss

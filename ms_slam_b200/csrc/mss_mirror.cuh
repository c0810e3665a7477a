// mss_mirror.cuh -- device side of the persistent mirror of the keyframe x map-point incidence (include/mss.h
// "Persistent device mirror").
//
// What it replaces in the reference: the three pointer-chasing passes of MapSparsification::Sparsifying
// (/root/reference/src/MapSparsification.cc:67-151: a copy of mvpMapPoints and of mGrid per keyframe, isBad() and
// Observations() per slot, a copy of mObservations per variable, two std::map keyed by keyframe pointer) and the host-side
// FlattenWindow that round 1 put in their place.  The incidence lives in HBM, keyframe-major:
//     slot_mp[kf][i]    KeyFrame::mvpMapPoints[i]            (handle or -1)
//     slot_cell[kf][i]  cell of keypoint i in KeyFrame::mGrid (col*48+row, 0xFFFF = not in the grid)
//     obs_mp[kf][i]     the map point whose mObservations holds (kf -> i), or -1
//     mp_nobs[h], mp_bad[h]   MapPoint::nObs, MapPoint::mbBad
// and a window is assembled from K keyframe handles by a handful of streaming kernels:
//     mark     window keyframes get their window id
//     first    every valid slot: atomicMin of its position (k*S + i) on its map point -> first occurrence; owner; is-variable
//     count    per keyframe: first occurrences (= map points it discovers), valid slots; handle range, observer range
//     scan     per window: exclusive scans -> feat_ptr, first table index of every keyframe (discovery order,
//              mnIndexForSparsification of MapSparsification.cc:91-99)
//     obs      the observations of the keyframes in the observer range are streamed once to find the outside keyframes
//              (MapSparsification.cc:125-142) and a second time to emit the (map point, outside keyframe) pairs
//     number / slots / okf   table index of every map point, packed slots, outside keyframes ordered by sort key, GetNumberMPs
// The result is an MSS_LAYOUT_PACKED view in device memory, solved in place by the persistent kernel.  Everything here is
// coalesced streaming over the keyframe-major arrays plus one spread-address access per valid slot (the per-map-point
// scratch): HBM-bound integer work, no shared-memory staging needed beyond the block scans.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mssm {

constexpr int kT = 256;
constexpr int kMaxOutside = 4095;              // MSS_LAYOUT_PACKED: 12-bit outside-keyframe index
constexpr unsigned kCellNone16 = 0xFFFFu;
enum : int { C_M = 0, C_F, C_H, C_O, C_HLO, C_HHI, C_KFLO, C_KFHI, C_ERR, C_PAIRCUR, C_NDEL, C_NMAXOBS, C_COUNT = 16 };
enum : int { ME_KF_RANGE = 1, ME_KF_TWICE = 2, ME_MP_RANGE = 4, ME_MP_SHARED = 8, ME_DEPENDENT = 16, ME_OUTSIDE_OVERFLOW = 32,
             ME_NOBS_RANGE = 64 };

// Everything the kernels read or write per map point sits in ONE 32-byte record (one sector): the window assembly touches
// a map point through a handle found in a slot, i.e. at a spread address, several times per pass.
struct __align__(32) MpRec {
    int nobs;                // MapPoint::nObs
    int obs_lo, obs_hi;      // lowest / highest keyframe handle that ever observed the point (never shrinks: a superset filter
                             // for the observation scan; the scan itself is exact); INT_MAX / -1 when none
    int loc;                 // scratch: rank among the points its first keyframe discovers (table index = that keyframe's
                             // first index + rank); -1 when idle
    unsigned long long fo;   // scratch: (window id + 1) << 32 | position k*S + i of the first occurrence, one atomicMin per
                             // valid slot; ~0 when idle.  Two windows meeting at one point: the loser sees a foreign id
    uint8_t bad;             // MapPoint::mbBad
    uint8_t isvar;           // scratch
    uint8_t pad_[6];
};
constexpr unsigned long long kFoIdle = ~0ull;
__device__ __forceinline__ unsigned long long fo_key(int w, int pos) { return ((unsigned long long)(unsigned)(w + 1) << 32) | (unsigned)pos; }

struct MirrorDev {
    int* slot_mp;            // [kf_cap * S]
    int* obs_mp;             // [kf_cap * S]
    uint16_t* slot_cell;     // [kf_cap * S]
    int* kf_n;               // [kf_cap] slots in use
    unsigned* kf_key;        // [kf_cap] sort key of the outside rows (KeyFrame::mnId)
    int* kf_win;             // [kf_cap] scratch: window id + 1 while a call is in flight, else 0
    uint8_t* okf_mark;       // [kf_cap] scratch
    int* okf_idx;            // [kf_cap] scratch: index of an outside keyframe in its window's table
    MpRec* mp;               // [mp_cap]
    int S, n_kf, n_mp;
};

struct MWin {
    const int* kf;           // [K] window keyframe handles (device)
    int K, w, n_max_floor, apply;
    int* kf_first;           // [K + 1] first occurrences per keyframe -> exclusive scan (first table index)
    int* feat_ptr;           // [K + 1] valid slots per keyframe -> exclusive scan
    uint32_t* slots;         // [F]
    uint16_t* nobs16;        // [M]
    int* mp_handle;          // [M]
    uint32_t* pairs;         // [O]
    int* okf_list;           // [4096] outside keyframes as collected, then [4096] ordered by sort key
    int* okf_total;          // [H]
    int* cnt;                // [C_COUNT]
    const uint32_t* keep;    // keep bits of the solve (result slot), set before mk_deleted
    unsigned* del;           // bits of the dropped variables over this window's handle range: bit (h - 32 * (h_lo / 32))
};

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += y;
    }
    return v;
}
// exclusive scan over the block's threads; total to everybody.  s_w: 8 ints of shared scratch, reusable after the call
__device__ __forceinline__ int block_excl(int v, int& total, int* s_w) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v);
    __syncthreads();
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < kT / 32; ++q) { const int t = s_w[q]; if (q < wid) off += t; tot += t; }
    total = tot;
    return off + inc - v;
}

// ---- maintenance ----------------------------------------------------------------------------------------------------
__global__ void mk_fill_i32(int* p, int v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void mk_init_mp(MpRec* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        MpRec r;
        r.nobs = 0; r.obs_lo = 0x7FFFFFFF; r.obs_hi = -1; r.loc = -1; r.fo = kFoIdle; r.bad = 0; r.isvar = 0;
        for (int k = 0; k < 6; ++k) r.pad_[k] = 0;
        p[i] = r;
    }
}
__global__ void mk_set_mp(MpRec* p, const int* nobs, const uint8_t* bad, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { p[i].nobs = nobs[i]; p[i].bad = bad ? bad[i] : (uint8_t)0; }
}

// observer ranges of the map points named by the obs_mp entries of keyframes [kf0, kf0 + n)
__global__ void mk_obs_ranges(MirrorDev D, int kf0, int n) {
    const size_t total = (size_t)n * D.S;
    for (size_t it = blockIdx.x * (size_t)blockDim.x + threadIdx.x; it < total; it += (size_t)gridDim.x * blockDim.x) {
        const int kf = kf0 + (int)(it / D.S);
        const int h = D.obs_mp[(size_t)kf0 * D.S + it];
        if (h >= 0 && h < D.n_mp) { atomicMin(&D.mp[h].obs_lo, kf); atomicMax(&D.mp[h].obs_hi, kf); }
    }
}

struct DevOp { int kind, a, b, c; };
// resolved stores (the host has removed duplicates: one op per address)
__global__ void mk_apply_ops(MirrorDev D, const DevOp* ops, int n, int* err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevOp o = ops[i];
    if (o.kind == 1 || o.kind == 2) {
        if (o.a < 0 || o.a >= D.n_kf || o.b < 0 || o.b >= D.S || o.c < -1 || o.c >= D.n_mp) { atomicOr(err, 1); return; }
        const size_t pos = (size_t)o.a * D.S + o.b;
        if (o.kind == 1) {
            D.slot_mp[pos] = o.c;
            if (o.b >= D.kf_n[o.a]) atomicMax(&D.kf_n[o.a], o.b + 1);
        } else {
            D.obs_mp[pos] = o.c;
            if (o.c >= 0) { atomicMin(&D.mp[o.c].obs_lo, o.a); atomicMax(&D.mp[o.c].obs_hi, o.a); }
        }
    } else if (o.kind == 3) {
        if (o.a < 0 || o.a >= D.n_mp) { atomicOr(err, 1); return; }
        D.mp[o.a].nobs = o.b;
        D.mp[o.a].bad = o.c ? 1 : 0;
    }
}

// KeyFrame::EraseBadDescriptor (src/KeyFrame.cc:311-361): keep the non-empty slots, in order; every kept point observes
// the keyframe at its new index (MapPoint::UpdateObservation); the grid is dropped (cells -> none).  One CTA.
__global__ void mk_kf_compact(MirrorDev D, int kf) {
    __shared__ int s_w[kT / 32];
    __shared__ int s_mp[2048];
    const size_t base = (size_t)kf * D.S;
    const int n = D.kf_n[kf];
    int out = 0;
    for (int c0 = 0; c0 < n; c0 += 2048) {          // S may exceed 2048: chunks of 2048 staged in shared memory
        const int m = min(2048, n - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += kT) s_mp[i] = D.slot_mp[base + c0 + i];
        __syncthreads();
        for (int b0 = 0; b0 < m; b0 += kT) {
            const int i = b0 + (int)threadIdx.x;
            const int h = i < m ? s_mp[i] : -1;
            int tot;
            const int p = block_excl(h >= 0 ? 1 : 0, tot, s_w);
            if (h >= 0) {                               // out + p <= c0 + i: never overtakes the part still to be read
                D.slot_mp[base + out + p] = h;
                D.obs_mp[base + out + p] = h;
                if (h < D.n_mp) { atomicMin(&D.mp[h].obs_lo, kf); atomicMax(&D.mp[h].obs_hi, kf); }
            }
            out += tot;
        }
    }
    __syncthreads();
    for (int i = out + (int)threadIdx.x; i < n; i += kT) { D.slot_mp[base + i] = -1; D.obs_mp[base + i] = -1; }
    for (int i = threadIdx.x; i < n; i += kT) D.slot_cell[base + i] = (uint16_t)kCellNone16;
    if (threadIdx.x == 0) D.kf_n[kf] = out;
}

// ---- window assembly --------------------------------------------------------------------------------------------------
__global__ void mk_mark(MirrorDev D, const MWin* W) {
    const MWin w = W[blockIdx.y];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= w.K) return;
    const int kf = w.kf[k];
    if (kf < 0 || kf >= D.n_kf) { atomicOr(&w.cnt[C_ERR], ME_KF_RANGE); return; }
    if (atomicCAS(&D.kf_win[kf], 0, w.w + 1) != 0) atomicOr(&w.cnt[C_ERR], ME_KF_TWICE);
}

// one CTA per window keyframe: first occurrence of every map point, owner, is-variable, valid-slot count
__global__ void mk_first(MirrorDev D, const MWin* W) {
    __shared__ int s_w[kT / 32];
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR]) return;
    for (int k = blockIdx.x; k < w.K; k += gridDim.x) {
        const int kf = w.kf[k];
        const size_t base = (size_t)kf * D.S;
        const int n = D.kf_n[kf];
        int nvalid = 0;
        unsigned err = 0;
        for (int i = threadIdx.x; i < n; i += kT) {
            const int h = D.slot_mp[base + i];
            if (h < 0) continue;
            if (h >= D.n_mp) { err |= ME_MP_RANGE; continue; }
            const MpRec r = D.mp[h];                                        // one sector: bad, current first occurrence, isvar
            if (r.bad) continue;                                            // MapSparsification.cc:70,90
            ++nvalid;
            // keyframes are scheduled roughly in window order, so most later occurrences find a smaller key in place and
            // need no atomic (a stale read only costs a redundant atomicMin)
            const unsigned long long key = fo_key(w.w, k * D.S + i);
            if (key < r.fo) atomicMin(&D.mp[h].fo, key);
            if (!r.isvar && D.slot_cell[base + i] != (uint16_t)kCellNone16) D.mp[h].isvar = 1;
        }
        int tot;
        block_excl(nvalid, tot, s_w);
        if (threadIdx.x == 0) w.feat_ptr[k] = tot;
        if (err) atomicOr(&w.cnt[C_ERR], (int)err);
    }
}

// one CTA per window keyframe: the map points it discovers and their rank among them (slot order); handle range; observer
// range of the variables; a point first met by another window of the call = the windows are not independent
__global__ void mk_count(MirrorDev D, const MWin* W) {
    __shared__ int s_w[kT / 32];
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR]) return;
    for (int k = blockIdx.x; k < w.K; k += gridDim.x) {
        const int kf = w.kf[k];
        const size_t base = (size_t)kf * D.S;
        const int n = D.kf_n[kf];
        int run = 0, hlo = 0x7FFFFFFF, hhi = -1, klo = 0x7FFFFFFF, khi = -1, nmax_obs = 0;
        unsigned err = 0;
        for (int b0 = 0; b0 < n; b0 += kT) {
            const int i = b0 + (int)threadIdx.x;
            int h = -1;
            MpRec r;
            if (i < n) {
                h = D.slot_mp[base + i];
                if (h >= 0 && h < D.n_mp) {
                    r = D.mp[h];
                    if (r.bad) h = -1;
                    else if ((int)(r.fo >> 32) != w.w + 1) { err |= ME_MP_SHARED; h = -1; }
                    else if ((unsigned)r.fo != (unsigned)(k * D.S + i)) h = -1;
                } else h = -1;
            }
            int tot;
            const int p = block_excl(h >= 0 ? 1 : 0, tot, s_w);
            if (h >= 0) {
                D.mp[h].loc = run + p;
                hlo = min(hlo, h); hhi = max(hhi, h);
                nmax_obs = max(nmax_obs, r.nobs);
                if (r.isvar) { klo = min(klo, r.obs_lo); khi = max(khi, r.obs_hi); }
            }
            run += tot;
        }
        hlo = __reduce_min_sync(0xFFFFFFFFu, hlo); hhi = __reduce_max_sync(0xFFFFFFFFu, hhi);
        klo = __reduce_min_sync(0xFFFFFFFFu, klo); khi = __reduce_max_sync(0xFFFFFFFFu, khi);
        nmax_obs = __reduce_max_sync(0xFFFFFFFFu, nmax_obs);
        if ((threadIdx.x & 31) == 0) {
            if (hhi >= 0) { atomicMin(&w.cnt[C_HLO], hlo); atomicMax(&w.cnt[C_HHI], hhi + 1); }
            if (khi >= 0) { atomicMin(&w.cnt[C_KFLO], klo); atomicMax(&w.cnt[C_KFHI], khi); }
            if (nmax_obs > 65535) atomicOr(&w.cnt[C_ERR], ME_NOBS_RANGE);
        }
        if (err) atomicOr(&w.cnt[C_ERR], (int)err);
        if (threadIdx.x == 0) w.kf_first[k] = run;
    }
}

// one CTA per window: exclusive scans of the per-keyframe counts
__global__ void mk_scan(const MWin* W) {
    __shared__ int s_w[kT / 32];
    const MWin w = W[blockIdx.x];
    if (w.cnt[C_ERR]) return;
    int cm = 0, cf = 0;
    for (int b0 = 0; b0 <= w.K; b0 += kT) {
        const int k = b0 + (int)threadIdx.x;
        const int a = k < w.K ? w.kf_first[k] : 0, b = k < w.K ? w.feat_ptr[k] : 0;
        int ta, tb;
        const int ea = block_excl(a, ta, s_w);
        const int eb = block_excl(b, tb, s_w);
        if (k <= w.K) { w.kf_first[k] = cm + ea; w.feat_ptr[k] = cf + eb; }
        cm += ta; cf += tb;
    }
    if (threadIdx.x == 0) { w.cnt[C_M] = cm; w.cnt[C_F] = cf; }
}

// Observation scan over the keyframes of the observer range, one warp per keyframe at a time.  PASS 0 counts the pairs and
// marks the outside keyframes, PASS 1 emits the pairs; both skip the window's own keyframes before touching their
// observations (mnMapSaprsificationId == mnId, MapSparsification.cc:132).  PASS 2 applies the deletion (SetBadFlag,
// src/MapPoint.cc:227-255: every (kf, idx) of the point's observations is erased from the keyframe and the observation
// itself is dropped) and therefore visits every keyframe of the range.
template <int PASS>
__global__ void mk_obs_scan(MirrorDev D, const MWin* W) {
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR] || (PASS == 2 && !w.apply)) return;
    const int klo = w.cnt[C_KFLO], khi = w.cnt[C_KFHI];
    if (khi < klo) return;
    const int me = w.w + 1;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int hbase = w.cnt[C_HLO] & ~31;
    int npairs = 0;
    for (int kf = klo + gw; kf <= khi; kf += nw) {
        const int kw = D.kf_win[kf];
        if (PASS != 2 && kw == me) continue;
        const size_t base = (size_t)kf * D.S;
        const int n = D.kf_n[kf];
        bool any = false;
        for (int i = lane; i < n; i += 32) {
            const int h = D.obs_mp[base + i];
            if (h < 0 || h >= D.n_mp) continue;
            const MpRec r = D.mp[h];
            if ((int)(r.fo >> 32) != me) continue;
            if (PASS == 2) {
                const int b = h - hbase;
                if ((w.del[b >> 5] >> (b & 31)) & 1u) { D.obs_mp[base + i] = -1; D.slot_mp[base + i] = -1; }
                continue;
            }
            if (!r.isvar) continue;                         // MapSparsification.cc:127-142 walks the variables only
            if (kw != 0) { atomicOr(&w.cnt[C_ERR], ME_DEPENDENT); continue; }
            if (PASS == 0) { any = true; ++npairs; }
            else w.pairs[atomicAdd(&w.cnt[C_PAIRCUR], 1)] = ((uint32_t)(w.kf_first[(unsigned)r.fo / (unsigned)D.S] + r.loc) << 12) | (uint32_t)D.okf_idx[kf];
        }
        if (PASS == 0 && any) D.okf_mark[kf] = 1;
    }
    if (PASS == 0) {
        npairs = __reduce_add_sync(0xFFFFFFFFu, npairs);
        if (lane == 0 && npairs) atomicAdd(&w.cnt[C_O], npairs);
    }
}

__global__ void mk_okf_collect(MirrorDev D, const MWin* W) {
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR]) return;
    const int klo = w.cnt[C_KFLO], khi = w.cnt[C_KFHI];
    if (khi < klo) return;
    for (int kf = klo + blockIdx.x * blockDim.x + threadIdx.x; kf <= khi; kf += gridDim.x * blockDim.x) {
        if (!D.okf_mark[kf]) continue;
        const int p = atomicAdd(&w.cnt[C_H], 1);
        if (p < kMaxOutside + 1) w.okf_list[p] = kf;
    }
}

// one CTA per window: order the outside keyframes by (sort key, handle) -- FlattenWindow orders them by KeyFrame::mnId
__global__ void mk_okf_rank(MirrorDev D, const MWin* W) {
    const MWin w = W[blockIdx.x];
    if (w.cnt[C_ERR]) return;
    const int H = w.cnt[C_H];
    if (H > kMaxOutside) { if (threadIdx.x == 0) atomicOr(&w.cnt[C_ERR], ME_OUTSIDE_OVERFLOW); return; }
    int* sorted = w.okf_list + (kMaxOutside + 1);
    for (int j = threadIdx.x; j < H; j += kT) {
        const int kf = w.okf_list[j];
        const unsigned key = D.kf_key[kf];
        int r = 0;
        for (int i = 0; i < H; ++i) {
            const int kf2 = w.okf_list[i];
            const unsigned k2 = D.kf_key[kf2];
            r += (k2 < key || (k2 == key && kf2 < kf)) ? 1 : 0;
        }
        sorted[r] = kf;
        D.okf_idx[kf] = r;
    }
}

// one CTA per outside keyframe: KeyFrame::GetNumberMPs (src/KeyFrame.cc:286-297)
__global__ void mk_okf_total(MirrorDev D, const MWin* W) {
    __shared__ int s_w[kT / 32];
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR]) return;
    const int H = w.cnt[C_H];
    const int* sorted = w.okf_list + (kMaxOutside + 1);
    for (int j = blockIdx.x; j < H; j += gridDim.x) {
        const int kf = sorted[j];
        const size_t base = (size_t)kf * D.S;
        const int n = D.kf_n[kf];
        int c = 0;
        for (int i = threadIdx.x; i < n; i += kT) {
            const int h = D.slot_mp[base + i];
            c += (h >= 0 && h < D.n_mp && !D.mp[h].bad) ? 1 : 0;
        }
        int tot;
        block_excl(c, tot, s_w);
        if (threadIdx.x == 0) w.okf_total[j] = tot;
    }
}

// one CTA per window keyframe: its valid slots, in slot order, as (table index << 12) | cell; the table index of a point =
// first index of the keyframe that discovers it + its rank there (mk_count); that keyframe also writes the point's row of
// the handle and nObs tables (discovery order)
__global__ void mk_slots(MirrorDev D, const MWin* W) {
    __shared__ int s_w[kT / 32];
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR]) return;
    for (int k = blockIdx.x; k < w.K; k += gridDim.x) {
        const int kf = w.kf[k];
        const size_t base = (size_t)kf * D.S;
        const int n = D.kf_n[kf];
        int run = w.feat_ptr[k];
        for (int b0 = 0; b0 < n; b0 += kT) {
            const int i = b0 + (int)threadIdx.x;
            int h = -1, idx = 0;
            if (i < n) {
                h = D.slot_mp[base + i];
                if (h >= 0 && h < D.n_mp) {
                    const MpRec r = D.mp[h];
                    if (r.bad) h = -1;
                    else {
                        idx = w.kf_first[(unsigned)r.fo / (unsigned)D.S] + r.loc;
                        if ((unsigned)r.fo == (unsigned)(k * D.S + i)) {
                            w.mp_handle[idx] = h;
                            w.nobs16[idx] = (uint16_t)min(max(r.nobs, 0), 65535);
                        }
                    }
                } else h = -1;
            }
            int tot;
            const int p = block_excl(h >= 0 ? 1 : 0, tot, s_w);
            if (h >= 0) {
                const unsigned c = D.slot_cell[base + i];
                w.slots[run + p] = ((uint32_t)idx << 12) | (c == kCellNone16 ? 0xFFFu : c);
            }
            run += tot;
        }
    }
}

// bits of the dropped variables over the handle space; optionally the points become bad in the mirror
__global__ void mk_del_zero(const MWin* W) {
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR] || !w.del) return;
    const int n = ((w.cnt[C_HHI] + 31) >> 5) - (w.cnt[C_HLO] >> 5);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w.del[i] = 0u;
}
__global__ void mk_deleted(MirrorDev D, const MWin* W) {
    const MWin w = W[blockIdx.y];
    if (w.cnt[C_ERR] || !w.keep) return;
    const int M = w.cnt[C_M];
    int nd = 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < M; idx += gridDim.x * blockDim.x) {
        if ((w.keep[idx >> 5] >> (idx & 31)) & 1u) continue;
        const int h = w.mp_handle[idx];
        const int b = h - (w.cnt[C_HLO] & ~31);
        atomicOr(&w.del[b >> 5], 1u << (b & 31));
        if (w.apply) D.mp[h].bad = 1;
        ++nd;
    }
    nd = __reduce_add_sync(0xFFFFFFFFu, nd);
    if ((threadIdx.x & 31) == 0 && nd) atomicAdd(&w.cnt[C_NDEL], nd);
}

// leave the per-handle scratch idle again.  stage 0: map points (needs mp_handle: only after mk_number), window keyframes;
// stage 1: outside marks
__global__ void mk_reset(MirrorDev D, const MWin* W, int have_tables) {
    const MWin w = W[blockIdx.y];
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
    if (have_tables && !w.cnt[C_ERR]) {
        const int M = w.cnt[C_M];
        for (int idx = gt; idx < M; idx += gs) {
            const int h = w.mp_handle[idx];
            D.mp[h].fo = kFoIdle; D.mp[h].loc = -1; D.mp[h].isvar = 0;
        }
    } else {
        // no tables (error before the numbering pass): walk the window's slots again
        for (int k = 0; k < w.K; ++k) {
            const int kf = w.kf[k];
            if (kf < 0 || kf >= D.n_kf) continue;
            const size_t base = (size_t)kf * D.S;
            const int n = D.kf_n[kf];
            for (int i = gt; i < n; i += gs) {
                const int h = D.slot_mp[base + i];
                if (h >= 0 && h < D.n_mp && (int)(D.mp[h].fo >> 32) == w.w + 1) { D.mp[h].fo = kFoIdle; D.mp[h].loc = -1; D.mp[h].isvar = 0; }
            }
        }
    }
    for (int k = gt; k < w.K; k += gs) {
        const int kf = w.kf[k];
        if (kf >= 0 && kf < D.n_kf && D.kf_win[kf] == w.w + 1) D.kf_win[kf] = 0;
    }
    const int klo = w.cnt[C_KFLO], khi = w.cnt[C_KFHI];          // (INT_MAX, -1) when no variable was seen: nothing to clear
    if (khi >= klo)
        for (int kf = klo + gt; kf <= khi; kf += gs) D.okf_mark[kf] = 0;
}

}  // namespace mssm

// mss_components.cuh -- connected components of one flattened window (sm_100a).
//
// Why: the reference's final flush hands EVERY not-yet-sparsified keyframe to one ILP
// (/root/reference/src/MapSparsification.cc:38-47).  That model decomposes exactly along the connected components of the
// bipartite graph {keyframe rows (window + outside)} x {variables}: two variables interact only through a row they share
// (a keyframe row :119-122, one of its cell rows :111-116, or an outside-keyframe row :146-150), so every component is an
// independent window once it is given the window-wide nMax (:66-76) -- which is what lets one flush shard over several
// GPUs (BASELINE.json north_star).
//
// Nodes: rows 0..R-1 (R = K + H), then variables R..R+M-1.  Edges: every valid grid-listed slot (k, p) and every outside
// observation (K + j, p) of a variable p.  Lock-free union-find: roots are hooked larger-under-smaller with atomicMin-style
// CAS, so the final root of a component is its smallest node id = its first keyframe row; the result does not depend on
// scheduling.  Labels are then made dense in order of that first row.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mss_kernels.cuh"

namespace mss {

__device__ __forceinline__ int cc_find(int* parent, int v) {
    // path halving; parent[] only ever decreases, so concurrent readers stay inside the same tree
    int p = parent[v];
    while (p != v) {
        const int g = parent[p];
        if (g != p) parent[v] = g;
        v = p;
        p = g;
    }
    return v;
}

__device__ __forceinline__ void cc_union(int* parent, int a, int b) {
    while (true) {
        a = cc_find(parent, a);
        b = cc_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          // hook the larger root (a) under the smaller (b)
        const int old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
        a = old;                                               // somebody else hooked a meanwhile: retry from there
    }
}

__global__ void cc_init_kernel(int* parent, uint8_t* isvar, int R, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R + M) parent[i] = i;
    if (i < M) isvar[i] = 0;
}

// one CTA per keyframe row (grid-stride): union (k, p) for every valid grid-listed slot; marks the variables
__global__ void cc_slots_kernel(const WinDesc D, int* parent, uint8_t* isvar, unsigned* err, int* n_max) {
    const int R = D.K + D.H;
    int nmax = 0;
    for (int k = blockIdx.x; k < D.K; k += gridDim.x) {
        const int beg = ldv(D.feat_ptr + k), end = ldv(D.feat_ptr + k + 1);
        if (beg < 0 || end < beg || end > D.F) { if (threadIdx.x == 0) atomicOr(err, ERR_PTR); continue; }
        if (D.packed == 2) {
            // 16-bit tokens: the map-point index is a running sum over the keyframe's tokens (block scan per chunk)
            __shared__ int s_w[32];
            __shared__ int s_carry;
            const uint16_t* tk = reinterpret_cast<const uint16_t*>(D.feat_mp);
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
            __syncthreads();
            if (threadIdx.x == 0) s_carry = 0;
            __syncthreads();
            for (int base = beg; base < end; base += blockDim.x) {
                const int i = base + (int)threadIdx.x;
                int adv = 0;
                bool slot = false;
                unsigned c = kCellNone;
                if (i < end) adv = tok_decode(ldv(tk + i), slot, c);
                int x = adv;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
                    if (lane >= o) x += y;
                }
                if (lane == 31) s_w[wid] = x;
                __syncthreads();
                int woff = 0, tot = 0;
                for (int q = 0; q < nw; ++q) { const int t = s_w[q]; if (q < wid) woff += t; tot += t; }
                const int mp = s_carry + woff + x;
                __syncthreads();
                if (threadIdx.x == 0) s_carry += tot;
                __syncthreads();
                if (!slot) continue;
                if ((unsigned)mp >= (unsigned)D.M) { atomicOr(err, ERR_INDEX); continue; }     // (a wrapped running index is negative)
                nmax = max(nmax, ld_nobs(D, mp));
                if (c == kCellNone) continue;
                if (c >= (unsigned)kCells) { atomicOr(err, ERR_INDEX); continue; }
                isvar[mp] = 1;
                cc_union(parent, k, R + mp);
            }
            continue;
        }
        for (int i = beg + (int)threadIdx.x; i < end; i += blockDim.x) {
            int mp;
            unsigned c;
            ld_slot(D, i, mp, c);
            if (mp < 0) { if (mp < -1) atomicOr(err, ERR_INDEX); continue; }
            if (mp >= D.M) { atomicOr(err, ERR_INDEX); continue; }
            nmax = max(nmax, ld_nobs(D, mp));                   // every valid slot counts for nMax (:66-76)
            if (c == kCellNone) continue;                       // not in mGrid: no variable through this slot (:84-107)
            if (c >= (unsigned)kCells) { atomicOr(err, ERR_INDEX); continue; }
            isvar[mp] = 1;
            cc_union(parent, k, R + mp);
        }
    }
    nmax = __reduce_max_sync(0xFFFFFFFFu, nmax);
    if ((threadIdx.x & 31) == 0 && nmax > 0) atomicMax(n_max, nmax);
}

// packed layout: the outside observations are a flat pair list
__global__ void cc_pairs_kernel(const WinDesc D, int* parent, const uint8_t* isvar, unsigned* err) {
    const int R = D.K + D.H;
    const uint32_t* pairs = reinterpret_cast<const uint32_t*>(D.mp_obs_kf);
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < D.O; o += gridDim.x * blockDim.x) {
        const uint32_t pr = ldv(pairs + o);
        const int mp = (int)(pr >> kCellBits), j = (int)(pr & kCellCov);
        if (mp >= D.M || j >= D.H) { atomicOr(err, ERR_INDEX); continue; }
        if (isvar[mp]) cc_union(parent, D.K + j, R + mp);
    }
}

// one thread per map point: its outside observations (only variables have outside rows, :127-142)
__global__ void cc_outside_kernel(const WinDesc D, int* parent, const uint8_t* isvar, unsigned* err) {
    const int R = D.K + D.H;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= D.M || !isvar[p]) return;
    const int beg = ldv(D.mp_obs_ptr + p), end = ldv(D.mp_obs_ptr + p + 1);
    if (beg < 0 || end < beg || end > D.O) { atomicOr(err, ERR_PTR); return; }
    for (int o = beg; o < end; ++o) {
        const int kf = ld_obs_kf(D, o);
        if (kf < 0 || kf >= R) { atomicOr(err, ERR_INDEX); continue; }
        if (kf < D.K) continue;                                 // window keyframes are covered by their slots
        cc_union(parent, kf, R + p);
    }
}

// single CTA: dense component ids in order of the first row of each component; rows that hold no variable at all (empty
// keyframes, outside keyframes that see no variable) are components of their own
__global__ void cc_rows_kernel(int* parent, int* row_label, int* ncomp, int R) {
    __shared__ int s_scan[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < R; base += blockDim.x) {
        const int r = base + (int)threadIdx.x;
        int root = -1;
        if (r < R) root = cc_find(parent, r);
        const int isroot = (r < R && root == r) ? 1 : 0;
        int x = isroot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[wid] = x;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int q = 0; q < nw; ++q) { const int t = s_scan[q]; if (q < wid) woff += t; tot += t; }
        const int carry = s_carry;
        if (isroot) row_label[r] = carry + woff + x - 1;        // roots first: a root's id precedes every member's id
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    // members: the root is a smaller row index, already labelled (possibly by an earlier chunk; same chunk needs the barrier)
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int root = cc_find(parent, r);
        if (root != r) row_label[r] = row_label[root];
    }
    if (threadIdx.x == 0) *ncomp = s_carry;
}

__global__ void cc_vars_kernel(int* parent, const int* row_label, const uint8_t* isvar, int* mp_label, int R, int M) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    // a variable always hangs under a row (it has at least one grid-listed slot), so its root is a row
    mp_label[p] = isvar[p] ? row_label[cc_find(parent, R + p)] : -1;
}

}  // namespace mss

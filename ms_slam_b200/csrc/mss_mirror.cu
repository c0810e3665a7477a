// mss_mirror.cu -- host side of the persistent device mirror (include/mss.h "Persistent device mirror"):
// storage that grows on demand, delta application (duplicates resolved on the host: the last op on an address wins),
// window assembly from keyframe handles, solve in place, deleted-handle bitmask hand-back, optional application of the
// deletion to the mirror itself.  Kernels: mss_mirror.cuh.  Replaces the per-window pointer walk of
// /root/reference/src/MapSparsification.cc:67-151.
#define MSS_KERNELS_TYPES_ONLY
#include "mss_internal.h"
#include "mss_mirror.cuh"
#include "mss_compact.cuh"

#include <algorithm>
#include <chrono>
#include <cstring>

using mssi::DevBuf;
using mssi::align_up;
using mssi::ensure;
using mssi::ensure_pinned;
using mssi::release;
using namespace mssm;

struct mss_mirror {
    mss_handle* h = nullptr;
    int S = 0;
    int n_kf = 0, n_mp = 0;                   // handles in use: [0, n)
    size_t kf_cap = 0, mp_cap = 0;
    DevBuf<int> slot_mp, obs_mp, kf_n, kf_win, okf_idx;
    DevBuf<uint16_t> slot_cell;
    DevBuf<unsigned> kf_key;
    DevBuf<uint8_t> okf_mark;
    DevBuf<MpRec> mp;
    DevBuf<uint8_t> win;                      // per-call window buffers (descriptors, handle lists, counts, view arrays)
    DevBuf<uint8_t> upload;                   // staging of ops / bulk loads
    uint8_t* h_pin = nullptr; size_t h_pin_cap = 0;    // pinned: descriptors up, counters down
    uint8_t* h_del = nullptr; size_t h_del_cap = 0;    // pinned: the deleted-handle words of a call
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    mss_mirror_stats stats{};
};

namespace {

MirrorDev dev_of(const mss_mirror* m) {
    MirrorDev D;
    D.slot_mp = m->slot_mp.p; D.obs_mp = m->obs_mp.p; D.slot_cell = m->slot_cell.p; D.kf_n = m->kf_n.p; D.kf_key = m->kf_key.p;
    D.kf_win = m->kf_win.p; D.okf_mark = m->okf_mark.p; D.okf_idx = m->okf_idx.p; D.mp = m->mp.p;
    D.S = m->S; D.n_kf = m->n_kf; D.n_mp = m->n_mp;
    return D;
}

inline int blocks_for(size_t n, int sm) { return (int)std::max<size_t>(1, std::min<size_t>((n + kT - 1) / kT, (size_t)sm * 8)); }

// grow a per-handle array, keeping its contents, and give the new tail its idle value
template <class T>
int grow_fill(mss_mirror* m, DevBuf<T>& b, size_t old_n, size_t new_cap, int fill_byte) {
    mss_handle* h = m->h;
    const size_t old_cap = b.cap;
    int rc = ensure(h, b, new_cap, true);
    if (rc) return rc;
    if (b.cap > old_cap) {
        const size_t from = std::min(old_n, old_cap);
        MSS_CUDA(h, cudaMemsetAsync(b.p + from, fill_byte, (b.cap - from) * sizeof(T), h->stream));
    }
    return MSS_OK;
}

int ensure_kfs(mss_mirror* m, int n_kf) {
    if ((size_t)n_kf > m->kf_cap) {
        size_t cap = m->kf_cap ? m->kf_cap : 64;
        while (cap < (size_t)n_kf) cap = cap + cap / 2 + 64;
        const size_t S = (size_t)m->S, old = (size_t)m->n_kf;
        int rc;
        if ((rc = grow_fill(m, m->slot_mp, old * S, cap * S, 0xFF))) return rc;          // -1
        if ((rc = grow_fill(m, m->obs_mp, old * S, cap * S, 0xFF))) return rc;
        if ((rc = grow_fill(m, m->slot_cell, old * S, cap * S, 0xFF))) return rc;        // 0xFFFF = not in the grid
        if ((rc = grow_fill(m, m->kf_n, old, cap, 0))) return rc;
        if ((rc = grow_fill(m, m->kf_key, old, cap, 0))) return rc;
        if ((rc = grow_fill(m, m->kf_win, old, cap, 0))) return rc;
        if ((rc = grow_fill(m, m->okf_idx, old, cap, 0))) return rc;
        if ((rc = grow_fill(m, m->okf_mark, old, cap, 0))) return rc;
        m->kf_cap = std::min({m->slot_mp.cap / S, m->obs_mp.cap / S, m->slot_cell.cap / S, m->kf_n.cap, m->kf_key.cap, m->kf_win.cap,
                              m->okf_idx.cap, m->okf_mark.cap});
    }
    m->n_kf = std::max(m->n_kf, n_kf);
    return MSS_OK;
}

int ensure_mps(mss_mirror* m, int n_mp) {
    mss_handle* h = m->h;
    if ((size_t)n_mp > m->mp_cap) {
        size_t cap = m->mp_cap ? m->mp_cap : 4096;
        while (cap < (size_t)n_mp) cap = cap + cap / 2 + 4096;
        const size_t old_cap = m->mp.cap;
        int rc;
        if ((rc = ensure(h, m->mp, cap, true))) return rc;
        const size_t keep = std::min((size_t)m->n_mp, old_cap);          // records in use keep their contents
        mk_init_mp<<<blocks_for(m->mp.cap - keep, h->sm_count), kT, 0, h->stream>>>(m->mp.p + keep, m->mp.cap - keep);
        MSS_CUDA(h, cudaGetLastError());
        m->mp_cap = m->mp.cap;
    }
    m->n_mp = std::max(m->n_mp, n_mp);
    return MSS_OK;
}

struct WinLayout {              // byte offsets inside mss_mirror::win of one window's buffers
    size_t kf, kf_first, feat_ptr, okf_list, okf_total, cnt;      // phase A
    size_t slots, nobs16, mp_handle, pairs, del;                   // phase B (sized after the read-back)
};

// Assemble the views of nwin windows on the device.  On return (stream synchronised) h_cnt holds every window's counters
// and `lay` / `dW` describe the buffers; the per-handle scratch is still armed (caller runs the solve, then finish()).
struct Assembly {
    std::vector<WinLayout> lay;
    std::vector<MWin> hw;
    MWin* dW = nullptr;
    int* h_cnt = nullptr;       // pinned [nwin][C_COUNT]
    int Kmax = 0;
    bool tables = false;
    float build_ms = 0.f;
    size_t del_base = 0, del_bytes = 0;     // the block of all windows' deleted-handle words inside mss_mirror::win
};

int assemble(mss_mirror* m, int nwin, const mss_mirror_window* win, Assembly& A) {
    mss_handle* h = m->h;
    const MirrorDev D = dev_of(m);
    A.lay.assign(nwin, WinLayout{});
    A.hw.assign(nwin, MWin{});
    // uploaded part first (descriptors, keyframe handle lists, counters), device-only scratch behind it
    size_t off = align_up((size_t)nwin * sizeof(MWin), 256);
    const size_t cnt_base = off;                      // the counters of all windows are contiguous: one copy brings them back
    off += align_up((size_t)nwin * C_COUNT * 4, 256);
    for (int w = 0; w < nwin; ++w) {
        const int K = win[w].K;
        if (K < 0 || (K > 0 && !win[w].kf)) { h->err = "mirror window: negative K or NULL keyframe list"; return MSS_E_BADARG; }
        if (K > mss::kMaxWindowRows) { h->err = "mirror window: more than 65535 keyframes"; return MSS_E_BADARG; }
        A.Kmax = std::max(A.Kmax, K);
        WinLayout& L = A.lay[w];
        L.kf = off; off += align_up((size_t)std::max(K, 1) * 4, 64);
        L.cnt = cnt_base + (size_t)w * C_COUNT * 4;
    }
    const size_t upload_bytes = off = align_up(off, 256);
    for (int w = 0; w < nwin; ++w) {
        const int K = win[w].K;
        WinLayout& L = A.lay[w];
        L.kf_first = off; off += align_up((size_t)(K + 1) * 4, 256);
        L.feat_ptr = off; off += align_up((size_t)(K + 1) * 4, 256);
        L.okf_list = off; off += align_up((size_t)2 * (kMaxOutside + 1) * 4, 256);
        L.okf_total = off; off += align_up((size_t)(kMaxOutside + 1) * 4, 256);
    }
    const size_t phaseA_bytes = off;
    int rc;
    if ((rc = ensure(h, m->win, phaseA_bytes + 256))) return rc;
    size_t pin_bytes = upload_bytes + (size_t)nwin * C_COUNT * 4 + 256;
    if ((rc = ensure_pinned(h, (void**)&m->h_pin, &m->h_pin_cap, pin_bytes))) return rc;
    uint8_t* hp = m->h_pin;
    memset(hp, 0, upload_bytes);
    int64_t h2d = 0;
    for (int w = 0; w < nwin; ++w) {
        const WinLayout& L = A.lay[w];
        MWin& q = A.hw[w];
        q.kf = reinterpret_cast<const int*>(m->win.p + L.kf);
        q.K = win[w].K; q.w = w; q.n_max_floor = win[w].n_max_floor; q.apply = win[w].apply ? 1 : 0;
        q.kf_first = reinterpret_cast<int*>(m->win.p + L.kf_first);
        q.feat_ptr = reinterpret_cast<int*>(m->win.p + L.feat_ptr);
        q.okf_list = reinterpret_cast<int*>(m->win.p + L.okf_list);
        q.okf_total = reinterpret_cast<int*>(m->win.p + L.okf_total);
        q.cnt = reinterpret_cast<int*>(m->win.p + L.cnt);
        if (win[w].K) memcpy(hp + L.kf, win[w].kf, (size_t)win[w].K * 4);
        int* c = reinterpret_cast<int*>(hp + L.cnt);
        c[C_HLO] = 0x7FFFFFFF; c[C_KFLO] = 0x7FFFFFFF; c[C_KFHI] = -1;
        h2d += (int64_t)win[w].K * 4;
    }
    memcpy(hp, A.hw.data(), (size_t)nwin * sizeof(MWin));
    A.dW = reinterpret_cast<MWin*>(m->win.p);
    MSS_CUDA(h, cudaEventRecord(m->e0, h->stream));
    MSS_CUDA(h, cudaMemcpyAsync(m->win.p, hp, upload_bytes, cudaMemcpyHostToDevice, h->stream));
    h2d += (int64_t)nwin * (int64_t)(sizeof(MWin) + C_COUNT * 4);
    const int sm = h->sm_count;
    const int gk = std::max(1, std::min(A.Kmax, sm * 8));
    const dim3 g_kf(gk, nwin), g_flat(std::max(1, std::min(sm * 4, 2048)), nwin);
    if (A.Kmax > 0) {
        mk_mark<<<dim3((A.Kmax + kT - 1) / kT, nwin), kT, 0, h->stream>>>(D, A.dW);
        mk_first<<<g_kf, kT, 0, h->stream>>>(D, A.dW);
        mk_count<<<g_kf, kT, 0, h->stream>>>(D, A.dW);
    }
    mk_scan<<<nwin, kT, 0, h->stream>>>(A.dW);
    mk_obs_scan<0><<<g_flat, kT, 0, h->stream>>>(D, A.dW);
    mk_okf_collect<<<dim3(std::max(1, std::min(sm, 64)), nwin), kT, 0, h->stream>>>(D, A.dW);
    mk_okf_rank<<<nwin, kT, 0, h->stream>>>(D, A.dW);
    MSS_CUDA(h, cudaGetLastError());
    h->stats.kernel_launches += 4 + (A.Kmax > 0 ? 3 : 0);
    // ---- read-back of the counters: buffer sizes of phase B ------------------------------------------------------------
    A.h_cnt = reinterpret_cast<int*>(hp + upload_bytes);
    MSS_CUDA(h, cudaMemcpyAsync(A.h_cnt, m->win.p + A.lay[0].cnt, (size_t)nwin * C_COUNT * 4, cudaMemcpyDeviceToHost, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    off = phaseA_bytes;
    for (int w = 0; w < nwin; ++w) {
        const int* c = A.h_cnt + (size_t)w * C_COUNT;
        if (c[C_ERR]) continue;
        WinLayout& L = A.lay[w];
        L.slots = off; off += align_up((size_t)std::max(c[C_F], 1) * 4, 256);
        L.nobs16 = off; off += align_up((size_t)std::max(c[C_M], 1) * 2, 256);
        L.mp_handle = off; off += align_up((size_t)std::max(c[C_M], 1) * 4, 256);
        L.pairs = off; off += align_up((size_t)std::max(c[C_O], 1) * 4, 256);
    }
    // the deleted-handle words of all windows lie in ONE block: they come back with one copy (a copy per window costs the
    // DMA stream a few microseconds each, whatever its size)
    A.del_base = off;
    for (int w = 0; w < nwin; ++w) {
        const int* c = A.h_cnt + (size_t)w * C_COUNT;
        if (c[C_ERR]) continue;
        const int dwords = c[C_M] > 0 ? ((c[C_HHI] + 31) >> 5) - (c[C_HLO] >> 5) : 0;
        A.lay[w].del = off; off += align_up((size_t)std::max(dwords, 1) * 4, 16);
    }
    A.del_bytes = off - A.del_base;
    off = align_up(off, 256);
    if (off > m->win.cap) {
        // the phase-A buffers hold live data: grow with copy
        if ((rc = ensure(h, m->win, off + 256, true))) return rc;
        // pointers moved: rebuild the phase-A part of the descriptors
        for (int w = 0; w < nwin; ++w) {
            const WinLayout& L = A.lay[w];
            MWin& q = A.hw[w];
            q.kf = reinterpret_cast<const int*>(m->win.p + L.kf);
            q.kf_first = reinterpret_cast<int*>(m->win.p + L.kf_first);
            q.feat_ptr = reinterpret_cast<int*>(m->win.p + L.feat_ptr);
            q.okf_list = reinterpret_cast<int*>(m->win.p + L.okf_list);
            q.okf_total = reinterpret_cast<int*>(m->win.p + L.okf_total);
            q.cnt = reinterpret_cast<int*>(m->win.p + L.cnt);
        }
        A.dW = reinterpret_cast<MWin*>(m->win.p);
    }
    for (int w = 0; w < nwin; ++w) {
        const int* c = A.h_cnt + (size_t)w * C_COUNT;
        if (c[C_ERR]) continue;
        const WinLayout& L = A.lay[w];
        MWin& q = A.hw[w];
        q.slots = reinterpret_cast<uint32_t*>(m->win.p + L.slots);
        q.nobs16 = reinterpret_cast<uint16_t*>(m->win.p + L.nobs16);
        q.mp_handle = reinterpret_cast<int*>(m->win.p + L.mp_handle);
        q.pairs = reinterpret_cast<uint32_t*>(m->win.p + L.pairs);
        q.del = reinterpret_cast<unsigned*>(m->win.p + L.del);
    }
    memcpy(hp, A.hw.data(), (size_t)nwin * sizeof(MWin));
    MSS_CUDA(h, cudaMemcpyAsync(m->win.p, hp, (size_t)nwin * sizeof(MWin), cudaMemcpyHostToDevice, h->stream));
    if (A.Kmax > 0) mk_slots<<<g_kf, kT, 0, h->stream>>>(D, A.dW);
    mk_okf_total<<<dim3(std::max(1, std::min(sm * 2, kMaxOutside)), nwin), kT, 0, h->stream>>>(D, A.dW);
    mk_obs_scan<1><<<g_flat, kT, 0, h->stream>>>(D, A.dW);
    MSS_CUDA(h, cudaGetLastError());
    MSS_CUDA(h, cudaEventRecord(m->e1, h->stream));
    h->stats.kernel_launches += 2 + (A.Kmax > 0 ? 1 : 0);
    A.tables = true;
    m->stats.last_h2d_bytes = h2d;
    m->stats.windows_built += nwin;
    return MSS_OK;
}

std::string mirror_error_text(int e) {
    std::string s;
    if (e & ME_KF_RANGE) s += " keyframe handle out of range;";
    if (e & ME_KF_TWICE) s += " keyframe listed twice or in two windows;";
    if (e & ME_MP_RANGE) s += " map-point handle out of range in a slot;";
    if (e & ME_MP_SHARED) s += " two windows of the call share a map point;";
    if (e & ME_DEPENDENT) s += " a keyframe of one window observes a variable of another (windows are not independent);";
    if (e & ME_OUTSIDE_OVERFLOW) s += " more than 4095 outside keyframes;";
    if (e & ME_NOBS_RANGE) s += " Observations() above 65535;";
    return s;
}

mss_window_view view_of(const mss_mirror* m, const Assembly& A, int w) {
    const int* c = A.h_cnt + (size_t)w * C_COUNT;
    const MWin& q = A.hw[w];
    mss_window_view v{};
    v.K = q.K; v.H = c[C_H]; v.M = c[C_M]; v.F = c[C_F]; v.O = c[C_O];
    v.memory = MSS_MEM_DEVICE;
    v.layout = MSS_LAYOUT_PACKED;
    v.n_max_floor = q.n_max_floor;
    v.feat_ptr = q.feat_ptr; v.slots = q.slots; v.mp_nobs16 = q.nobs16; v.obs_pairs = q.pairs; v.okf_total = q.okf_total;
    (void)m;
    return v;
}

}  // namespace

extern "C" {

int mss_mirror_create(mss_handle* h, int32_t slots_per_kf, mss_mirror** out) {
    if (!h || !out) return MSS_E_BADARG;
    *out = nullptr;
    if (slots_per_kf <= 0 || slots_per_kf > 65536) { h->err = "mirror: slots_per_kf out of range"; return MSS_E_BADARG; }
    mss_mirror* m = new (std::nothrow) mss_mirror();
    if (!m) return MSS_E_NOMEM;
    m->h = h;
    m->S = slots_per_kf;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaEventCreate(&m->e0) != cudaSuccess || cudaEventCreate(&m->e1) != cudaSuccess) {
        h->err = "mirror: cudaEventCreate failed";
        delete m;
        return MSS_E_CUDA;
    }
    m->stats.slots_per_kf = slots_per_kf;
    *out = m;
    return MSS_OK;
}

void mss_mirror_destroy(mss_mirror* m) {
    if (!m) return;
    cudaSetDevice(m->h->device);
    cudaStreamSynchronize(m->h->stream);
    release(m->slot_mp); release(m->obs_mp); release(m->kf_n); release(m->kf_win); release(m->okf_idx); release(m->mp);
    release(m->slot_cell); release(m->kf_key); release(m->okf_mark); release(m->win);
    release(m->upload);
    if (m->h_pin) cudaFreeHost(m->h_pin);
    if (m->h_del) cudaFreeHost(m->h_del);
    if (m->e0) cudaEventDestroy(m->e0);
    if (m->e1) cudaEventDestroy(m->e1);
    delete m;
}

int mss_mirror_add_keyframes(mss_mirror* m, int32_t kf0, int32_t n, const uint32_t* sort_key, const int32_t* n_slots,
                             const uint16_t* cells, const int32_t* slot_mp, const int32_t* obs_mp) {
    if (!m) return MSS_E_BADARG;
    mss_handle* h = m->h;
    h->err.clear();
    if (kf0 < 0 || n < 0 || !n_slots || !cells || !slot_mp) { h->err = "mirror add_keyframes: bad arguments"; return MSS_E_BADARG; }
    if (n == 0) return MSS_OK;
    MSS_CUDA(h, cudaSetDevice(h->device));
    const size_t S = (size_t)m->S, tot = (size_t)n * S;
    int max_mp = -1;
    for (int i = 0; i < n; ++i)
        if (n_slots[i] < 0 || n_slots[i] > m->S) { h->err = "mirror add_keyframes: n_slots exceeds slots_per_kf"; return MSS_E_BADARG; }
    for (size_t i = 0; i < tot; ++i) {
        if (slot_mp[i] < -1 || (obs_mp && obs_mp[i] < -1)) { h->err = "mirror add_keyframes: map-point handle below -1"; return MSS_E_BADARG; }
        max_mp = std::max(max_mp, std::max(slot_mp[i], obs_mp ? obs_mp[i] : -1));
    }
    int rc;
    if ((rc = ensure_kfs(m, kf0 + n))) return rc;
    if ((rc = ensure_mps(m, max_mp + 1))) return rc;
    const size_t base = (size_t)kf0 * S;
    MSS_CUDA(h, cudaMemcpyAsync(m->slot_mp.p + base, slot_mp, tot * 4, cudaMemcpyHostToDevice, h->stream));
    MSS_CUDA(h, cudaMemcpyAsync(m->slot_cell.p + base, cells, tot * 2, cudaMemcpyHostToDevice, h->stream));
    if (obs_mp) MSS_CUDA(h, cudaMemcpyAsync(m->obs_mp.p + base, obs_mp, tot * 4, cudaMemcpyHostToDevice, h->stream));
    else MSS_CUDA(h, cudaMemsetAsync(m->obs_mp.p + base, 0xFF, tot * 4, h->stream));
    MSS_CUDA(h, cudaMemcpyAsync(m->kf_n.p + kf0, n_slots, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    if (sort_key) MSS_CUDA(h, cudaMemcpyAsync(m->kf_key.p + kf0, sort_key, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    else {
        std::vector<unsigned> keys(n);
        for (int i = 0; i < n; ++i) keys[i] = (unsigned)(kf0 + i);
        MSS_CUDA(h, cudaMemcpyAsync(m->kf_key.p + kf0, keys.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
        MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    if (obs_mp) {
        mk_obs_ranges<<<blocks_for(tot, h->sm_count), kT, 0, h->stream>>>(dev_of(m), kf0, n);
        MSS_CUDA(h, cudaGetLastError());
        h->stats.kernel_launches += 1;
    }
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));           // the caller's arrays may be pageable and reused at once
    m->stats.n_keyframes = m->n_kf; m->stats.n_map_points = m->n_mp;
    return MSS_OK;
}

int mss_mirror_add_keyframe(mss_mirror* m, int32_t kf, uint32_t sort_key, int32_t n_slots, const uint16_t* cells,
                            const int32_t* slot_mp, const int32_t* obs_mp) {
    if (!m) return MSS_E_BADARG;
    mss_handle* h = m->h;
    if (n_slots < 0 || n_slots > m->S || (n_slots > 0 && (!cells || !slot_mp))) { h->err = "mirror add_keyframe: bad arguments"; return MSS_E_BADARG; }
    const size_t S = (size_t)m->S;
    std::vector<uint16_t> c(S, (uint16_t)MSS_CELL_NONE);
    std::vector<int32_t> s(S, -1), o(S, -1);
    if (n_slots) {
        memcpy(c.data(), cells, (size_t)n_slots * 2);
        memcpy(s.data(), slot_mp, (size_t)n_slots * 4);
        if (obs_mp) memcpy(o.data(), obs_mp, (size_t)n_slots * 4);
    }
    return mss_mirror_add_keyframes(m, kf, 1, &sort_key, &n_slots, c.data(), s.data(), o.data());
}

int mss_mirror_set_map_points(mss_mirror* m, int32_t mp0, int32_t n, const int32_t* nobs, const uint8_t* bad) {
    if (!m) return MSS_E_BADARG;
    mss_handle* h = m->h;
    h->err.clear();
    if (mp0 < 0 || n < 0 || (n > 0 && !nobs)) { h->err = "mirror set_map_points: bad arguments"; return MSS_E_BADARG; }
    if (n == 0) return MSS_OK;
    MSS_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure_mps(m, mp0 + n))) return rc;
    const size_t off_bad = align_up((size_t)n * 4, 16);
    if ((rc = ensure(h, m->upload, off_bad + (size_t)n + 16))) return rc;
    MSS_CUDA(h, cudaMemcpyAsync(m->upload.p, nobs, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    if (bad) MSS_CUDA(h, cudaMemcpyAsync(m->upload.p + off_bad, bad, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    mk_set_mp<<<(n + kT - 1) / kT, kT, 0, h->stream>>>(m->mp.p + mp0, reinterpret_cast<const int*>(m->upload.p),
                                                     bad ? m->upload.p + off_bad : nullptr, n);
    MSS_CUDA(h, cudaGetLastError());
    h->stats.kernel_launches += 1;
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    m->stats.n_map_points = m->n_mp;
    return MSS_OK;
}

int mss_mirror_apply(mss_mirror* m, const mss_mirror_op* ops, int32_t n) {
    if (!m) return MSS_E_BADARG;
    mss_handle* h = m->h;
    h->err.clear();
    if (n < 0 || (n > 0 && !ops)) { h->err = "mirror apply: bad arguments"; return MSS_E_BADARG; }
    if (n == 0) return MSS_OK;
    MSS_CUDA(h, cudaSetDevice(h->device));
    // The device applies a segment of ops in parallel, so order is resolved here: inside a segment the last op on an
    // address wins; a keyframe compaction renumbers slots and therefore ends a segment.
    int rc;
    int max_kf = -1, max_mp = -1;
    for (int i = 0; i < n; ++i) {
        const mss_mirror_op& o = ops[i];
        if (o.kind == MSS_MOP_SLOT || o.kind == MSS_MOP_OBS) { max_kf = std::max(max_kf, o.a); max_mp = std::max(max_mp, o.c); }
        else if (o.kind == MSS_MOP_MP) max_mp = std::max(max_mp, o.a);
        else if (o.kind == MSS_MOP_KF_COMPACT) max_kf = std::max(max_kf, o.a);
        else { h->err = "mirror apply: unknown op kind"; return MSS_E_BADARG; }
        if (o.a < 0) { h->err = "mirror apply: negative handle"; return MSS_E_BADARG; }
    }
    if ((rc = ensure_kfs(m, max_kf + 1))) return rc;
    if ((rc = ensure_mps(m, max_mp + 1))) return rc;
    if ((rc = ensure(h, m->upload, (size_t)n * sizeof(DevOp) + 64))) return rc;
    int* d_err = reinterpret_cast<int*>(m->upload.p + align_up((size_t)n * sizeof(DevOp), 16));
    MSS_CUDA(h, cudaMemsetAsync(d_err, 0, 4, h->stream));
    std::vector<DevOp> seg;
    size_t up_off = 0;
    // last op per address: open addressing over (key -> index of the last op seen), sized for the segment
    std::vector<uint64_t> tkey;
    std::vector<int> tval;
    auto key_of = [](const DevOp& o) {
        return ((uint64_t)o.kind << 60) | ((uint64_t)(uint32_t)o.a << 24) | (uint64_t)(uint32_t)(o.kind == MSS_MOP_MP ? 0 : o.b);
    };
    auto flush = [&]() -> int {
        if (seg.empty()) return MSS_OK;
        size_t cap = 16;
        while (cap < seg.size() * 2) cap <<= 1;
        tkey.assign(cap, ~0ull);
        tval.assign(cap, -1);
        for (size_t i = 0; i < seg.size(); ++i) {
            const uint64_t key = key_of(seg[i]);
            size_t s = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (cap - 1);
            while (tkey[s] != ~0ull && tkey[s] != key) s = (s + 1) & (cap - 1);
            tkey[s] = key;
            tval[s] = (int)i;
        }
        std::vector<DevOp> uniq;
        uniq.reserve(seg.size());
        for (size_t i = 0; i < seg.size(); ++i) {
            const uint64_t key = key_of(seg[i]);
            size_t s = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (cap - 1);
            while (tkey[s] != key) s = (s + 1) & (cap - 1);
            if (tval[s] == (int)i) uniq.push_back(seg[i]);
        }
        MSS_CUDA(h, cudaMemcpyAsync(m->upload.p + up_off, uniq.data(), uniq.size() * sizeof(DevOp), cudaMemcpyHostToDevice, h->stream));
        MSS_CUDA(h, cudaStreamSynchronize(h->stream));           // `uniq` is pageable and dies with this scope
        mk_apply_ops<<<(int)((uniq.size() + kT - 1) / kT), kT, 0, h->stream>>>(dev_of(m), reinterpret_cast<const DevOp*>(m->upload.p + up_off),
                                                                            (int)uniq.size(), d_err);
        MSS_CUDA(h, cudaGetLastError());
        h->stats.kernel_launches += 1;
        m->stats.last_h2d_bytes += (int64_t)(uniq.size() * sizeof(DevOp));
        up_off += align_up(uniq.size() * sizeof(DevOp), 16);
        seg.clear();
        return MSS_OK;
    };
    m->stats.last_h2d_bytes = 0;
    for (int i = 0; i < n; ++i) {
        const mss_mirror_op& o = ops[i];
        if (o.kind == MSS_MOP_KF_COMPACT) {
            if ((rc = flush())) return rc;
            mk_kf_compact<<<1, kT, 0, h->stream>>>(dev_of(m), o.a);
            MSS_CUDA(h, cudaGetLastError());
            h->stats.kernel_launches += 1;
            continue;
        }
        seg.push_back(DevOp{o.kind, o.a, o.b, o.c});
    }
    if ((rc = flush())) return rc;
    int err = 0;
    MSS_CUDA(h, cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    m->stats.ops_applied += n;
    m->stats.n_keyframes = m->n_kf; m->stats.n_map_points = m->n_mp;
    if (err) { h->err = "mirror apply: an op addressed a slot index or handle out of range (ignored)"; return MSS_E_BADARG; }
    return MSS_OK;
}

int mss_mirror_solve(mss_mirror* m, int32_t nwin, mss_mirror_window* windows, mss_result* results) {
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    if (!m) return MSS_E_BADARG;
    mss_handle* h = m->h;
    h->err.clear();
    if (nwin < 0 || (nwin > 0 && (!windows || !results))) { h->err = "mirror solve: bad arguments"; return MSS_E_BADARG; }
    if (nwin == 0) return MSS_OK;
    if (h->nranks > 1) { h->err = "mirror solve: the handle has a communicator attached; mirrors are per device"; return MSS_E_BADARG; }
    MSS_CUDA(h, cudaSetDevice(h->device));
    for (int w = 0; w < nwin; ++w)
        if (windows[w].del_bits && windows[w].del_words < (m->n_mp + 31) / 32) { h->err = "mirror solve: del_bits is smaller than the handle space"; return MSS_E_BADARG; }
    Assembly A;
    int rc = assemble(m, nwin, windows, A);
    if (rc != MSS_OK) return rc;
    const MirrorDev D = dev_of(m);
    // ---- solve in place (device views) -----------------------------------------------------------------------------------
    std::vector<mss_window_view> views(nwin);
    int first_err = -1;
    for (int w = 0; w < nwin; ++w) {
        const int* c = A.h_cnt + (size_t)w * C_COUNT;
        if (c[C_ERR]) {
            if (first_err < 0) first_err = w;
            views[w] = mss_window_view{};                  // an empty window: solved trivially, reported as rejected below
            views[w].memory = MSS_MEM_DEVICE;
            views[w].layout = MSS_LAYOUT_PACKED;
            views[w].feat_ptr = A.hw[w].feat_ptr;          // K = 0: only feat_ptr[0] is touched by nobody
        } else {
            views[w] = view_of(m, A, w);
        }
        views[w].result_memory = MSS_RESULT_HOST;          // keep_bits / kf_cov / kf_slack are host buffers (or NULL)
    }
    float build_ms = 0.f;
    rc = mssi::solve_batch_impl(h, nwin, views.data(), results);
    const std::string solve_err = h->err;
    const double solve_ms = h->stats.last_device_ms;
    cudaEventElapsedTime(&build_ms, m->e0, m->e1);
    const bool solved = rc == MSS_OK || rc == MSS_E_NOCONVERGE;
    // ---- deleted-handle bitmask, optional application to the mirror, scratch back to idle ----------------------------------
    const int sm = h->sm_count;
    const dim3 g_flat(std::max(1, std::min(sm * 4, 2048)), nwin);
    if (solved) {
        for (int w = 0; w < nwin; ++w) {
            const int* c = A.h_cnt + (size_t)w * C_COUNT;
            A.hw[w].keep = (c[C_ERR] || results[w].status == MSS_E_BADARG) ? nullptr : h->out.p + h->last_out_off[w] + mss::kHdrWords;
        }
        memcpy(m->h_pin, A.hw.data(), (size_t)nwin * sizeof(MWin));
        MSS_CUDA(h, cudaMemcpyAsync(m->win.p, m->h_pin, (size_t)nwin * sizeof(MWin), cudaMemcpyHostToDevice, h->stream));
        mk_del_zero<<<dim3(std::max(1, std::min(sm, 64)), nwin), kT, 0, h->stream>>>(A.dW);
        mk_deleted<<<dim3(std::max(1, std::min(sm, 64)), nwin), kT, 0, h->stream>>>(D, A.dW);
        bool any_apply = false;
        for (int w = 0; w < nwin; ++w) any_apply = any_apply || windows[w].apply;
        if (any_apply) mk_obs_scan<2><<<g_flat, kT, 0, h->stream>>>(D, A.dW);        // (windows with apply == 0 have no bit set... see below)
        h->stats.kernel_launches += 2 + (any_apply ? 1 : 0);
    }
    mk_reset<<<dim3(std::max(1, std::min(sm, 64)), nwin), kT, 0, h->stream>>>(D, A.dW, A.tables ? 1 : 0);
    MSS_CUDA(h, cudaGetLastError());
    h->stats.kernel_launches += 1;
    int64_t d2h = (int64_t)nwin * C_COUNT * 4;
    // (the sizes read below are the ones of the first read-back; this copy adds the number of deleted points)
    std::vector<int> c0(A.h_cnt, A.h_cnt + (size_t)nwin * C_COUNT);
    bool want_del = false;
    int rc2 = MSS_OK;
    MSS_CUDA(h, cudaMemcpyAsync(A.h_cnt, m->win.p + A.lay[0].cnt, (size_t)nwin * C_COUNT * 4, cudaMemcpyDeviceToHost, h->stream));
    for (int w = 0; w < nwin; ++w) {
        const int* c = c0.data() + (size_t)w * C_COUNT;
        mss_mirror_window& q = windows[w];
        q.M = c[C_ERR] ? 0 : c[C_M]; q.H = c[C_ERR] ? 0 : c[C_H]; q.F = c[C_ERR] ? 0 : c[C_F]; q.O = c[C_ERR] ? 0 : c[C_O];
        const bool has = !c[C_ERR] && c[C_M] > 0;
        q.h_lo = has ? c[C_HLO] : 0; q.h_hi = has ? c[C_HHI] : 0;
        if (q.del_bits && has && solved) { want_del = true; d2h += (int64_t)(((q.h_hi + 31) >> 5) - (q.h_lo >> 5)) * 4; }
        if (q.mp_handle && has) {
            if (q.mp_cap < c[C_M]) { h->err = "mirror solve: mp_handle buffer smaller than the window's map-point table"; rc = MSS_E_BADARG; }
            else {
                MSS_CUDA(h, cudaMemcpyAsync(q.mp_handle, m->win.p + A.lay[w].mp_handle, (size_t)c[C_M] * 4, cudaMemcpyDeviceToHost, h->stream));
                d2h += (int64_t)c[C_M] * 4;
            }
        }
    }
    if (want_del && A.del_bytes > 0) {
        if ((rc2 = ensure_pinned(h, (void**)&m->h_del, &m->h_del_cap, A.del_bytes))) return rc2;
        MSS_CUDA(h, cudaMemcpyAsync(m->h_del, m->win.p + A.del_base, A.del_bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (want_del && A.del_bytes > 0)
        for (int w = 0; w < nwin; ++w) {
            const int* c = c0.data() + (size_t)w * C_COUNT;
            const mss_mirror_window& q = windows[w];
            if (!q.del_bits || c[C_ERR] || c[C_M] <= 0 || !solved) continue;
            const int lo = q.h_lo >> 5, hi = (q.h_hi + 31) >> 5;
            memcpy(q.del_bits + lo, m->h_del + (A.lay[w].del - A.del_base), (size_t)(hi - lo) * 4);
        }
    for (int w = 0; w < nwin; ++w) {
        const int* c = A.h_cnt + (size_t)w * C_COUNT;
        windows[w].n_deleted = c[C_ERR] ? 0 : c[C_NDEL];
        if (c[C_ERR]) {
            results[w].status = MSS_E_BADARG;
            results[w].n_kept = 0; results[w].n_vars = 0;
        }
    }
    if (first_err >= 0) {
        h->err = "mirror window " + std::to_string(first_err) + " rejected:" + mirror_error_text(A.h_cnt[(size_t)first_err * C_COUNT + C_ERR]) +
                 " nothing is deleted for it";
        if (rc == MSS_OK || rc == MSS_E_NOCONVERGE) rc = MSS_E_BADARG;
    } else if (!solved) {
        h->err = solve_err;
    }
    m->stats.last_build_ms = build_ms;
    m->stats.last_solve_ms = solve_ms;
    m->stats.last_d2h_bytes = d2h + h->stats.last_d2h_bytes;
    m->stats.last_h2d_bytes += h->stats.last_h2d_bytes;
    m->stats.last_total_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
    m->stats.device_bytes = h->device_bytes;
    return rc;
}

int mss_mirror_build_view(mss_mirror* m, const mss_mirror_window* window, int32_t* sizes5, int32_t* feat_ptr, uint32_t* slots,
                          uint16_t* mp_nobs16, uint32_t* obs_pairs, int32_t* okf_total, int32_t* mp_handle, int32_t* okf_handle) {
    if (!m) return MSS_E_BADARG;
    mss_handle* h = m->h;
    h->err.clear();
    if (!window || !sizes5) { h->err = "mirror build_view: bad arguments"; return MSS_E_BADARG; }
    MSS_CUDA(h, cudaSetDevice(h->device));
    Assembly A;
    int rc = assemble(m, 1, window, A);
    if (rc != MSS_OK) return rc;
    const int* c = A.h_cnt;
    const int err = c[C_ERR];
    if (!err) {
        sizes5[0] = window->K; sizes5[1] = c[C_H]; sizes5[2] = c[C_M]; sizes5[3] = c[C_F]; sizes5[4] = c[C_O];
        const WinLayout& L = A.lay[0];
        const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
        if (feat_ptr) MSS_CUDA(h, cudaMemcpyAsync(feat_ptr, m->win.p + L.feat_ptr, (size_t)(window->K + 1) * 4, k, h->stream));
        if (slots && c[C_F]) MSS_CUDA(h, cudaMemcpyAsync(slots, m->win.p + L.slots, (size_t)c[C_F] * 4, k, h->stream));
        if (mp_nobs16 && c[C_M]) MSS_CUDA(h, cudaMemcpyAsync(mp_nobs16, m->win.p + L.nobs16, (size_t)c[C_M] * 2, k, h->stream));
        if (obs_pairs && c[C_O]) MSS_CUDA(h, cudaMemcpyAsync(obs_pairs, m->win.p + L.pairs, (size_t)c[C_O] * 4, k, h->stream));
        if (okf_total && c[C_H]) MSS_CUDA(h, cudaMemcpyAsync(okf_total, m->win.p + L.okf_total, (size_t)c[C_H] * 4, k, h->stream));
        if (mp_handle && c[C_M]) MSS_CUDA(h, cudaMemcpyAsync(mp_handle, m->win.p + L.mp_handle, (size_t)c[C_M] * 4, k, h->stream));
        if (okf_handle && c[C_H]) MSS_CUDA(h, cudaMemcpyAsync(okf_handle, m->win.p + L.okf_list + (size_t)(kMaxOutside + 1) * 4, (size_t)c[C_H] * 4, k, h->stream));
    }
    mk_reset<<<dim3(std::max(1, std::min(h->sm_count, 64)), 1), kT, 0, h->stream>>>(dev_of(m), A.dW, A.tables ? 1 : 0);
    MSS_CUDA(h, cudaGetLastError());
    h->stats.kernel_launches += 1;
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (err) { h->err = "mirror window rejected:" + mirror_error_text(err); return MSS_E_BADARG; }
    return MSS_OK;
}

int mss_mirror_components(mss_mirror* m, const mss_mirror_window* window, int32_t* kf_label, int32_t* ncomp, int32_t* n_max) {
    if (!m) return MSS_E_BADARG;
    mss_handle* h = m->h;
    h->err.clear();
    if (!window || !kf_label || !ncomp) { h->err = "mirror components: bad arguments"; return MSS_E_BADARG; }
    MSS_CUDA(h, cudaSetDevice(h->device));
    Assembly A;
    int rc = assemble(m, 1, window, A);
    if (rc != MSS_OK) return rc;
    const int err = A.h_cnt[C_ERR];
    if (!err) {
        const mss_window_view v = view_of(m, A, 0);
        DevBuf<int> lab;
        rc = ensure(h, lab, (size_t)v.K + v.H + 16);
        if (rc == MSS_OK) rc = mss_components(h, &v, lab.p, nullptr, ncomp, n_max);          // labels in device memory, like the view
        if (rc == MSS_OK && v.K) {
            const cudaError_t e = cudaMemcpy(kf_label, lab.p, (size_t)v.K * 4, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { h->err = std::string("mirror components: ") + cudaGetErrorString(e); rc = MSS_E_CUDA; }
        }
        release(lab);
    }
    const std::string keep_err = h->err;
    mk_reset<<<dim3(std::max(1, std::min(h->sm_count, 64)), 1), kT, 0, h->stream>>>(dev_of(m), A.dW, A.tables ? 1 : 0);
    MSS_CUDA(h, cudaGetLastError());
    h->stats.kernel_launches += 1;
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    h->err = keep_err;
    if (err) { h->err = "mirror window rejected:" + mirror_error_text(err); return MSS_E_BADARG; }
    return rc;
}

static int compact_impl(mss_handle* h, mss_mirror* m, int32_t nkf, const mss_kf_payload* kfs, int32_t* n_out) {
    h->err.clear();
    if (nkf < 0 || (nkf > 0 && (!kfs || !n_out))) { h->err = "compact_keyframes: bad arguments"; return MSS_E_BADARG; }
    if (nkf == 0) return MSS_OK;
    MSS_CUDA(h, cudaSetDevice(h->device));
    std::vector<mssc::KfPayload> hp(nkf);
    for (int i = 0; i < nkf; ++i) {
        const mss_kf_payload& k = kfs[i];
        if (k.n < 0 || (!m && k.n > 0 && !k.keep) || (m && (k.kf < 0 || k.kf >= m->n_kf || k.n > m->S))) {
            h->err = "compact_keyframes: keyframe " + std::to_string(i) + ": bad row count, flags or handle";
            return MSS_E_BADARG;
        }
        hp[i] = mssc::KfPayload{k.n, m ? k.kf : -1, m ? nullptr : k.keep, static_cast<uint4*>(k.descriptors),
                                static_cast<uint32_t*>(k.keypoints), k.uright, k.depth};
    }
    DevBuf<uint8_t>& up = m ? m->upload : h->stage;
    const size_t off_out = align_up((size_t)nkf * sizeof(mssc::KfPayload), 16);
    int rc = ensure(h, up, off_out + (size_t)nkf * 4 + 16);
    if (rc) return rc;
    MSS_CUDA(h, cudaMemcpyAsync(up.p, hp.data(), (size_t)nkf * sizeof(mssc::KfPayload), cudaMemcpyHostToDevice, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));                  // hp is pageable
    MSS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    // TMA-staged variant when every array (and the mirror's slot rows) is 16-byte aligned; MSS_COMPACT_TMA=0 forces the register variant
    bool tma = !(getenv("MSS_COMPACT_TMA") && atoi(getenv("MSS_COMPACT_TMA")) == 0) && (!m || m->S % 4 == 0);
    for (int i = 0; i < nkf && tma; ++i) {
        const mssc::KfPayload& k = hp[i];
        tma = ((reinterpret_cast<uintptr_t>(k.desc) | reinterpret_cast<uintptr_t>(k.keys) | reinterpret_cast<uintptr_t>(k.uright) |
                reinterpret_cast<uintptr_t>(k.depth) | reinterpret_cast<uintptr_t>(k.keep)) & 15u) == 0;
    }
    if (tma) {
        static_assert(sizeof(mssc::TmaSmem) <= 48 * 1024, "the TMA variant stays within the default dynamic shared memory limit");
        mssc::compact_keyframes_tma_kernel<<<std::min(nkf, h->sm_count * 8), mssc::kT, sizeof(mssc::TmaSmem), h->stream>>>(
            reinterpret_cast<const mssc::KfPayload*>(up.p), nkf, m ? m->slot_mp.p : nullptr, m ? m->S : 0, reinterpret_cast<int*>(up.p + off_out));
    } else {
        mssc::compact_keyframes_kernel<<<std::min(nkf, h->sm_count * 8), mssc::kT, 0, h->stream>>>(
            reinterpret_cast<const mssc::KfPayload*>(up.p), nkf, m ? m->slot_mp.p : nullptr, m ? m->S : 0, reinterpret_cast<int*>(up.p + off_out));
    }
    MSS_CUDA(h, cudaGetLastError());
    h->stats.kernel_launches += 1;
    if (m) {                                                        // the mirror's own rows follow (EraseBadDescriptor semantics)
        const MirrorDev D = dev_of(m);
        for (int i = 0; i < nkf; ++i) mk_kf_compact<<<1, kT, 0, h->stream>>>(D, kfs[i].kf);
        MSS_CUDA(h, cudaGetLastError());
        h->stats.kernel_launches += nkf;
    }
    MSS_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    MSS_CUDA(h, cudaMemcpyAsync(n_out, up.p + off_out, (size_t)nkf * 4, cudaMemcpyDeviceToHost, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->stats.last_device_ms = ms;      // the compaction kernel(s)
    return MSS_OK;
}

int mss_compact_keyframes(mss_handle* h, int32_t nkf, const mss_kf_payload* kfs, int32_t* n_out) {
    if (!h) return MSS_E_BADARG;
    return compact_impl(h, nullptr, nkf, kfs, n_out);
}

int mss_mirror_compact_keyframes(mss_mirror* m, int32_t nkf, const mss_kf_payload* kfs, int32_t* n_out) {
    if (!m) return MSS_E_BADARG;
    return compact_impl(m->h, m, nkf, kfs, n_out);
}

int mss_mirror_get_stats(const mss_mirror* m, mss_mirror_stats* out) {
    if (!m || !out) return MSS_E_BADARG;
    *out = m->stats;
    out->n_keyframes = m->n_kf; out->n_map_points = m->n_mp; out->slots_per_kf = m->S;
    out->device_bytes = m->h->device_bytes;
    return MSS_OK;
}

}  // extern "C"

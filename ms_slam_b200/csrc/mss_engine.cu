// mss_engine.cu -- host side of libmss.so: device arena, batch layout, ONE cooperative launch per call (host views are
// copied on a second stream while it runs, per-window ready flags), result hand-back (only what was asked for crosses
// PCIe), NCCL all-gather of the result slots, connected components, and the extern "C" entry points of include/mss.h.
//
// Replaces the GUROBI environment/model objects of the reference (GRBEnv mGRBEnv, GRBModel model:
// /root/reference/include/MapSparsification.h:59, /root/reference/src/MapSparsification.cc:6,20,61,153-157).
#include "mss_internal.h"
#include "mss_components.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <mutex>
#include <vector>

namespace {

using mss::Ctrl;
using mss::Params;
using mss::WinDesc;
using mss::WinState;
using mssi::DevBuf;
using mssi::align_up;
using mssi::ensure;
using mssi::ensure_pinned;
using mssi::release;

// ---- NCCL through dlopen: no link-time dependency, single-GPU users never touch it ------------------------------------
struct NcclUniqueId { char internal[128]; };
using mssi::NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
const int kNcclUint32 = 3;   // ncclUint32 (nccl.h ncclDataType_t)

bool load_nccl(std::string& err) {
    if (g_nccl.ok) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
    g_nccl.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(NcclComm))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.CommDestroy && g_nccl.GetErrorString;
    if (!g_nccl.ok) err = "libnccl is missing required symbols";
    return g_nccl.ok;
}

}  // namespace

// device-resident result arrays are filled by one small kernel (one CTA per window) instead of three copies per window
struct ResCopy {
    const uint32_t* slot;
    uint32_t* keep;
    int32_t* cov;
    int32_t* slack;
    int words_keep, rows;
};
__global__ void scatter_results_kernel(const ResCopy* rc) {
    const ResCopy r = rc[blockIdx.x];
    const uint32_t* src = r.slot + mss::kHdrWords;
    if (r.keep) for (int i = threadIdx.x; i < r.words_keep; i += blockDim.x) r.keep[i] = src[i];
    if (r.cov) for (int i = threadIdx.x; i < r.rows; i += blockDim.x) r.cov[i] = (int32_t)src[r.words_keep + i];
    if (r.slack) for (int i = threadIdx.x; i < r.rows; i += blockDim.x) r.slack[i] = (int32_t)src[r.words_keep + r.rows + i];
}

namespace {

inline const void* slots_ptr(const mss_window_view& v) {
    return v.layout == MSS_LAYOUT_PACKED16 ? (const void*)v.slots16 : v.layout == MSS_LAYOUT_PACKED ? (const void*)v.slots : (const void*)v.feat_mp;
}
inline size_t slots_bytes(const mss_window_view& v) { return (size_t)v.F * (v.layout == MSS_LAYOUT_PACKED16 ? 2 : 4); }
inline int packed_code(const mss_window_view& v) { return v.layout == MSS_LAYOUT_PACKED16 ? 2 : v.layout == MSS_LAYOUT_PACKED ? 1 : 0; }

// where a window's result arrays live: by default where its view lives; mss_window_view::result_memory overrides
inline bool res_on_device(const mss_window_view& v) {
    return v.result_memory == MSS_RESULT_DEVICE || (v.result_memory == MSS_RESULT_SAME && v.memory == MSS_MEM_DEVICE);
}

struct SlotLayout { int words_keep, rows, total; };
inline SlotLayout slot_of(const mss_window_view& v) {
    SlotLayout s;
    s.words_keep = (v.M + 31) / 32;
    s.rows = v.K + v.H;
    s.total = mss::kHdrWords + s.words_keep + 2 * s.rows;
    return s;
}

void fill_failsafe(const mss_window_view& v, mss_result& r, int status, bool host_mem, mss_handle* h) {
    // deleting map points is irreversible: on any error keep everything (bitmask all ones)
    const size_t words = (size_t)(v.M + 31) / 32;
    if (r.keep_bits) {
        if (host_mem) memset(r.keep_bits, 0xFF, words * 4);
        else cudaMemsetAsync(r.keep_bits, 0xFF, words * 4, h->stream);
    }
    const size_t rows = (size_t)v.K + v.H;
    if (host_mem) {
        if (r.kf_cov) memset(r.kf_cov, 0, rows * 4);
        if (r.kf_slack) memset(r.kf_slack, 0, rows * 4);
    } else {
        if (r.kf_cov) cudaMemsetAsync(r.kf_cov, 0, rows * 4, h->stream);
        if (r.kf_slack) cudaMemsetAsync(r.kf_slack, 0, rows * 4, h->stream);
    }
    r.objective = NAN; r.dual_bound = NAN; r.sum_cost = 0; r.uncovered_cells = 0; r.total_slack = 0;
    r.n_max = 0; r.n_vars = 0; r.n_cells = 0; r.nnz = 0; r.n_kept = 0; r.rounds = 0;
    r.status = status; r.time_build_us = 0; r.time_solve_us = 0;
}

int validate_view(mss_handle* h, const mss_window_view& v, bool owned) {
    if (v.K < 0 || v.H < 0 || v.M < 0) { h->err = "view: negative K/H/M"; return MSS_E_BADARG; }
    if (!owned) return MSS_OK;
    if (v.F < 0 || v.O < 0) { h->err = "view: negative F/O"; return MSS_E_BADARG; }
    if (v.memory != MSS_MEM_HOST && v.memory != MSS_MEM_DEVICE) { h->err = "view: bad memory kind"; return MSS_E_BADARG; }
    if (v.result_memory != MSS_RESULT_SAME && v.result_memory != MSS_RESULT_HOST && v.result_memory != MSS_RESULT_DEVICE) { h->err = "view: bad result_memory"; return MSS_E_BADARG; }
    if (v.layout != MSS_LAYOUT_SOA && v.layout != MSS_LAYOUT_PACKED && v.layout != MSS_LAYOUT_PACKED16) { h->err = "view: bad layout"; return MSS_E_BADARG; }
    const bool pkv = v.layout != MSS_LAYOUT_SOA;
    const void* pk_slots = v.layout == MSS_LAYOUT_PACKED16 ? (const void*)v.slots16 : (const void*)v.slots;
    if (!v.feat_ptr || (!pkv && !v.mp_obs_ptr)) { h->err = "view: feat_ptr / mp_obs_ptr is NULL"; return MSS_E_BADARG; }
    if (pkv && v.H > 4095) { h->err = "view: more than 4095 outside keyframes in the packed layout (use MSS_LAYOUT_SOA)"; return MSS_E_BADARG; }
    const bool null_arr = pkv
        ? ((v.F > 0 && !pk_slots) || (v.M > 0 && !v.mp_nobs16) || (v.O > 0 && !v.obs_pairs))
        : ((v.F > 0 && (!v.feat_mp || !v.feat_cell)) || (v.M > 0 && !v.mp_nobs) || (v.O > 0 && !v.mp_obs_kf));
    if (null_arr || (v.H > 0 && !v.okf_total)) { h->err = "view: NULL array with non-zero size"; return MSS_E_BADARG; }
    if (v.memory == MSS_MEM_HOST) {
        if (v.feat_ptr[0] != 0 || v.feat_ptr[v.K] != v.F) { h->err = "view: feat_ptr must start at 0 and end at F"; return MSS_E_BADARG; }
        if (!pkv && (v.mp_obs_ptr[0] != 0 || v.mp_obs_ptr[v.M] != v.O)) { h->err = "view: mp_obs_ptr must start at 0 and end at O"; return MSS_E_BADARG; }
    }
    return MSS_OK;
}

}  // namespace

int mssi::solve_batch_impl(mss_handle* h, int nwin, const mss_window_view* views, mss_result* results) {
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    h->err.clear();
    if (!views || !results || nwin < 0) { h->err = "NULL views/results or negative nwin"; return MSS_E_BADARG; }
    if (nwin == 0) return MSS_OK;
    MSS_CUDA(h, cudaSetDevice(h->device));
    const int nranks = h->nranks, rank = h->rank;

    // ---- layout ---------------------------------------------------------------------------------------------------
    std::vector<int> local;                 // windows solved on this rank
    std::vector<int> out_off(nwin, 0);
    int slot_stride = 0;
    std::vector<int> rejected;              // owned windows whose view failed host-side validation (multi-rank only)
    std::string first_reject;
    for (int w = 0; w < nwin; ++w) {
        const bool owned = (w % nranks) == rank;
        int rc = validate_view(h, views[w], owned);
        // size limits depend on K / H / M only, which every rank knows: all ranks fail alike, before anything collective
        if (rc == MSS_OK && views[w].M > mss::kMaxWindowMps) { h->err = "view: more than 2^20 map points in one window"; return MSS_E_BADARG; }
        if (rc == MSS_OK && views[w].K + views[w].H > mss::kMaxWindowRows) { h->err = "view: more than 65535 keyframe rows in one window"; return MSS_E_BADARG; }
        if (rc != MSS_OK) {
            // single rank: nothing has been started, fail the call.  With a communicator the other ranks are already on their
            // way into the all-gather: this rank must follow, so the window is only left out (its slot stays unwritten and
            // every rank reports MSS_E_BADARG for it, all map points kept).  Sizes must be sane on every rank, though.
            if (nranks == 1 || !owned || views[w].K < 0 || views[w].H < 0 || views[w].M < 0) return rc;
            rejected.push_back(w);
            if (first_reject.empty()) first_reject = h->err;
        } else if (owned) {
            local.push_back(w);
        }
        slot_stride = std::max(slot_stride, slot_of(views[w]).total);
    }
    const int spr = (nwin + nranks - 1) / nranks;          // slots per rank
    size_t out_words = 0;
    if (nranks == 1) {
        for (int w = 0; w < nwin; ++w) { out_off[w] = (int)out_words; out_words += (size_t)slot_of(views[w]).total; }
    } else {
        for (int w = 0; w < nwin; ++w) out_off[w] = ((w % nranks) * spr + w / nranks) * slot_stride;
        out_words = (size_t)nranks * spr * slot_stride;
    }
    const int nl = (int)local.size();
    long long Ktot = 0, Htot = 0, Mpad = 0, Ftot = 0, Otot = 0;
    size_t stage_bytes = 0;
    for (int w : local) {
        const mss_window_view& v = views[w];
        Ktot += v.K; Htot += v.H; Mpad += (long long)align_up((size_t)std::max(v.M, 1), mss::kVarTile); Ftot += v.F; Otot += v.O;
        if (v.memory == MSS_MEM_HOST) {
            const bool pk = v.layout != MSS_LAYOUT_SOA;
            stage_bytes = align_up(stage_bytes, 128);     // a window's staging region starts on its own cache line (see `stage`)
            stage_bytes += align_up((size_t)(v.K + 1) * 4, 16) + align_up(slots_bytes(v), 16) + (pk ? 0 : align_up((size_t)v.F * 2, 16)) +
                           align_up((size_t)v.M * (pk ? (v.nobs8 ? 1 : 2) : 4), 16) + (pk ? 0 : align_up((size_t)(v.M + 1) * 4, 16)) +
                           align_up((size_t)v.O * 4, 16) + align_up((size_t)v.H * 4, 16) +
                           (v.mp_tie ? align_up((size_t)v.M * 4, 16) : 0);
        }
    }
    if (Mpad > 0x7FFFFF00LL || Ktot + Htot > 0x7FFFFF00LL || Ftot + Otot > 0x7FFFFF00LL) { h->err = "batch too large for 32-bit indices"; return MSS_E_BADARG; }
    const int Rtot = (int)(Ktot + Htot);

    // ---- host views travel while the kernel runs: the copies go to the copy stream in queue order, each window followed
    //      by a 4-byte copy that sets its ready flag; the persistent kernel is launched at once and a group waits for the flag
    //      of the window it draws.  Device-resident views: no flags.
    bool any_host = false;
    for (int w : local) any_host = any_host || views[w].memory == MSS_MEM_HOST;
    const bool gated = any_host && h->overlap_copy && nl > 1;
    // ---- groups: the grid is cut into equal groups of CTAs; every group pulls windows from one queue (largest first), so a
    //      group that draws a short solve simply takes the next window.  One window -> one group, whole grid. --------------
    const int max_grid = std::max(1, h->max_ctas_per_sm * h->sm_count);
    auto work = [&](int i) { const mss_window_view& v = views[local[i]]; return 1.0 + (double)v.F + (double)v.O + 0.25 * (double)v.M; };
    struct Plan { int gsize, ngroups, grid; size_t off_grp, off_cta, off_gwin; std::vector<int> order; } plan;
    size_t meta_bytes = align_up((size_t)std::max(nl, 1) * sizeof(WinDesc), 16), sync_words = 0;
    {
        plan.order.resize(nl);
        for (int i = 0; i < nl; ++i) plan.order[i] = i;
        std::stable_sort(plan.order.begin(), plan.order.end(), [&](int x, int y) { return work(x) > work(y); });
        int cap = 1;        // a window cannot use more CTAs than it has rows or variable tiles
        for (int i : plan.order) { const mss_window_view& v = views[local[i]]; cap = std::max(cap, std::max(v.K + v.H, (v.M + mss::kVarTile - 1) / mss::kVarTile)); }
        int gsize = h->group_ctas > 0 ? h->group_ctas : std::max(16, max_grid / std::max(nl, 1));
        gsize = std::max(1, std::min(std::min(gsize, cap), max_grid));
        int ngroups = std::max(1, std::min(std::max(nl, 1), max_grid / gsize));
        if (h->group_ctas <= 0 && nl > ngroups) {
            // several windows per group: even out the number of windows per group, then give the groups all the CTAs
            const int waves = (nl + ngroups - 1) / ngroups;
            ngroups = (nl + waves - 1) / waves;
            gsize = std::max(1, std::min(cap, max_grid / ngroups));
        }
        plan.gsize = gsize;
        plan.ngroups = ngroups;
        plan.grid = ngroups * gsize;
        plan.off_grp = meta_bytes;
        plan.off_cta = plan.off_grp + align_up((size_t)ngroups * sizeof(mss::GroupDesc), 16);
        plan.off_gwin = plan.off_cta + align_up((size_t)plan.grid * 4, 16);
        meta_bytes = plan.off_gwin + align_up((size_t)std::max(nl, 1) * 4, 16);
        sync_words = 32 + (size_t)ngroups * 32;
    }
    const size_t ready_off = sync_words;
    sync_words += align_up((size_t)std::max(nl, 1), 32);
    int grid = plan.grid;

    // result scatter table (device-resident result arrays), appended to the descriptor blob
    int n_dev_res = 0;
    for (int w = 0; w < nwin; ++w)
        if (res_on_device(views[w]) && (results[w].keep_bits || results[w].kf_cov || results[w].kf_slack)) ++n_dev_res;
    const size_t off_res = meta_bytes;
    meta_bytes += align_up((size_t)std::max(n_dev_res, 1) * sizeof(ResCopy), 16);
    int rc;
    if ((rc = ensure(h, h->meta, meta_bytes))) return rc;
    if ((rc = ensure(h, h->ws, (size_t)std::max(nl, 1)))) return rc;
    if ((rc = ensure(h, h->st, (size_t)Mpad + 16))) return rc;
    if ((rc = ensure(h, h->acc, (size_t)Mpad + 16))) return rc;
    if ((rc = ensure(h, h->gain, (size_t)Mpad + 16))) return rc;
    if ((rc = ensure(h, h->deg, (size_t)Mpad + 16))) return rc;
    if ((rc = ensure(h, h->seen, (size_t)Mpad + 16))) return rc;
    if ((rc = ensure(h, h->vlist, (size_t)2 * Mpad + 16))) return rc;
    if (h->trace_on && (rc = ensure(h, h->trace, (size_t)std::max(nl, 1) * mss::kTraceCap))) return rc;
    if ((rc = ensure(h, h->ent, (size_t)(Ftot + Otot) + 16))) return rc;
    if ((rc = ensure(h, h->live, (size_t)(Ftot + Otot) + 16))) return rc;
    if ((rc = ensure(h, h->rows, (size_t)7 * ((size_t)Rtot + 16)))) return rc;
    if (h->want_bound) {
        if ((rc = ensure(h, h->b_snap, (size_t)(Ftot + Otot) + 16))) return rc;
        if ((rc = ensure(h, h->b_rows, (size_t)2 * ((size_t)Rtot + 16)))) return rc;
        if ((rc = ensure(h, h->b_vars, (size_t)2 * ((size_t)Mpad + 16)))) return rc;
    }
    if ((rc = ensure(h, h->out, out_words + 16))) return rc;
    if ((rc = ensure(h, h->stage, stage_bytes + 16))) return rc;
    if ((rc = ensure(h, h->sync, sync_words))) return rc;
    h->ctrl = reinterpret_cast<Ctrl*>(h->sync.p);
    if ((rc = ensure_pinned(h, (void**)&h->h_meta, &h->h_meta_cap, meta_bytes))) return rc;
    if ((rc = ensure_pinned(h, (void**)&h->h_out, &h->h_out_cap, (out_words + 16) * 4))) return rc;

    // ---- descriptors -----------------------------------------------------------------------------------------------------
    // Staging: the arrays of a window lie back to back at 16-byte boundaries (so a blob travels with one copy); every
    // window's region starts at a 128-byte boundary, so no cache line holds data of two windows -- in the gated mode a
    // window may still be in flight while the kernel already reads its neighbour.
    cudaStream_t cstream = gated ? h->copy_stream : h->stream;
    WinDesc* hd = reinterpret_cast<WinDesc*>(h->h_meta);
    int64_t h2d = 0;
    size_t soff = 0;
    auto stage = [&](const void* src, size_t bytes) -> const void* {
        uint8_t* dst = h->stage.p + soff;
        soff += align_up(bytes, 16);
        h2d += (int64_t)bytes;
        return dst;
    };
    int row_base = 0, slot_base = 0, obs_base = 0, var_base = 0;
    for (int i = 0; i < nl; ++i) {
        const mss_window_view& v = views[local[i]];
        WinDesc d;
        memset(&d, 0, sizeof(d));
        const bool pk = v.layout != MSS_LAYOUT_SOA;
        d.packed = packed_code(v);
        d.n_max_floor = v.n_max_floor;
        d.nobs8 = (pk && v.nobs8) ? 1 : 0;
        if (v.memory == MSS_MEM_HOST) {
            soff = align_up(soff, 128);
            d.feat_ptr = (const int*)stage(v.feat_ptr, (size_t)(v.K + 1) * 4);
            d.feat_mp = (const int*)stage(slots_ptr(v), slots_bytes(v));
            d.feat_cell = pk ? nullptr : (const uint16_t*)stage(v.feat_cell, (size_t)v.F * 2);
            d.mp_nobs = (const int*)stage(pk ? (const void*)v.mp_nobs16 : (const void*)v.mp_nobs, (size_t)v.M * (pk ? (v.nobs8 ? 1 : 2) : 4));
            d.mp_obs_ptr = pk ? nullptr : (const int*)stage(v.mp_obs_ptr, (size_t)(v.M + 1) * 4);
            d.mp_obs_kf = (const int*)stage(pk ? (const void*)v.obs_pairs : (const void*)v.mp_obs_kf, (size_t)v.O * 4);
            d.okf_total = (const int*)stage(v.okf_total, (size_t)v.H * 4);
            d.mp_tie = v.mp_tie ? (const unsigned*)stage(v.mp_tie, (size_t)v.M * 4) : nullptr;      // (behind the blob: its own copy)
        } else if (pk) {
            d.feat_ptr = v.feat_ptr; d.feat_mp = (const int*)slots_ptr(v); d.feat_cell = nullptr; d.mp_nobs = (const int*)v.mp_nobs16;
            d.mp_obs_ptr = nullptr; d.mp_obs_kf = (const int*)v.obs_pairs; d.okf_total = v.okf_total;
        } else {
            d.feat_ptr = v.feat_ptr; d.feat_mp = v.feat_mp; d.feat_cell = v.feat_cell; d.mp_nobs = v.mp_nobs;
            d.mp_obs_ptr = v.mp_obs_ptr; d.mp_obs_kf = v.mp_obs_kf; d.okf_total = v.okf_total;
        }
        if (v.memory != MSS_MEM_HOST) d.mp_tie = v.mp_tie;
        d.K = v.K; d.H = v.H; d.M = v.M; d.F = v.F; d.O = v.O;
        d.row_base = row_base; d.slot_base = slot_base; d.obs_base = obs_base; d.var_base = var_base;
        d.out_off = out_off[local[i]];
        hd[i] = d;
        row_base += v.K + v.H; slot_base += v.F; obs_base += v.O;
        var_base += (int)align_up((size_t)std::max(v.M, 1), mss::kVarTile);
    }
    {
        mss::GroupDesc* h_grp = reinterpret_cast<mss::GroupDesc*>(h->h_meta + plan.off_grp);
        int* h_cta_grp = reinterpret_cast<int*>(h->h_meta + plan.off_cta);
        int* h_gwin = reinterpret_cast<int*>(h->h_meta + plan.off_gwin);
        int cta = 0;
        for (int g = 0; g < plan.ngroups; ++g) {
            h_grp[g].cta0 = cta; h_grp[g].ncta = plan.gsize; h_grp[g].pad_[0] = h_grp[g].pad_[1] = 0;
            for (int k = 0; k < plan.gsize; ++k) h_cta_grp[cta++] = g;
        }
        for (int i = 0; i < nl; ++i) h_gwin[i] = plan.order[i];
    }
    // descriptors first, on the compute stream (the kernel needs them at once)
    {
        ResCopy* hr = reinterpret_cast<ResCopy*>(h->h_meta + off_res);
        int q = 0;
        for (int w = 0; w < nwin; ++w) {
            const mss_result& r = results[w];
            if (!res_on_device(views[w]) || !(r.keep_bits || r.kf_cov || r.kf_slack)) continue;
            const SlotLayout sl = slot_of(views[w]);
            hr[q++] = ResCopy{h->out.p + out_off[w], r.keep_bits, r.kf_cov, r.kf_slack, sl.words_keep, sl.rows};
        }
    }
    MSS_CUDA(h, cudaMemcpyAsync(h->meta.p, h->h_meta, meta_bytes, cudaMemcpyHostToDevice, h->stream));
    h2d += (int64_t)meta_bytes;
    // (cleared by a DMA copy of pinned zeros: a memset kernel would have to wait for a persistent kernel of another handle
    // on this device to drain, and with it this call's staging copies)
    if (sync_words * 4 > h->h_zero_cap) {
        const size_t old = h->h_zero_cap;
        if ((rc = ensure_pinned(h, (void**)&h->h_zero, &h->h_zero_cap, sync_words * 4))) return rc;
        (void)old;
        memset(h->h_zero, 0, h->h_zero_cap);
    }
    MSS_CUDA(h, cudaMemcpyAsync(h->sync.p, h->h_zero, sync_words * 4, cudaMemcpyHostToDevice, h->stream));
    for (int w : rejected) MSS_CUDA(h, cudaMemsetAsync(h->out.p + out_off[w], 0, (size_t)mss::kHdrWords * 4, h->stream));
    if (gated) MSS_CUDA(h, cudaEventRecord(h->ev_ready, h->stream));      // (the staging stream waits for it right before the copies)
    // ---- staging copies, in queue order ------------------------------------------------------------------------------------
    cudaError_t copy_err = cudaSuccess;
    auto copy_window = [&](int i) {
        const mss_window_view& v = views[local[i]];
        if (v.memory != MSS_MEM_HOST) return;
        const WinDesc& d = hd[i];
        const bool pk = v.layout != MSS_LAYOUT_SOA;
        auto put = [&](const void* dst, const void* src, size_t bytes) {
            if (!bytes || copy_err != cudaSuccess) return;
            copy_err = cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, cstream);
        };
        // a view whose arrays lie back to back on the host, in staging order and each at the next 16-byte boundary (one
        // pinned blob per window, as FlattenWindow lays them out), travels with ONE copy
        {
            const void* src[7] = {v.feat_ptr, slots_ptr(v), pk ? nullptr : (const void*)v.feat_cell,
                                  pk ? (const void*)v.mp_nobs16 : (const void*)v.mp_nobs, pk ? nullptr : (const void*)v.mp_obs_ptr,
                                  pk ? (const void*)v.obs_pairs : (const void*)v.mp_obs_kf, v.okf_total};
            const size_t len[7] = {(size_t)(v.K + 1) * 4, slots_bytes(v), pk ? 0 : (size_t)v.F * 2, (size_t)v.M * (pk ? (v.nobs8 ? 1 : 2) : 4),
                                   pk ? 0 : (size_t)(v.M + 1) * 4, (size_t)v.O * 4, (size_t)v.H * 4};
            const uint8_t* base = static_cast<const uint8_t*>(src[0]);
            size_t off = 0, end = 0;
            bool blob = (reinterpret_cast<uintptr_t>(base) & 15u) == 0;
            for (int a = 0; a < 7 && blob; ++a) {
                if (pk && (a == 2 || a == 4)) continue;
                if (len[a]) { if (static_cast<const uint8_t*>(src[a]) != base + off) blob = false; end = off + len[a]; }
                off += align_up(len[a], 16);
            }
            if (blob) { put(d.feat_ptr, base, end); if (v.mp_tie) put(d.mp_tie, v.mp_tie, (size_t)v.M * 4); return; }
        }
        put(d.feat_ptr, v.feat_ptr, (size_t)(v.K + 1) * 4);
        put(d.feat_mp, slots_ptr(v), slots_bytes(v));
        if (!pk) put(d.feat_cell, v.feat_cell, (size_t)v.F * 2);
        put(d.mp_nobs, pk ? (const void*)v.mp_nobs16 : (const void*)v.mp_nobs, (size_t)v.M * (pk ? (v.nobs8 ? 1 : 2) : 4));
        if (!pk) put(d.mp_obs_ptr, v.mp_obs_ptr, (size_t)(v.M + 1) * 4);
        put(d.mp_obs_kf, pk ? (const void*)v.obs_pairs : (const void*)v.mp_obs_kf, (size_t)v.O * 4);
        put(d.okf_total, v.okf_total, (size_t)v.H * 4);
        if (v.mp_tie) put(d.mp_tie, v.mp_tie, (size_t)v.M * 4);
    };
    unsigned* d_ready = h->sync.p + ready_off;
    if (!gated) for (int q = 0; q < nl; ++q) copy_window(plan.order[q]);
    MSS_CUDA(h, copy_err);
    MSS_CUDA(h, cudaGetLastError());

    // ---- launch -----------------------------------------------------------------------------------------------------------
    Params P;
    memset(&P, 0, sizeof(P));
    P.win = reinterpret_cast<const WinDesc*>(h->meta.p);
    P.ws = h->ws.p;
    P.st = h->st.p; P.acc = h->acc.p; P.gain = h->gain.p; P.deg = h->deg.p; P.seen = h->seen.p;
    P.vlist = h->vlist.p; P.Mpad = (int)Mpad;
    P.trace = h->trace_on ? h->trace.p : nullptr;
    P.ent = h->ent.p; P.live = h->live.p;
    {
        const size_t rs = (size_t)Rtot + 16;
        P.row_off = h->rows.p; P.ent_n = h->rows.p + rs; P.live_n = h->rows.p + 2 * rs; P.row_need = h->rows.p + 3 * rs;
        P.row_cov = h->rows.p + 4 * rs; P.row_ncell = h->rows.p + 5 * rs; P.ocursor = h->rows.p + 6 * rs;
    }
    if (h->want_bound) {
        P.b_snap = h->b_snap.p;
        P.b_snap_n = h->b_rows.p; P.b_snap_d = h->b_rows.p + ((size_t)Rtot + 16);
        P.b_share = h->b_vars.p; P.b_red = h->b_vars.p + ((size_t)Mpad + 16);
    }
    P.out = h->out.p;
    P.Ftot = (int)Ftot;
    P.N = h->cfg.min_points;
    P.max_rounds = h->cfg.max_rounds; P.all_rule_steps = h->cfg.all_rule_steps; P.max_drop_rounds = h->cfg.max_drop_rounds;
    P.stall_den = h->cfg.stall_den;
    P.lam = (double)h->cfg.lambda; P.glam = (double)h->cfg.grid_lambda;
    P.watchdog_ns = h->watchdog_ns;
    P.tail_vars = std::min(h->tail_vars, mss::kTailVars); P.tail_ents = std::min(h->tail_ents, mss::kTailEnts);
    P.ready = gated ? d_ready : nullptr;
    P.w1_tma = h->w1_tma;

    // Once the persistent kernel is launched it may be spinning on ready flags that a failed copy will never set: any
    // error after the launch raises the device-side abort flag from the host and drains both streams before returning
    // (the staging buffers and pinned mirrors are reused by the next call).
    bool launched = false;
    auto abort_launch = [&]() {
        if (!launched) return;
        cudaMemcpyAsync(&h->ctrl->abort, h->h_one, 4, cudaMemcpyHostToDevice, h->copy_stream);
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
    };
#define MSS_CUDA_LAUNCHED(h, expr)                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            abort_launch();                                                                        \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                         \
            return MSS_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

    float dev_ms = 0.f;
    if (nl > 0) {
        if (h->trace_on) MSS_CUDA(h, cudaMemsetAsync(h->trace.p, 0, (size_t)nl * mss::kTraceCap * sizeof(uint2), h->stream));
        MSS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
        P.grp = reinterpret_cast<const mss::GroupDesc*>(h->meta.p + plan.off_grp);
        P.cta_grp = reinterpret_cast<const int*>(h->meta.p + plan.off_cta);
        P.gwin = reinterpret_cast<const int*>(h->meta.p + plan.off_gwin);
        P.ctrl = h->ctrl;
        P.gbar = h->sync.p + 32;
        P.nwin = nl; P.ngroups = plan.ngroups;
        void* args[] = {(void*)&P};
        MSS_CUDA(h, cudaLaunchCooperativeKernel((const void*)mss::mss_persistent_kernel, dim3(plan.grid), dim3(mss::kThreads), args, mss::kSmemBytes, h->stream));
        h->stats.kernel_launches += 1;
        launched = true;
        MSS_CUDA_LAUNCHED(h, cudaEventRecord(h->ev1, h->stream));
        if (gated) {
            // the kernel is running (or queued); feed it: window after window in queue order, flag after data.  The whole
            // batch is enqueued under the device's staging mutex: batches of different handles travel one after the other
            std::lock_guard<std::mutex> lock(*static_cast<std::mutex*>(h->copy_mutex));
            copy_err = cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0);
            // (the ready flags of kFlagBatch consecutive windows of the queue are raised by ONE small copy behind the last of
            // them: every tiny copy between two blobs costs the DMA stream ~2 us)
            constexpr int kFlagBatch = 4;
            for (int q = 0; q < nl && copy_err == cudaSuccess; ++q) {
                copy_window(plan.order[q]);
                if (copy_err == cudaSuccess && ((q + 1) % kFlagBatch == 0 || q + 1 == nl)) {
                    const int q0 = q - (q % kFlagBatch);
                    copy_err = cudaMemcpyAsync(d_ready + q0, h->h_one, (size_t)(q + 1 - q0) * 4, cudaMemcpyHostToDevice, h->copy_stream);
                }
            }
            if (copy_err == cudaSuccess) copy_err = cudaEventRecord(h->ev_copied, h->copy_stream);
            MSS_CUDA_LAUNCHED(h, copy_err);
            h2d += 4ll * nl;
        }
    } else {
        grid = 0;
    }
    // ---- all-gather of the result slots (keep bits + row coverage only) ---------------------------------------------------
    if (nranks > 1) {
        const size_t count = (size_t)spr * slot_stride;
        const int nrc = g_nccl.AllGather(h->out.p + (size_t)rank * count, h->out.p, count, kNcclUint32, h->comm, h->stream);
        if (nrc != 0) { abort_launch(); h->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(nrc); return MSS_E_NCCL; }
    }
    // ---- hand-back ------------------------------------------------------------------------------------------------------
    // Only what the caller asked for crosses PCIe: the whole slot of a window whose result arrays are host buffers, the
    // 64-byte header (status + counters) of every other window (device-resident results, or windows of other ranks the
    // caller did not ask arrays for -- their bitmasks stay in the all-gathered device buffer).
    int64_t d2h = 0;
    {
        auto wants_full = [&](int w) {
            const mss_result& r = results[w];
            return !res_on_device(views[w]) && (r.keep_bits || r.kf_cov || r.kf_slack);
        };
        int n_full = 0;
        bool uniform = true;
        for (int w = 0; w < nwin; ++w) {
            const mss_result& r = results[w];
            if (!res_on_device(views[w]) && (r.keep_bits || r.kf_cov || r.kf_slack)) ++n_full;
            if (w > 0 && out_off[w] - out_off[w - 1] != out_off[1] - out_off[0]) uniform = false;
        }
        if (nranks > 1) uniform = true;                        // rank-major slots with one stride
        const size_t stride = nranks > 1 ? (size_t)slot_stride : (nwin > 1 ? (size_t)(out_off[1] - out_off[0]) : (size_t)slot_of(views[0]).total);
        const size_t nslots = nranks > 1 ? (size_t)nranks * spr : (size_t)nwin;
        // several ranks, and the caller asks arrays exactly for the windows this rank solved (the usual case: the others'
        // bitmasks stay in the all-gathered device buffer): the rank's slots are contiguous, one copy brings them all
        bool own_only = nranks > 1 && nl > 0 && n_full == nl;
        for (int w = 0; w < nwin && own_only; ++w) own_only = wants_full(w) == ((w % nranks) == rank);
        if (n_full == nwin) {
            MSS_CUDA_LAUNCHED(h, cudaMemcpyAsync(h->h_out, h->out.p, out_words * 4, cudaMemcpyDeviceToHost, h->stream));
            d2h += (int64_t)(out_words * 4);
        } else if (own_only) {
            const size_t region = (size_t)spr * slot_stride, off = (size_t)rank * region;
            MSS_CUDA_LAUNCHED(h, cudaMemcpy2DAsync(h->h_out, stride * 4, h->out.p, stride * 4, (size_t)mss::kHdrWords * 4, nslots,
                                                   cudaMemcpyDeviceToHost, h->stream));
            MSS_CUDA_LAUNCHED(h, cudaMemcpyAsync(h->h_out + off, h->out.p + off, region * 4, cudaMemcpyDeviceToHost, h->stream));
            d2h += (int64_t)(nslots * mss::kHdrWords * 4 + region * 4);
        } else {
            if (uniform) {      // headers of all slots in one strided copy
                MSS_CUDA_LAUNCHED(h, cudaMemcpy2DAsync(h->h_out, stride * 4, h->out.p, stride * 4, (size_t)mss::kHdrWords * 4, nslots,
                                              cudaMemcpyDeviceToHost, h->stream));
                d2h += (int64_t)(nslots * mss::kHdrWords * 4);
            }
            for (int w = 0; w < nwin; ++w) {
                const bool full = wants_full(w);
                if (!full && uniform) continue;
                const size_t words = full ? (size_t)slot_of(views[w]).total : (size_t)mss::kHdrWords;
                MSS_CUDA_LAUNCHED(h, cudaMemcpyAsync(h->h_out + out_off[w], h->out.p + out_off[w], words * 4, cudaMemcpyDeviceToHost, h->stream));
                d2h += (int64_t)(words * 4);
            }
        }
    }
    MSS_CUDA_LAUNCHED(h, cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
    d2h += (int64_t)sizeof(Ctrl);
    bool aborted = false;
    if (n_dev_res > 0) {                   // device-resident result buffers are filled device-to-device, all windows in one launch
        scatter_results_kernel<<<n_dev_res, 256, 0, h->stream>>>(reinterpret_cast<const ResCopy*>(h->meta.p + off_res));
        MSS_CUDA_LAUNCHED(h, cudaGetLastError());
        h->stats.kernel_launches += 1;
    }
    MSS_CUDA_LAUNCHED(h, cudaStreamSynchronize(h->stream));
    if (gated) MSS_CUDA_LAUNCHED(h, cudaEventSynchronize(h->ev_copied));       // (only an aborted launch can finish before its copies;
                                                                               //  the stream itself may already carry another handle's batch)
    if (nl > 0) MSS_CUDA(h, cudaEventElapsedTime(&dev_ms, h->ev0, h->ev1));
    unsigned long long row_entries = 0, var_visits = 0;
    if (nl > 0) {
        aborted = h->h_ctrl->abort != 0;
        row_entries = h->h_ctrl->row_entries;
        var_visits = h->h_ctrl->var_visits;
    }
    if (h->trace_on && nl > 0) {
        h->h_trace.resize((size_t)nl * mss::kTraceCap);
        h->h_trace_nwin = nl;
        MSS_CUDA(h, cudaMemcpy(h->h_trace.data(), h->trace.p, h->h_trace.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
    }

    int ret = MSS_OK;
    const double t_build_us = nl > 0 ? (double)(h->h_ctrl->t_build - h->h_ctrl->t_start) * 1e-3 : 0.0;
    const double t_solve_us = nl > 0 ? (double)(h->h_ctrl->t_end - h->h_ctrl->t_build) * 1e-3 : 0.0;
    for (int w = 0; w < nwin; ++w) {
        const mss_window_view& v = views[w];
        mss_result& r = results[w];
        const SlotLayout s = slot_of(v);
        const uint32_t* slot = h->h_out + out_off[w];
        const bool host_mem = !res_on_device(v);
        if (aborted) {
            fill_failsafe(v, r, MSS_E_INTERNAL, host_mem, h);
            ret = MSS_E_INTERNAL;
            h->err = "device watchdog: a group barrier or ready-flag wait exceeded the limit (MSS_WATCHDOG_MS); launch aborted, all map points kept";
            continue;
        }
        if (slot[14] != 0x4D535331u) {
            fill_failsafe(v, r, MSS_E_BADARG, host_mem, h);
            if (ret == MSS_OK) { ret = MSS_E_BADARG; h->err = "window " + std::to_string(w) + ": view failed validation (index out of range, bad pointer table or more than 65535 points in one grid cell); all map points kept"; }
            continue;
        }
        if (host_mem) {
            if (r.keep_bits) memcpy(r.keep_bits, slot + mss::kHdrWords, (size_t)s.words_keep * 4);
            if (r.kf_cov) memcpy(r.kf_cov, slot + mss::kHdrWords + s.words_keep, (size_t)s.rows * 4);
            if (r.kf_slack) memcpy(r.kf_slack, slot + mss::kHdrWords + s.words_keep + s.rows, (size_t)s.rows * 4);
        }
        r.status = (int32_t)slot[0];
        r.rounds = (int32_t)slot[1];
        r.n_max = (int32_t)slot[2];
        r.n_vars = (int32_t)slot[3];
        r.n_cells = (int32_t)slot[4];
        r.nnz = (int32_t)slot[5];
        r.n_kept = (int32_t)slot[6];
        r.uncovered_cells = (int32_t)slot[7];
        r.total_slack = (int32_t)slot[8];
        r.sum_cost = (int64_t)(((uint64_t)slot[10] << 32) | slot[9]);
        r.objective = (double)r.sum_cost + (double)h->cfg.grid_lambda * (double)r.uncovered_cells +
                      (double)h->cfg.lambda * (double)r.total_slack;
        // lower bound proven on the device (mss_bound.cuh): flag 1 = counters of the snapshot, 2 = dominance alone decided
        // every point (the selection is optimal)
        r.dual_bound = NAN;
        if (slot[16] == 2u) r.dual_bound = r.objective;
        else if (slot[16] == 1u) {
            const int64_t cost_in = (int64_t)(((uint64_t)slot[21] << 32) | slot[20]);
            const uint64_t zsum = ((uint64_t)slot[23] << 32) | slot[22], drows = ((uint64_t)slot[25] << 32) | slot[24];
            const int64_t u0 = (int64_t)r.uncovered_cells - (int64_t)slot[18];
            r.dual_bound = (double)cost_in + (double)h->cfg.grid_lambda * (double)u0 + (double)h->cfg.lambda * (double)slot[17] +
                           (double)(zsum + drows) / (double)(1 << mss::kBndScBits);
        }
        r.time_build_us = (float)t_build_us;
        r.time_solve_us = (float)t_solve_us;
        if (r.status != MSS_OK && ret == MSS_OK) { ret = r.status; h->err = "window " + std::to_string(w) + ": round cap reached (selection is feasible but may be loose)"; }
    }
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));    // fail-safe memsets on device results
    h->stats.solves += nl;
    h->stats.last_device_ms = dev_ms;
    h->stats.last_h2d_bytes = h2d;
    h->stats.last_d2h_bytes = d2h;
    h->stats.device_bytes = h->device_bytes;
    h->stats.grid_ctas = grid;
    h->stats.last_row_entries = (int64_t)row_entries;
    h->stats.last_var_visits = (int64_t)var_visits;
    h->stats.last_total_ms = std::chrono::duration<double, std::milli>(clk::now() - t_begin).count();
    h->last_out_off = out_off;
    return ret;
}

// =====================================================================================================================
// C-ABI
// =====================================================================================================================
extern "C" {

int mss_version(void) { return MSS_VERSION; }

// One staging (host -> device) stream per DEVICE, shared by every handle on it.  Handles are single-threaded, so an
// application that wants the copies of one batch to travel while another batch is being solved runs two handles from two
// threads (bench.py `e2e`); their staging copies must then queue FIFO, batch after batch -- copies of two batches that
// interleave would feed both persistent kernels at half the PCIe rate.  The mutex orders the enqueueing of whole batches.
namespace {
struct DeviceCopyStream { cudaStream_t stream = nullptr; std::mutex enqueue; };
std::mutex g_copy_table_mutex;
std::map<int, DeviceCopyStream*> g_copy_table;
DeviceCopyStream* device_copy_stream(int device) {
    std::lock_guard<std::mutex> lock(g_copy_table_mutex);
    auto it = g_copy_table.find(device);
    if (it != g_copy_table.end()) return it->second;
    DeviceCopyStream* d = new DeviceCopyStream();
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) { delete d; return nullptr; }
    g_copy_table[device] = d;                 // lives as long as the process (handles come and go)
    return d;
}
}  // namespace

int mss_create(const mss_config* cfg, mss_handle** out) {
    if (!cfg || !out) return MSS_E_BADARG;
    *out = nullptr;
    if (cfg->min_points < 0 || !(cfg->lambda >= 0.f) || !(cfg->grid_lambda >= 0.f)) return MSS_E_BADARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        fprintf(stderr, "libmss: no usable CUDA device (requested %d of %d); there is no CPU fallback\n", cfg->device, ndev);
        return MSS_E_CUDA;
    }
    mss_handle* h = new (std::nothrow) mss_handle();
    if (!h) return MSS_E_NOMEM;
    h->cfg = *cfg;
    if (h->cfg.max_rounds <= 0) h->cfg.max_rounds = 256;
    if (h->cfg.all_rule_steps < 0) h->cfg.all_rule_steps = 0;
    if (h->cfg.stall_den == 0) h->cfg.stall_den = 8;
    else if (h->cfg.stall_den < 0) h->cfg.stall_den = 0;
    if (h->cfg.max_drop_rounds < 0) h->cfg.max_drop_rounds = 0;
    else if (h->cfg.max_drop_rounds == 0) h->cfg.max_drop_rounds = 16;
    h->device = cfg->device;
    auto fail = [&](const char* what, cudaError_t e) {
        fprintf(stderr, "libmss: %s: %s\n", what, cudaGetErrorString(e));
        mss_destroy(h);
        return MSS_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(h->device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    h->sm_count = prop.multiProcessorCount;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
    if (!coop) { fprintf(stderr, "libmss: device does not support cooperative launch\n"); mss_destroy(h); return MSS_E_CUDA; }
    if ((e = cudaFuncSetAttribute((const void*)mss::mss_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mss::kSmemBytes)) != cudaSuccess)
        return fail("cudaFuncSetAttribute(max dynamic shared memory) (was libmss built for this GPU? it ships sm_100a code only)", e);
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->max_ctas_per_sm, mss::mss_persistent_kernel, mss::kThreads, mss::kSmemBytes)) != cudaSuccess)
        return fail("occupancy query (was libmss built for this GPU? it ships sm_100a code only)", e);
    if (h->max_ctas_per_sm <= 0) { fprintf(stderr, "libmss: kernel does not fit on an SM\n"); mss_destroy(h); return MSS_E_CUDA; }
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return fail("cudaEventCreate", e);
    if (const char* gc = getenv("MSS_GROUP_CTAS")) h->group_ctas = atoi(gc);
    if (const char* tv = getenv("MSS_TAIL_VARS")) h->tail_vars = atoi(tv);
    if (const char* te = getenv("MSS_TAIL_ENTS")) h->tail_ents = atoi(te);
    if (const char* wd = getenv("MSS_WATCHDOG_MS")) { const long long ms = atoll(wd); if (ms > 0) h->watchdog_ns = (unsigned long long)ms * 1000000ull; }
    if ((e = cudaHostAlloc((void**)&h->h_ctrl, sizeof(Ctrl), cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
    {
        DeviceCopyStream* dcs = device_copy_stream(h->device);
        if (!dcs) return fail("cudaStreamCreate (staging stream)", cudaErrorUnknown);
        h->copy_stream = dcs->stream;
        h->copy_mutex = &dcs->enqueue;
    }
    if ((e = cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaHostAlloc((void**)&h->h_one, 64, cudaHostAllocDefault)) != cudaSuccess) return fail("cudaHostAlloc", e);
    for (int i = 0; i < 16; ++i) h->h_one[i] = 1u;
    if (const char* oc = getenv("MSS_OVERLAP_COPY")) h->overlap_copy = atoi(oc);
    if (const char* db = getenv("MSS_DUAL_BOUND")) h->want_bound = atoi(db) != 0;
    if (const char* wt = getenv("MSS_W1_TMA")) h->w1_tma = atoi(wt) != 0;
    h->stats.sm_count = h->sm_count;
    *out = h;
    return MSS_OK;
}

void mss_destroy(mss_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
    release(h->meta); release(h->ws); release(h->st); release(h->acc); release(h->gain); release(h->deg);
    release(h->seen); release(h->vlist); release(h->trace); release(h->ent); release(h->live); release(h->rows); release(h->out); release(h->stage); release(h->sync); release(h->cc);
    release(h->b_snap); release(h->b_rows); release(h->b_vars);
    if (h->h_meta) cudaFreeHost(h->h_meta);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_ready) cudaEventDestroy(h->ev_ready);
    if (h->h_one) cudaFreeHost(h->h_one);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    if (h->h_zero) cudaFreeHost(h->h_zero);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* mss_last_error(const mss_handle* h) { return h ? h->err.c_str() : "null handle"; }

int mss_set_params(mss_handle* h, int32_t min_points, float lambda, float grid_lambda) {
    if (!h) return MSS_E_BADARG;
    if (min_points < 0 || !(lambda >= 0.f) || !(grid_lambda >= 0.f)) { h->err = "bad parameters"; return MSS_E_BADARG; }
    h->cfg.min_points = min_points; h->cfg.lambda = lambda; h->cfg.grid_lambda = grid_lambda;
    return MSS_OK;
}

int mss_set_dual_bound(mss_handle* h, int32_t enable) {
    if (!h) return MSS_E_BADARG;
    h->want_bound = enable != 0;
    return MSS_OK;
}

int mss_solve(mss_handle* h, const mss_window_view* view, mss_result* result) {
    if (!h) return MSS_E_BADARG;
    if (h->nranks > 1) { h->err = "mss_solve on a handle with a communicator: use mss_solve_batch"; return MSS_E_BADARG; }
    return mssi::solve_batch_impl(h, 1, view, result);
}

int mss_components(mss_handle* h, const mss_window_view* view, int32_t* row_label, int32_t* mp_label, int32_t* ncomp, int32_t* n_max) {
    if (!h) return MSS_E_BADARG;
    h->err.clear();
    if (!view || !ncomp) { h->err = "NULL view / ncomp"; return MSS_E_BADARG; }
    const mss_window_view& v = *view;
    int rc = validate_view(h, v, true);
    if (rc != MSS_OK) return rc;
    if (v.M > mss::kMaxWindowMps || v.K + v.H > mss::kMaxWindowRows) { h->err = "view: window too large"; return MSS_E_BADARG; }
    MSS_CUDA(h, cudaSetDevice(h->device));
    const bool pk = v.layout != MSS_LAYOUT_SOA, host = v.memory == MSS_MEM_HOST;
    const int R = v.K + v.H, M = v.M;
    if ((rc = ensure(h, h->cc, (size_t)2 * (R + M) + 16))) return rc;
    if ((rc = ensure(h, h->seen, (size_t)M + 16))) return rc;
    WinDesc d;
    memset(&d, 0, sizeof(d));
    d.K = v.K; d.H = v.H; d.M = v.M; d.F = v.F; d.O = v.O; d.packed = packed_code(v);
    d.nobs8 = (pk && v.nobs8) ? 1 : 0;
    if (host) {
        const void* src[7] = {v.feat_ptr, slots_ptr(v), pk ? nullptr : (const void*)v.feat_cell,
                              pk ? (const void*)v.mp_nobs16 : (const void*)v.mp_nobs, pk ? nullptr : (const void*)v.mp_obs_ptr,
                              pk ? (const void*)v.obs_pairs : (const void*)v.mp_obs_kf, v.okf_total};
        const size_t len[7] = {(size_t)(v.K + 1) * 4, slots_bytes(v), pk ? 0 : (size_t)v.F * 2, (size_t)v.M * (pk ? (v.nobs8 ? 1 : 2) : 4),
                               pk ? 0 : (size_t)(v.M + 1) * 4, (size_t)v.O * 4, (size_t)v.H * 4};
        size_t total = 0;
        for (int a = 0; a < 7; ++a) total += align_up(len[a], 16);
        if ((rc = ensure(h, h->stage, total + 16))) return rc;
        const void* dst[7];
        size_t off = 0;
        for (int a = 0; a < 7; ++a) {
            dst[a] = h->stage.p + off;
            if (len[a]) MSS_CUDA(h, cudaMemcpyAsync(h->stage.p + off, src[a], len[a], cudaMemcpyHostToDevice, h->stream));
            off += align_up(len[a], 16);
        }
        d.feat_ptr = (const int*)dst[0]; d.feat_mp = (const int*)dst[1]; d.feat_cell = pk ? nullptr : (const uint16_t*)dst[2];
        d.mp_nobs = (const int*)dst[3]; d.mp_obs_ptr = pk ? nullptr : (const int*)dst[4]; d.mp_obs_kf = (const int*)dst[5];
        d.okf_total = (const int*)dst[6];
    } else if (pk) {
        d.feat_ptr = v.feat_ptr; d.feat_mp = (const int*)slots_ptr(v); d.mp_nobs = (const int*)v.mp_nobs16; d.mp_obs_kf = (const int*)v.obs_pairs;
        d.okf_total = v.okf_total;
    } else {
        d.feat_ptr = v.feat_ptr; d.feat_mp = v.feat_mp; d.feat_cell = v.feat_cell; d.mp_nobs = v.mp_nobs;
        d.mp_obs_ptr = v.mp_obs_ptr; d.mp_obs_kf = v.mp_obs_kf; d.okf_total = v.okf_total;
    }
    int* parent = h->cc.p;
    int* d_row = h->cc.p + (R + M);
    int* d_mp = d_row + R;
    int* d_misc = d_mp + M;            // ncomp, n_max, err
    uint8_t* isvar = h->seen.p;
    MSS_CUDA(h, cudaMemsetAsync(d_misc, 0, 3 * sizeof(int), h->stream));
    const int T = 256;
    if (R + M > 0) mss::cc_init_kernel<<<(R + M + T - 1) / T, T, 0, h->stream>>>(parent, isvar, R, M);
    if (v.K > 0) mss::cc_slots_kernel<<<std::min(v.K, 4 * h->sm_count), T, 0, h->stream>>>(d, parent, isvar, (unsigned*)(d_misc + 2), d_misc + 1);
    if (v.O > 0) {
        if (pk) mss::cc_pairs_kernel<<<std::min((v.O + T - 1) / T, 4 * h->sm_count), T, 0, h->stream>>>(d, parent, isvar, (unsigned*)(d_misc + 2));
        else mss::cc_outside_kernel<<<(M + T - 1) / T, T, 0, h->stream>>>(d, parent, isvar, (unsigned*)(d_misc + 2));
    }
    mss::cc_rows_kernel<<<1, 1024, 0, h->stream>>>(parent, d_row, d_misc, R);
    if (M > 0) mss::cc_vars_kernel<<<(M + T - 1) / T, T, 0, h->stream>>>(parent, d_row, isvar, d_mp, R, M);
    MSS_CUDA(h, cudaGetLastError());
    h->stats.kernel_launches += 3 + (v.K > 0) + (v.O > 0);
    int misc[3] = {0, 0, 0};
    MSS_CUDA(h, cudaMemcpyAsync(misc, d_misc, sizeof(misc), cudaMemcpyDeviceToHost, h->stream));
    const cudaMemcpyKind back = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (row_label && R) MSS_CUDA(h, cudaMemcpyAsync(row_label, d_row, (size_t)R * 4, back, h->stream));
    if (mp_label && M) MSS_CUDA(h, cudaMemcpyAsync(mp_label, d_mp, (size_t)M * 4, back, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    if (misc[2]) { h->err = "view failed device-side validation (index out of range or bad pointer table)"; return MSS_E_BADARG; }
    *ncomp = misc[0];
    if (n_max) *n_max = misc[1];
    return MSS_OK;
}

int mss_solve_batch(mss_handle* h, int32_t nwin, const mss_window_view* views, mss_result* results) {
    if (!h) return MSS_E_BADARG;
    return mssi::solve_batch_impl(h, nwin, views, results);
}

int mss_comm_unique_id(void* out_id128) {
    if (!out_id128) return MSS_E_BADARG;
    std::string err;
    if (!load_nccl(err)) { fprintf(stderr, "libmss: %s\n", err.c_str()); return MSS_E_NCCL; }
    NcclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return MSS_E_NCCL;
    memcpy(out_id128, &id, sizeof(id));
    return MSS_OK;
}

int mss_comm_init(mss_handle* h, const void* id128, int32_t rank, int32_t nranks) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return MSS_E_BADARG;
    if (!load_nccl(h->err)) return MSS_E_NCCL;
    if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    MSS_CUDA(h, cudaSetDevice(h->device));
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (rc != 0) { h->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc); h->comm = nullptr; return MSS_E_NCCL; }
    h->rank = rank; h->nranks = nranks;
    return MSS_OK;
}

int mss_comm_destroy(mss_handle* h) {
    if (!h) return MSS_E_BADARG;
    if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
    h->comm = nullptr; h->rank = 0; h->nranks = 1;
    return MSS_OK;
}

void* mss_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void mss_host_free(void* p) { if (p) cudaFreeHost(p); }

void* mss_device_alloc(mss_handle* h, size_t bytes) {
    if (!h) return nullptr;
    void* p = nullptr;
    if (cudaSetDevice(h->device) != cudaSuccess) return nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { h->err = "cudaMalloc failed"; return nullptr; }
    return p;
}
void mss_device_free(mss_handle* h, void* p) { if (h && p) { cudaSetDevice(h->device); cudaFree(p); } }

int mss_memcpy_h2d(mss_handle* h, void* dst, const void* src, size_t bytes) {
    if (!h) return MSS_E_BADARG;
    MSS_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    return MSS_OK;
}
int mss_memcpy_d2h(mss_handle* h, void* dst, const void* src, size_t bytes) {
    if (!h) return MSS_E_BADARG;
    MSS_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    MSS_CUDA(h, cudaStreamSynchronize(h->stream));
    return MSS_OK;
}

int mss_get_stats(const mss_handle* h, mss_stats* out) {
    if (!h || !out) return MSS_E_BADARG;
    *out = h->stats;
    out->sm_count = h->sm_count;
    out->device_bytes = h->device_bytes;
    return MSS_OK;
}

void* mss_stream(mss_handle* h) { return h ? (void*)h->stream : nullptr; }

int mss_debug_trace(mss_handle* h, int32_t enable) {
    if (!h) return MSS_E_BADARG;
    h->trace_on = enable != 0;
    return MSS_OK;
}

int mss_debug_get_trace(const mss_handle* h, int32_t local_window, uint32_t* out_pairs, int32_t cap_pairs) {
    if (!h || !out_pairs || local_window < 0 || local_window >= h->h_trace_nwin) return MSS_E_BADARG;
    const uint2* t = h->h_trace.data() + (size_t)local_window * mss::kTraceCap;
    const int n = std::min((int)t[mss::kTraceCap - 1].x, std::min((int)cap_pairs, mss::kTraceCap - 1));
    for (int i = 0; i < n; ++i) { out_pairs[2 * i] = t[i].x; out_pairs[2 * i + 1] = t[i].y; }
    return n;
}

}  // extern "C"

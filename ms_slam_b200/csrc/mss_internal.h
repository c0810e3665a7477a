// mss_internal.h -- shared between the translation units of libmss.so (mss_engine.cu, mss_mirror.cu): the engine handle,
// device-buffer helpers and the internal entry point of the batch solve.  Not part of the C-ABI.
#pragma once
#include "../../include/mss.h"
#include "mss_kernels.cuh"

#include <cstdint>
#include <string>
#include <vector>

namespace mssi {

using mss::Ctrl;
using mss::Params;
using mss::WinDesc;
using mss::WinState;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;     // elements
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

typedef void* NcclComm;

}  // namespace mssi

struct mss_handle {
    mss_config cfg{};
    int device = 0;
    int sm_count = 0;
    int max_ctas_per_sm = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_ready = nullptr;  // compute stream -> copy stream: the ready flags of this call have been zeroed
    cudaEvent_t ev_copied = nullptr; // copy stream: the last staging copy of this call
    void* copy_mutex = nullptr;      // std::mutex of the device's shared copy stream (see device_copy_stream in mss_engine.cu)
    uint8_t* h_zero = nullptr; size_t h_zero_cap = 0;   // pinned zeros: the sync words are cleared by a DMA copy, not by a kernel
    int overlap_copy = 1;            // host views: copy on the copy stream while the kernel runs (per-window ready flags); 0 = copy first
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    // device arena
    mssi::DevBuf<uint8_t> meta;            // WinDesc[] | GroupDesc[] | cta_grp[] | gwin[]
    mssi::DevBuf<mss::WinState> ws;
    mssi::DevBuf<uint8_t> st;
    mssi::DevBuf<unsigned long long> acc;
    mssi::DevBuf<float> gain;
    mssi::DevBuf<unsigned> deg;
    mssi::DevBuf<uint8_t> seen;
    mssi::DevBuf<int> vlist;               // [2][Mpad] FREE lists
    mssi::DevBuf<uint2> trace;             // [nwin][kTraceCap], only when tracing is on
    bool trace_on = false;
    std::vector<uint2> h_trace;
    int h_trace_nwin = 0;
    mssi::DevBuf<uint32_t> ent, live;      // CSR entries / live lists (keyframe-row segments, then outside-row segments)
    mssi::DevBuf<int> rows;                // 7 per-row int arrays: row_off | ent_n | live_n | row_need | row_cov | row_ncell | ocursor
    // dual bound (mss_set_dual_bound): copy of the live lists at the snapshot, per-row and per-point scratch
    bool want_bound = false;
    int w1_tma = 1;                        // token rows of PACKED16 views staged by bulk copies in W1 (MSS_W1_TMA)
    mssi::DevBuf<uint32_t> b_snap;
    mssi::DevBuf<int> b_rows;              // snap_n | snap_d
    mssi::DevBuf<unsigned> b_vars;         // share | red
    mssi::DevBuf<uint32_t> out;
    mssi::DevBuf<uint8_t> stage;           // host views staged here
    mssi::DevBuf<unsigned> sync;           // Ctrl (first 128 B) | one barrier counter per group, 128 B apart | ready flags
    unsigned* h_one = nullptr;             // pinned constant 1: source of the ready-flag copies
    mssi::DevBuf<int> cc;                  // mss_components: parent[R + M] | row_label[R] | mp_label[M] | ncomp, n_max, err
    mss::Ctrl* ctrl = nullptr;             // = sync.p
    // pinned host mirrors
    uint8_t* h_meta = nullptr; size_t h_meta_cap = 0;
    uint32_t* h_out = nullptr; size_t h_out_cap = 0;
    mss::Ctrl* h_ctrl = nullptr;
    // comm
    mssi::NcclComm comm = nullptr;
    int rank = 0, nranks = 1;
    unsigned long long watchdog_ns = 20000000000ull;
    int tail_vars = 64, tail_ents = 256;
    int group_ctas = 0;              // CTAs per window group; 0 = heuristic
    // stats
    mss_stats stats{};
    int64_t device_bytes = 0;
    // result slots of the last batch (mirror post-processing reads the keep bits on the device)
    std::vector<int> last_out_off;
};

#define MSS_CUDA(h, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                         \
            return MSS_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

namespace mssi {

template <class T>
int ensure(mss_handle* h, DevBuf<T>& b, size_t n, bool keep = false) {
    if (n <= b.cap && b.p) return MSS_OK;
    size_t ncap = b.cap ? b.cap : 1024;
    while (ncap < n) ncap = ncap + ncap / 2 + 1024;
    T* np = nullptr;
    MSS_CUDA(h, cudaMalloc((void**)&np, ncap * sizeof(T)));
    if (b.p) {
        if (keep) MSS_CUDA(h, cudaMemcpyAsync(np, b.p, b.cap * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
        if (keep) MSS_CUDA(h, cudaStreamSynchronize(h->stream));
        MSS_CUDA(h, cudaFree(b.p));
        h->device_bytes -= (int64_t)(b.cap * sizeof(T));
    }
    b.p = np;
    b.cap = ncap;
    h->device_bytes += (int64_t)(ncap * sizeof(T));
    return MSS_OK;
}

inline int ensure_pinned(mss_handle* h, void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap && *p) return MSS_OK;
    size_t ncap = *cap ? *cap : 4096;
    while (ncap < bytes) ncap = ncap + ncap / 2 + 4096;
    if (*p) { MSS_CUDA(h, cudaFreeHost(*p)); *p = nullptr; *cap = 0; }
    MSS_CUDA(h, cudaHostAlloc(p, ncap, cudaHostAllocDefault));
    *cap = ncap;
    return MSS_OK;
}

template <class T>
void release(DevBuf<T>& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

// the batch solve behind mss_solve / mss_solve_batch (mss_engine.cu)
int solve_batch_impl(mss_handle* h, int nwin, const mss_window_view* views, mss_result* results);

}  // namespace mssi

// mss_multi.cu -- several GPUs from ONE process (include/mss.h mss_multi_*).
//
// The reference is one process with one sparsifier thread (/root/reference/src/System.cc:159-160), so the class that
// replaces it can never be "rank r of 8" of an NCCL job; what it can do is drive every GPU of the box itself.  A multi
// handle owns one engine handle per device and one worker thread per device; mss_multi_solve_batch deals the windows
// of a batch out (window w -> device w % n, the same rule as the multi-process path), every worker runs the ordinary
// mss_solve_batch on its share with its own stream and arena, and the results land directly in the caller's buffers.
// The host is the only consumer of the bitmasks here, so nothing has to be gathered between the GPUs; the NCCL
// all-gather of the result slots belongs to the one-process-per-GPU path (mss_comm_init), where every rank needs them.
#define MSS_KERNELS_TYPES_ONLY
#include "mss_internal.h"

#include <condition_variable>
#include <mutex>
#include <thread>

struct mss_multi {
    struct Worker {
        mss_handle* h = nullptr;
        std::thread th;
        std::vector<mss_window_view> views;
        std::vector<mss_result> results;
        std::vector<int> index;             // position of each of its windows in the caller's arrays
        int rc = MSS_OK;
        double ms = 0.0;
    };
    std::vector<Worker> workers;
    std::mutex mtx;
    std::condition_variable cv_go, cv_done;
    long generation = 0;
    int pending = 0;
    bool quit = false;
    std::string err;
};

namespace {
void worker_loop(mss_multi* m, int i) {
    mss_multi::Worker& w = m->workers[i];
    cudaSetDevice(w.h->device);
    long seen = 0;
    while (true) {
        {
            std::unique_lock<std::mutex> lock(m->mtx);
            m->cv_go.wait(lock, [&] { return m->quit || m->generation != seen; });
            if (m->quit) return;
            seen = m->generation;
        }
        w.rc = w.views.empty() ? MSS_OK : mssi::solve_batch_impl(w.h, (int)w.views.size(), w.views.data(), w.results.data());
        w.ms = w.h->stats.last_device_ms;
        {
            std::unique_lock<std::mutex> lock(m->mtx);
            if (--m->pending == 0) m->cv_done.notify_all();
        }
    }
}
}  // namespace

extern "C" {

int mss_multi_create(const mss_config* cfg, const int32_t* devices, int32_t n, mss_multi** out) {
    if (!cfg || !out || n < 1 || n > 64) return MSS_E_BADARG;
    *out = nullptr;
    mss_multi* m = new (std::nothrow) mss_multi();
    if (!m) return MSS_E_NOMEM;
    m->workers.resize(n);
    for (int i = 0; i < n; ++i) {
        mss_config c = *cfg;
        c.device = devices ? devices[i] : i;
        const int rc = mss_create(&c, &m->workers[i].h);
        if (rc != MSS_OK) {
            for (int j = 0; j < i; ++j) mss_destroy(m->workers[j].h);
            delete m;
            return rc;
        }
    }
    for (int i = 0; i < n; ++i) m->workers[i].th = std::thread(worker_loop, m, i);
    *out = m;
    return MSS_OK;
}

void mss_multi_destroy(mss_multi* m) {
    if (!m) return;
    {
        std::unique_lock<std::mutex> lock(m->mtx);
        m->quit = true;
    }
    m->cv_go.notify_all();
    for (auto& w : m->workers) if (w.th.joinable()) w.th.join();
    for (auto& w : m->workers) mss_destroy(w.h);
    delete m;
}

int mss_multi_device_count(const mss_multi* m) { return m ? (int)m->workers.size() : 0; }
const char* mss_multi_last_error(const mss_multi* m) { return m ? m->err.c_str() : "null handle"; }

int mss_multi_set_params(mss_multi* m, int32_t min_points, float lambda, float grid_lambda) {
    if (!m) return MSS_E_BADARG;
    int rc = MSS_OK;
    for (auto& w : m->workers) { const int r = mss_set_params(w.h, min_points, lambda, grid_lambda); if (r != MSS_OK) rc = r; }
    return rc;
}

int mss_multi_solve_batch(mss_multi* m, int32_t nwin, const mss_window_view* views, mss_result* results) {
    if (!m || nwin < 0 || (nwin > 0 && (!views || !results))) return MSS_E_BADARG;
    m->err.clear();
    if (nwin == 0) return MSS_OK;
    const int n = (int)m->workers.size();
    for (int w = 0; w < nwin; ++w)
        if (views[w].memory != MSS_MEM_HOST || views[w].result_memory == MSS_RESULT_DEVICE) {
            m->err = "mss_multi_solve_batch takes host views and host result buffers (device memory belongs to one GPU)";
            return MSS_E_BADARG;
        }
    for (auto& w : m->workers) { w.views.clear(); w.results.clear(); w.index.clear(); }
    for (int w = 0; w < nwin; ++w) {
        mss_multi::Worker& k = m->workers[w % n];
        k.views.push_back(views[w]);
        k.results.push_back(results[w]);
        k.index.push_back(w);
    }
    {
        std::unique_lock<std::mutex> lock(m->mtx);
        m->pending = n;
        ++m->generation;
    }
    m->cv_go.notify_all();
    {
        std::unique_lock<std::mutex> lock(m->mtx);
        m->cv_done.wait(lock, [&] { return m->pending == 0; });
    }
    int rc = MSS_OK;
    for (auto& k : m->workers) {
        for (size_t i = 0; i < k.index.size(); ++i) results[k.index[i]] = k.results[i];
        if (k.rc != MSS_OK && (rc == MSS_OK || rc == MSS_E_NOCONVERGE)) {
            rc = k.rc;
            m->err = std::string("device ") + std::to_string(k.h->device) + ": " + mss_last_error(k.h);
        }
    }
    return rc;
}

int mss_multi_get_stats(const mss_multi* m, int32_t device_index, mss_stats* out) {
    if (!m || !out || device_index < 0 || device_index >= (int)m->workers.size()) return MSS_E_BADARG;
    return mss_get_stats(m->workers[device_index].h, out);
}

}  // extern "C"

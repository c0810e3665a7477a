// host_capi.cc -- extern "C" test / benchmark harness around the C++ MapSparsification mirror.  It builds a small
// ORB-SLAM3-shaped world (SlamShims.h) from a flat window view, drives the sparsifier thread through the same calls
// System / LocalMapping / LoopClosing make upstream, and exposes what happened as plain arrays so that Python tests can
// compare it with the engine called directly.  Not part of the product surface.
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "MapSparsification.h"

using namespace ORB_SLAM3;

namespace {
struct World {
    Atlas atlas;
    LoopClosing* loop = nullptr;
    MapSparsification* ms = nullptr;
    std::vector<std::shared_ptr<KeyFrame>> kfs;        // K window keyframes, then H outside keyframes
    std::vector<std::shared_ptr<MapPoint>> mps;        // M map points of the view, then filler points
    int K = 0, H = 0, M = 0;
    std::thread thread;
    bool running = false;
    WindowSnapshot flat;                               // msh_flatten_only
    long unsigned int flat_id = 1000000;
    ~World() {
        if (running) { ms->RequestFinish(); thread.join(); }
        delete ms;
        delete loop;
    }
};
}  // namespace

extern "C" {

void* msh_create(const char* settings_path, int inertial) {
    World* w = new World();
    w->loop = new LoopClosing(&w->atlas);
    w->ms = new MapSparsification(settings_path ? settings_path : "", &w->atlas, inertial != 0);
    w->ms->SetLoopClosing(w->loop);
    SparsificationSettings s;
    if (settings_path && ReadSparsificationSettings(settings_path, s) && s.NonLocalKF > 0) KeyFrame::mnNonLocalKF = s.NonLocalKF;
    return w;
}

void msh_destroy(void* h) { delete static_cast<World*>(h); }

int msh_engine_ready(void* h) { return static_cast<World*>(h)->ms->EngineReady() ? 1 : 0; }

// Build the pointer graph a flat view describes (inverse of FlattenWindow).
int msh_build_world(void* h, int K, int H, int M, const int32_t* feat_ptr, const int32_t* feat_mp, const uint16_t* feat_cell,
                    const int32_t* mp_nobs, const int32_t* mp_obs_ptr, const int32_t* mp_obs_kf, const int32_t* okf_total) {
    World* w = static_cast<World*>(h);
    Map* map = w->atlas.GetCurrentMap();
    w->K = K; w->H = H; w->M = M;
    for (int p = 0; p < M; ++p) {
        w->mps.push_back(std::make_shared<MapPoint>((long unsigned)p, map));
        map->AddMapPoint(w->mps.back());
    }
    for (int k = 0; k < K; ++k) {
        const int n = feat_ptr[k + 1] - feat_ptr[k];
        auto kf = std::make_shared<KeyFrame>((long unsigned)k, map, (size_t)n);
        for (int i = 0; i < n; ++i) {
            const int s = feat_ptr[k] + i;
            const int p = feat_mp[s];
            if (feat_cell[s] != MSS_CELL_NONE) kf->SetGridCell(feat_cell[s] / MSS_GRID_ROWS, feat_cell[s] % MSS_GRID_ROWS, (size_t)i);
            if (p < 0 || p >= M) continue;
            kf->AddMapPoint(w->mps[p], (size_t)i);
            w->mps[p]->AddObservation(kf, i);
        }
        w->kfs.push_back(kf);
        map->AddKeyFrame(kf);
    }
    // outside keyframes: the window points that observe them first, then filler points up to GetNumberMPs() == okf_total
    std::vector<std::vector<int>> outside(H);
    for (int p = 0; p < M; ++p)
        for (int o = mp_obs_ptr[p]; o < mp_obs_ptr[p + 1]; ++o)
            if (mp_obs_kf[o] >= K && mp_obs_kf[o] < K + H) outside[mp_obs_kf[o] - K].push_back(p);
    long unsigned next_mp = (long unsigned)M;
    for (int j = 0; j < H; ++j) {
        const int n = std::max((int)outside[j].size(), okf_total[j]);
        auto kf = std::make_shared<KeyFrame>((long unsigned)(K + j), map, (size_t)n);
        for (int i = 0; i < n; ++i) {
            std::shared_ptr<MapPoint> mp;
            if (i < (int)outside[j].size()) mp = w->mps[outside[j][i]];
            else { mp = std::make_shared<MapPoint>(next_mp++, map); map->AddMapPoint(mp); w->mps.push_back(mp); mp->nObs = 3; }
            // GetNumberMPs() must come out as okf_total[j]: observers beyond that number observe the keyframe without
            // sitting in one of its slots (mObservations and mvpMapPoints are separate structures upstream too)
            if (i < okf_total[j]) { kf->AddMapPoint(mp, (size_t)i); mp->AddObservation(kf, i); }
        }
        kf->mbSparsified = true;            // processed by an earlier window: the final flush must not pick it up again
        w->kfs.push_back(kf);
        map->AddKeyFrame(kf);
        // (slot-less observations are added once the keyframe is in the map, like any later map operation: a keyframe
        // enters the map with observations only where it holds the point, LocalMapping::ProcessNewKeyFrame)
        for (int i = okf_total[j]; i < (int)outside[j].size(); ++i) w->mps[outside[j][i]]->AddObservation(kf, i);
    }
    for (int p = 0; p < M; ++p) {
        w->mps[p]->nObs = mp_nobs[p];      // the view's Observations(), whatever the stereo mix was
        if (map->mpMirror) map->mpMirror->OnMapPoint(w->mps[p].get(), mp_nobs[p], false);     // (a direct member write has no hook)
    }
    return 0;
}

// FlattenWindow on the K window keyframes without solving (works without a GPU).
int msh_flatten_only(void* h) {
    World* w = static_cast<World*>(h);
    std::vector<std::shared_ptr<KeyFrame>> win(w->kfs.begin(), w->kfs.begin() + w->K);
    FlattenWindow(win, ++w->flat_id, w->flat);         // a fresh window id per call, like Sparsifying() (mnId++)
    return 0;
}
// host time of the last msh_flatten_only in microseconds (FlattenWindow incl. packing; WindowSnapshot::flatten_ms)
int msh_flatten_us(void* h) { return (int)(static_cast<World*>(h)->flat.flatten_ms * 1000.0); }

static const WindowSnapshot& snap(World* w, int which) { return which == 0 ? w->flat : w->ms->LastSnapshot(); }

// which: 0 = msh_flatten_only result, 1 = snapshot of the last window the thread processed
void msh_snapshot_sizes(void* h, int which, int32_t* out5) {
    const WindowSnapshot& s = snap(static_cast<World*>(h), which);
    out5[0] = s.K; out5[1] = s.H; out5[2] = (int32_t)s.mp_nobs.size(); out5[3] = (int32_t)s.feat_mp.size(); out5[4] = (int32_t)s.mp_obs_kf.size();
}

void msh_snapshot_copy(void* h, int which, int32_t* feat_ptr, int32_t* feat_mp, uint16_t* feat_cell, int32_t* mp_nobs,
                       int32_t* mp_obs_ptr, int32_t* mp_obs_kf, int32_t* okf_total, int64_t* mp_ids, int64_t* okf_ids, uint8_t* is_var) {
    const WindowSnapshot& s = snap(static_cast<World*>(h), which);
    memcpy(feat_ptr, s.feat_ptr.data(), s.feat_ptr.size() * 4);
    memcpy(feat_mp, s.feat_mp.data(), s.feat_mp.size() * 4);
    memcpy(feat_cell, s.feat_cell.data(), s.feat_cell.size() * 2);
    memcpy(mp_nobs, s.mp_nobs.data(), s.mp_nobs.size() * 4);
    memcpy(mp_obs_ptr, s.mp_obs_ptr.data(), s.mp_obs_ptr.size() * 4);
    memcpy(mp_obs_kf, s.mp_obs_kf.data(), s.mp_obs_kf.size() * 4);
    memcpy(okf_total, s.okf_total.data(), s.okf_total.size() * 4);
    for (size_t p = 0; p < s.mp_ids.size(); ++p) { mp_ids[p] = (int64_t)s.mp_ids[p]; is_var[p] = s.is_var[p]; }
    for (size_t j = 0; j < s.okf_ids.size(); ++j) okf_ids[j] = (int64_t)s.okf_ids[j];
}

// the packed transport blob of a snapshot (MSS_LAYOUT_PACKED16): sizes first (out4 = tokens, pairs, packed flag, blob bytes),
// then the arrays; tok_ptr[K+1], tokens[n_tokens], nobs16[M], pairs[n_pairs]
void msh_snapshot_packed_sizes(void* h, int which, int64_t* out4) {
    const WindowSnapshot& s = snap(static_cast<World*>(h), which);
    const mss_window_view v = s.View();
    out4[0] = (int64_t)s.n_tokens; out4[1] = (int64_t)s.n_pairs; out4[2] = s.packed ? 1 : 0;
    out4[3] = (v.layout == MSS_LAYOUT_PACKED16) ? (int64_t)(s.off_okf + (size_t)s.H * 4) : 0;
}
void msh_snapshot_packed_copy(void* h, int which, int32_t* tok_ptr, uint16_t* tokens, uint16_t* nobs16, uint32_t* pairs) {
    const WindowSnapshot& s = snap(static_cast<World*>(h), which);
    const mss_window_view v = s.View();
    if (v.layout != MSS_LAYOUT_PACKED16) return;
    memcpy(tok_ptr, v.feat_ptr, (size_t)(v.K + 1) * 4);
    memcpy(tokens, v.slots16, (size_t)v.F * 2);
    if (v.nobs8) { const uint8_t* b = reinterpret_cast<const uint8_t*>(v.mp_nobs16); for (int p = 0; p < v.M; ++p) nobs16[p] = b[p]; }   // widened
    else memcpy(nobs16, v.mp_nobs16, (size_t)v.M * 2);
    memcpy(pairs, v.obs_pairs, (size_t)v.O * 4);
}
int msh_snapshot_nobs8(void* h, int which) { return snap(static_cast<World*>(h), which).View().nobs8; }

// The calls System makes at start-up (src/System.cc:160): run the sparsifier on its own thread.
int msh_start(void* h) {
    World* w = static_cast<World*>(h);
    if (w->running) return -1;
    w->thread = std::thread(&MapSparsification::Run, w->ms);
    w->running = true;
    return 0;
}

// LocalMapping's producer hook (src/LocalMapping.cc:252-274) for the first `count` window keyframes, in order.
int msh_feed(void* h, int first, int count) {
    World* w = static_cast<World*>(h);
    for (int k = first; k < first + count && k < w->K; ++k) w->ms->InsertKeyFrame(w->kfs[k]);
    return 0;
}

// the non-local detector of LocalMapping (KeyFrame::UpdateCountInLocalMapping): returns how many updates it took until
// keyframe k reported "non-local"
int msh_nonlocal_after(void* h, int k, int max_updates) {
    World* w = static_cast<World*>(h);
    for (int i = 1; i <= max_updates; ++i)
        if (w->kfs[k]->UpdateCountInLocalMapping(false)) return i;
    return -1;
}

int msh_forwarded_count(void* h) { return (int)static_cast<World*>(h)->loop->mvForwardedIds.size(); }

int msh_wait_forwarded(void* h, int n, int timeout_ms) {
    World* w = static_cast<World*>(h);
    const auto t0 = std::chrono::steady_clock::now();
    while (true) {
        // forwarding is the last thing Sparsifying does; isStopped() turns true right after it
        if (w->loop->SparsifiedQueueSize() >= (size_t)n && w->ms->isStopped()) return 0;
        if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > timeout_ms) return -1;
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
}

// LoopClosing::CorrectLoop's handshake (src/LoopClosing.cc:930,955-958,1162)
int msh_stop_handshake(void* h, int timeout_ms) {
    World* w = static_cast<World*>(h);
    w->ms->RequestStop();
    const auto t0 = std::chrono::steady_clock::now();
    while (!w->ms->isStopped()) {
        if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > timeout_ms) return -1;
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    const int blocked = w->ms->CheckNewKeyFrames() ? 0 : 1;      // while stopped, no new window may start
    w->ms->Release();
    return blocked;
}

// LoopClosing::Run's consumer step (src/LoopClosing.cc:104,318-328)
int msh_consume(void* h) { static_cast<World*>(h)->loop->DeleteOutdatedInfo(); return 0; }

// System::Shutdown (src/System.cc:460-471)
int msh_finish(void* h, int timeout_ms) {
    World* w = static_cast<World*>(h);
    if (!w->running) return -1;
    w->ms->RequestFinish();
    const auto t0 = std::chrono::steady_clock::now();
    while (!w->ms->isFinished()) {
        if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > timeout_ms) return -2;
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    w->thread.join();
    w->running = false;
    return 0;
}

void msh_bad_flags(void* h, uint8_t* out_m) {
    World* w = static_cast<World*>(h);
    for (int p = 0; p < w->M; ++p) out_m[p] = w->mps[p]->isBad() ? 1 : 0;
}

int msh_forwarded_ids(void* h, int64_t* out, int cap) {
    World* w = static_cast<World*>(h);
    const int n = std::min(cap, (int)w->loop->mvForwardedIds.size());
    for (int i = 0; i < n; ++i) out[i] = (int64_t)w->loop->mvForwardedIds[i];
    return (int)w->loop->mvForwardedIds.size();
}

// per keyframe: [valid slots now, mbSparsified, EraseBadDescriptor calls]
void msh_keyframe_state(void* h, int32_t* out3) {
    World* w = static_cast<World*>(h);
    for (size_t k = 0; k < w->kfs.size(); ++k) {
        out3[3 * k] = w->kfs[k]->GetNumberMPs();
        out3[3 * k + 1] = w->kfs[k]->mbSparsified ? 1 : 0;
        out3[3 * k + 2] = w->kfs[k]->mnEraseBadDescriptorCalls;
    }
}

int msh_map_counts(void* h, int64_t* out3) {
    Map* m = static_cast<World*>(h)->atlas.GetCurrentMap();
    out3[0] = (int64_t)m->MapPointsInMap(); out3[1] = (int64_t)m->SparsifiedMapPointsInMap(); out3[2] = (int64_t)m->GetAllSparsifiedKeyFrames().size();
    return 0;
}

// reports of the windows processed so far: 13 doubles each (msh_reports) / 19 with the mirror fields (msh_reports2)
int msh_reports2(void* h, double* out, int cap_windows) {
    World* w = static_cast<World*>(h);
    const auto reps = w->ms->GetReports();
    const int n = std::min(cap_windows, (int)reps.size());
    for (int i = 0; i < n; ++i) {
        const auto& r = reps[i];
        double* o = out + 20 * i;
        o[0] = r.status; o[1] = r.K; o[2] = r.H; o[3] = r.M; o[4] = r.n_vars; o[5] = r.n_kept; o[6] = r.n_deleted; o[7] = r.rounds;
        o[8] = r.objective; o[9] = r.flatten_ms; o[10] = r.solve_ms; o[11] = r.apply_ms; o[12] = r.components;
        o[13] = r.mirror; o[14] = (double)r.delta_ops; o[15] = r.build_ms; o[16] = (double)r.h2d_bytes; o[17] = (double)r.d2h_bytes; o[18] = r.devices;
        o[19] = r.dual_bound;
    }
    return (int)reps.size();
}
int msh_mirror_active(void* h) { return static_cast<World*>(h)->ms->MirrorActive() ? 1 : 0; }

int msh_reports(void* h, double* out, int cap_windows) {
    World* w = static_cast<World*>(h);
    const auto reps = w->ms->GetReports();
    const int n = std::min(cap_windows, (int)reps.size());
    for (int i = 0; i < n; ++i) {
        const auto& r = reps[i];
        double* o = out + 13 * i;
        o[0] = r.status; o[1] = r.K; o[2] = r.H; o[3] = r.M; o[4] = r.n_vars; o[5] = r.n_kept; o[6] = r.n_deleted; o[7] = r.rounds;
        o[8] = r.objective; o[9] = r.flatten_ms; o[10] = r.solve_ms; o[11] = r.apply_ms; o[12] = r.components;
    }
    return (int)reps.size();
}

int msh_set_min_points(void* h, int n) { static_cast<World*>(h)->ms->mnMinNum = n; return 0; }

}  // extern "C"

// MapSparsification.cc -- the sparsifier thread of MS-SLAM on the B200 engine.
//
// Behaviour kept from the reference (/root/reference/src/MapSparsification.cc), by line:
//   :4-21    constructor reads Sparsification.N / Lambda / GridLambda / WindowLength and starts the solver environment
//   :23-56   Run(): poll every 3 ms; when more than 10 keyframes wait, take the oldest <= WindowLength and sparsify them;
//            mbStopped is false only while a window is being processed; on finish, sparsify every keyframe that is not
//            sparsified yet and call EraseBadDescriptor on each
//   :159-166 every variable map point whose solution value is 0 gets SetBadFlag()
//   :168-170 every window keyframe is forwarded to LoopClosing::InsertSparsifiedKeyFrame
//   :173-199 FIFO queue under mMutexNewKFs, trigger size > 10, inertial gate on Map::GetIniertialBA2
//   :205-241 stop / release / finish handshakes under mMutexStop / mMutexFinish
// New: Sparsifying() = FlattenWindow (one walk of the pointer graph) -> mss_solve (libmss, CUDA) -> apply the bitmask.
// Fail-safe: if the engine is unavailable or a solve fails, no map point is deleted (deleting is irreversible) and the
// keyframes are still forwarded, so the SLAM pipeline keeps running with an unsparsified window.
#include "MapSparsification.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>
#include <unordered_map>

namespace ORB_SLAM3 {

using std::shared_ptr;
using std::vector;
typedef std::chrono::steady_clock Clock;
static double MsSince(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

// ---------------------------------------------------------------------------------------------------------------------
// settings
// ---------------------------------------------------------------------------------------------------------------------
bool ReadSparsificationSettings(const std::string& path, SparsificationSettings& s) {
    std::ifstream in(path.c_str());
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        auto trim = [](std::string t) {
            const char* ws = " \t\r\n\"";
            const size_t a = t.find_first_not_of(ws), b = t.find_last_not_of(ws);
            return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
        };
        const std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (val.empty()) continue;
        const double v = std::atof(val.c_str());
        if (key == "Sparsification.N") s.N = (int)v;
        else if (key == "Sparsification.Lambda") s.Lambda = (float)v;
        else if (key == "Sparsification.GridLambda") s.GridLambda = (float)v;
        else if (key == "Sparsification.WindowLength") s.WindowLength = (int)v;
        else if (key == "Sparsification.NonLocalKF") s.NonLocalKF = (int)v;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// flatten: pointer graph -> mss_window_view
// ---------------------------------------------------------------------------------------------------------------------
// Worker threads of the flatten pass: the accessors of the SLAM classes take per-object mutexes (KeyFrame.cc:434-444,
// MapPoint.cc:215-225,325-331), so keyframes and map points can be read from several threads; everything that fixes an
// ORDER (map-point numbering, outside-keyframe table) stays sequential.  MSS_FLATTEN_THREADS overrides the default.
static int FlattenThreads() {
    if (const char* e = std::getenv("MSS_FLATTEN_THREADS")) { const int n = std::atoi(e); if (n > 0) return std::min(n, 64); }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hw ? hw : 1u, 8u));
}

template <class F>
static void ParallelChunks(size_t n, int nthreads, size_t min_per_thread, F f) {   // f(thread, begin, end), contiguous ascending chunks
    nthreads = (int)std::max<size_t>(1, std::min<size_t>((size_t)nthreads, (n + min_per_thread - 1) / min_per_thread));
    if (nthreads <= 1) { f(0, (size_t)0, n); return; }
    vector<std::thread> pool;
    const size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        pool.emplace_back([=, &f] { f(t, lo, hi); });
    }
    for (std::thread& th : pool) th.join();
}

WindowSnapshot::Blob::~Blob() {
    if (!p) return;
    if (pinned) mss_host_free(p); else std::free(p);
}

void WindowSnapshot::Blob::Reserve(size_t bytes) {
    if (bytes <= cap && p) return;
    if (p) { if (pinned) mss_host_free(p); else std::free(p); p = nullptr; cap = 0; }
    const size_t ncap = bytes + bytes / 2 + 4096;
    p = static_cast<uint8_t*>(mss_host_alloc(ncap));          // NULL without a CUDA device: plain memory still works
    pinned = p != nullptr;
    if (!p) p = static_cast<uint8_t*>(std::malloc(ncap));
    cap = p ? ncap : 0;
}

void WindowSnapshot::Pack() {
    // MSS_LAYOUT_PACKED16 (include/mss.h): valid slots only, sorted by map-point index inside every keyframe (the order of
    // the slots inside a keyframe carries no meaning for the model; sorted, the entries a warp handles touch neighbouring
    // map points) and delta-coded as 16-bit tokens; u16 nObs; the outside observations as one pair list.
    packed = false;
    const size_t M = mp_nobs.size(), F = feat_mp.size(), O = mp_obs_kf.size();
    if (M > (1u << 20) || H > 4095) return;
    int32_t nobs_max = 0;
    for (int32_t n : mp_nobs) { if (n < 0 || n > 65535) return; nobs_max = std::max(nobs_max, n); }
    nobs8 = nobs_max <= 255;                   // one byte per map point on the wire whenever Observations() fits
    auto up = [](size_t x) { return (x + 15) / 16 * 16; };
    std::vector<uint32_t> slots(F);
    for (size_t i = 0; i < F; ++i) {
        const uint32_t cell = feat_cell[i] == (uint16_t)MSS_CELL_NONE ? MSS_SLOT_CELL_NONE : (uint32_t)feat_cell[i];
        slots[i] = feat_mp[i] < 0 ? MSS_SLOT_EMPTY : (((uint32_t)feat_mp[i] << 12) | cell);
    }
    std::vector<int32_t> tok_ptr(K + 1, 0);
    const int nthreads = FlattenThreads();
    ParallelChunks((size_t)K, nthreads, 4, [&](int, size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            std::sort(slots.begin() + feat_ptr[k], slots.begin() + feat_ptr[k + 1]);   // empty slots sort to the end
            uint32_t prev = 0;
            int32_t nt = 0;
            for (int32_t i = feat_ptr[k]; i < feat_ptr[k + 1] && slots[i] != MSS_SLOT_EMPTY; ++i) {
                const uint32_t mp = slots[i] >> 12, units = (mp - prev) / 15u;
                nt += 1 + (int32_t)((units + 4095u) / 4096u);
                prev = mp;
            }
            tok_ptr[k + 1] = nt;
        }
    });
    for (int k = 0; k < K; ++k) tok_ptr[k + 1] += tok_ptr[k];
    n_tokens = (size_t)tok_ptr[K];
    off_slots = up((size_t)(K + 1) * 4);
    off_nobs = off_slots + up(n_tokens * 2);
    n_pairs = 0;
    for (size_t o = 0; o < O; ++o) n_pairs += mp_obs_kf[o] >= K ? 1 : 0;       // FlattenWindow emits outside observations only
    off_pairs = off_nobs + up(M * (nobs8 ? 1 : 2));
    off_okf = off_pairs + up(n_pairs * 4);
    const size_t total = off_okf + up((size_t)H * 4);
    if (!blob) blob = std::make_shared<Blob>();
    blob->Reserve(total);
    if (!blob->p) return;
    memcpy(blob->p, tok_ptr.data(), (size_t)(K + 1) * 4);
    uint16_t* tok = reinterpret_cast<uint16_t*>(blob->p + off_slots);
    ParallelChunks((size_t)K, nthreads, 4, [&](int, size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            size_t t = (size_t)tok_ptr[k];
            uint32_t prev = 0;
            for (int32_t i = feat_ptr[k]; i < feat_ptr[k + 1] && slots[i] != MSS_SLOT_EMPTY; ++i) {
                const uint32_t mp = slots[i] >> 12, gap = mp - prev;
                for (uint32_t units = gap / 15u; units > 0;) {                  // 15 * (low + 1) per escape token
                    const uint32_t u = std::min(units, 4096u);
                    tok[t++] = (uint16_t)((15u << 12) | (u - 1u));
                    units -= u;
                }
                tok[t++] = (uint16_t)(((gap % 15u) << 12) | (slots[i] & 0xFFFu));
                prev = mp;
            }
        }
    });
    if (nobs8) {
        uint8_t* nobs = blob->p + off_nobs;
        for (size_t p = 0; p < M; ++p) nobs[p] = (uint8_t)mp_nobs[p];
    } else {
        uint16_t* nobs = reinterpret_cast<uint16_t*>(blob->p + off_nobs);
        for (size_t p = 0; p < M; ++p) nobs[p] = (uint16_t)mp_nobs[p];
    }
    uint32_t* pairs = reinterpret_cast<uint32_t*>(blob->p + off_pairs);
    size_t np = 0;
    for (size_t p = 0; p < M; ++p)
        for (int32_t o = mp_obs_ptr[p]; o < mp_obs_ptr[p + 1]; ++o)
            if (mp_obs_kf[o] >= K) pairs[np++] = ((uint32_t)p << 12) | (uint32_t)(mp_obs_kf[o] - K);
    if (H) memcpy(blob->p + off_okf, okf_total.data(), (size_t)H * 4);
    packed = true;
}

void SplitSnapshot(const WindowSnapshot& in, const vector<int32_t>& row_label, const vector<int32_t>& mp_label, int n_max,
                   vector<WindowSnapshot>& parts) {
    parts.clear();
    const int K = in.K, H = in.H;
    const size_t M = in.mp_nobs.size();
    // components in order of their first window keyframe (labels are dense in that order already)
    vector<int32_t> comp_of;                     // label -> part index or -1
    for (int k = 0; k < K; ++k) {
        const int32_t c = row_label[k];
        if ((size_t)c >= comp_of.size()) comp_of.resize(c + 1, -1);
        if (comp_of[c] < 0) { comp_of[c] = (int32_t)parts.size(); parts.emplace_back(); }
    }
    vector<int32_t> mp_new(M, -1), okf_new(H, -1);
    for (size_t p = 0; p < M; ++p) {
        const int32_t c = mp_label[p];
        if (c < 0 || (size_t)c >= comp_of.size() || comp_of[c] < 0) continue;
        WindowSnapshot& q = parts[comp_of[c]];
        mp_new[p] = (int32_t)q.mp_nobs.size();
        q.mp_nobs.push_back(in.mp_nobs[p]);
        if (!in.mp_tie.empty()) q.mp_tie.push_back(in.mp_tie[p]);        // (ranks stay ordered: that is all a tie-break needs)
        q.is_var.push_back(1);
        q.vpMapPoints.push_back(in.vpMapPoints[p]);
        q.part_mp.push_back((int32_t)p);
    }
    for (int j = 0; j < H; ++j) {
        const int32_t c = row_label[K + j];
        if ((size_t)c >= comp_of.size() || comp_of[c] < 0) continue;      // sees no variable of any window keyframe
        WindowSnapshot& q = parts[comp_of[c]];
        okf_new[j] = q.H++;
        q.okf_total.push_back(in.okf_total[j]);
        q.vpOutsideKFs.push_back(in.vpOutsideKFs[j]);
    }
    for (WindowSnapshot& q : parts) { q.feat_ptr.assign(1, 0); q.mp_obs_ptr.assign(1, 0); q.n_max_floor = n_max; }
    for (int k = 0; k < K; ++k) {
        WindowSnapshot& q = parts[comp_of[row_label[k]]];
        for (int32_t s = in.feat_ptr[k]; s < in.feat_ptr[k + 1]; ++s) {
            const int32_t p = in.feat_mp[s];
            if (p < 0 || mp_new[p] < 0) continue;                        // not a variable: counts for nMax only (carried)
            q.feat_mp.push_back(mp_new[p]);
            q.feat_cell.push_back(in.feat_cell[s]);
        }
        q.feat_ptr.push_back((int32_t)q.feat_mp.size());
        q.K++;
    }
    for (size_t p = 0; p < M; ++p) {
        if (mp_new[p] < 0) continue;
        WindowSnapshot& q = parts[comp_of[mp_label[p]]];
        for (int32_t o = in.mp_obs_ptr[p]; o < in.mp_obs_ptr[p + 1]; ++o) {
            const int32_t kf = in.mp_obs_kf[o];
            if (kf >= K && okf_new[kf - K] >= 0) q.mp_obs_kf.push_back(q.K + okf_new[kf - K]);
        }
        q.mp_obs_ptr.push_back((int32_t)q.mp_obs_kf.size());
    }
    for (WindowSnapshot& q : parts) q.Pack();
}

mss_window_view WindowSnapshot::View() const {
    mss_window_view v{};
    v.n_max_floor = n_max_floor;
    if (packed && blob && blob->p) {
        v.K = K; v.H = H; v.M = (int32_t)mp_nobs.size(); v.F = (int32_t)n_tokens; v.O = (int32_t)n_pairs;
        v.memory = MSS_MEM_HOST;
        v.layout = MSS_LAYOUT_PACKED16;
        v.feat_ptr = reinterpret_cast<const int32_t*>(blob->p);
        v.slots16 = reinterpret_cast<const uint16_t*>(blob->p + off_slots);
        v.mp_nobs16 = reinterpret_cast<const uint16_t*>(blob->p + off_nobs);
        v.nobs8 = nobs8 ? 1 : 0;
        v.obs_pairs = reinterpret_cast<const uint32_t*>(blob->p + off_pairs);
        v.okf_total = reinterpret_cast<const int32_t*>(blob->p + off_okf);
        v.mp_tie = mp_tie.size() == mp_nobs.size() && !mp_tie.empty() ? mp_tie.data() : nullptr;
        return v;
    }
    v.K = K; v.H = H; v.M = (int32_t)mp_nobs.size(); v.F = (int32_t)feat_mp.size(); v.O = (int32_t)mp_obs_kf.size();
    v.memory = MSS_MEM_HOST;
    v.feat_ptr = feat_ptr.data(); v.feat_mp = feat_mp.data(); v.feat_cell = feat_cell.data();
    v.mp_nobs = mp_nobs.data(); v.mp_obs_ptr = mp_obs_ptr.data(); v.mp_obs_kf = mp_obs_kf.data();
    v.okf_total = okf_total.data();
    v.mp_tie = mp_tie.size() == mp_nobs.size() && !mp_tie.empty() ? mp_tie.data() : nullptr;
    return v;
}

void FlattenWindow(const vector<shared_ptr<KeyFrame>>& vpKFs, long unsigned int nId, WindowSnapshot& out) {
    const Clock::time_point t0 = Clock::now();
    const int K = (int)vpKFs.size();
    const int nthreads = FlattenThreads();
    const bool trace = std::getenv("MSS_FLATTEN_TRACE") != nullptr;
    double tA = 0, tB = 0, tC = 0, tD = 0;
    out.K = K; out.H = 0;
    out.feat_ptr.assign(1, 0);
    out.feat_mp.clear(); out.feat_cell.clear(); out.mp_nobs.clear(); out.is_var.clear();
    out.mp_obs_ptr.assign(1, 0); out.mp_obs_kf.clear(); out.okf_total.clear();
    out.vpMapPoints.clear(); out.vpOutsideKFs.clear();

    // window membership first (the reference stamps inside its second pass, :81; observations are classified with it, :132)
    for (int k = 0; k < K; ++k) vpKFs[k]->mnMapSaprsificationId = nId;

    // keyframe side, parallel over keyframes: valid slots and their cells (passes 1 and 2, :67-123).  Only valid slots are
    // kept (empty slots and bad points contribute nothing to the model, :70,90).
    struct KfLocal { vector<shared_ptr<MapPoint>> mps; vector<uint16_t> cell; };
    vector<KfLocal> local(K);
    ParallelChunks((size_t)K, nthreads, 4, [&](int, size_t lo, size_t hi) {
        vector<int32_t> slotPos;
        for (size_t k = lo; k < hi; ++k) {
            const vector<shared_ptr<MapPoint>> vMPs = vpKFs[k]->GetMapPointMatches();
            const size_t n = vMPs.size();
            KfLocal& L = local[k];
            slotPos.assign(n, -1);
            for (size_t i = 0; i < n; ++i) {
                const shared_ptr<MapPoint>& pMP = vMPs[i];
                if (!pMP || pMP->isBad()) continue;
                slotPos[i] = (int32_t)L.mps.size();
                L.mps.push_back(pMP);
                L.cell.push_back((uint16_t)MSS_CELL_NONE);
            }
            const auto& grid = vpKFs[k]->GetFeatureGrids();
            for (size_t col = 0; col < grid.size(); ++col)
                for (size_t row = 0; row < grid[col].size(); ++row)
                    for (size_t i : grid[col][row]) {
                        if (i >= n || slotPos[i] < 0) continue;
                        L.cell[slotPos[i]] = (uint16_t)(col * MSS_GRID_ROWS + row);
                    }
        }
    });
    tA = MsSince(t0);
    // numbering, sequential: map points in discovery order = mnIndexForSparsification of the reference (:91-99)
    for (int k = 0; k < K; ++k) {
        const KfLocal& L = local[k];
        for (size_t i = 0; i < L.mps.size(); ++i) {
            MapPoint* pMP = L.mps[i].get();
            if (pMP->mnMapSparsificationId != nId) {
                pMP->mnMapSparsificationId = nId;
                pMP->mnIndexForSparsification = out.vpMapPoints.size();
                out.vpMapPoints.push_back(L.mps[i]);
                out.is_var.push_back(0);
            }
            out.feat_mp.push_back((int32_t)pMP->mnIndexForSparsification);
            out.feat_cell.push_back(L.cell[i]);
            if (L.cell[i] != (uint16_t)MSS_CELL_NONE) out.is_var[pMP->mnIndexForSparsification] = 1;
        }
        out.feat_ptr.push_back((int32_t)out.feat_mp.size());
    }
    vector<KfLocal>().swap(local);
    tB = MsSince(t0);

    // map-point side, parallel over map points: Observations() of every point and, for the variables, their observations
    // by keyframes OUTSIDE the window (pass 3, :125-142); observations by window keyframes are not needed (the engine only
    // builds the outside rows).  Chunks are contiguous, so concatenating the threads' lists keeps map-point order.
    const size_t M = out.vpMapPoints.size();
    out.mp_nobs.assign(M, 0);
    typedef std::pair<int32_t, shared_ptr<KeyFrame>> OutObs;
    vector<vector<OutObs>> found((size_t)std::max(nthreads, 1));
    ParallelChunks(M, nthreads, 1024, [&](int t, size_t lo, size_t hi) {
        vector<OutObs>& mine = found[t];
        for (size_t p = lo; p < hi; ++p) {
            out.mp_nobs[p] = out.vpMapPoints[p]->Observations();
            if (!out.is_var[p]) continue;
            const auto obs = out.vpMapPoints[p]->GetObservations();
            for (const auto& kv : obs)
                if (kv.first->mnMapSaprsificationId != nId) mine.emplace_back((int32_t)p, kv.first);
        }
    });
    tC = MsSince(t0);
    // outside-keyframe table, sequential (discovery order for now)
    std::unordered_map<const KeyFrame*, int> outsideIndex;
    vector<int32_t> cnt(M + 1, 0);
    for (const vector<OutObs>& lst : found)
        for (const OutObs& o : lst) {
            auto it = outsideIndex.find(o.second.get());
            if (it == outsideIndex.end()) {
                it = outsideIndex.emplace(o.second.get(), (int)out.vpOutsideKFs.size()).first;
                out.vpOutsideKFs.push_back(o.second);
            }
            out.mp_obs_kf.push_back(K + it->second);
            ++cnt[o.first + 1];
        }
    out.mp_obs_ptr.assign(M + 1, 0);
    for (size_t p = 0; p < M; ++p) out.mp_obs_ptr[p + 1] = out.mp_obs_ptr[p] + cnt[p + 1];
    // deterministic order of the outside rows: by keyframe id (the reference orders by pointer value, :125-126)
    const int H = (int)out.vpOutsideKFs.size();
    vector<int> order(H), rank(H);
    for (int j = 0; j < H; ++j) order[j] = j;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return out.vpOutsideKFs[a]->mnId < out.vpOutsideKFs[b]->mnId; });
    vector<shared_ptr<KeyFrame>> sorted(H);
    for (int j = 0; j < H; ++j) { rank[order[j]] = j; sorted[j] = out.vpOutsideKFs[order[j]]; }
    out.vpOutsideKFs.swap(sorted);
    for (int32_t& kf : out.mp_obs_kf) if (kf >= K) kf = K + rank[kf - K];
    out.H = H;
    out.okf_total.resize(H);
    ParallelChunks((size_t)H, nthreads, 4, [&](int, size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; ++j) out.okf_total[j] = out.vpOutsideKFs[j]->GetNumberMPs();      // :146
    });
    out.mp_ids.resize(M);
    for (size_t p = 0; p < M; ++p) out.mp_ids[p] = out.vpMapPoints[p]->mnId;
    // MSS_TIE_GID=1: ties between equally good candidates break on MapPoint::mnId instead of on the table index,
    // i.e. independently of the order in which this walk met the points (costs 4 bytes per map point on the wire)
    out.mp_tie.clear();
    static const bool tie_gid = std::getenv("MSS_TIE_GID") && std::atoi(std::getenv("MSS_TIE_GID")) != 0;
    if (tie_gid) {
        std::vector<uint32_t> order(M);
        for (size_t p = 0; p < M; ++p) order[p] = (uint32_t)p;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return out.mp_ids[a] < out.mp_ids[b]; });
        out.mp_tie.resize(M);
        for (size_t r = 0; r < M; ++r) out.mp_tie[order[r]] = (uint32_t)r;
    }
    out.okf_ids.resize(H);
    for (int j = 0; j < H; ++j) out.okf_ids[j] = out.vpOutsideKFs[j]->mnId;
    tD = MsSince(t0);
    out.Pack();
    out.flatten_ms = MsSince(t0);
    if (trace)
        std::cerr << "FlattenWindow: " << nthreads << " threads, keyframes " << tA << " ms, numbering " << (tB - tA) << ", map points "
                  << (tC - tB) << ", outside table " << (tD - tC) << ", pack " << (out.flatten_ms - tD) << ", total " << out.flatten_ms << std::endl;
}

// ---------------------------------------------------------------------------------------------------------------------
// the thread object
// ---------------------------------------------------------------------------------------------------------------------
MapSparsification::MapSparsification(const std::string& strSettingsFile, Atlas* pAtlas, bool bInertial)
    : mnMinNum(0), mbFinishRequested(false), mbFinished(true), mnId(0), mbStopRequested(false), mbStopped(true),
      mpEngine(nullptr), mfLambda(0.f), mfGridLambda(0.f), mnWindowLength(0), mpLoopClosing(nullptr), mpAtlas(pAtlas),
      mbInertial(bInertial) {
    SparsificationSettings s;
    if (!ReadSparsificationSettings(strSettingsFile, s))
        std::cerr << "MapSparsification: cannot read settings file " << strSettingsFile << std::endl;
    mnMinNum = s.N;
    mfLambda = s.Lambda;
    mfGridLambda = s.GridLambda;
    mnWindowLength = s.WindowLength;
    std::cout << std::endl << "*****************************************" << std::endl;
    std::cout << "Map Sparsification settings: " << std::endl;
    std::cout << "Sparsification.N: " << mnMinNum << std::endl;
    std::cout << "Sparsification.Lambda: " << mfLambda << std::endl;
    std::cout << "Sparsification.GridLambda: " << mfGridLambda << std::endl;
    std::cout << "Sparsification.WindowLength: " << mnWindowLength << std::endl;
    std::cout << "*****************************************" << std::endl;
    mss_config cfg{};
    const char* dev = std::getenv("MSS_DEVICE");
    cfg.device = dev ? std::atoi(dev) : 0;
    cfg.min_points = mnMinNum;
    cfg.lambda = mfLambda;
    cfg.grid_lambda = mfGridLambda;
    const int rc = mss_create(&cfg, &mpEngine);
    if (rc != MSS_OK) {
        mpEngine = nullptr;
        std::cerr << "MapSparsification: mss_create failed (" << rc << "): no usable CUDA device; windows will be forwarded "
                     "unsparsified (there is no CPU solver)" << std::endl;
    }
    // Persistent device mirror (MSS_MIRROR=0 turns it off: every window is then flattened from the pointer graph).  The
    // recorder is attached to the map before any other thread runs (System constructs the sparsifier before it starts the
    // threads, src/System.cc:159-160); keyframes already in the map are registered here.
    if (const char* bh = std::getenv("MSS_BATCHED_HANDBACK")) mbBatchedHandback = std::atoi(bh) != 0;
    // MSS_DEVICES=n (or a list "0,2,3"): the independent components of the final flush are dealt out over n GPUs of the box,
    // driven from this one process (mss_multi_*); live windows stay on the first device
    if (const char* dv = std::getenv("MSS_DEVICES")) {
        vector<int32_t> devs;
        const std::string str(dv);
        if (str.find(',') == std::string::npos) { for (int i = 0; i < std::atoi(dv); ++i) devs.push_back(i); }
        else { size_t p = 0; while (p < str.size()) { devs.push_back(std::atoi(str.c_str() + p)); p = str.find(',', p); if (p == std::string::npos) break; ++p; } }
        if (mpEngine && devs.size() > 1) {
            if (mss_multi_create(&cfg, devs.data(), (int32_t)devs.size(), &mpMulti) != MSS_OK) {
                mpMulti = nullptr;
                std::cerr << "MapSparsification: MSS_DEVICES=" << dv << ": not all of these devices are usable; the flush stays on one GPU" << std::endl;
            }
        }
    }
    const char* mir = std::getenv("MSS_MIRROR");
    if (mpEngine && mpAtlas && !(mir && std::atoi(mir) == 0)) {
        const char* sl = std::getenv("MSS_MIRROR_SLOTS");
        const int slots = sl && std::atoi(sl) > 0 ? std::atoi(sl) : 2048;
        if (mss_mirror_create(mpEngine, slots, &mpMirror) == MSS_OK) {
            mpRecorder = new MirrorRecorder(slots);
            mpAtlas->GetCurrentMap()->SetMirror(mpRecorder);
            for (const shared_ptr<KeyFrame>& pKF : mpAtlas->GetAllKeyFrames()) mpRecorder->OnKeyFrameAdded(pKF);
        } else {
            mpMirror = nullptr;
            std::cerr << "MapSparsification: device mirror unavailable: " << mss_last_error(mpEngine) << std::endl;
        }
    }
}

MapSparsification::~MapSparsification() {
    if (mGraveThread.joinable()) mGraveThread.join();
    if (mpAtlas && mpRecorder) mpAtlas->GetCurrentMap()->SetMirror(nullptr);
    if (mpMirror) mss_mirror_destroy(mpMirror);
    if (mpMulti) mss_multi_destroy(mpMulti);
    delete mpRecorder;
    if (mpEngine) mss_destroy(mpEngine);
}

void MapSparsification::Run() {
    {
        std::unique_lock<std::mutex> lock(mMutexFinish);
        mbFinished = false;
    }
    while (true) {
        if (CheckNewKeyFrames()) {
            {
                std::unique_lock<std::mutex> lock(mMutexStop);
                mbStopped = false;
            }
            vector<shared_ptr<KeyFrame>> vpKFs = GetLastestKeyFrames();
            Sparsifying(vpKFs);
            {
                std::unique_lock<std::mutex> lock(mMutexStop);
                mbStopped = true;
            }
        }
        if (CheckFinish()) {
            // final flush: one window with every keyframe that has not been sparsified yet
            vector<shared_ptr<KeyFrame>> vRemain;
            for (const shared_ptr<KeyFrame>& pKF : mpAtlas->GetAllKeyFrames())
                if (!pKF->mbSparsified) vRemain.push_back(pKF);
            mbFlushing = true;          // the flush window is split into its independent components (one batch launch)
            Sparsifying(vRemain);
            mbFlushing = false;
            for (const shared_ptr<KeyFrame>& pKF : vRemain) pKF->EraseBadDescriptor();
            break;
        }
        std::this_thread::sleep_for(std::chrono::microseconds(3000));
    }
    SetFinish();
}

int MapSparsification::EraseBatched(vector<shared_ptr<MapPoint>>& vpDrop) {
    // Everything that only frees memory (the dropped points themselves, the tree nodes of their observation maps) is
    // collected and destroyed by a helper thread: the window is done when the map no longer shows the points.
    struct Grave {
        vector<shared_ptr<MapPoint>> points;
        vector<MapPoint::ObsMap> maps;
    };
    std::unique_ptr<Grave> grave(new Grave());
    // first half of SetBadFlag for every point (bad, observations dropped), collecting the slots they sat in ...
    vector<std::pair<KeyFrame*, int>> vSlots;
    vSlots.reserve(vpDrop.size() * 4);
    vector<shared_ptr<MapPoint>> vDone;
    vDone.reserve(vpDrop.size());
    grave->maps.reserve(vpDrop.size());
    for (shared_ptr<MapPoint>& pMP : vpDrop)
        if (pMP && pMP->SetBadFlagBatched(vSlots, &grave->maps)) vDone.emplace_back(std::move(pMP));
    // ... then one lock per keyframe for all of its slots, and one lock of the map for its sets
    std::unordered_map<KeyFrame*, vector<int>> byKF;
    byKF.reserve(1024);
    for (const std::pair<KeyFrame*, int>& s : vSlots) byKF[s.first].push_back(s.second);
    grave->points.reserve(vSlots.size() + vDone.size());
    for (auto& kv : byKF) kv.first->EraseMapPointMatches(kv.second, &grave->points);
    if (!vDone.empty() && vDone[0]->GetMap()) vDone[0]->GetMap()->EraseMapPoints(vDone);
    const int n = (int)vDone.size();
    for (shared_ptr<MapPoint>& p : vDone) grave->points.emplace_back(std::move(p));
    if (mGraveThread.joinable()) mGraveThread.join();
    Grave* raw = grave.release();
    mGraveThread = std::thread([raw]() { delete raw; });
    return n;
}

bool MapSparsification::SparsifyingFromMirror(vector<shared_ptr<KeyFrame>>& vpKFs, WindowReport& rep) {
    const int K = (int)vpKFs.size();
    vector<int32_t> handles(K);
    for (int k = 0; k < K; ++k) {
        handles[k] = vpKFs[k]->mnMirrorHandle;
        if (handles[k] < 0) return false;                  // a keyframe the mirror does not hold (too many slots): flatten path
    }
    Clock::time_point t0 = Clock::now();
    int rc = mpRecorder->Flush(mpMirror);
    rep.flatten_ms = MsSince(t0);
    rep.delta_ops = mpRecorder->LastFlushOps();
    if (rc != MSS_OK) {
        std::cerr << "MapSparsification: device mirror out of sync (" << mss_last_error(mpEngine) << "): switched off" << std::endl;
        mpAtlas->GetCurrentMap()->SetMirror(nullptr);
        mss_mirror_destroy(mpMirror);
        mpMirror = nullptr;
        return false;
    }
    for (int k = 0; k < K; ++k) vpKFs[k]->mnMapSaprsificationId = mnId;       // window membership stamp (:81)
    t0 = Clock::now();
    mss_set_params(mpEngine, mnMinNum, mfLambda, mfGridLambda);
    // the final flush (:38-47) is block diagonal along the connected components: one independent window each
    vector<vector<int32_t>> groups;
    int32_t nmax = 0;
    if (mbFlushing && K > 1) {
        vector<int32_t> label(K);
        int32_t ncomp = 0;
        mss_mirror_window whole{};
        whole.K = K; whole.kf = handles.data();
        if (mss_mirror_components(mpMirror, &whole, label.data(), &ncomp, &nmax) == MSS_OK && ncomp > 1) {
            // only components that hold a window keyframe become windows (an outside keyframe that sees no variable is a
            // component of its own and constrains nothing); order of first appearance
            vector<int> dense(ncomp, -1);
            for (int k = 0; k < K; ++k) {
                if (label[k] < 0 || label[k] >= ncomp) { groups.clear(); break; }
                if (dense[label[k]] < 0) { dense[label[k]] = (int)groups.size(); groups.emplace_back(); }
                groups[dense[label[k]]].push_back(handles[k]);
            }
            if (groups.size() <= 1) groups.clear();
        }
    }
    if (groups.empty()) { groups.emplace_back(handles); nmax = 0; }
    const size_t n = groups.size();
    const size_t words = ((size_t)mpRecorder->MapPointHandles() + 31) / 32 + 1;
    vector<mss_mirror_window> wins(n);
    vector<mss_result> results(n);
    vector<vector<uint32_t>> bits(n);
    for (size_t i = 0; i < n; ++i) {
        wins[i] = mss_mirror_window{};
        wins[i].K = (int32_t)groups[i].size();
        wins[i].kf = groups[i].data();
        wins[i].n_max_floor = nmax;
        wins[i].apply = 1;                                // the device performs the deletion on its own copy of the map
        bits[i].assign(words, 0u);
        wins[i].del_bits = bits[i].data();
        wins[i].del_words = (int32_t)words;
        results[i] = mss_result{};
    }
    rc = mss_mirror_solve(mpMirror, (int32_t)n, wins.data(), results.data());
    rep.status = rc;
    rep.solve_ms = MsSince(t0);
    rep.components = (int)n;
    rep.mirror = 1;
    mss_mirror_stats ms{};
    if (mss_mirror_get_stats(mpMirror, &ms) == MSS_OK) { rep.build_ms = ms.last_build_ms; rep.h2d_bytes = (long)ms.last_h2d_bytes; rep.d2h_bytes = (long)ms.last_d2h_bytes; }
    if (rc != MSS_OK && rc != MSS_E_NOCONVERGE)
        std::cerr << "MapSparsification: window " << mnId << " not (fully) sparsified: " << mss_last_error(mpEngine) << std::endl;
    // hand-back: whatever bits came back were applied on the device; do the same to the map, without echoing it as deltas
    t0 = Clock::now();
    vector<shared_ptr<MapPoint>> vpDrop;
    for (size_t i = 0; i < n; ++i) {
        rep.H += wins[i].H; rep.M += wins[i].M;
        rep.n_vars += results[i].n_vars; rep.n_kept += results[i].n_kept; rep.objective += results[i].objective;
        rep.dual_bound = i == 0 ? results[i].dual_bound : rep.dual_bound + results[i].dual_bound;      // (NaN stays NaN)
        rep.rounds = std::max(rep.rounds, results[i].rounds);
        for (int32_t wd = wins[i].h_lo >> 5; wd < ((wins[i].h_hi + 31) >> 5); ++wd)
            for (uint32_t b = bits[i][wd]; b; b &= b - 1u) {
                shared_ptr<MapPoint> pMP = mpRecorder->PointOf(wd * 32 + __builtin_ctz(b));
                if (pMP) vpDrop.emplace_back(std::move(pMP));
            }
    }
    {
        MirrorRecorder::Suppress quiet;
        rep.n_deleted = EraseBatched(vpDrop);
    }
    rep.K = K;
    if (mpLoopClosing)
        for (const shared_ptr<KeyFrame>& pKF : vpKFs) mpLoopClosing->InsertSparsifiedKeyFrame(pKF);
    rep.apply_ms = MsSince(t0);
    return true;
}

// With MSS_DUAL_BOUND=1 every window carries the lower bound the device proved for it: the certified gap is logged the way
// GUROBI would report its MIPGap (the reference runs with OutputFlag = 0, :154, and accepts 0.2 % unseen).
static void LogCertificate(long unsigned int id, const MapSparsification::WindowReport& rep) {
    if (!(rep.dual_bound > 0.0) || rep.status != 0) return;
    std::cout << "MapSparsification: window " << id << ": F = " << rep.objective << ", certified lower bound " << rep.dual_bound
              << " (gap " << 100.0 * (rep.objective / rep.dual_bound - 1.0) << " %)" << std::endl;
}

void MapSparsification::Sparsifying(vector<shared_ptr<KeyFrame>>& vpKFs) {
    mnId++;
    WindowReport rep;
    // (a flush that is to be dealt out over several GPUs is flattened on the host: a mirror lives on one device)
    if (mpMirror && mpRecorder && !vpKFs.empty() && !(mbFlushing && mpMulti) && SparsifyingFromMirror(vpKFs, rep)) {
        LogCertificate(mnId, rep);
        std::unique_lock<std::mutex> lock(mMutexReports);
        if (mReports.size() >= kMaxReports) mReports.erase(mReports.begin());
        mReports.push_back(rep);
        return;
    }
    rep = WindowReport();
    FlattenWindow(vpKFs, mnId, mLast);
    const size_t M = mLast.vpMapPoints.size();
    rep.K = mLast.K; rep.H = mLast.H; rep.M = (int)M; rep.flatten_ms = mLast.flatten_ms;
    mKeepBits.assign((M + 31) / 32, 0xFFFFFFFFu);

    Clock::time_point t0 = Clock::now();
    int rc = MSS_E_CUDA;
    mss_result res{};
    bool solved = false;
    if (mpEngine && mbFlushing && mLast.K > 1) {
        // final flush (MapSparsification.cc:38-47): the model is block diagonal along the connected components of the
        // keyframe x variable graph -> solve the components as one batch of independent windows (same objective)
        mss_set_params(mpEngine, mnMinNum, mfLambda, mfGridLambda);
        const mss_window_view whole = mLast.View();
        vector<int32_t> rowLabel(mLast.K + mLast.H), mpLabel(M);
        int32_t ncomp = 0, nmax = 0;
        if (mss_components(mpEngine, &whole, rowLabel.data(), mpLabel.data(), &ncomp, &nmax) == MSS_OK) {
            vector<WindowSnapshot> parts;
            SplitSnapshot(mLast, rowLabel, mpLabel, nmax, parts);
            if (parts.size() > 1) {
                const size_t n = parts.size();
                vector<mss_window_view> views(n);
                vector<mss_result> results(n);
                vector<vector<uint32_t>> bits(n);
                for (size_t i = 0; i < n; ++i) {
                    views[i] = parts[i].View();
                    bits[i].assign((parts[i].mp_nobs.size() + 31) / 32, 0xFFFFFFFFu);
                    results[i] = mss_result{};
                    results[i].keep_bits = bits[i].data();
                }
                if (mpMulti) {
                    mss_multi_set_params(mpMulti, mnMinNum, mfLambda, mfGridLambda);
                    rc = mss_multi_solve_batch(mpMulti, (int32_t)n, views.data(), results.data());
                    rep.devices = mss_multi_device_count(mpMulti);
                } else {
                    rc = mss_solve_batch(mpEngine, (int32_t)n, views.data(), results.data());
                }
                if (rc == MSS_OK || rc == MSS_E_NOCONVERGE) {
                    for (size_t i = 0; i < n; ++i) {
                        for (size_t q = 0; q < parts[i].part_mp.size(); ++q)
                            if (!((bits[i][q >> 5] >> (q & 31)) & 1u)) {
                                const size_t p = (size_t)parts[i].part_mp[q];
                                mKeepBits[p >> 5] &= ~(1u << (p & 31));
                            }
                        res.n_vars += results[i].n_vars; res.n_kept += results[i].n_kept; res.objective += results[i].objective;
                        res.dual_bound = i == 0 ? results[i].dual_bound : res.dual_bound + results[i].dual_bound;
                        res.rounds = std::max(res.rounds, results[i].rounds);
                    }
                    rep.components = (int)n;
                } else {
                    std::cerr << "MapSparsification: window " << mnId << " not sparsified: "
                              << (mpMulti ? mss_multi_last_error(mpMulti) : mss_last_error(mpEngine)) << std::endl;
                }
                solved = true;
            }
        }
    }
    if (mpEngine && !solved) {
        // the yaml values can be changed through the public member between windows, like upstream
        mss_set_params(mpEngine, mnMinNum, mfLambda, mfGridLambda);
        const mss_window_view view = mLast.View();
        res.keep_bits = mKeepBits.data();
        rc = mss_solve(mpEngine, &view, &res);
        if (rc != MSS_OK && rc != MSS_E_NOCONVERGE) {
            std::cerr << "MapSparsification: window " << mnId << " not sparsified: " << mss_last_error(mpEngine) << std::endl;
            std::fill(mKeepBits.begin(), mKeepBits.end(), 0xFFFFFFFFu);
        }
    }
    rep.status = rc; rep.solve_ms = MsSince(t0);
    rep.n_vars = res.n_vars; rep.n_kept = res.n_kept; rep.rounds = res.rounds; rep.objective = res.objective;
    rep.dual_bound = res.dual_bound;

    // hand-back: delete what the selection dropped (only variables can have a 0 bit)
    t0 = Clock::now();
    if (!mbBatchedHandback) {
        for (size_t p = 0; p < M; ++p) {                 // the reference's loop (:159-166): one SetBadFlag fan-out per point
            if (!mLast.is_var[p]) continue;
            if (!((mKeepBits[p >> 5] >> (p & 31)) & 1u)) {
                mLast.vpMapPoints[p]->SetBadFlag();
                ++rep.n_deleted;
            }
        }
    } else {
        vector<shared_ptr<MapPoint>> vpDrop;
        for (size_t p = 0; p < M; ++p)
            if (mLast.is_var[p] && !((mKeepBits[p >> 5] >> (p & 31)) & 1u)) vpDrop.push_back(mLast.vpMapPoints[p]);
        rep.n_deleted = EraseBatched(vpDrop);          // (with a mirror attached the hooks record the erasures as deltas)
    }
    if (mpLoopClosing)
        for (const shared_ptr<KeyFrame>& pKF : vpKFs) mpLoopClosing->InsertSparsifiedKeyFrame(pKF);
    rep.apply_ms = MsSince(t0);
    // nothing of the window is kept alive between calls (the reference holds no state either): only the flat arrays stay
    // for LastSnapshot()
    vector<shared_ptr<MapPoint>>().swap(mLast.vpMapPoints);
    vector<shared_ptr<KeyFrame>>().swap(mLast.vpOutsideKFs);
    LogCertificate(mnId, rep);
    std::unique_lock<std::mutex> lock(mMutexReports);
    if (mReports.size() >= kMaxReports) mReports.erase(mReports.begin());
    mReports.push_back(rep);
}

vector<MapSparsification::WindowReport> MapSparsification::GetReports() {
    std::unique_lock<std::mutex> lock(mMutexReports);
    return mReports;
}

vector<shared_ptr<KeyFrame>> MapSparsification::GetLastestKeyFrames() {
    std::unique_lock<std::mutex> lock(mMutexNewKFs);
    const size_t take = std::min(mvpNewKeyFrames.size(), (size_t)std::max(mnWindowLength, 0));
    vector<shared_ptr<KeyFrame>> vKFs(mvpNewKeyFrames.begin(), mvpNewKeyFrames.begin() + take);
    mvpNewKeyFrames.erase(mvpNewKeyFrames.begin(), mvpNewKeyFrames.begin() + take);
    return vKFs;
}

void MapSparsification::InsertKeyFrame(shared_ptr<KeyFrame> pKF) {
    std::unique_lock<std::mutex> lock(mMutexNewKFs);
    mvpNewKeyFrames.push_back(pKF);
}

bool MapSparsification::CheckNewKeyFrames() {
    std::unique_lock<std::mutex> lock2(mMutexStop);
    std::unique_lock<std::mutex> lock(mMutexNewKFs);
    return mvpNewKeyFrames.size() > 10 && !mbStopRequested &&
           (!mbInertial || mvpNewKeyFrames[0]->GetMap()->GetIniertialBA2());
}

void MapSparsification::SetLoopClosing(LoopClosing* pLoopClosing) { mpLoopClosing = pLoopClosing; }

void MapSparsification::RequestStop() {
    std::unique_lock<std::mutex> lock2(mMutexStop);
    mbStopRequested = true;
    std::cout << "Map Sparsification STOP" << std::endl;
}

void MapSparsification::Release() {
    std::unique_lock<std::mutex> lock2(mMutexStop);
    mbStopRequested = false;
    std::cout << "Map Sparsification RELEASE" << std::endl;
}

bool MapSparsification::isStopped() {
    std::unique_lock<std::mutex> lock2(mMutexStop);
    return mbStopped;
}

void MapSparsification::RequestFinish() {
    std::unique_lock<std::mutex> lock(mMutexFinish);
    mbFinishRequested = true;
}

bool MapSparsification::CheckFinish() {
    std::unique_lock<std::mutex> lock(mMutexFinish);
    return mbFinishRequested;
}

void MapSparsification::SetFinish() {
    std::unique_lock<std::mutex> lock(mMutexFinish);
    mbFinished = true;
}

bool MapSparsification::isFinished() {
    std::unique_lock<std::mutex> lock(mMutexFinish);
    return mbFinished;
}

}  // namespace ORB_SLAM3

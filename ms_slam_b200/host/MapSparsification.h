// MapSparsification.h -- host-side mirror of MS-SLAM's sparsifier thread, B200 engine underneath.
//
// Same class, same namespace, same public members as /root/reference/include/MapSparsification.h:26-70, so the call
// sites in System (src/System.cc:159-162,460-466), LocalMapping (src/LocalMapping.cc:264), Tracking
// (src/Tracking.cc:3626) and LoopClosing (src/LoopClosing.cc:930,955,1162,2441) compile unchanged.  What differs is
// private: the GUROBI environment member (`GRBEnv mGRBEnv`, :59) is an opaque `mss_handle*` of libmss (include/mss.h),
// and Sparsifying() snapshots the window into a flat view, calls mss_solve and applies the returned keep-bitmask instead
// of building a GRBModel.
#pragma once

#include <mutex>
#include <string>
#include <thread>
#include <limits>
#include <vector>

#include "../../include/mss.h"
#ifdef MSS_WITH_ORBSLAM3_HEADERS      // building inside an MS-SLAM checkout
#include "KeyFrame.h"
#include "Map.h"
#include "Atlas.h"
#include "LoopClosing.h"
#else
#include "SlamShims.h"
#endif
#include "MirrorRecorder.h"

namespace ORB_SLAM3 {

// Flat snapshot of one window: exactly the arrays of mss_window_view plus the objects the bits refer to.
struct WindowSnapshot {
    std::vector<int32_t> feat_ptr, feat_mp, mp_nobs, mp_obs_ptr, mp_obs_kf, okf_total;
    std::vector<uint16_t> feat_cell;
    std::vector<uint8_t> is_var;                              // map point reachable through a grid cell (an ILP variable)
    std::vector<std::shared_ptr<MapPoint>> vpMapPoints;       // table order = mnIndexForSparsification = bit position
    std::vector<std::shared_ptr<KeyFrame>> vpOutsideKFs;      // ordered by KeyFrame::mnId
    std::vector<long unsigned int> mp_ids, okf_ids;           // mnId of the above (stay valid after the objects are released)
    std::vector<uint32_t> mp_tie;                             // rank of mp_ids (MSS_TIE_GID=1: mss_window_view::mp_tie), else empty
    int K = 0, H = 0;
    double flatten_ms = 0.0;
    // MSS_LAYOUT_PACKED16 transport form of the same arrays in ONE host blob (pinned when a CUDA device is present), laid
    // out back to back at 16-byte boundaries so the engine moves the window with a single copy.  Built by FlattenWindow
    // whenever the window fits the packed ranges (M <= 2^20, nObs <= 65535, H <= 4095); View() then returns it.
    struct Blob {
        uint8_t* p = nullptr;
        size_t cap = 0;
        bool pinned = false;
        ~Blob();
        void Reserve(size_t bytes);
    };
    std::shared_ptr<Blob> blob;
    size_t off_slots = 0, off_nobs = 0, off_pairs = 0, off_okf = 0, n_pairs = 0, n_tokens = 0;
    bool packed = false;
    bool nobs8 = false;                 // the blob carries Observations() as one byte per map point (all <= 255)
    int n_max_floor = 0;       // window-wide nMax carried by a component of a larger window (mss.h)
    std::vector<int32_t> part_mp;   // component only: index of each of its map points in the parent snapshot
    void Pack();
    mss_window_view View() const;
};

// Passes 1-3 of the reference (MapSparsification.cc:66-151) as ONE walk over the pointer graph that only records what it
// sees; stamps mnMapSaprsificationId / mnMapSparsificationId / mnIndexForSparsification like the reference (:81,93,98).
void FlattenWindow(const std::vector<std::shared_ptr<KeyFrame>>& vpKFs, long unsigned int nId, WindowSnapshot& out);

// Independent sub-windows of a snapshot from the labels of mss_components: one per component that contains a window
// keyframe, each carrying the window-wide nMax.  The reference solves the final flush as ONE model
// (MapSparsification.cc:38-47); the model is block diagonal along these components, so solving them as a batch gives the
// same objective and lets a flush shard over GPUs.
void SplitSnapshot(const WindowSnapshot& in, const std::vector<int32_t>& row_label, const std::vector<int32_t>& mp_label,
                   int n_max, std::vector<WindowSnapshot>& parts);

struct SparsificationSettings {
    int N = 0, WindowLength = 0, NonLocalKF = 0;
    float Lambda = 0.f, GridLambda = 0.f;
};
// Reads the `Sparsification.*` keys of an ORB-SLAM3 settings file (OpenCV-YAML subset: `key: value` lines); missing keys
// stay 0 exactly like cv::FileNode -> int/float does upstream (MapSparsification.cc:8-12).
bool ReadSparsificationSettings(const std::string& path, SparsificationSettings& s);

class MapSparsification {
public:
    MapSparsification(const std::string& strSettingsFile, Atlas* pAtlas, bool bInertial);
    ~MapSparsification();

    void Run();

    bool CheckNewKeyFrames();

    void InsertKeyFrame(std::shared_ptr<KeyFrame> pKF);

    void SetLoopClosing(LoopClosing* pLoopClosing);

    std::vector<std::shared_ptr<KeyFrame>> GetLastestKeyFrames();

    bool isStopped();
    void RequestStop();
    void Release();
    void RequestFinish();
    bool isFinished();
    int mnMinNum;

    // ---- additions (not in the reference; read-only diagnostics) -------------------------------------------------------
    struct WindowReport {
        int status = 0;                 // mss_status of the solve (0 ok); on error every map point was kept
        int K = 0, H = 0, M = 0, n_vars = 0, n_kept = 0, n_deleted = 0, rounds = 0;
        int components = 1;             // independent sub-windows the window was solved as (one batch launch)
        double objective = 0.0, flatten_ms = 0.0, solve_ms = 0.0, apply_ms = 0.0;
        int mirror = 0;                 // 1 = solved from the device mirror (flatten_ms is then the delta drain, not a graph walk)
        long delta_ops = 0;             // records drained into the mirror before this window
        double build_ms = 0.0;          // device time of the view assembly (mirror) inside solve_ms
        long h2d_bytes = 0, d2h_bytes = 0;
        int devices = 1;                // GPUs the window's components were dealt out to
        double dual_bound = std::numeric_limits<double>::quiet_NaN();   // lower bound of the reference ILP proven on the device
                                        // (sum over the components; NaN unless MSS_DUAL_BOUND=1): objective / dual_bound - 1 is
                                        // a certified optimality gap, the counterpart of GUROBI's MIPGap (:155-156)
    };
    std::vector<WindowReport> GetReports();           // the last kMaxReports windows
    static constexpr size_t kMaxReports = 256;
    bool MirrorActive() const { return mpMirror != nullptr; }
    const WindowSnapshot& LastSnapshot() const { return mLast; }      // valid while the thread is stopped
    bool EngineReady() const { return mpEngine != nullptr; }

private:
    void Sparsifying(std::vector<std::shared_ptr<KeyFrame>>& vpKFs);
    // Sparsifying() from the device mirror: K keyframe handles up, bitmask over map-point handles down
    bool SparsifyingFromMirror(std::vector<std::shared_ptr<KeyFrame>>& vpKFs, WindowReport& rep);
    // hand-back of many map points at once: one lock per keyframe and one for the map instead of the per-point
    // SetBadFlag fan-out (src/MapSparsification.cc:159-166, src/MapPoint.cc:227-255, src/Map.cc:109-113)
    int EraseBatched(std::vector<std::shared_ptr<MapPoint>>& vpDrop);
    bool CheckFinish();
    void SetFinish();

    bool mbFinishRequested;
    bool mbFinished;
    std::mutex mMutexFinish;

    long unsigned int mnId;
    bool mbStopRequested;
    bool mbStopped;
    bool mbFlushing = false;            // Sparsifying() is running the final flush

    mss_handle* mpEngine;               // replaces GRBEnv mGRBEnv (include/MapSparsification.h:59)
    mss_multi* mpMulti = nullptr;       // MSS_DEVICES > 1: the components of the final flush are dealt out over several GPUs
    mss_mirror* mpMirror = nullptr;     // persistent device mirror of the incidence (include/mss.h), fed by mpRecorder
    MirrorRecorder* mpRecorder = nullptr;
    std::thread mGraveThread;           // frees what a batched hand-back released (see EraseBatched)
    bool mbBatchedHandback = true;      // MSS_BATCHED_HANDBACK=0: per-point SetBadFlag like the reference (flatten path only)
    float mfLambda;
    float mfGridLambda;
    int mnWindowLength;
    std::vector<std::shared_ptr<KeyFrame>> mvpNewKeyFrames;
    std::mutex mMutexNewKFs;
    std::mutex mMutexStop;
    LoopClosing* mpLoopClosing;
    Atlas* mpAtlas;
    bool mbInertial;

    WindowSnapshot mLast;
    std::vector<uint32_t> mKeepBits;
    std::vector<WindowReport> mReports;
    std::mutex mMutexReports;
};

}  // namespace ORB_SLAM3

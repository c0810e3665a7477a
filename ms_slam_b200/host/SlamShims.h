// SlamShims.h -- the touch-set of the ORB-SLAM3 data model that MapSparsification uses, and nothing else.
//
// MS-SLAM's sparsifier reads and writes a pointer graph of KeyFrame / MapPoint / Map / Atlas / LoopClosing objects
// (SURVEY.md section 2.1 "touch-set only" rows).  Rebuilding ORB-SLAM3 is out of scope, so this header declares just
// the members the hot path calls, with the reference's names and signatures, so that host/MapSparsification.cc compiles
// unchanged against either these shims (tests, benchmarks) or the real headers of an MS-SLAM checkout.
//
//   KeyFrame   /root/reference/include/KeyFrame.h:110-120,148,161-168,191,272-274,298,307
//   MapPoint   /root/reference/include/MapPoint.h:59-70,101,118,122,150
//   Map        /root/reference/include/Map.h:51-53,104,137-138
//   Atlas      /root/reference/include/Atlas.h:82
//   LoopClosing/root/reference/include/LoopClosing.h:67,126,173-176
//
// No OpenCV / Eigen / DBoW2: descriptors, poses and the BoW vocabulary do not influence the selection.
#pragma once

#include <atomic>
#include <cstddef>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

namespace ORB_SLAM3 {

using std::shared_ptr;

constexpr int FRAME_GRID_ROWS = 48;   // include/Frame.h:44
constexpr int FRAME_GRID_COLS = 64;   // include/Frame.h:45

class KeyFrame;
class MapPoint;
class Map;
class MirrorRecorder;      // MirrorRecorder.h: the map tells the device mirror what changed (hooks in SlamShims.cc = the
                           // patch INTEGRATION.md lists for MapPoint.cc / KeyFrame.cc / Map.cc of an MS-SLAM checkout)

typedef std::vector<std::vector<std::vector<size_t>>> FeatureGrid;   // [col][row] -> slot indices (KeyFrame::mGrid)

class MapPoint : public std::enable_shared_from_this<MapPoint> {
public:
    MapPoint(long unsigned int id, Map* pMap);

    std::map<shared_ptr<KeyFrame>, std::tuple<int, int>> GetObservations();
    int Observations();
    void AddObservation(shared_ptr<KeyFrame> pKF, int idx);      // += 2 for a stereo keypoint (src/MapPoint.cc:155-158)
    void UpdateObservation(shared_ptr<KeyFrame> pKF, int idx);
    void EraseObservation(shared_ptr<KeyFrame> pKF);
    void SetBadFlag();
    bool isBad();
    Map* GetMap() { return mpMap; }
    std::tuple<int, int> GetIndexInKeyFrame(shared_ptr<KeyFrame> pKF);       // src/MapPoint.cc:437-444
    // Batched hand-back (not upstream): the first half of SetBadFlag (src/MapPoint.cc:227-243) -- mark bad, drop
    // the observations -- returning the (keyframe, slot) pairs instead of erasing them one by one; the caller clears the
    // slots with one lock per keyframe (KeyFrame::EraseMapPointMatches) and the map sets with one lock (Map::EraseMapPoints).
    typedef std::map<shared_ptr<KeyFrame>, std::tuple<int, int>> ObsMap;
    bool SetBadFlagBatched(std::vector<std::pair<KeyFrame*, int>>& vSlots, std::vector<ObsMap>* pGrave = nullptr);

    long unsigned int mnId;
    int nObs;
    std::atomic<int> mnMirrorHandle{-1};          // handle in the device mirror (MirrorRecorder), -1 = none yet
    long unsigned int mnMapSparsificationId;      // window stamp (include/MapPoint.h:118)
    long unsigned int mnIndexForSparsification;   // bit position in the keep mask (include/MapPoint.h:122)

protected:
    std::map<shared_ptr<KeyFrame>, std::tuple<int, int>> mObservations;
    bool mbBad;
    Map* mpMap;
    std::mutex mMutexFeatures;
};

class KeyFrame : public std::enable_shared_from_this<KeyFrame> {
public:
    // nSlots feature slots; uRight[i] >= 0 marks a stereo keypoint
    KeyFrame(long unsigned int id, Map* pMap, size_t nSlots);

    int GetNumberMPs();
    void AddMapPoint(shared_ptr<MapPoint> pMP, const size_t& idx);
    void EraseMapPointMatch(const int& idx);
    // batched hand-back: one lock for all of them (not upstream); the emptied slots' references go to pGrave when given
    void EraseMapPointMatches(const std::vector<int>& vIdx, std::vector<shared_ptr<MapPoint>>* pGrave = nullptr);
    void EraseBadDescriptor();
    std::vector<shared_ptr<MapPoint>> GetMapPointMatches();
    shared_ptr<MapPoint> GetMapPoint(const size_t& idx);
    float GetuRight(int idx);
    Map* GetMap() { return mpMap; }
    // The reference returns mGrid by value (a deep copy of 3072 vectors, include/KeyFrame.h:161); a const reference
    // is source-compatible with every caller and lets the flatten pass read it in place.
    const FeatureGrid& GetFeatureGrids() { return mGrid; }
    void SetGridCell(int col, int row, size_t idx);        // what Frame::AssignFeaturesToGrid does (src/Frame.cc:385-416)
    bool UpdateCountInLocalMapping(bool bLocal);
    bool UpdateCountInTracking(bool bLocal);
    bool isNonLocal();

    long unsigned int mnId;
    int N;
    std::atomic<int> mnMirrorHandle{-1};         // handle in the device mirror (MirrorRecorder), -1 = not registered
    long unsigned int mnMapSaprsificationId;     // [sic] include/KeyFrame.h:191
    bool mbSparsified;                           // include/KeyFrame.h:272
    static int mnNonLocalKF;                     // include/KeyFrame.h:274
    std::vector<float> mvuRight;
    int mnEraseBadDescriptorCalls;               // test instrumentation (may run twice)

protected:
    std::vector<shared_ptr<MapPoint>> mvpMapPoints;
    FeatureGrid mGrid;
    Map* mpMap;
    int mnCountInLocal;
    bool mbNonLocalKF;
    std::mutex mMutexFeatures;
};

class Map {
public:
    Map() : mbIMU_BA2(false) {}
    void AddKeyFrame(shared_ptr<KeyFrame> pKF);
    void AddMapPoint(shared_ptr<MapPoint> pMP);
    void EraseMapPoint(shared_ptr<MapPoint> pMP);
    void EraseMapPoints(const std::vector<shared_ptr<MapPoint>>& vpMPs);     // batched hand-back: one lock (not upstream)
    void SetMirror(MirrorRecorder* pRec) { mpMirror = pRec; }
    MirrorRecorder* mpMirror = nullptr;          // set once, before the threads start
    void AddSparsifiedMapPoint(shared_ptr<MapPoint> pMP);
    void AddSparsifiedKeyFrame(shared_ptr<KeyFrame> pKF);
    std::vector<shared_ptr<KeyFrame>> GetAllKeyFrames();
    std::vector<shared_ptr<MapPoint>> GetAllMapPoints();
    long unsigned int MapPointsInMap();
    long unsigned int SparsifiedMapPointsInMap();
    std::vector<shared_ptr<KeyFrame>> GetAllSparsifiedKeyFrames();     // include/Map.h:60
    void SetIniertialBA2() { std::unique_lock<std::mutex> l(mMutexMap); mbIMU_BA2 = true; }
    bool GetIniertialBA2() { std::unique_lock<std::mutex> l(mMutexMap); return mbIMU_BA2; }

protected:
    struct ById {
        template <class T> bool operator()(const shared_ptr<T>& a, const shared_ptr<T>& b) const { return a->mnId < b->mnId; }
    };
    std::set<shared_ptr<KeyFrame>, ById> mspKeyFrames;
    std::set<shared_ptr<MapPoint>, ById> mspMapPoints;
    std::set<shared_ptr<MapPoint>, ById> mspSparsifiedMapPoints;
    std::set<shared_ptr<KeyFrame>, ById> mspSparsifiedKeyFrames;
    bool mbIMU_BA2;
    std::mutex mMutexMap;
};

class Atlas {
public:
    Atlas() : mpCurrentMap(new Map()) {}
    ~Atlas() { delete mpCurrentMap; }
    Map* GetCurrentMap() { return mpCurrentMap; }
    std::vector<shared_ptr<KeyFrame>> GetAllKeyFrames();

protected:
    Map* mpCurrentMap;
    std::mutex mMutexAtlas;
};

class LoopClosing {
public:
    explicit LoopClosing(Atlas* pAtlas) : mpAtlas(pAtlas) {}
    void InsertSparsifiedKeyFrame(shared_ptr<KeyFrame> pKF);
    // the consumer side of the queue (src/LoopClosing.cc:318-328) minus the KeyFrameDatabase insert
    void DeleteOutdatedInfo();
    size_t SparsifiedQueueSize();
    std::vector<long unsigned int> mvForwardedIds;     // test instrumentation: every KF id ever forwarded, in order

protected:
    Atlas* mpAtlas;
    std::list<shared_ptr<KeyFrame>> mlpSparsifiedKeyFrameQueue;
    std::mutex mMutexLoopQueue2;
};

}  // namespace ORB_SLAM3

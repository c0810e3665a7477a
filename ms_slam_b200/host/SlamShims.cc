// SlamShims.cc -- behaviour of the touch-set declared in SlamShims.h.  Written from the behaviour the sparsifier relies on
// (SURVEY.md sections 2.1 and 8a, rows a7-a10), not from the reference sources:
//   * a stereo observation counts twice in MapPoint::nObs            (/root/reference/src/MapPoint.cc:155-158)
//   * a map point with two or fewer observations left is discarded   (/root/reference/src/MapPoint.cc:201-203)
//   * SetBadFlag empties the point's slots in every observing keyframe and removes it from the map  (:227-255)
//   * GetNumberMPs counts the slots that hold a good point           (/root/reference/src/KeyFrame.cc:286-297)
//   * EraseBadDescriptor compacts a keyframe to its surviving slots, re-indexes their observations, drops the feature
//     grid and marks the keyframe sparsified                         (/root/reference/src/KeyFrame.cc:311-361)
//   * a keyframe becomes "non-local" after mnNonLocalKF consecutive updates in which it was not local  (:980-1016)
#include "SlamShims.h"
#include "MirrorRecorder.h"

#include <algorithm>

namespace ORB_SLAM3 {

int KeyFrame::mnNonLocalKF = 30;     // Sparsification.NonLocalKF (Examples/Stereo/KITTI00-02.yaml:74)

// ---- MapPoint ---------------------------------------------------------------------------------------------------------
MapPoint::MapPoint(long unsigned int id, Map* pMap)
    : mnId(id), nObs(0), mnMapSparsificationId(0), mnIndexForSparsification(0), mbBad(false), mpMap(pMap) {}

std::map<shared_ptr<KeyFrame>, std::tuple<int, int>> MapPoint::GetObservations() {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return mObservations;
}

int MapPoint::Observations() {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return nObs;
}

// ---- device-mirror hooks (MirrorRecorder.h): one line where the map changes, under the object's own mutex ------------------
static inline MirrorRecorder* Rec(Map* pMap) { return pMap ? pMap->mpMirror : nullptr; }

void MapPoint::AddObservation(shared_ptr<KeyFrame> pKF, int idx) {
    const bool stereo = pKF->GetuRight(idx) >= 0.f;
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    auto it = mObservations.find(pKF);
    int old = -1;
    if (it == mObservations.end()) mObservations.emplace(pKF, std::make_tuple(idx, -1));
    else { old = std::get<0>(it->second); std::get<0>(it->second) = idx; }
    nObs += stereo ? 2 : 1;
    if (MirrorRecorder* r = Rec(mpMap)) {
        if (old != -1 && old != idx) r->OnObservation(pKF.get(), old, nullptr);
        r->OnObservation(pKF.get(), idx, this);
        r->OnMapPoint(this, nObs, mbBad);
    }
}

void MapPoint::UpdateObservation(shared_ptr<KeyFrame> pKF, int idx) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    auto it = mObservations.find(pKF);
    int old = -1;
    if (it == mObservations.end()) mObservations.emplace(pKF, std::make_tuple(idx, -1));
    else { old = std::get<0>(it->second); std::get<0>(it->second) = idx; }
    if (MirrorRecorder* r = Rec(mpMap)) {
        if (old != -1 && old != idx) r->OnObservation(pKF.get(), old, nullptr);
        r->OnObservation(pKF.get(), idx, this);
    }
}

std::tuple<int, int> MapPoint::GetIndexInKeyFrame(shared_ptr<KeyFrame> pKF) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    auto it = mObservations.find(pKF);
    return it == mObservations.end() ? std::make_tuple(-1, -1) : it->second;
}

bool MapPoint::SetBadFlagBatched(std::vector<std::pair<KeyFrame*, int>>& vSlots, std::vector<ObsMap>* pGrave) {
    ObsMap obs;
    {
        std::unique_lock<std::mutex> lock(mMutexFeatures);
        if (mbBad) return false;
        mbBad = true;
        obs.swap(mObservations);
        if (MirrorRecorder* r = Rec(mpMap)) r->OnMapPoint(this, nObs, true);        // (suppressed by the batched hand-back)
    }
    for (auto& kv : obs) {
        if (std::get<0>(kv.second) != -1) vSlots.emplace_back(kv.first.get(), std::get<0>(kv.second));
        if (std::get<1>(kv.second) != -1) vSlots.emplace_back(kv.first.get(), std::get<1>(kv.second));
    }
    if (pGrave) pGrave->emplace_back(std::move(obs));       // the tree nodes are freed off the critical path
    return true;
}

void MapPoint::EraseObservation(shared_ptr<KeyFrame> pKF) {
    bool discard = false;
    {
        std::unique_lock<std::mutex> lock(mMutexFeatures);
        auto it = mObservations.find(pKF);
        if (it == mObservations.end()) return;
        const int left = std::get<0>(it->second);
        if (left != -1) nObs -= (pKF->GetuRight(left) >= 0.f) ? 2 : 1;
        if (std::get<1>(it->second) != -1) nObs -= 1;
        if (MirrorRecorder* r = Rec(mpMap)) {
            r->OnObservation(pKF.get(), left != -1 ? left : std::get<1>(it->second), nullptr);
            r->OnMapPoint(this, nObs, mbBad);
        }
        mObservations.erase(it);
        discard = nObs <= 2;
    }
    if (discard) SetBadFlag();
}

void MapPoint::SetBadFlag() {
    std::map<shared_ptr<KeyFrame>, std::tuple<int, int>> obs;
    {
        std::unique_lock<std::mutex> lock(mMutexFeatures);
        if (mbBad) return;
        mbBad = true;
        obs.swap(mObservations);
        if (MirrorRecorder* r = Rec(mpMap)) r->OnMapPoint(this, nObs, true);
    }
    for (auto& kv : obs) {
        const int left = std::get<0>(kv.second), right = std::get<1>(kv.second);
        if (MirrorRecorder* r = Rec(mpMap)) r->OnObservation(kv.first.get(), left != -1 ? left : right, nullptr);
        if (left != -1) kv.first->EraseMapPointMatch(left);
        if (right != -1) kv.first->EraseMapPointMatch(right);
    }
    if (mpMap) mpMap->EraseMapPoint(shared_from_this());
}

bool MapPoint::isBad() {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return mbBad;
}

// ---- KeyFrame ---------------------------------------------------------------------------------------------------------
KeyFrame::KeyFrame(long unsigned int id, Map* pMap, size_t nSlots)
    : mnId(id), N((int)nSlots), mnMapSaprsificationId(0), mbSparsified(false), mvuRight(nSlots, -1.f),
      mnEraseBadDescriptorCalls(0), mvpMapPoints(nSlots), mpMap(pMap), mnCountInLocal(0), mbNonLocalKF(false) {
    mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
}

int KeyFrame::GetNumberMPs() {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    int n = 0;
    for (const auto& p : mvpMapPoints)
        if (p && !p->isBad()) ++n;
    return n;
}

void KeyFrame::AddMapPoint(shared_ptr<MapPoint> pMP, const size_t& idx) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    mvpMapPoints[idx] = pMP;
    if (MirrorRecorder* r = Rec(mpMap)) r->OnSlot(this, (int)idx, pMP.get());
}

void KeyFrame::EraseMapPointMatch(const int& idx) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    if (idx >= 0 && idx < (int)mvpMapPoints.size()) {
        mvpMapPoints[idx].reset();
        if (MirrorRecorder* r = Rec(mpMap)) r->OnSlot(this, idx, nullptr);
    }
}

void KeyFrame::EraseMapPointMatches(const std::vector<int>& vIdx, std::vector<shared_ptr<MapPoint>>* pGrave) {
    std::vector<shared_ptr<MapPoint>> local;               // the last references die outside the lock
    std::vector<shared_ptr<MapPoint>>& released = pGrave ? *pGrave : local;
    released.reserve(released.size() + vIdx.size());
    {
        std::unique_lock<std::mutex> lock(mMutexFeatures);
        for (int idx : vIdx)
            if (idx >= 0 && idx < (int)mvpMapPoints.size() && mvpMapPoints[idx]) {
                released.emplace_back(std::move(mvpMapPoints[idx]));
                mvpMapPoints[idx].reset();
                if (MirrorRecorder* r = Rec(mpMap)) r->OnSlot(this, idx, nullptr);
            }
    }
}

std::vector<shared_ptr<MapPoint>> KeyFrame::GetMapPointMatches() {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return mvpMapPoints;
}

shared_ptr<MapPoint> KeyFrame::GetMapPoint(const size_t& idx) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return idx < mvpMapPoints.size() ? mvpMapPoints[idx] : shared_ptr<MapPoint>();
}

float KeyFrame::GetuRight(int idx) {
    return (idx >= 0 && idx < (int)mvuRight.size()) ? mvuRight[idx] : -1.f;
}

void KeyFrame::SetGridCell(int col, int row, size_t idx) {
    if (col >= 0 && col < FRAME_GRID_COLS && row >= 0 && row < FRAME_GRID_ROWS) mGrid[col][row].push_back(idx);
}

void KeyFrame::EraseBadDescriptor() {
    shared_ptr<KeyFrame> self = shared_from_this();
    // the mirror gets ONE compaction record instead of one observation update per surviving slot
    struct Note { KeyFrame* kf; MirrorRecorder::Suppress quiet; ~Note() { if (MirrorRecorder* r = Rec(kf->GetMap())) r->OnCompact(kf); } } note{this, {}};
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    ++mnEraseBadDescriptorCalls;
    std::vector<shared_ptr<MapPoint>> kept;
    std::vector<float> keptRight;
    for (size_t i = 0; i < mvpMapPoints.size(); ++i) {
        if (!mvpMapPoints[i]) continue;
        mvpMapPoints[i]->UpdateObservation(self, (int)kept.size());
        kept.push_back(mvpMapPoints[i]);
        keptRight.push_back(i < mvuRight.size() ? mvuRight[i] : -1.f);
    }
    mvpMapPoints.swap(kept);
    mvuRight.swap(keptRight);
    N = (int)mvpMapPoints.size();
    FeatureGrid().swap(mGrid);          // the grid (and, upstream, the raw keypoints) are not kept for sparsified keyframes
    mbSparsified = true;
}

static bool CountNonLocal(bool bLocal, int& counter, bool& flag, int limit) {
    if (bLocal) { counter = 0; flag = false; return false; }
    ++counter;
    flag = counter >= limit;
    return flag;
}

bool KeyFrame::UpdateCountInLocalMapping(bool bLocal) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return CountNonLocal(bLocal, mnCountInLocal, mbNonLocalKF, mnNonLocalKF);
}

bool KeyFrame::UpdateCountInTracking(bool bLocal) {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return CountNonLocal(bLocal, mnCountInLocal, mbNonLocalKF, mnNonLocalKF);
}

bool KeyFrame::isNonLocal() {
    std::unique_lock<std::mutex> lock(mMutexFeatures);
    return mbNonLocalKF;
}

// ---- Map / Atlas ------------------------------------------------------------------------------------------------------
void Map::AddKeyFrame(shared_ptr<KeyFrame> pKF) {
    {
        std::unique_lock<std::mutex> l(mMutexMap);
        mspKeyFrames.insert(pKF);
    }
    if (mpMirror) mpMirror->OnKeyFrameAdded(pKF);
}
void Map::AddMapPoint(shared_ptr<MapPoint> pMP) { std::unique_lock<std::mutex> l(mMutexMap); mspMapPoints.insert(pMP); }

void Map::EraseMapPoint(shared_ptr<MapPoint> pMP) {
    std::unique_lock<std::mutex> l(mMutexMap);
    mspMapPoints.erase(pMP);
    mspSparsifiedMapPoints.erase(pMP);
}

void Map::EraseMapPoints(const std::vector<shared_ptr<MapPoint>>& vpMPs) {
    std::unique_lock<std::mutex> l(mMutexMap);
    if (vpMPs.size() * 8 < mspMapPoints.size()) {          // few: one lookup each
        for (const shared_ptr<MapPoint>& p : vpMPs) {
            mspMapPoints.erase(p);
            mspSparsifiedMapPoints.erase(p);
        }
        return;
    }
    // many (a sparsified window loses ~90 % of its points): one sweep over the sets instead of a tree search per point
    std::vector<const MapPoint*> victims;
    victims.reserve(vpMPs.size());
    for (const shared_ptr<MapPoint>& p : vpMPs) victims.push_back(p.get());
    std::sort(victims.begin(), victims.end());
    auto sweep = [&](std::set<shared_ptr<MapPoint>, ById>& st) {
        for (auto it = st.begin(); it != st.end();)
            if (std::binary_search(victims.begin(), victims.end(), (const MapPoint*)it->get())) it = st.erase(it);
            else ++it;
    };
    sweep(mspMapPoints);
    if (!mspSparsifiedMapPoints.empty()) sweep(mspSparsifiedMapPoints);
}

void Map::AddSparsifiedMapPoint(shared_ptr<MapPoint> pMP) {
    if (pMP) mspSparsifiedMapPoints.insert(pMP);       // caller holds mMutexMap (AddSparsifiedKeyFrame)
}

void Map::AddSparsifiedKeyFrame(shared_ptr<KeyFrame> pKF) {
    const std::vector<shared_ptr<MapPoint>> vMPs = pKF->GetMapPointMatches();
    std::unique_lock<std::mutex> l(mMutexMap);
    mspSparsifiedKeyFrames.insert(pKF);
    for (const auto& p : vMPs) AddSparsifiedMapPoint(p);
}

std::vector<shared_ptr<KeyFrame>> Map::GetAllKeyFrames() {
    std::unique_lock<std::mutex> l(mMutexMap);
    return std::vector<shared_ptr<KeyFrame>>(mspKeyFrames.begin(), mspKeyFrames.end());
}

std::vector<shared_ptr<MapPoint>> Map::GetAllMapPoints() {
    std::unique_lock<std::mutex> l(mMutexMap);
    return std::vector<shared_ptr<MapPoint>>(mspMapPoints.begin(), mspMapPoints.end());
}

long unsigned int Map::MapPointsInMap() { std::unique_lock<std::mutex> l(mMutexMap); return mspMapPoints.size(); }
long unsigned int Map::SparsifiedMapPointsInMap() { std::unique_lock<std::mutex> l(mMutexMap); return mspSparsifiedMapPoints.size(); }
std::vector<shared_ptr<KeyFrame>> Map::GetAllSparsifiedKeyFrames() {
    std::unique_lock<std::mutex> l(mMutexMap);
    return std::vector<shared_ptr<KeyFrame>>(mspSparsifiedKeyFrames.begin(), mspSparsifiedKeyFrames.end());
}

std::vector<shared_ptr<KeyFrame>> Atlas::GetAllKeyFrames() {
    std::unique_lock<std::mutex> l(mMutexAtlas);
    return mpCurrentMap->GetAllKeyFrames();
}

// ---- LoopClosing ------------------------------------------------------------------------------------------------------
void LoopClosing::InsertSparsifiedKeyFrame(shared_ptr<KeyFrame> pKF) {
    std::unique_lock<std::mutex> l(mMutexLoopQueue2);
    mlpSparsifiedKeyFrameQueue.push_back(pKF);
    mvForwardedIds.push_back(pKF->mnId);
}

void LoopClosing::DeleteOutdatedInfo() {
    std::unique_lock<std::mutex> l(mMutexLoopQueue2);
    Map* pMap = mpAtlas->GetCurrentMap();
    while (!mlpSparsifiedKeyFrameQueue.empty()) {
        shared_ptr<KeyFrame> pKF = mlpSparsifiedKeyFrameQueue.front();
        mlpSparsifiedKeyFrameQueue.pop_front();
        pKF->EraseBadDescriptor();
        pMap->AddSparsifiedKeyFrame(pKF);       // upstream also re-inserts pKF into the KeyFrameDatabase here (out of scope)
    }
}

size_t LoopClosing::SparsifiedQueueSize() {
    std::unique_lock<std::mutex> l(mMutexLoopQueue2);
    return mlpSparsifiedKeyFrameQueue.size();
}

}  // namespace ORB_SLAM3

// MirrorRecorder.h -- host side of the persistent device mirror (include/mss.h mss_mirror_*).
//
// The sparsifier of the reference re-reads the whole neighbourhood of a window from the pointer graph every time
// (/root/reference/src/MapSparsification.cc:67-151).  Here the map tells the mirror what changed, where it changes:
//   MapPoint::AddObservation / UpdateObservation / EraseObservation / SetBadFlag   (/root/reference/src/MapPoint.cc:133-255)
//   KeyFrame::AddMapPoint / EraseMapPointMatch / EraseBadDescriptor                (/root/reference/src/KeyFrame.cc:299-361)
//   Map::AddKeyFrame                                                               (/root/reference/src/Map.cc)
// call the On* hooks below (one line each, under the object's own mutex: INTEGRATION.md lists the patch; SlamShims.cc
// carries it for the tests).  A hook only appends a 16-byte record to a queue; the sparsifier thread drains the queue into
// mss_mirror_add_keyframe / mss_mirror_apply right before it solves a window (Flush).  Records are absolute stores
// ("slot i of keyframe k now holds point p"), so the queue order of one object's records -- which its mutex fixes -- is all
// that matters.  Objects are named by small dense handles kept in the objects themselves (mnMirrorHandle).
#pragma once

#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/mss.h"

namespace ORB_SLAM3 {

class KeyFrame;
class MapPoint;

class MirrorRecorder {
public:
    explicit MirrorRecorder(int nSlotsPerKF) : mnSlots(nSlotsPerKF), mnNextKF(0), mnNextMP(0), mnDropped(0) {}

    // ---- hooks (any thread) ----------------------------------------------------------------------------------------------
    // a keyframe enters the map: snapshot of its slots, grid cells and the observations that already point at it
    void OnKeyFrameAdded(const std::shared_ptr<KeyFrame>& pKF);
    void OnSlot(KeyFrame* pKF, int idx, MapPoint* pMP);            // mvpMapPoints[idx] = pMP (nullptr: emptied)
    void OnObservation(KeyFrame* pKF, int idx, MapPoint* pMP);     // pMP->mObservations[pKF] = idx (nullptr: that observation is gone)
    void OnMapPoint(MapPoint* pMP, int nObs, bool bBad);           // nObs / mbBad changed
    void OnCompact(KeyFrame* pKF);                                 // EraseBadDescriptor ran
    // the calling thread's own map changes are not recorded while a Suppress object lives (the batched hand-back: the
    // device has applied the deletion to the mirror already; EraseBadDescriptor: one compaction record instead of N updates)
    struct Suppress {
        Suppress() { ++Depth(); }
        ~Suppress() { --Depth(); }
        static int& Depth() { static thread_local int d = 0; return d; }
    };

    // ---- sparsifier thread -----------------------------------------------------------------------------------------------
    int Flush(mss_mirror* m);                                      // mss_status; drains everything queued so far
    int HandleOf(MapPoint* pMP);                                   // assigns a handle at first use
    std::shared_ptr<MapPoint> PointOf(int handle);                 // nullptr when the point is gone
    int MapPointHandles() const { return mnNextMP.load(); }
    int SlotsPerKF() const { return mnSlots; }
    long Dropped() const { return mnDropped.load(); }              // records refused (keyframe with more slots than the mirror holds)
    size_t Queued();
    double LastFlushMs() const { return mLastFlushMs; }
    long LastFlushOps() const { return mLastFlushOps; }

private:
    struct KfAdd {
        int handle;
        uint32_t key;
        int n;
        std::vector<uint16_t> cells;
        std::vector<int32_t> slot, obs;
    };
    struct Rec { int32_t kind, a, b, c; };                         // kind 0 = "keyframe mvAdds[a]", else an mss_mirror_op
    void Push(int kind, int a, int b, int c);

    const int mnSlots;
    std::atomic<int> mnNextKF, mnNextMP;
    std::atomic<long> mnDropped;
    std::mutex mMutexRegister;                                     // held for a whole keyframe registration and by Flush
    std::mutex mMutexQueue;
    std::vector<Rec> mvQueue;
    std::vector<KfAdd> mvAdds;
    std::mutex mMutexPoints;
    std::vector<std::weak_ptr<MapPoint>> mvPoints;                 // handle -> object
    // nObs / mbBad change with almost every map operation: only the latest values per point travel (one op per dirty point)
    std::vector<int32_t> mvMpNobs;
    std::vector<uint8_t> mvMpBad, mvMpDirty;
    std::vector<int32_t> mvDirty;
    double mLastFlushMs = 0.0;
    long mLastFlushOps = 0;
};

}  // namespace ORB_SLAM3

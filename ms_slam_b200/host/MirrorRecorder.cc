// MirrorRecorder.cc -- see MirrorRecorder.h
#include "MirrorRecorder.h"

#include <chrono>

#ifdef MSS_WITH_ORBSLAM3_HEADERS
#include "KeyFrame.h"
#include "MapPoint.h"
#else
#include "SlamShims.h"
#endif

namespace ORB_SLAM3 {

void MirrorRecorder::Push(int kind, int a, int b, int c) {
    std::unique_lock<std::mutex> lock(mMutexQueue);
    mvQueue.push_back(Rec{kind, a, b, c});
}

size_t MirrorRecorder::Queued() {
    std::unique_lock<std::mutex> lock(mMutexQueue);
    return mvQueue.size();
}

int MirrorRecorder::HandleOf(MapPoint* pMP) {
    if (!pMP) return -1;
    int h = pMP->mnMirrorHandle.load(std::memory_order_acquire);
    if (h >= 0) return h;
    std::unique_lock<std::mutex> lock(mMutexPoints);
    h = pMP->mnMirrorHandle.load(std::memory_order_relaxed);
    if (h >= 0) return h;
    h = mnNextMP.fetch_add(1);
    if ((size_t)h >= mvPoints.size()) {
        const size_t n = (size_t)h + 1 + mvPoints.size() / 2;
        mvPoints.resize(n);
        mvMpNobs.resize(n, 0); mvMpBad.resize(n, 0); mvMpDirty.resize(n, 0);
    }
    mvPoints[h] = pMP->weak_from_this();
    pMP->mnMirrorHandle.store(h, std::memory_order_release);
    return h;
}

std::shared_ptr<MapPoint> MirrorRecorder::PointOf(int handle) {
    std::unique_lock<std::mutex> lock(mMutexPoints);
    if (handle < 0 || (size_t)handle >= mvPoints.size()) return std::shared_ptr<MapPoint>();
    return mvPoints[handle].lock();
}

void MirrorRecorder::OnKeyFrameAdded(const std::shared_ptr<KeyFrame>& pKF) {
    if (!pKF) return;
    std::unique_lock<std::mutex> reg(mMutexRegister);              // one registration at a time; Flush waits for it
    if (pKF->mnMirrorHandle >= 0) return;
    const int n = (int)pKF->GetMapPointMatches().size();
    if (n > mnSlots) { ++mnDropped; return; }                      // stays unknown to the mirror: windows with it use the flatten path
    // The snapshot takes its place in the queue BEFORE it is taken and before the keyframe's own hooks start recording:
    // whatever they record lands behind it and is replayed on top of it (stores are absolute, so replaying one that the
    // snapshot already saw changes nothing).
    size_t at;
    {
        std::unique_lock<std::mutex> lock(mMutexQueue);
        at = mvAdds.size();
        mvAdds.emplace_back();
        mvQueue.push_back(Rec{0, (int)at, 0, 0});
    }
    KfAdd add;
    add.handle = mnNextKF.fetch_add(1);
    add.key = (uint32_t)pKF->mnId;
    add.n = n;
    pKF->mnMirrorHandle = add.handle;
    add.cells.assign(n, (uint16_t)MSS_CELL_NONE);
    add.slot.assign(n, -1);
    add.obs.assign(n, -1);
    const auto& grid = pKF->GetFeatureGrids();
    for (size_t col = 0; col < grid.size(); ++col)
        for (size_t row = 0; row < grid[col].size(); ++row)
            for (size_t i : grid[col][row])
                if (i < (size_t)n) add.cells[i] = (uint16_t)(col * MSS_GRID_ROWS + row);
    const std::vector<std::shared_ptr<MapPoint>> now = pKF->GetMapPointMatches();
    for (int i = 0; i < n && i < (int)now.size(); ++i) {
        const std::shared_ptr<MapPoint>& pMP = now[i];
        if (!pMP) continue;
        const int h = HandleOf(pMP.get());
        add.slot[i] = h;
        const std::tuple<int, int> idx = pMP->GetIndexInKeyFrame(pKF);
        if (std::get<0>(idx) == i || (std::get<0>(idx) == -1 && std::get<1>(idx) == i)) add.obs[i] = h;
        {
            // a point possibly first met through this keyframe: attributes as of now (read before mMutexPoints is taken: the
            // hooks take it while they hold the point's own mutex)
            const int nObs = pMP->Observations();
            const bool bad = pMP->isBad();
            std::unique_lock<std::mutex> lock(mMutexPoints);
            if (!mvMpDirty[h]) { mvMpNobs[h] = nObs; mvMpBad[h] = bad ? 1 : 0; mvMpDirty[h] = 1; mvDirty.push_back(h); }
        }
    }
    std::unique_lock<std::mutex> lock(mMutexQueue);
    mvAdds[at] = std::move(add);
}

void MirrorRecorder::OnSlot(KeyFrame* pKF, int idx, MapPoint* pMP) {
    if (Suppress::Depth() || !pKF || pKF->mnMirrorHandle < 0) return;
    if (idx < 0 || idx >= mnSlots) { ++mnDropped; return; }
    Push(MSS_MOP_SLOT, pKF->mnMirrorHandle, idx, HandleOf(pMP));
}

void MirrorRecorder::OnObservation(KeyFrame* pKF, int idx, MapPoint* pMP) {
    if (Suppress::Depth() || !pKF || pKF->mnMirrorHandle < 0 || idx < 0) return;
    if (idx >= mnSlots) { ++mnDropped; return; }
    Push(MSS_MOP_OBS, pKF->mnMirrorHandle, idx, HandleOf(pMP));
}

void MirrorRecorder::OnMapPoint(MapPoint* pMP, int nObs, bool bBad) {
    if (Suppress::Depth() || !pMP) return;
    const int h = HandleOf(pMP);
    std::unique_lock<std::mutex> lock(mMutexPoints);
    mvMpNobs[h] = nObs;
    mvMpBad[h] = bBad ? 1 : 0;
    if (!mvMpDirty[h]) { mvMpDirty[h] = 1; mvDirty.push_back(h); }
}

void MirrorRecorder::OnCompact(KeyFrame* pKF) {
    if (!pKF || pKF->mnMirrorHandle < 0) return;
    Push(MSS_MOP_KF_COMPACT, pKF->mnMirrorHandle, 0, 0);
}

int MirrorRecorder::Flush(mss_mirror* m) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<Rec> q;
    std::vector<KfAdd> adds;
    {
        std::unique_lock<std::mutex> reg(mMutexRegister);
        std::unique_lock<std::mutex> lock(mMutexQueue);
        q.swap(mvQueue);
        adds.swap(mvAdds);
    }
    int rc = MSS_OK;
    std::vector<mss_mirror_op> ops;
    {
        std::unique_lock<std::mutex> lock(mMutexPoints);
        ops.reserve(mvDirty.size() + q.size());
        for (int32_t h : mvDirty) {
            ops.push_back(mss_mirror_op{MSS_MOP_MP, h, mvMpNobs[h], (int32_t)mvMpBad[h]});
            mvMpDirty[h] = 0;
        }
        mvDirty.clear();
    }
    const long nAttr = (long)ops.size();
    auto send = [&]() {
        if (ops.empty() || rc != MSS_OK) { ops.clear(); return; }
        rc = mss_mirror_apply(m, ops.data(), (int32_t)ops.size());
        ops.clear();
    };
    const size_t S = (size_t)mnSlots;
    for (size_t i = 0; i < q.size() && rc == MSS_OK; ++i) {
        const Rec& r = q[i];
        if (r.kind != 0) { ops.push_back(mss_mirror_op{r.kind, r.a, r.b, r.c}); continue; }
        send();                                              // records queued before the keyframe existed in the mirror
        if (rc != MSS_OK) break;
        // a run of snapshots with consecutive handles (a map loaded at once) travels as one bulk call
        size_t j = i;
        while (j + 1 < q.size() && q[j + 1].kind == 0 && adds[q[j + 1].a].handle == adds[q[j].a].handle + 1) ++j;
        const size_t n = j - i + 1;
        std::vector<uint32_t> key(n);
        std::vector<int32_t> ns(n), slot(n * S, -1), obs(n * S, -1);
        std::vector<uint16_t> cells(n * S, (uint16_t)MSS_CELL_NONE);
        for (size_t t = 0; t < n; ++t) {
            const KfAdd& a = adds[q[i + t].a];
            key[t] = a.key; ns[t] = a.n;
            std::copy(a.slot.begin(), a.slot.end(), slot.begin() + t * S);
            std::copy(a.obs.begin(), a.obs.end(), obs.begin() + t * S);
            std::copy(a.cells.begin(), a.cells.end(), cells.begin() + t * S);
        }
        rc = mss_mirror_add_keyframes(m, adds[r.a].handle, (int32_t)n, key.data(), ns.data(), cells.data(), slot.data(), obs.data());
        i = j;
    }
    send();
    mLastFlushOps = (long)q.size() + nAttr;
    mLastFlushMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

}  // namespace ORB_SLAM3

"""Persistent device mirror (include/mss.h mss_mirror_*).

CPU part: the numpy model of the mirror (oracle/mirror_model.py) is pinned against the host C++ FlattenWindow -- both must
flatten the same map to the same arrays.  GPU part: the device-assembled view equals the model's view array for array, a
mirror solve equals the flatten-path solve bit for bit, and the mirror follows deltas (observations erased, keyframes
added / compacted, the deletion of a window applied on the device) exactly like the model.
"""
import numpy as np
import pytest

from conftest import LAM, GLAM
from ms_slam_b200 import msgen, make_view
from ms_slam_b200.window import pack_view, CELL_NONE
from oracle import emulate as em, mirror_model as mm

CASES = [("c1", 0, {}), ("live", 3, {}), ("live", 0, dict(M=1500, H=20)), ("c4", 0, dict(M=3000)), ("c3", 1, dict(K=40, M=9000))]


def packed_from_model(view):
    """the MSS_LAYOUT_PACKED arrays the device mirror emits for a model view (slots in slot order)"""
    cell = np.where(view.feat_cell == CELL_NONE, 0xFFF, view.feat_cell).astype(np.uint32)
    slots = (view.feat_mp.astype(np.uint32) << 12) | cell
    owner = np.repeat(np.arange(view.M, dtype=np.int64), np.diff(view.mp_obs_ptr))
    pairs = ((owner << 12) | (view.mp_obs_kf.astype(np.int64) - view.K)).astype(np.uint32)
    return slots, pairs


@pytest.mark.parametrize("name,seed,over", CASES)
def test_model_matches_host_flatten(build_native, name, seed, over):
    """oracle pin: MirrorModel.build == the C++ FlattenWindow on the same map (works without a GPU)"""
    from ms_slam_b200.host_mirror import World
    view, N = msgen.make_config(name, seed, **over)
    w = World(view, N=N)
    snap, mp_ids, okf_ids, is_var = w.flatten_only()
    w.close()
    L = mm.load_view(view, seed=5)
    model = mm.MirrorModel(L["S"])
    model.add_keyframes(0, None, L["n_slots"], L["cells"], L["slot_mp"], L["obs_mp"])
    model.set_map_points(0, L["nobs"])
    mv, mp_handle, okf_handle = model.build(L["window"])
    assert (mv.K, mv.H, mv.M, mv.F) == (snap.K, snap.H, snap.M, snap.F)
    assert np.array_equal(mv.feat_ptr, snap.feat_ptr) and np.array_equal(mv.feat_mp, snap.feat_mp)
    assert np.array_equal(mv.feat_cell, snap.feat_cell) and np.array_equal(mv.mp_nobs, snap.mp_nobs)
    assert np.array_equal(mv.okf_total, snap.okf_total) and np.array_equal(okf_handle, okf_ids)
    assert np.array_equal(L["mp_of_table"][mp_ids], mp_handle)          # same discovery order, through the handle permutation
    # FlattenWindow lists outside observations only: same CSR up to the order inside a map point
    assert np.array_equal(mv.mp_obs_ptr, snap.mp_obs_ptr)
    for a, b in ((mv, snap),):
        ka = np.lexsort((a.mp_obs_kf, np.repeat(np.arange(a.M), np.diff(a.mp_obs_ptr))))
        kb = np.lexsort((b.mp_obs_kf, np.repeat(np.arange(b.M), np.diff(b.mp_obs_ptr))))
        assert np.array_equal(a.mp_obs_kf[ka], b.mp_obs_kf[kb])


def test_model_ops_follow_the_reference_semantics():
    """SetBadFlag / EraseBadDescriptor restated on the arrays (src/MapPoint.cc:227-255, src/KeyFrame.cc:311-361)"""
    model = mm.MirrorModel(6)
    model.apply([(mm.MOP_SLOT, 0, i, h) for i, h in enumerate([4, -1, 2, 7])] + [(mm.MOP_OBS, 0, 0, 4), (mm.MOP_OBS, 0, 2, 2)] +
                [(mm.MOP_SLOT, 1, 0, 2), (mm.MOP_OBS, 1, 0, 2), (mm.MOP_MP, 2, 5, 0), (mm.MOP_MP, 4, 3, 0), (mm.MOP_MP, 7, 9, 0)])
    assert model.kf_n.tolist() == [4, 1]
    model.delete([2])
    assert model.slot_mp[0, :4].tolist() == [4, -1, -1, 7] and model.slot_mp[1, 0] == -1 and model.mp_bad[2]
    model.apply([(mm.MOP_KF_COMPACT, 0, 0, 0)])
    assert model.kf_n[0] == 2 and model.slot_mp[0, :3].tolist() == [4, 7, -1]
    assert model.obs_mp[0, :3].tolist() == [4, 7, -1]                   # every kept point observes the keyframe at its new index
    v, mp_handle, okf = model.build([0])
    assert v.M == 2 and mp_handle.tolist() == [4, 7] and v.feat_cell.tolist() == [CELL_NONE, CELL_NONE]


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng(build_native):
    from ms_slam_b200.engine import Engine
    e = Engine(N=100, lam=LAM, grid_lam=GLAM, device=0)
    yield e
    e.close()


def load_both(eng, L):
    from ms_slam_b200.mirror import Mirror
    mir = Mirror(eng, L["S"])
    mir.load(L)
    model = mm.MirrorModel(L["S"])
    model.add_keyframes(L["kf0"], None, L["n_slots"], L["cells"], L["slot_mp"], L["obs_mp"])
    model.set_map_points(L["mp0"], L["nobs"])
    return mir, model


def assert_same_view(mir, model, kfs):
    pv, mp_handle, okf_handle = mir.build_view(kfs)
    mv, m_handle, m_okf = model.build(kfs)
    assert (pv.K, pv.H, pv.M, pv.F, pv.O) == (mv.K, mv.H, mv.M, mv.F, mv.O)
    slots, pairs = packed_from_model(mv)
    assert np.array_equal(pv.feat_ptr, mv.feat_ptr) and np.array_equal(pv.slots, slots)
    assert np.array_equal(pv.mp_nobs16, mv.mp_nobs.astype(np.uint16)) and np.array_equal(pv.okf_total, mv.okf_total)
    assert np.array_equal(mp_handle, m_handle) and np.array_equal(okf_handle, m_okf)
    assert np.array_equal(np.sort(pv.obs_pairs), np.sort(pairs))        # the pair list is unordered
    return mv, m_handle


def check_solve(eng, mir, model, kfs, N, apply=False):
    mv, m_handle = model.build(kfs)[:2]
    eng.set_params(N, LAM, GLAM)
    ref = em.solve(mv, N, LAM, GLAM)
    flat = eng.solve(pack_view(mv))                                     # the flatten path on the same window
    r = mir.solve([kfs], apply=apply)[0]
    for got in (flat, r.result):
        assert np.array_equal(got.keep, ref["keep"]) and np.array_equal(got.kf_cov, ref["cov"]) and np.array_equal(got.kf_slack, ref["slack"])
        assert (got.objective, got.n_kept, got.n_vars, got.rounds, got.n_max) == (ref["objective"], ref["n_kept"], ref["n_vars"], ref["rounds"], ref["n_max"])
    assert np.array_equal(r.mp_handle, m_handle)
    want = np.sort(m_handle[~ref["keep"]])
    assert np.array_equal(r.deleted, want) and r.n_deleted == want.size
    return r, want


@pytest.mark.gpu
@pytest.mark.parametrize("name,seed,over", CASES)
def test_device_view_and_solve_match_the_flatten_path(eng, name, seed, over):
    view, N = msgen.make_config(name, seed, **over)
    L = mm.load_view(view, seed=seed + 1)
    mir, model = load_both(eng, L)
    assert_same_view(mir, model, L["window"])
    check_solve(eng, mir, model, L["window"], N)
    assert_same_view(mir, model, L["window"])                            # the solve left the mirror (and its scratch) untouched
    # a sub-window: the other window keyframes become outside keyframes
    sub = L["window"][3:3 + max(2, view.K // 2)]
    assert_same_view(mir, model, sub)
    check_solve(eng, mir, model, sub, N)
    mir.close()


@pytest.mark.gpu
def test_mirror_follows_deltas_and_applies_the_deletion(eng):
    """window 1 solved with apply=1 (the device performs SetBadFlag on its own copy), then deltas as the map moves on
    (observations erased, points re-observed, nObs changes, a sparsified keyframe compacted, a new keyframe), then a second
    window over later keyframes: view and result must equal the model's at every step."""
    view, N = msgen.make_config("live", 7, K=40, M=7000, H=12)
    L = mm.load_view(view, seed=3)
    mir, model = load_both(eng, L)
    rng = np.random.default_rng(0)
    w1 = L["window"][:20]
    r, deleted = check_solve(eng, mir, model, w1, N, apply=True)
    model.delete(deleted)
    assert_same_view(mir, model, w1)                                     # rebuilt after the deletion: fewer points, same arrays
    ops = []
    for _ in range(300):                                                 # EraseObservation + EraseMapPointMatch on random slots
        kf = int(rng.integers(0, view.K + view.H)); i = int(rng.integers(0, max(1, model.kf_n[kf])))
        h = int(model.slot_mp[kf, i])
        if h < 0:
            continue
        ops += [(mm.MOP_OBS, kf, i, -1), (mm.MOP_SLOT, kf, i, -1), (mm.MOP_MP, h, max(0, int(model.mp_nobs[h]) - 2), 0)]
    for kf in w1[:6]:                                                    # LoopClosing::DeleteOutdatedInfo -> EraseBadDescriptor
        ops.append((mm.MOP_KF_COMPACT, int(kf), 0, 0))
    new_kf = view.K + view.H                                             # a new keyframe that re-observes 200 surviving points
    alive = np.nonzero(~model.mp_bad[:view.M])[0]
    pick = rng.choice(alive, size=200, replace=False)
    for i, h in enumerate(pick):
        ops += [(mm.MOP_SLOT, new_kf, i, int(h)), (mm.MOP_OBS, new_kf, i, int(h)), (mm.MOP_MP, int(h), int(model.mp_nobs[h]) + 2, 0)]
    ops += [(mm.MOP_SLOT, 2, 1, int(pick[0])), (mm.MOP_SLOT, 2, 1, -1)]  # two ops on one address: the later one wins
    mir.apply(ops)
    model.apply(ops)
    w2 = L["window"][14:40]
    assert_same_view(mir, model, w2)
    r2, deleted2 = check_solve(eng, mir, model, w2, N, apply=True)
    model.delete(deleted2)
    assert_same_view(mir, model, w2)
    assert_same_view(mir, model, np.concatenate([w2, [new_kf]]).astype(np.int32))
    mir.close()


@pytest.mark.gpu
def test_batch_of_independent_windows_and_rejection_of_dependent_ones(eng):
    from ms_slam_b200.mirror import Mirror
    from ms_slam_b200.engine import MSS_E_BADARG
    specs = [("live", 21, {}), ("c1", 1, {}), ("live", 22, dict(M=1500, H=20)), ("c4", 1003, dict(M=4000, K=40)), ("live", 23, {})]
    N = 100
    views = [msgen.make_config(n, s, **o)[0] for n, s, o in specs]
    S = 2000
    mir, model = Mirror(eng, S), mm.MirrorModel(S)
    wins, kf0, mp0 = [], 0, 0
    for v in views:                                                       # disjoint handle ranges: independent components
        L = mm.load_view(v, S=S, seed=kf0, kf0=kf0, mp0=mp0)
        L["slot_mp"] = np.where(L["slot_mp"] >= 0, L["slot_mp"], -1)
        mir.load(L)
        model.add_keyframes(kf0, None, L["n_slots"], L["cells"], L["slot_mp"], L["obs_mp"])
        model.set_map_points(mp0, L["nobs"])
        wins.append(L["window"])
        kf0 += v.K + v.H
        mp0 += L["n_mp"]
    eng.set_params(N, LAM, GLAM)
    l0 = eng.stats()["kernel_launches"]
    res = mir.solve(wins)
    launches = eng.stats()["kernel_launches"] - l0
    for kfs, r in zip(wins, res):
        mv, mh = model.build(kfs)[:2]
        ref = em.solve(mv, N, LAM, GLAM)
        assert np.array_equal(r.result.keep, ref["keep"]) and r.result.objective == ref["objective"] and r.result.rounds == ref["rounds"]
        assert np.array_equal(r.deleted, np.sort(mh[~ref["keep"]]))
    assert launches <= 20                                                 # a fixed number of launches, not one set per window
    # two windows that share map points are not independent: rejected, nothing deleted, and the mirror stays usable
    bad = mir.solve([wins[0][:10], wins[0][8:20]], raise_on_status=False)
    assert any(b.result.status == MSS_E_BADARG for b in bad) and all(b.deleted.size == 0 for b in bad if b.result.status == MSS_E_BADARG)
    again = mir.solve([wins[0]])[0]
    assert np.array_equal(again.deleted, res[0].deleted)
    mir.close()

"""The C++ MapSparsification mirror (ms_slam_b200/host/): drop-in boundary of SURVEY.md section 8b.

CPU tests: the pointer-graph -> flat-view walk (FlattenWindow = passes 1-3 of /root/reference/src/MapSparsification.cc:66-151)
reproduces the view the graph was built from; settings reader; non-local counter; fail-safe behaviour without a GPU.
GPU tests: the sparsifier thread driven like System / LocalMapping / LoopClosing drive it upstream deletes exactly the map
points the engine's bitmask says, forwards every keyframe, honours the stop / finish handshakes."""
import os
import tempfile

import numpy as np
import pytest

from conftest import LAM, GLAM
from ms_slam_b200 import make_view, msgen, CELL_NONE


@pytest.fixture(scope="module")
def hm(build_native):
    from ms_slam_b200 import host_mirror
    host_mirror.load_library()
    return host_mirror


def check_roundtrip(view, flat, mp_ids, okf_ids, is_var):
    """flat view == original view up to the renumbering of the map-point table / outside keyframes"""
    K = view.K
    comp = view.compact()                  # FlattenWindow emits valid slots only and outside observations only
    assert flat.K == K and np.array_equal(flat.feat_ptr, comp.feat_ptr)
    valid = view.feat_mp >= 0
    assert (flat.feat_mp >= 0).all()
    assert np.array_equal(mp_ids[flat.feat_mp], comp.feat_mp)
    assert np.array_equal(flat.feat_cell, comp.feat_cell)
    assert np.unique(mp_ids).size == mp_ids.size                       # each map point has exactly one table entry
    assert set(mp_ids.tolist()) == set(view.feat_mp[valid].tolist())   # and the table holds exactly the points seen in valid slots
    assert np.array_equal(flat.mp_nobs, view.mp_nobs[mp_ids])
    # discovery order: keyframes in window order, slots in index order
    first = {}
    for s in np.nonzero(valid)[0]:
        first.setdefault(int(view.feat_mp[s]), len(first))
    assert [first[int(p)] for p in mp_ids] == list(range(mp_ids.size))
    grid = valid & (view.feat_cell != CELL_NONE)
    var_orig = np.zeros(view.M, bool)
    var_orig[view.feat_mp[grid]] = True
    assert np.array_equal(is_var, var_orig[mp_ids])
    # observations: complete for variables, omitted for the rest (the engine only reads those of variables)
    cnt_all = np.zeros(view.H, np.int64)
    for p in range(view.M):
        for kf in view.mp_obs_kf[view.mp_obs_ptr[p]:view.mp_obs_ptr[p + 1]]:
            if kf >= K:
                cnt_all[kf - K] += 1
    seen_outside = set()
    for q, p in enumerate(mp_ids):
        got = flat.mp_obs_kf[flat.mp_obs_ptr[q]:flat.mp_obs_ptr[q + 1]]
        if not is_var[q]:
            assert got.size == 0
            continue
        assert (got >= K).all()
        got = sorted(int(okf_ids[k - K]) for k in got)                           # outside keyframe ids are K + original j
        exp = sorted(int(k) for k in comp.mp_obs_kf[comp.mp_obs_ptr[p]:comp.mp_obs_ptr[p + 1]])
        assert got == exp
        seen_outside.update(k for k in got if k >= K)
    assert sorted(okf_ids.tolist()) == sorted(seen_outside) and list(okf_ids) == sorted(okf_ids)    # ordered by keyframe id
    for j, kid in enumerate(okf_ids):
        jj = int(kid) - K
        assert flat.okf_total[j] == max(int(view.okf_total[jj]), int(cnt_all[jj]))              # GetNumberMPs() of that keyframe


@pytest.mark.parametrize("name,seed,over", [("c1", 0, {}), ("live", 0, {}), ("live", 3, dict(M=1500, H=20)), ("c4", 1000, dict(M=3000))])
def test_flatten_roundtrip(hm, name, seed, over):
    view, N = msgen.make_config(name, seed, **over)
    w = hm.World(view, N=N)
    try:
        flat, mp_ids, okf_ids, is_var = w.flatten_only()
        flat.validate()
        check_roundtrip(view, flat, mp_ids, okf_ids, is_var)
    finally:
        w.close()


def decode_tokens(tok_ptr, tokens):
    """MSS_LAYOUT_PACKED16 tokens -> (kf, mp, cell) of every slot, the way the device decodes them (include/mss.h)"""
    out = []
    for k in range(tok_ptr.size - 1):
        mp = 0
        for t in tokens[tok_ptr[k]:tok_ptr[k + 1]].tolist():
            d, low = t >> 12, t & 0xFFF
            if d < 15:
                mp += d
                out.append((k, mp, 0xFFFF if low == 0xFFF else low))
            else:
                mp += 15 * (low + 1)
    return out


@pytest.mark.parametrize("name,seed,over", [("c1", 0, {}), ("live", 3, dict(M=1500, H=20)), ("c3", 0, {})])
def test_flatten_packed_blob_matches_the_python_packer(hm, name, seed, over):
    """The blob FlattenWindow ships (u16 tokens, u16 nObs, outside pair list) decodes to exactly the snapshot's slots, and is
    token for token what ms_slam_b200.window.pack_view(tokens16=True) produces for the same snapshot."""
    from ms_slam_b200 import pack_view
    view, N = msgen.make_config(name, seed, **over)
    w = hm.World(view, N=N)
    try:
        flat, mp_ids, okf_ids, is_var = w.flatten_only()
        blob = w.snapshot_packed(0)
        assert blob is not None
        ref = pack_view(flat, tokens16=True)
        assert np.array_equal(blob["tok_ptr"], ref.feat_ptr) and np.array_equal(blob["tokens"], ref.slots)
        assert blob["nobs8"] == bool(ref.meta["nobs8"]) == (int(flat.mp_nobs.max()) <= 255)      # one byte per map point when it fits
        assert ref.mp_nobs16.dtype == (np.uint8 if blob["nobs8"] else np.uint16)
        assert np.array_equal(blob["nobs16"], ref.mp_nobs16) and np.array_equal(np.sort(blob["pairs"]), np.sort(ref.obs_pairs))
        kf = np.repeat(np.arange(flat.K), np.diff(flat.feat_ptr))
        want = sorted(zip(kf.tolist(), flat.feat_mp.tolist(), flat.feat_cell.tolist()))
        assert sorted(decode_tokens(blob["tok_ptr"], blob["tokens"])) == want
        assert blob["blob_bytes"] < 0.40 * flat.input_bytes()
    finally:
        w.close()


def test_token_coding_escapes():
    """gaps of 15 and more use escape tokens (15 * (low + 1) each, several when the gap exceeds 61440)"""
    from ms_slam_b200 import pack_view
    M = 400000
    slots = [[(0, 5), (14, 6), (15, 7), (29, None), (30, 9), (100000, 10), (399999, 11)], [(61439, 1), (61440, 2), (122881, 3)], []]
    v = make_view(3, slots, [5] * M)
    pv = pack_view(v, tokens16=True)
    got = decode_tokens(pv.feat_ptr, pv.slots)
    want = sorted((k, p, 0xFFFF if c is None else c) for k, row in enumerate(slots) for p, c in row)
    assert sorted(got) == want
    assert pv.feat_ptr[-1] == pv.slots.size and pv.feat_ptr[3] == pv.feat_ptr[2]


def test_flatten_quirks(hm):
    # empty / bad slots, keypoints outside the grid, a point in two slots of one keyframe, an outside keyframe
    view = make_view(2, [[(0, 0), (None, 3), (1, None), (2, 7), (2, 9)], [(2, 100), (3, 3071), (None, None)]],
                     [4, 6, 8, 5], outside=[[2, 3], [1]], okf_total=[5, 2])
    w = hm.World(view, N=2)
    try:
        flat, mp_ids, okf_ids, is_var = w.flatten_only()
        check_roundtrip(view, flat, mp_ids, okf_ids, is_var)
        assert mp_ids.tolist() == [0, 1, 2, 3] and is_var.tolist() == [True, False, True, True]
        assert okf_ids.tolist() == [2]                 # the second outside keyframe is only seen by a non-variable: no row
    finally:
        w.close()


def test_settings_reader_and_nonlocal_counter(hm):
    view, N = msgen.make_config("c1", 0)
    w = hm.World(view, N=75, lam=437.5, grid_lam=9.5, window_length=30, non_local=15)
    try:
        assert w.nonlocal_after(0) == 15              # KeyFrame::UpdateCountInLocalMapping (src/KeyFrame.cc:980-997)
        assert w.nonlocal_after(1) == 15
    finally:
        w.close()


def test_without_gpu_keeps_every_point(hm):
    """No CUDA device -> no engine -> fail-safe: nothing is deleted, keyframes are still forwarded (SURVEY 8b error row)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    view, N = msgen.make_config("c1", 1)
    w = hm.World(view, N=N)
    try:
        assert not w.engine_ready()
        w.start()
        w.feed(0, view.K)
        assert w.wait_forwarded(view.K) == 0
        assert w.forwarded_ids() == list(range(view.K))
        assert not w.bad_flags().any()
        assert w.reports()[0]["status"] < 0 and w.reports()[0]["n_deleted"] == 0
        w.consume()
        assert w.finish() == 0
    finally:
        w.close()


# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,seed,over", [("c1", 0, {}), ("live", 0, {}), ("live", 5, dict(M=1500, H=20))])
def test_thread_flow_matches_engine(hm, name, seed, over):
    from ms_slam_b200.engine import Engine
    from oracle import ilp_model as om
    view, N = msgen.make_config(name, seed, **over)
    w = hm.World(view, N=N, mirror=False)
    eng = Engine(N=N, lam=LAM, grid_lam=GLAM)
    try:
        assert w.engine_ready() and not w.mirror_active()
        w.start()
        w.feed(0, 10)                                   # trigger is "more than 10 queued" (MapSparsification.cc:197)
        assert w.wait_forwarded(1, timeout_ms=200) == -1 and w.forwarded_ids() == []
        w.feed(10, view.K - 10)
        assert w.wait_forwarded(view.K) == 0
        assert w.forwarded_ids() == list(range(view.K))                  # every window keyframe, in window order (:168-170)
        flat, mp_ids, okf_ids, is_var = w.snapshot(1)
        check_roundtrip(view, flat, mp_ids, okf_ids, is_var)
        ref = eng.solve(flat)                                            # the engine called directly on the same snapshot
        bad = w.bad_flags()
        exp_bad = np.zeros(view.M, bool)
        exp_bad[mp_ids[~ref.keep]] = True
        assert np.array_equal(bad, exp_bad)                              # SetBadFlag exactly where the bit is 0 (:159-166)
        rep = w.reports()[0]
        assert rep["status"] == 0 and rep["objective"] == ref.objective and rep["n_deleted"] == int((~ref.keep).sum())
        # quality on the ORIGINAL view: rows satisfied, within 1 % of the LP bound of the reference ILP
        model = om.build_model(view, N)
        x = om.keep_to_x(model, ~bad)
        assert om.rows_satisfied(model, x, N)[0]
        assert om.objective(model, x, N, LAM, GLAM) <= 1.01 * om.solve_lp(view, N, LAM, GLAM, model=model).objective
        # deleted points are gone from their keyframes and from the map (MapPoint.cc:227-255)
        st = w.keyframe_state()
        kept_slots = np.array([int((~bad[view.feat_mp[view.feat_ptr[k]:view.feat_ptr[k + 1]][view.feat_mp[view.feat_ptr[k]:view.feat_ptr[k + 1]] >= 0]]).sum())
                               for k in range(view.K)])
        assert np.array_equal(st[:view.K, 0], kept_slots)
        # stop handshake of LoopClosing::CorrectLoop, consumer step, shutdown flush
        assert w.stop_handshake() == 1
        w.consume()
        st = w.keyframe_state()
        assert st[:view.K, 1].all() and (st[:view.K, 2] == 1).all()      # EraseBadDescriptor ran once per keyframe, mbSparsified
        assert w.map_counts()["sparsified_keyframes"] == view.K
        assert w.finish() == 0
        assert len(w.reports()) == 2 and w.reports()[1]["K"] == 0        # nothing left for the final flush
        flat_outcome = (bad.copy(), rep["objective"], rep["n_deleted"], rep["n_kept"], rep["rounds"], w.keyframe_state().copy(),
                        w.map_counts(), w.forwarded_ids())
    finally:
        eng.close()
        w.close()
    # ---- the same flow from the persistent device mirror: K keyframe handles up, a bitmask over map-point handles down; the
    #      map must end up in exactly the same state (bit-identical to the flatten path)
    w = hm.World(view, N=N, mirror=True)
    try:
        assert w.mirror_active()
        w.start()
        w.feed(0, 10)
        assert w.wait_forwarded(1, timeout_ms=200) == -1
        w.feed(10, view.K - 10)
        assert w.wait_forwarded(view.K) == 0
        rep = w.reports()[0]
        assert rep["mirror"] == 1 and rep["status"] == 0
        assert np.array_equal(w.bad_flags(), flat_outcome[0])
        assert (rep["objective"], rep["n_deleted"], rep["n_kept"], rep["rounds"]) == flat_outcome[1:5]
        assert rep["h2d_bytes"] < 64 * 1024 + 8 * view.K                 # handles and descriptors, not a flattened view
        assert w.stop_handshake() == 1
        w.consume()
        assert w.finish() == 0
        assert np.array_equal(w.keyframe_state(), flat_outcome[5]) and w.map_counts() == flat_outcome[6]
        assert w.forwarded_ids() == flat_outcome[7]
    finally:
        w.close()


@pytest.mark.gpu
def test_final_flush_takes_all_unsparsified_keyframes(hm):
    """RequestFinish with keyframes still queued: one window over every keyframe with !mbSparsified, then EraseBadDescriptor
    (MapSparsification.cc:38-52); the outside keyframes of the fixture are already sparsified and stay outside."""
    from ms_slam_b200.engine import Engine
    view, N = msgen.make_config("live", 2)
    w = hm.World(view, N=N, window_length=8, mirror=False)
    eng = Engine(N=N, lam=LAM, grid_lam=GLAM)
    try:
        w.start()
        w.feed(0, 5)                                    # below the trigger: nothing happens until shutdown
        assert w.finish() == 0
        reps = w.reports()
        assert len(reps) == 1 and reps[0]["K"] == view.K and reps[0]["H"] > 0
        flat, mp_ids, okf_ids, is_var = w.snapshot(1)
        ref = eng.solve(flat)
        exp_bad = np.zeros(view.M, bool)
        exp_bad[mp_ids[~ref.keep]] = True
        assert np.array_equal(w.bad_flags(), exp_bad)
        st = w.keyframe_state()
        assert st[:view.K, 1].all() and (st[:view.K, 2] == 1).all()
        assert w.forwarded_ids() == list(range(view.K))
        w.consume()                                     # quirk A.5.5: the same keyframes reach LoopClosing too -> second call
        assert (w.keyframe_state()[:view.K, 2] == 2).all()
        st_flat = w.keyframe_state().copy()
    finally:
        eng.close()
        w.close()
    w = hm.World(view, N=N, window_length=8, mirror=True)     # the same shutdown from the device mirror
    try:
        w.start()
        w.feed(0, 5)
        assert w.finish() == 0
        reps = w.reports()
        assert len(reps) == 1 and reps[0]["mirror"] == 1 and reps[0]["K"] == view.K and reps[0]["objective"] == ref.objective
        assert np.array_equal(w.bad_flags(), exp_bad)
        w.consume()
        assert np.array_equal(w.keyframe_state(), st_flat)
    finally:
        w.close()


@pytest.mark.gpu
def test_final_flush_splits_into_components(hm):
    """A flush over an atlas with three disjoint covisibility components: the C++ class finds the components on the device
    (mss_components), solves them as ONE batch of independent windows carrying the window-wide nMax, and deletes exactly
    what the engine selects for each component; the union satisfies every row of the undivided model and its objective is
    the sum of the parts (the flush model of MapSparsification.cc:38-47 is block diagonal)."""
    from ms_slam_b200 import merge_views, split_components, pack_view
    from ms_slam_b200.engine import Engine
    from oracle import ilp_model as om, components as oc
    N = 100
    parts = [msgen.make_config("live", 41)[0], msgen.make_config("c1", 6)[0], msgen.make_config("live", 42, M=1500, H=20)[0]]
    whole = merge_views(parts, interleave=True)
    w = hm.World(whole, N=N, window_length=8, mirror=False)
    eng = Engine(N=N, lam=LAM, grid_lam=GLAM)
    try:
        w.start()
        w.feed(0, 5)
        assert w.finish() == 0
        reps = w.reports()
        assert len(reps) == 1 and reps[0]["K"] == whole.K and reps[0]["status"] == 0
        flat, mp_ids, okf_ids, is_var = w.snapshot(1)
        rows, mps, nc, nmax = oc.components(flat)
        split = split_components(flat, rows, mps, nmax)
        assert reps[0]["components"] == len(split) >= 3
        res = eng.solve_batch([pack_view(p.compact()) for p, _, _ in split])
        exp_bad = np.zeros(whole.M, bool)
        for (p, _, mp_idx), r in zip(split, res):
            exp_bad[mp_ids[mp_idx[~r.keep]]] = True
        assert np.array_equal(w.bad_flags(), exp_bad)
        assert reps[0]["objective"] == sum(r.objective for r in res)
        model = om.build_model(whole, N)
        x = om.keep_to_x(model, ~exp_bad)
        assert om.rows_satisfied(model, x, N)[0]
        assert om.objective(model, x, N, LAM, GLAM) == reps[0]["objective"]
        assert w.forwarded_ids() == list(range(whole.K))
    finally:
        eng.close()
        w.close()
    # from the device mirror: components found on the device view (mss_mirror_components), each solved from its keyframe handles
    w = hm.World(whole, N=N, window_length=8, mirror=True)
    try:
        w.start()
        w.feed(0, 5)
        assert w.finish() == 0
        r2 = w.reports()
        assert len(r2) == 1 and r2[0]["mirror"] == 1 and r2[0]["status"] == 0 and r2[0]["components"] == reps[0]["components"]
        assert np.array_equal(w.bad_flags(), exp_bad) and r2[0]["objective"] == reps[0]["objective"]
    finally:
        w.close()


@pytest.mark.gpu
def test_c2_sized_flush_from_the_mirror_equals_the_flatten_path(hm):
    """The north-star window (500 KF x 200k MP) through the C++ class as a final flush, once flattened from the pointer graph
    and once from the persistent device mirror: the same map points are deleted; the mirror path ships a few KB."""
    view, N = msgen.make_config("c2", 3)
    out = {}
    for mirror in (False, True):
        w = hm.World(view, N=N, mirror=mirror)
        try:
            w.start()
            assert w.finish(timeout_ms=120000) == 0
            rep = [r for r in w.reports() if r["K"] > 0][0]
            out[mirror] = (w.bad_flags(), rep)
        finally:
            w.close()
    assert out[True][1]["mirror"] == 1 and out[False][1]["mirror"] == 0
    assert np.array_equal(out[True][0], out[False][0])
    for k in ("objective", "n_deleted", "n_kept", "n_vars", "rounds"):
        assert out[True][1][k] == out[False][1][k], k
    assert out[True][1]["h2d_bytes"] < 16384 and out[True][1]["d2h_bytes"] < 65536


@pytest.mark.gpu
@pytest.mark.parametrize("mirror", [False, True])
def test_windows_carry_a_certificate(hm, mirror):
    """MSS_DUAL_BOUND=1: every window the class solves reports the lower bound the device proved for it (the counterpart of
    GUROBI's MIPGap, MapSparsification.cc:155-156) -- never above the LP optimum of the reference model, within 1 % of F."""
    from oracle import ilp_model as om
    view, N = msgen.make_config("live", 2)
    w = hm.World(view, N=N, mirror=mirror, dual_bound=True)
    try:
        w.start()
        w.feed(0, view.K)
        assert w.wait_forwarded(view.K) == 0
        rep = w.reports()[0]
        lp = om.solve_lp(view, N, LAM, GLAM).objective
        assert rep["status"] == 0 and np.isfinite(rep["dual_bound"])
        assert rep["dual_bound"] <= lp + 1e-6 and rep["objective"] <= 1.01 * rep["dual_bound"]
        assert w.finish() == 0
    finally:
        w.close()
    w = hm.World(view, N=N, mirror=mirror)                   # off by default: no bound, same selection
    try:
        w.start()
        w.feed(0, view.K)
        assert w.wait_forwarded(view.K) == 0
        rep2 = w.reports()[0]
        assert np.isnan(rep2["dual_bound"]) and rep2["objective"] == rep["objective"] and rep2["n_deleted"] == rep["n_deleted"]
        assert w.finish() == 0
    finally:
        w.close()

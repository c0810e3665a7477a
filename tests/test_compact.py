"""Keyframe compaction on the device (include/mss.h mss_compact_keyframes / mss_mirror_compact_keyframes).

Oracle: oracle/mirror_model.py erase_bad_descriptor_rows = KeyFrame::EraseBadDescriptor (/root/reference/src/KeyFrame.cc:311-361)
restated for mDescriptors / mvKeysUn / mvuRight / mvDepth.  Pure data movement: bit-exact.
"""
import numpy as np
import pytest

from conftest import LAM, GLAM
from ms_slam_b200 import msgen
from oracle import mirror_model as mm


def test_oracle_keeps_surviving_rows_in_order():
    keep = np.array([1, 0, 1, 1, 0, 0, 1], bool)
    desc = np.arange(7 * 32, dtype=np.uint8).reshape(7, 32)
    keys = np.arange(7 * 7, dtype=np.uint32).reshape(7, 7)
    ur = np.arange(7, dtype=np.float32)
    d, k, u, z = mm.erase_bad_descriptor_rows(keep, desc, keys, ur, ur * 2)
    assert d.shape == (4, 32) and d[:, 0].tolist() == [0, 64, 96, 192] and k[:, 0].tolist() == [0, 14, 21, 42]
    assert u.tolist() == [0, 2, 3, 6] and z.tolist() == [0, 4, 6, 12]
    assert mm.erase_bad_descriptor_rows(np.zeros(0, bool), desc[:0])[0].shape == (0, 32)


@pytest.fixture(scope="module")
def eng(build_native):
    from ms_slam_b200.engine import Engine
    e = Engine(N=100, lam=LAM, grid_lam=GLAM, device=0)
    yield e
    e.close()


def random_payload(rng, n):
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    keys = rng.integers(0, 2**32, size=(n, 7), dtype=np.uint64).astype(np.uint32)   # any bit pattern: the kernel moves words
    return desc, keys, rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32) * 50


@pytest.mark.gpu
@pytest.mark.parametrize("tma", ["1", "0"])
def test_compaction_matches_the_reference_semantics(eng, tma, monkeypatch):
    """ragged keyframes (0, 1, 255, 256, 257, 2000, 5000 rows), all / none / random survivors, missing arrays"""
    from ms_slam_b200.mirror import KeyframePayload, compact_keyframes
    monkeypatch.setenv("MSS_COMPACT_TMA", tma)           # TMA-staged chunks (default) / register-staged kernel
    rng = np.random.default_rng(0)
    cases = []
    for n, frac in [(0, 0.5), (1, 1.0), (1, 0.0), (255, 0.3), (256, 0.5), (257, 0.9), (2000, 0.15), (2000, 1.0), (2000, 0.0), (5000, 0.5), (777, 0.5)]:
        keep = rng.random(n) < frac
        cases.append((keep,) + random_payload(rng, n))
    pay = []
    for i, (keep, desc, keys, ur, dp) in enumerate(cases):
        if i == len(cases) - 1:
            pay.append(KeyframePayload(eng, keep.size, keep, descriptors=desc, depth=dp))          # any array may be missing
        else:
            pay.append(KeyframePayload(eng, keep.size, keep, desc, keys, ur, dp))
    l0 = eng.stats()["kernel_launches"]
    n_out = compact_keyframes(eng, pay)
    assert eng.stats()["kernel_launches"] - l0 == 1                       # one launch for the whole batch of keyframes
    for (keep, desc, keys, ur, dp), p, n in zip(cases, pay, n_out):
        assert n == int(keep.sum())
        want = mm.erase_bad_descriptor_rows(keep, desc, keys, ur, dp)
        got = p.fetch(int(n))
        for name, w in zip(("descriptors", "keypoints", "uright", "depth"), want):
            if got[name] is not None:
                assert np.array_equal(got[name].view(np.uint8), np.ascontiguousarray(w).view(np.uint8).reshape(got[name].view(np.uint8).shape)), name
        p.free()
    assert compact_keyframes(eng, []).size == 0


@pytest.mark.gpu
def test_mirror_drives_the_compaction_after_a_window(eng):
    """A window is solved on the mirror with apply=1 (the device empties the slots of the deleted points), then the window's
    keyframes are compacted with the flags taken from the mirror: payload rows == the rows of the surviving slots, and the
    mirror's own rows are compacted as MSS_MOP_KF_COMPACT would (model.compact)."""
    from ms_slam_b200.mirror import Mirror, KeyframePayload, compact_keyframes
    view, N = msgen.make_config("live", 5, M=4000, H=10)
    L = mm.load_view(view, seed=2)
    mir = Mirror(eng, L["S"])
    mir.load(L)
    model = mm.MirrorModel(L["S"])
    model.add_keyframes(L["kf0"], None, L["n_slots"], L["cells"], L["slot_mp"], L["obs_mp"])
    model.set_map_points(L["mp0"], L["nobs"])
    eng.set_params(N, LAM, GLAM)
    r = mir.solve([L["window"]], apply=True)[0]
    assert r.deleted.size > 0
    model.delete(r.deleted)
    rng = np.random.default_rng(1)
    pay, want = [], []
    for kf in L["window"]:
        n = int(model.kf_n[kf])
        arrs = random_payload(rng, n)
        pay.append(KeyframePayload(eng, n, None, *arrs, kf=int(kf)))
        want.append(mm.erase_bad_descriptor_rows(model.slot_mp[kf, :n] >= 0, *arrs))
    n_out = compact_keyframes(eng, pay, mirror=mir)
    for kf, p, w, n in zip(L["window"], pay, want, n_out):
        assert n == w[0].shape[0]
        got = p.fetch(int(n))
        assert np.array_equal(got["descriptors"], w[0]) and np.array_equal(got["keypoints"], w[1].view(np.uint32))
        assert np.array_equal(got["uright"], w[2]) and np.array_equal(got["depth"], w[3])
        model.compact(int(kf))
        p.free()
    # the mirror followed: the view it assembles now equals the model's
    pv, mp_handle, _ = mir.build_view(L["window"])
    mv, m_handle, _ = model.build(L["window"])
    assert (pv.K, pv.M, pv.F) == (mv.K, mv.M, mv.F) and np.array_equal(mp_handle, m_handle) and np.array_equal(pv.feat_ptr, mv.feat_ptr)
    mir.close()

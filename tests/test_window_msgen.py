"""Window view container + msgen-v1 generator (CPU)."""
import numpy as np
import pytest

from ms_slam_b200 import WindowView, make_view, msgen, CELL_NONE, N_CELLS


def test_generator_is_deterministic_and_valid():
    a, N = msgen.make_config("live", seed=3)
    b, _ = msgen.make_config("live", seed=3)
    c, _ = msgen.make_config("live", seed=4)
    a.validate()
    for name in WindowView._ARRAYS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert not np.array_equal(a.feat_mp, c.feat_mp)
    assert N == 100 and a.K == 30 and a.H == 10 and a.M == 6000


@pytest.mark.parametrize("name", ["c1", "c3", "c4"])
def test_config_shapes(name):
    v, N = msgen.make_config(name, 0)
    v.validate()
    cfg = msgen.CONFIGS[name]
    assert v.K == cfg["K"] and v.M == cfg["M"] and v.H == cfg["H"] and v.F == cfg["K"] * cfg["n_feat"]
    valid = v.feat_mp >= 0
    # every slot's map point lists the keyframe among its observations (view self-consistency)
    kf_of_slot = np.repeat(np.arange(v.K), np.diff(v.feat_ptr))
    p = v.feat_mp[valid][:200]
    k = kf_of_slot[valid][:200]
    for pi, ki in zip(p, k):
        assert ki in v.mp_obs_kf[v.mp_obs_ptr[pi]:v.mp_obs_ptr[pi + 1]]
    # one map point at most once per keyframe, cells in range
    key = kf_of_slot[valid].astype(np.int64) * v.M + v.feat_mp[valid]
    assert np.unique(key).size == key.size
    cells = v.feat_cell[valid]
    assert ((cells < N_CELLS) | (cells == CELL_NONE)).all()
    # generator contract: max cost < Lambda
    assert v.mp_nobs[v.feat_mp[valid]].max() - v.mp_nobs.min() < msgen.LAMBDA
    assert (v.mp_nobs >= 3).all()


def test_validate_rejects_bad_views():
    v = make_view(1, [[(0, 0), (1, 1)]], [4, 6])
    v.validate()
    bad = make_view(1, [[(0, 0), (1, 1)]], [4, 6])
    bad.feat_mp[0] = 7
    with pytest.raises(ValueError):
        bad.validate()
    bad = make_view(1, [[(0, 0), (1, 1)]], [4, 6])
    bad.feat_cell[0] = 4000
    with pytest.raises(ValueError):
        bad.validate()
    bad = make_view(1, [[(0, 0), (1, 1)]], [4, 6])
    bad.feat_ptr[-1] = 5
    with pytest.raises(ValueError):
        bad.validate()


def test_save_load_roundtrip(tmp_path):
    v, _ = msgen.make_config("c1", 1)
    path = str(tmp_path / "w.npz")
    v.save(path)
    w = WindowView.load(path)
    for name in WindowView._ARRAYS:
        assert np.array_equal(getattr(v, name), getattr(w, name))
    assert w.input_bytes() == v.input_bytes()


# ---- transport forms (what FlattenWindow emits / bench.py ships) -------------------------------------------------------
def _unpack(pv):
    """PackedView -> (kf, mp, cell) triples of the valid slots and (mp, outside kf) pairs"""
    kf = np.repeat(np.arange(pv.K), np.diff(pv.feat_ptr))
    ok = pv.slots != 0xFFFFFFFF
    cell = (pv.slots & 0xFFF).astype(np.int64)
    cell = np.where(cell == 0xFFF, CELL_NONE, cell)
    return (kf[ok], (pv.slots >> 12).astype(np.int64)[ok], cell[ok]), ((pv.obs_pairs >> 12).astype(np.int64), (pv.obs_pairs & 0xFFF).astype(np.int64))


@pytest.mark.parametrize("name,seed", [("c1", 0), ("live", 1), ("c3", 0)])
def test_packed_view_carries_the_same_window(name, seed):
    from ms_slam_b200 import pack_view
    v, _ = msgen.make_config(name, seed)
    valid = v.feat_mp >= 0
    kf = np.repeat(np.arange(v.K), np.diff(v.feat_ptr))
    want = sorted(zip(kf[valid].tolist(), v.feat_mp[valid].tolist(), v.feat_cell[valid].tolist()))
    owner = np.repeat(np.arange(v.M), np.diff(v.mp_obs_ptr))
    out = v.mp_obs_kf >= v.K
    want_pairs = sorted(zip(owner[out].tolist(), (v.mp_obs_kf[out] - v.K).tolist()))
    for pv in (pack_view(v), pack_view(v.compact()), pack_view(v.compact(), sort_slots=True)):
        (k, m, c), (pm, pj) = _unpack(pv)
        assert sorted(zip(k.tolist(), m.tolist(), c.tolist())) == want
        assert sorted(zip(pm.tolist(), pj.tolist())) == want_pairs
        assert np.array_equal(pv.mp_nobs16, v.mp_nobs) and np.array_equal(pv.okf_total, v.okf_total)
    srt = pack_view(v.compact(), sort_slots=True)
    for k in range(srt.K):                                   # sorted inside every keyframe, keyframe boundaries untouched
        seg = srt.slots[srt.feat_ptr[k]:srt.feat_ptr[k + 1]]
        assert np.all(np.diff(seg.astype(np.int64)) >= 0)
    assert pack_view(v.compact()).input_bytes() < 0.62 * v.compact().input_bytes()


def test_packed_view_range_checks():
    from ms_slam_b200 import pack_view
    v = make_view(1, [[(0, 0)]], [70000])
    with pytest.raises(ValueError):
        pack_view(v)


@pytest.mark.parametrize("name,seed", [("c1", 1), ("live", 2)])
def test_discovery_order_is_a_renumbering(name, seed):
    v, _ = msgen.make_config(name, seed)
    d = v.compact().discovery_order()
    d.validate()
    perm = d.meta["mp_perm"]
    assert np.array_equal(np.sort(perm), np.arange(v.M))
    assert np.array_equal(d.mp_nobs, v.mp_nobs[perm])
    # first appearances (keyframes in window order, slots in slot order) are 0, 1, 2, ...
    seen = d.feat_mp[d.feat_mp >= 0]
    _, first = np.unique(seen, return_index=True)
    assert np.all(np.diff(first) > 0)
    # same incidences after mapping back
    kf = np.repeat(np.arange(d.K), np.diff(d.feat_ptr))
    a = sorted(zip(kf.tolist(), perm[d.feat_mp].tolist(), d.feat_cell.tolist()))
    c = v.compact()
    kfc = np.repeat(np.arange(c.K), np.diff(c.feat_ptr))
    assert a == sorted(zip(kfc.tolist(), c.feat_mp.tolist(), c.feat_cell.tolist()))


def test_merge_views_is_block_diagonal():
    from ms_slam_b200 import merge_views
    parts = [msgen.make_config("c1", 7)[0], msgen.make_config("live", 9, M=1500, H=20)[0]]
    for inter in (True, False):
        w = merge_views(parts, interleave=inter)
        w.validate()
        assert (w.K, w.H, w.M, w.F, w.O) == tuple(sum(getattr(p, a) for p in parts) for a in ("K", "H", "M", "F", "O"))
        # no slot of a keyframe of part 0 holds a map point of part 1 and vice versa
        kf = np.repeat(np.arange(w.K), np.diff(w.feat_ptr))
        ok = w.feat_mp >= 0
        part_of_mp = (w.feat_mp[ok] >= parts[0].M).astype(int)
        nslots = np.diff(w.feat_ptr)
        # a keyframe belongs to the part whose slot count it has at its position in the interleaving / concatenation
        kf_part = np.zeros(w.K, int)
        for k in range(w.K):
            seg = w.feat_mp[w.feat_ptr[k]:w.feat_ptr[k + 1]]
            seg = seg[seg >= 0]
            kf_part[k] = int(seg[0] >= parts[0].M) if seg.size else 0
        assert np.array_equal(part_of_mp, kf_part[kf[ok]])


# ---- property tests (hypothesis) ---------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.tuples(st.integers(0, 300000), st.one_of(st.none(), st.integers(0, 3071))), max_size=40), min_size=1, max_size=5),
       st.integers(0, 3))
def test_token_coding_roundtrip_property(rows, extra):
    """any set of slots (arbitrary gaps, duplicates, off-grid keypoints, empty keyframes) survives MSS_LAYOUT_PACKED16"""
    from ms_slam_b200 import pack_view
    M = 300001 + extra
    v = make_view(len(rows), rows, [5] * M)
    pv = pack_view(v, tokens16=True)
    got = []
    for k in range(pv.K):
        mp = 0
        for t in pv.slots[pv.feat_ptr[k]:pv.feat_ptr[k + 1]].tolist():
            d, low = t >> 12, t & 0xFFF
            if d < 15:
                mp += d
                got.append((k, mp, None if low == 0xFFF else low))
            else:
                mp += 15 * (low + 1)
    want = sorted(((k, p, c) for k, row in enumerate(rows) for p, c in row), key=lambda x: (x[0], x[1], -1 if x[2] is None else x[2]))
    assert sorted(got, key=lambda x: (x[0], x[1], -1 if x[2] is None else x[2])) == want
    assert int(pv.feat_ptr[-1]) == pv.slots.size


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.integers(1, 30), st.integers(0, 2 ** 32 - 1))
def test_components_oracle_vs_naive_union_find(K, M, seed):
    """oracle/components.py (scipy) against a plain union-find over the same edges, canonical labels included"""
    from oracle import components as oc
    rng = np.random.default_rng(seed)
    H = int(rng.integers(0, 3))
    rows = [[(int(p), None if rng.random() < 0.2 else int(rng.integers(0, 3072))) for p in rng.choice(M, size=int(rng.integers(0, min(M, 6) + 1)), replace=False)]
            for _ in range(K)]
    outside = [rng.choice(M, size=int(rng.integers(0, min(M, 4) + 1)), replace=False).tolist() for _ in range(H)]
    v = make_view(K, rows, [5] * M, outside=outside, okf_total=[9] * H if H else None)
    rl, ml, nc, _ = oc.components(v)
    parent = list(range(K + H + M))
    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    isvar = [False] * M
    for k, row in enumerate(rows):
        for p, c in row:
            if c is not None:
                isvar[p] = True
                a, b = find(k), find(K + H + p)
                parent[max(a, b)] = min(a, b)
    for j, lst in enumerate(outside):
        for p in lst:
            if isvar[p]:
                a, b = find(K + j), find(K + H + p)
                parent[max(a, b)] = min(a, b)
    roots = sorted({find(r) for r in range(K + H)})
    dense = {r: i for i, r in enumerate(roots)}
    assert nc == len(roots)
    assert rl.tolist() == [dense[find(r)] for r in range(K + H)]
    assert ml.tolist() == [dense[find(K + H + p)] if isvar[p] else -1 for p in range(M)]


def test_one_byte_nobs_table_is_chosen_only_when_every_value_fits():
    """mss_window_view::nobs8: pack_view (like FlattenWindow) sends Observations() as bytes with the u16-token layout whenever
    max <= 255, keeps u16 otherwise, and refuses an explicit nobs8 that would truncate"""
    import pytest
    from ms_slam_b200 import make_view
    from ms_slam_b200.window import pack_view
    slots = [[(0, 5), (1, 6), (2, 7)], [(1, 8), (2, 9), (3, 10)]]
    small = make_view(2, slots, [4, 8, 255, 3])
    big = make_view(2, slots, [4, 8, 256, 3])
    a, b = pack_view(small, tokens16=True), pack_view(big, tokens16=True)
    assert a.meta["nobs8"] and a.mp_nobs16.dtype == np.uint8 and a.mp_nobs16.tolist() == [4, 8, 255, 3]
    assert not b.meta["nobs8"] and b.mp_nobs16.dtype == np.uint16 and b.mp_nobs16.tolist() == [4, 8, 256, 3]
    assert not pack_view(small).meta["nobs8"]                      # the u32-slot layout keeps u16 unless asked
    assert pack_view(small, nobs8=True).mp_nobs16.dtype == np.uint8
    with pytest.raises(ValueError):
        pack_view(big, nobs8=True)
    assert a.input_bytes() == b.input_bytes() - 4


def test_tie_ranks_follow_the_gids_not_the_numbering():
    from ms_slam_b200.window import tie_ranks, pack_view
    v, _ = msgen.make_config("c1", 3)
    d = v.compact().discovery_order()
    assert np.array_equal(tie_ranks(v), np.arange(v.M, dtype=np.uint32))           # generated gids are ascending
    assert np.array_equal(tie_ranks(d), d.meta["mp_perm"].astype(np.uint32))       # rank = the original index, wherever the point went
    assert np.array_equal(pack_view(d, tokens16=True, tie=True).mp_tie, tie_ranks(d)) and pack_view(d).mp_tie is None

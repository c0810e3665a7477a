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

"""Device-side lower bound of the reference ILP (mss_result.dual_bound, include/mss.h mss_set_dual_bound, csrc/mss_bound.cuh).

CPU: the integer twin oracle/dual_bound.py device_twin() is a valid bound (never above the committed LP optimum of the
reference model, tests/golden/config_bounds.json) and certifies the 1 % bar of north_star on its own.
GPU: the device reports exactly the twin's number (integer fixed point: bit-identical), the selection is unchanged by
switching the bound on, and objective <= 1.01 * dual_bound on every committed configuration.
"""
import json

import numpy as np
import pytest

from conftest import LAM, GLAM, view_from_fixture
from ms_slam_b200 import msgen
from ms_slam_b200.window import pack_view
from oracle import dual_bound as db, emulate as em, ilp_model as om

SMALL = ["c1:0:{}", "c1:3:{}", "live:0:{}", "live:1:{}", 'c4:0:{"M": 3000}', 'live:0:{"H": 20, "M": 1500}', "c3:0:{}", "c4:1000:{}"]
BIG = ["c2:0:{}", "c2:176:{}", "c3:4:{}", "c5:0:{}"]


def make(key):
    name, seed, over = key.split(":", 2)
    return msgen.make_config(name, int(seed), **json.loads(over))


@pytest.mark.parametrize("key", SMALL)
def test_twin_is_a_valid_and_tight_bound(key, config_bounds):
    view, N = make(key)
    t = db.device_twin(view, N, LAM, GLAM)
    lp = config_bounds[key]["lp"]
    assert t["flag"] == 1
    assert t["bound"] <= lp + 1e-6                        # validity: LP* >= bound
    assert t["objective"] <= 1.01 * t["bound"]            # certifies the 1 % bar without a solver
    assert t["u0"] >= 0 and t["res_unc"] >= 0 and t["bound"] >= 0.99 * lp


def test_twin_on_known_answers(known_answers):
    """micro windows: the bound never exceeds the brute-force optimum; when dominance decides everything it equals it"""
    for name, rec in known_answers.items():
        view = view_from_fixture(rec)
        t = db.device_twin(view, rec["N"], rec["lam"], rec["grid_lam"])
        assert t["flag"] in (1, 2)
        assert t["bound"] <= rec["F_opt"] + 1e-9, name
        if t["flag"] == 2:
            assert t["bound"] == pytest.approx(rec["F_opt"], abs=1e-9)


@pytest.fixture(scope="module")
def eng(build_native):
    from ms_slam_b200.engine import Engine
    e = Engine(N=100, lam=LAM, grid_lam=GLAM, device=0)
    yield e
    e.close()


def check(eng, view, N, lp=None):
    eng.set_params(N, LAM, GLAM)
    eng.set_dual_bound(False)
    plain = eng.solve(view)
    assert np.isnan(plain.dual_bound)
    eng.set_dual_bound(True)
    outs = [eng.solve(view), eng.solve(pack_view(view.compact(), tokens16=True))]    # SoA view, packed16 transport form
    eng.set_dual_bound(False)
    t = db.device_twin(view, N, LAM, GLAM)
    for i, r in enumerate(outs):
        if i == 0:
            assert np.array_equal(r.keep, plain.keep)                     # the certificate does not change the selection
        assert (r.objective, r.rounds, r.n_kept) == (plain.objective, plain.rounds, plain.n_kept)
        assert r.objective == t["objective"]
        assert r.dual_bound == t["bound"], (r.dual_bound, t)             # integer fixed point on both sides: exact
        assert r.objective <= 1.01 * r.dual_bound
        if lp is not None:
            assert r.dual_bound <= lp + 1e-6
    return outs[0], t


@pytest.mark.gpu
@pytest.mark.parametrize("key", SMALL)
def test_device_bound_equals_the_twin(eng, key, config_bounds):
    view, N = make(key)
    check(eng, view, N, config_bounds[key]["lp"])


@pytest.mark.gpu
@pytest.mark.parametrize("key", BIG)
def test_device_bound_full_size(eng, key, config_bounds):
    view, N = make(key)
    r, t = check(eng, view, N, config_bounds[key]["lp"])
    assert t["flag"] == 1


@pytest.mark.gpu
def test_device_bound_known_answers_and_batches(eng, known_answers):
    for name, rec in known_answers.items():
        view = view_from_fixture(rec)
        eng.set_params(rec["N"], rec["lam"], rec["grid_lam"])
        eng.set_dual_bound(True)
        r = eng.solve(view)
        eng.set_dual_bound(False)
        t = db.device_twin(view, rec["N"], rec["lam"], rec["grid_lam"])
        assert r.dual_bound == t["bound"] and r.dual_bound <= rec["F_opt"] + 1e-9, name
    # a batch (several windows per launch, groups of CTAs): every window carries its own certificate
    eng.set_params(100, LAM, GLAM)
    views = [msgen.make_config("live", s)[0] for s in range(6)] + [msgen.make_config("c4", 1000 + s, M=4000, K=40)[0] for s in range(3)]
    eng.set_dual_bound(True)
    res = eng.solve_batch([pack_view(v.compact()) for v in views])
    eng.set_dual_bound(False)
    for v, r in zip(views, res):
        t = db.device_twin(v, 100, LAM, GLAM)
        assert r.dual_bound == t["bound"] and r.objective == t["objective"] and r.objective <= 1.01 * r.dual_bound

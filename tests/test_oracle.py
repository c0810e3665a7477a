"""CPU tests of the oracle: the ILP restatement (oracle/ilp_model.py) against hand-derived known answers, brute force
and HiGHS, and the emulation of the device algorithm (oracle/emulate.py) against the ILP."""
import hashlib
import json

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from conftest import view_from_fixture, fixture_key, LAM, GLAM
from ms_slam_b200 import make_view, msgen
from oracle import ilp_model as om, emulate as em


def test_known_answers_bruteforce_and_highs(known_answers):
    # fixtures carry the hand-derived optimum; brute force and HiGHS must both reproduce it
    for name, rec in known_answers.items():
        view = view_from_fixture(rec)
        view.validate()
        F_bf, args = om.brute_force(view, rec["N"], rec["lam"], rec["grid_lam"])
        assert F_bf == pytest.approx(rec["F_opt"], abs=1e-9), name
        sol = om.solve_ilp(view, rec["N"], rec["lam"], rec["grid_lam"], mip_rel_gap=0.0)
        assert sol.objective == pytest.approx(rec["F_opt"], abs=1e-6), name
        model = om.build_model(view, rec["N"])
        assert model.n_max == rec["n_max"], name
        assert model.out_need.tolist() == rec["out_need"], name
        keeps = [sorted(int(model.var_mp[i]) for i in np.nonzero(a)[0]) for a in args]
        assert keeps == rec["optimal_keep_sets"], name


def test_hand_values():
    # KA-1: keep {A,B,C}, delete D, F* = 6
    v = make_view(2, [[(0, 0), (1, 1), (2, 1)], [(1, 5), (2, 6), (3, 6)]], [4, 8, 6, 4])
    m = om.build_model(v, 2)
    assert m.n_max == 8 and m.cost.tolist() == [4, 0, 2, 4]
    assert om.objective(m, [1, 1, 1, 0], 2, LAM, GLAM) == 6
    assert om.objective(m, [0, 0, 0, 0], 2, LAM, GLAM) == 4 * 10 + 500 * 4
    # outside-row rhs: cnt=2,total=7,N=100 -> fl32(2/7)*100 = 28.57 -> 29
    assert int(om.outside_need(2, 7, 100)) == 29
    assert int(om.outside_need(50, 100, 100)) == 50
    assert int(em.outside_need(np.array([2]), np.array([7]), 100)[0]) == 29


def test_discovery_order_and_multiplicity():
    # variables are indexed in discovery order: KF order, cells col-major, features in slot order (MapSparsification.cc:78-99)
    v = make_view(1, [[(2, 100), (0, 5), (1, 5), (0, 7)]], [5, 5, 5])
    m = om.build_model(v, 1)
    assert m.var_mp.tolist() == [0, 1, 2]
    # MP 0 sits in two slots of the keyframe -> coefficient 2 in the keyframe row
    kf_cov, _, _ = om.coverage(m, [1, 0, 0])
    assert kf_cov.tolist() == [2]


def test_emulator_on_known_answers(known_answers):
    for name, rec in known_answers.items():
        view = view_from_fixture(rec)
        r = em.solve(view, rec["N"], rec["lam"], rec["grid_lam"])
        model = om.build_model(view, rec["N"])
        x = om.keep_to_x(model, r["keep"])
        F = om.objective(model, x, rec["N"], rec["lam"], rec["grid_lam"])
        assert F == r["objective"], name
        assert F == pytest.approx(rec["F_opt"], abs=1e-9), name          # exact on every micro fixture
        assert om.rows_satisfied(model, x, rec["N"])[0], name
        # non-variables are never deleted
        nonvar = np.ones(view.M, bool)
        nonvar[model.var_mp] = False
        assert r["keep"][nonvar].all(), name


def test_emulation_checksums(emulation_golden):
    for key, rec in emulation_golden.items():
        name, seed, over = key.split(":", 2)
        if name == "c2":
            continue                     # full-size: covered by the gpu suite
        view, N = msgen.make_config(name, int(seed), **json.loads(over))
        vh = hashlib.sha256(view.feat_mp.tobytes() + view.feat_cell.tobytes() + view.mp_nobs.tobytes()
                            + view.mp_obs_kf.tobytes() + view.okf_total.tobytes()).hexdigest()
        assert vh == rec["view_sha256"], f"msgen-v1 drifted for {key}"
        r = em.solve(view, N, LAM, GLAM)
        assert r["objective"] == rec["objective"] and r["n_kept"] == rec["n_kept"] and r["rounds"] == rec["rounds"], key
        assert hashlib.sha256(em.pack_bits(r["keep"]).tobytes()).hexdigest() == rec["keep_sha256"], key


@pytest.mark.parametrize("name,seed,over", [("c1", 0, {}), ("c1", 3, {}), ("live", 0, {}), ("c4", 0, dict(M=3000)),
                                            ("live", 0, dict(M=1500, H=20))])
def test_emulator_within_one_percent_of_ilp(config_bounds, name, seed, over):
    rec = config_bounds[fixture_key(name, seed, over)]
    view, N = msgen.make_config(name, seed, **over)
    r = em.solve(view, N, LAM, GLAM)
    model = om.build_model(view, N)
    x = om.keep_to_x(model, r["keep"])
    assert om.rows_satisfied(model, x, N)[0]
    assert rec["lp"] - 1e-6 <= r["objective"] <= 1.01 * rec["ilp"]
    assert r["objective"] <= 1.001 * rec["ilp"]          # in practice two orders of magnitude tighter than the bar


def test_ilp_matches_committed_bounds(config_bounds):
    # the oracle's own solve is reproducible: re-solve one small config and compare with the committed value
    rec = config_bounds[fixture_key("c1", 1)]
    view, N = msgen.make_config("c1", 1)
    s = om.solve_ilp(view, N, LAM, GLAM)
    model = om.build_model(view, N)
    assert om.objective(model, s.x, N, LAM, GLAM) == pytest.approx(rec["ilp"], rel=2e-3)
    assert om.solve_lp(view, N, LAM, GLAM, model=model).objective == pytest.approx(rec["lp"], rel=1e-6)


# ---- property tests on random micro windows ---------------------------------------------------------------------------
@st.composite
def micro_windows(draw):
    K = draw(st.integers(1, 3))
    M = draw(st.integers(1, 9))
    H = draw(st.integers(0, 2))
    nobs = [draw(st.integers(3, 40)) for _ in range(M)]
    slots = []
    for _ in range(K):
        n = draw(st.integers(0, 6))
        row = []
        for _ in range(n):
            p = draw(st.one_of(st.none(), st.integers(0, M - 1)))
            cell = draw(st.one_of(st.none(), st.integers(0, 3)))
            row.append((p, cell))
        slots.append(row)
    outside = [sorted(set(draw(st.lists(st.integers(0, M - 1), max_size=4)))) for _ in range(H)]
    total = [max(1, len(o) + draw(st.integers(0, 5))) for o in outside]
    N = draw(st.integers(0, 4))
    return make_view(K, slots, nobs, outside=outside, okf_total=total), N


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(micro_windows())
def test_property_bruteforce_vs_highs_vs_emulator(case):
    view, N = case
    view.validate()
    F_bf, _ = om.brute_force(view, N, LAM, GLAM)
    sol = om.solve_ilp(view, N, LAM, GLAM, mip_rel_gap=0.0)
    assert sol.objective == pytest.approx(F_bf, abs=1e-6)
    r = em.solve(view, N, LAM, GLAM)
    model = om.build_model(view, N)
    x = om.keep_to_x(model, r["keep"])
    assert om.objective(model, x, N, LAM, GLAM) == r["objective"]
    assert r["objective"] >= F_bf - 1e-9
    assert om.rows_satisfied(model, x, N)[0]


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(micro_windows())
def test_property_dominance_rules_are_exact(case):
    """PROP only fixes variables that some optimal solution agrees with: brute force over the remaining FREE variables
    must still reach the global optimum."""
    view, N = case
    F_bf, _ = om.brute_force(view, N, LAM, GLAM)
    e = em.Emulator(view, N, LAM, GLAM)
    while True:
        changed, nfree = e._prop_round()
        if changed == 0:
            break
    model = om.build_model(view, N)
    st_var = e.st[model.var_mp]
    free = np.nonzero(st_var == em.FREE)[0]
    base = (st_var == em.IN).astype(np.int64)
    best = None
    for mask in range(1 << free.size):
        x = base.copy()
        x[free] = (mask >> np.arange(free.size)) & 1
        f = om.objective(model, x, N, LAM, GLAM)
        best = f if best is None else min(best, f)
    assert best == pytest.approx(F_bf, abs=1e-9)


@pytest.mark.parametrize("name,seed,over", [("c1", 0, {}), ("live", 0, {}), ("c4", 0, dict(M=3000)), ("live", 0, dict(M=1500, H=20))])
def test_dual_bound_is_valid_and_tight(name, seed, over):
    """The solver-free lower bound planned for mss_result.dual_bound (oracle/dual_bound.py): never above the LP optimum of
    the reference model, and close enough to certify the 1 % bar on its own."""
    from oracle import dual_bound as db
    view, N = msgen.make_config(name, seed, **over)
    b = db.dual_bound(view, N, LAM, GLAM)
    lp = om.solve_lp(view, N, LAM, GLAM).objective
    F = em.solve(view, N, LAM, GLAM)["objective"]
    assert b["bound"] <= lp + 1e-6
    assert F <= 1.01 * b["bound"]
    assert b["bound"] >= 0.99 * lp

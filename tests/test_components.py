"""Connected components of a window (mss_components) and the exact decomposition of the flush model along them."""
import numpy as np
import pytest

from conftest import LAM, GLAM
from ms_slam_b200 import make_view, msgen, merge_views, split_components, pack_view
from oracle import components as oc, emulate as em, ilp_model as om


def three_part_window():
    parts = [msgen.make_config("live", 31)[0], msgen.make_config("c1", 5)[0], msgen.make_config("live", 32, M=1500, H=20)[0]]
    return parts, merge_views(parts, interleave=True)


def test_oracle_components_hand_cases():
    # two keyframes sharing nothing, one empty keyframe, one map point only seen off-grid
    v = make_view(3, [[(0, 1), (1, 2)], [(2, 5), (3, None)], []], [5, 6, 7, 9])
    rows, mps, nc, nmax = oc.components(v)
    assert rows.tolist() == [0, 1, 2] and mps.tolist() == [0, 0, 1, -1] and nc == 3 and nmax == 9
    # an outside keyframe that observes one point of each keyframe ties them together
    v = make_view(2, [[(0, 1)], [(1, 5)]], [5, 6], outside=[[0, 1]], okf_total=[4])
    rows, mps, nc, _ = oc.components(v)
    assert rows.tolist() == [0, 0, 0] and mps.tolist() == [0, 0] and nc == 1
    # ... but not through a map point that is no variable
    v = make_view(2, [[(0, 1), (2, None)], [(1, 5)]], [5, 6, 7], outside=[[0], [2, 1]], okf_total=[4, 4])
    rows, mps, nc, _ = oc.components(v)
    assert rows.tolist() == [0, 1, 0, 1] and mps.tolist() == [0, 1, -1] and nc == 2


def test_merged_window_splits_back_and_objective_decomposes():
    parts, whole = three_part_window()
    whole.validate()
    rows, mps, nc, nmax = oc.components(whole)
    assert nc >= 3
    N = 100
    split = split_components(whole, rows, mps, nmax)
    assert sum(p.K for p, _, _ in split) == whole.K
    keep = np.ones(whole.M, bool)
    F_parts = 0.0
    for p, kf_idx, mp_idx in split:
        p.validate()
        r = em.solve(p, N, LAM, GLAM)
        assert r["n_max"] == nmax
        keep[mp_idx] = r["keep"]
        F_parts += r["objective"]
    model = om.build_model(whole, N)
    x = om.keep_to_x(model, keep)
    F_union = om.objective(model, x, N, LAM, GLAM)
    assert F_union == F_parts                                  # the model is block diagonal: F is the sum over components
    assert om.rows_satisfied(model, x, N)[0]
    F_whole = em.solve(whole, N, LAM, GLAM)["objective"]      # same heuristic on the undivided window: different schedule,
    assert abs(F_union - F_whole) <= 1e-3 * F_whole           # same quality (both within the 1 % bar of the same optimum)
    lp = om.solve_lp(whole, N, LAM, GLAM, model=model).objective
    assert lp - 1e-6 <= F_union <= 1.01 * lp


@pytest.mark.gpu
def test_gpu_components_match_oracle(build_native):
    from ms_slam_b200.engine import Engine
    eng = Engine(N=100, lam=LAM, grid_lam=GLAM)
    _, whole = three_part_window()
    cases = [whole, msgen.make_config("c1", 0)[0], msgen.make_config("live", 0, M=1500, H=20)[0], msgen.make_config("c3", 0)[0],
             make_view(3, [[(0, 1), (1, 2)], [(2, 5), (3, None)], []], [5, 6, 7, 9]),
             make_view(2, [[(0, 1), (2, None)], [(1, 5)]], [5, 6, 7], outside=[[0], [2, 1]], okf_total=[4, 4]),
             make_view(0, [], [5, 6])]
    for v in cases:
        ref = oc.components(v)
        for view in (v, pack_view(v.compact()), pack_view(v, tokens16=True)):
            rows, mps, nc, nmax = eng.components(view)
            assert np.array_equal(rows, ref[0]) and np.array_equal(mps, ref[1]) and (nc, nmax) == (ref[2], ref[3])
    eng.close()


@pytest.mark.gpu
def test_gpu_flush_as_batch_of_components(build_native):
    """A flush window with several components solved as one batch of independent windows (each carrying the window-wide
    nMax): per-component parity with the emulation, union feasible on the whole model, objective = sum of the parts."""
    from ms_slam_b200.engine import Engine
    N = 100
    eng = Engine(N=N, lam=LAM, grid_lam=GLAM)
    _, whole = three_part_window()
    rows, mps, nc, nmax = eng.components(pack_view(whole.compact()))
    split = split_components(whole, rows, mps, nmax)
    res = eng.solve_batch([pack_view(p.compact()) for p, _, _ in split])
    keep = np.ones(whole.M, bool)
    for (p, _, mp_idx), r in zip(split, res):
        ref = em.solve(p, N, LAM, GLAM)
        assert np.array_equal(r.keep, ref["keep"]) and r.objective == ref["objective"] and r.n_max == nmax
        keep[mp_idx] = r.keep
    model = om.build_model(whole, N)
    x = om.keep_to_x(model, keep)
    assert om.objective(model, x, N, LAM, GLAM) == sum(r.objective for r in res)
    assert om.rows_satisfied(model, x, N)[0]
    eng.close()

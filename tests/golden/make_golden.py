"""Generates tests/golden/*.json.  Run from the repo root:  python tests/golden/make_golden.py [--slow]

The reference (fishmarch/MS-SLAM) ships no golden vectors for MapSparsification::Sparsifying and cannot be built
here (GUROBI/OpenCV/Eigen absent), so the fixtures are produced by the oracle itself and pinned three ways:
  known_answers.json  hand-derived micro windows (SURVEY Appendix A.6 KA-1..4 + quirk cases); every optimum is
                      enumerated by 2^V brute force AND solved by HiGHS at generation time, both must agree.
  config_bounds.json  HiGHS ILP optimum (MIPGap 0.002 as MapSparsification.cc:155) and LP bound of seeded msgen-v1
                      windows, so GPU tests can check "within 1 % of the ILP optimum" without re-solving big models.
  emulation.json      checksums of the CPU emulation of the device algorithm on seeded windows (drift detector).
"""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ms_slam_b200 import make_view, msgen
from oracle import ilp_model as om, emulate as em

HERE = os.path.dirname(os.path.abspath(__file__))
LAM, GLAM = 500.0, 10.0


def ka_windows():
    """name -> (kf_slots, nobs, outside, okf_total, N, expected F*, expected kept set or None)"""
    A, B, C_, D = 0, 1, 2, 3
    topo = [[(A, 0), (B, 1), (C_, 1)], [(B, 5), (C_, 6), (D, 6)]]
    W = {}
    W["KA-1"] = dict(slots=topo, nobs=[4, 8, 6, 4], outside=[], total=[], N=2, F=6.0, keep=[A, B, C_])
    W["KA-2"] = dict(slots=topo, nobs=[4, 20, 6, 4], outside=[], total=[], N=2, F=24.0, keep=[B, C_])
    W["KA-3"] = dict(slots=[[(A, 0), (B, 1)]], nobs=[4, 6], outside=[], total=[], N=3, F=502.0, keep=[A, B])
    W["KA-4"] = dict(slots=[[(A, 0), (B, 1), (C_, 2)]], nobs=[6, 6, 4], outside=[[A, C_]], total=[1], N=1, F=2.0,
                     keep=[A, B, C_])
    # (i) one KF, three MPs in three cells, N=2: two cheapest kept; third cell bought iff its point costs < GridLambda
    W["Q-cheap-third"] = dict(slots=[[(0, 0), (1, 1), (2, 2)]], nobs=[30, 29, 25], outside=[], total=[], N=2, F=6.0,
                              keep=[0, 1, 2])
    W["Q-dear-third"] = dict(slots=[[(0, 0), (1, 1), (2, 2)]], nobs=[30, 29, 5], outside=[], total=[], N=2, F=11.0,
                             keep=[0, 1])
    # (ii) two KFs sharing one long-track MP of cost 0
    W["Q-shared"] = dict(slots=[[(0, 0), (1, 7)], [(0, 3), (2, 9)]], nobs=[40, 8, 8], outside=[], total=[], N=1, F=20.0,
                         keep=[0])
    # (iv) outside row with fractional rhs: cnt=2,total=7,N=100 -> 28.57 -> need 29 > cnt: both kept, slack 27
    W["Q-outside-ceil"] = dict(slots=[[(0, 0), (1, 1)]], nobs=[6, 6], outside=[[0, 1]], total=[7], N=100,
                               F=500.0 * (98 + 27), keep=[0, 1])
    # (v) duplicate MP in two slots of one keyframe -> coefficient 2
    W["Q-duplicate"] = dict(slots=[[(0, 0), (0, 1), (1, 2)]], nobs=[5, 9], outside=[], total=[], N=2, F=None, keep=None)
    # empty / degenerate rows: a keyframe with no valid slot pays Lambda*N
    W["Q-empty-kf"] = dict(slots=[[(None, None), (None, 3)], [(0, 0)]], nobs=[5], outside=[], total=[], N=2,
                           F=500.0 * 2 + 500.0 * 1, keep=[0])
    # slot whose keypoint is outside the grid: counts for nMax but is not a variable through that slot
    W["Q-offgrid"] = dict(slots=[[(0, None), (1, 4)]], nobs=[50, 6], outside=[], total=[], N=1, F=None, keep=None)
    return W


def main():
    slow = "--slow" in sys.argv
    ka = {}
    for name, w in ka_windows().items():
        view = make_view(len(w["slots"]), w["slots"], w["nobs"], outside=w["outside"], okf_total=w["total"] or None)
        F_bf, args = om.brute_force(view, w["N"], LAM, GLAM)
        sol = om.solve_ilp(view, w["N"], LAM, GLAM, mip_rel_gap=0.0)
        assert abs(sol.objective - F_bf) < 1e-6, (name, sol.objective, F_bf)
        if w["F"] is not None:
            assert abs(F_bf - w["F"]) < 1e-9, (name, F_bf, w["F"])
        model = om.build_model(view, w["N"])
        opt_keeps = [sorted(int(model.var_mp[i]) for i in np.nonzero(a)[0]) for a in args]
        if w["keep"] is not None:
            assert sorted(w["keep"]) in opt_keeps, (name, opt_keeps)
        ka[name] = dict(K=view.K, H=view.H, N=w["N"], lam=LAM, grid_lam=GLAM,
                        feat_ptr=view.feat_ptr.tolist(), feat_mp=view.feat_mp.tolist(), feat_cell=view.feat_cell.tolist(),
                        mp_nobs=view.mp_nobs.tolist(), mp_obs_ptr=view.mp_obs_ptr.tolist(), mp_obs_kf=view.mp_obs_kf.tolist(),
                        okf_total=view.okf_total.tolist(), F_opt=F_bf, optimal_keep_sets=opt_keeps,
                        n_max=model.n_max, out_need=model.out_need.tolist(), n_vars=int(model.var_mp.size))
        print(name, "F*", F_bf, "optima", opt_keeps)
    json.dump(ka, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)

    # ---- ILP / LP values of seeded generator windows -----------------------------------------------------------------
    bounds_path = os.path.join(HERE, "config_bounds.json")
    bounds = json.load(open(bounds_path)) if os.path.exists(bounds_path) else {}
    jobs = [("c1", s, {}, True) for s in range(5)] + [("live", s, {}, True) for s in range(2)] + \
           [("c4", 0, dict(M=3000), True), ("live", 0, dict(M=1500, H=20), True)]
    jobs += [("c3", 0, {}, False), ("c4", 1000, {}, False), ("c4", 1001, {}, False), ("c2", 0, {}, False),
             ("c2", 7, {}, False), ("c2", 14, {}, False), ("c2", 176, {}, False)]
    if slow:
        jobs += [("c3", 0, {}, True), ("c5", 0, {}, False)]
    for name, seed, over, do_ilp in jobs:
        key = f"{name}:{seed}:{json.dumps(over, sort_keys=True)}"
        if key in bounds and (bounds[key].get("ilp") is not None or not do_ilp):
            continue
        view, N = msgen.make_config(name, seed, **over)
        model = om.build_model(view, N)
        t = time.time()
        lp = om.solve_lp(view, N, LAM, GLAM, model=model)
        rec = dict(N=N, lp=lp.objective, lp_seconds=lp.seconds, ilp=None, n_vars=int(model.var_mp.size), G=model.G,
                   nnz=int(model.ent_var.size))
        if do_ilp:
            s = om.solve_ilp(view, N, LAM, GLAM, model=model, time_limit=1200)
            F = om.objective(model, s.x, N, LAM, GLAM)
            rec.update(ilp=F, ilp_seconds=s.seconds, ilp_status=int(s.status), ilp_gap=s.mip_gap)
        bounds[key] = rec
        print(key, rec, f"{time.time()-t:.1f}s", flush=True)
        json.dump(bounds, open(bounds_path, "w"), indent=1)

    # ---- emulation checksums ------------------------------------------------------------------------------------------
    emu = {}
    emu_jobs = [("c1", 0, {}), ("c1", 1, {}), ("live", 0, {}), ("c3", 0, {}), ("c4", 1000, {}), ("c4", 0, dict(M=3000)),
                ("c2", 0, {}), ("c2", 7, {}), ("c2", 14, {}), ("c2", 176, {})]      # c2:176 = nMax outlier, stall-triggered greedy
    if slow:
        emu_jobs += [("c5", 0, {})]
    else:
        old = json.load(open(os.path.join(HERE, "emulation.json"))) if os.path.exists(os.path.join(HERE, "emulation.json")) else {}
        emu = {k: v for k, v in old.items() if k.startswith("c5:")}            # kept from the last --slow run
    for name, seed, over in emu_jobs:
        view, N = msgen.make_config(name, seed, **over)
        r = em.solve(view, N, LAM, GLAM)
        bits = em.pack_bits(r["keep"])
        emu[f"{name}:{seed}:{json.dumps(over, sort_keys=True)}"] = dict(
            N=N, objective=r["objective"], n_kept=r["n_kept"], n_vars=r["n_vars"], n_cells=r["n_cells"], nnz=r["nnz"],
            rounds=r["rounds"], n_max=r["n_max"], keep_sha256=hashlib.sha256(bits.tobytes()).hexdigest(),
            view_sha256=hashlib.sha256(view.feat_mp.tobytes() + view.feat_cell.tobytes() + view.mp_nobs.tobytes()
                                       + view.mp_obs_kf.tobytes() + view.okf_total.tobytes()).hexdigest())
    json.dump(emu, open(os.path.join(HERE, "emulation.json"), "w"), indent=1)
    print("wrote fixtures")


def _extra_job(job):
    """one window of the round-2 additions: LP bound of the reference model + checksum of the CPU emulation"""
    name, seed, over = job
    view, N = msgen.make_config(name, seed, **over)
    model = om.build_model(view, N)
    lp = om.solve_lp(view, N, LAM, GLAM, model=model)
    r = em.solve(view, N, LAM, GLAM)
    bits = em.pack_bits(r["keep"])
    key = f"{name}:{seed}:{json.dumps(over, sort_keys=True)}"
    bound = dict(N=N, lp=lp.objective, lp_seconds=lp.seconds, ilp=None, n_vars=int(model.var_mp.size), G=model.G,
                 nnz=int(model.ent_var.size))
    emu = dict(N=N, objective=r["objective"], n_kept=r["n_kept"], n_vars=r["n_vars"], n_cells=r["n_cells"], nnz=r["nnz"],
               rounds=r["rounds"], n_max=r["n_max"], keep_sha256=hashlib.sha256(bits.tobytes()).hexdigest(),
               view_sha256=hashlib.sha256(view.feat_mp.tobytes() + view.feat_cell.tobytes() + view.mp_nobs.tobytes()
                                          + view.mp_obs_kf.tobytes() + view.okf_total.tobytes()).hexdigest())
    return key, bound, emu


def extra(procs=6):
    """seeds 0-4 of c2 and c3, and all 64 windows of config 4 (seeds 1000..1063): LP bound + emulation
    checksum each, merged into config_bounds.json / emulation.json.  Run: python tests/golden/make_golden.py --extra"""
    from concurrent.futures import ProcessPoolExecutor
    bounds_path, emu_path = os.path.join(HERE, "config_bounds.json"), os.path.join(HERE, "emulation.json")
    bounds, emu = json.load(open(bounds_path)), json.load(open(emu_path))
    jobs = [("c4", 1000 + w, {}) for w in range(64)] + [("c3", s, {}) for s in range(1, 5)] + [("c2", s, {}) for s in range(1, 5)]
    jobs = [j for j in jobs if f"{j[0]}:{j[1]}:{json.dumps(j[2], sort_keys=True)}" not in emu
            or f"{j[0]}:{j[1]}:{json.dumps(j[2], sort_keys=True)}" not in bounds]
    with ProcessPoolExecutor(max_workers=procs) as pool:
        for key, b, e in pool.map(_extra_job, jobs):
            if key not in bounds:
                bounds[key] = b
            emu[key] = e
            print(key, b["lp"], e["objective"], f"gap {e['objective']/b['lp']-1:.5f}", flush=True)
            json.dump(bounds, open(bounds_path, "w"), indent=1)
            json.dump(emu, open(emu_path, "w"), indent=1)


if __name__ == "__main__":
    if "--extra" in sys.argv:
        extra()
    else:
        main()

import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ms_slam_b200 import msgen, engine as E, mirror as MR
from ms_slam_b200.window import pack_view
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = msgen.CONFIGS[name]
eng = E.Engine(N=cfg["N"], lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA)
mir = MR.Mirror(eng, cfg["n_feat"])
span = cfg["K"] + cfg["H"]
kfs, mp0, views = [], 0, []
for i in range(3):
    v = msgen.make_config(name, seed=i)[0]
    views.append(v)
    L = MR.arrays_from_view(v, S=cfg["n_feat"], seed=i, kf0=i * span, mp0=0, shuffle=False)
    for key in ("slot_mp", "obs_mp"):
        L[key] = np.where(L[key] >= 0, L[key] + mp0, -1).astype(np.int32)
    L["mp0"] = mp0
    mir.load(L)
    kfs.append(L["window"]); mp0 += L["n_mp"]
res = mir.solve(kfs)
for i, (v, r) in enumerate(zip(views, res)):
    ref = eng.solve(pack_view(v.compact().discovery_order(), tokens16=True))
    a, b = r.result, ref
    print(i, "mirror", (a.objective, a.n_kept, a.n_vars, a.rounds, a.n_max, a.nnz, a.n_cells, r.M, r.H, r.F, r.O, r.n_deleted, r.deleted.size),
          "view", (b.objective, b.n_kept, b.n_vars, b.rounds, b.n_max, b.nnz, b.n_cells, v.M, v.H), flush=True)
    pv, mh, okf = mir.build_view(kfs[i])
    dv = v.compact().discovery_order()
    print("   H", pv.H, dv.H, "okf_total eq", np.array_equal(pv.okf_total, dv.okf_total) if pv.H == dv.H else (pv.okf_total[:5], dv.okf_total[:5]),
          "nobs eq", np.array_equal(pv.mp_nobs16[:dv.M], dv.mp_nobs.astype(np.uint16)[:pv.M]), "O", pv.O, dv.O)

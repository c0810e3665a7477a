"""GPU-box check: every config through the C-ABI vs the CPU emulation (bit-exact) + quick timings.
Writes gpurun_out/gpu_check.json.  Run: python tests/gpu_check.py [--big]  (a checker, like the tests: it uses the oracle)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)
from ms_slam_b200 import msgen
from ms_slam_b200.engine import Engine, DeviceView
from oracle import emulate as em

lam, glam = msgen.LAMBDA, msgen.GRID_LAMBDA
cases = [("c1", {}), ("live", {}), ("c3", {}), ("c4", {}), ("c4", dict(M=3000)), ("c4", dict(M=10000, tau=12.0)),
         ("live", dict(M=1500, H=20)), ("c2", {})]
if "--big" in sys.argv:
    cases.append(("c5", {}))
out = []
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for name, over in cases:
    view, N = msgen.make_config(name, 0, **over)
    eng = Engine(N=N, lam=lam, grid_lam=glam)
    t0 = time.time(); res = eng.solve(view); t_first = time.time() - t0
    ref = em.solve(view, N, lam, glam)
    same = bool(np.array_equal(res.keep, ref["keep"]))
    rec = dict(case=name, over=over, same_bitmask=same, F_gpu=res.objective, F_emu=ref["objective"],
               rounds_gpu=res.rounds, rounds_emu=ref["rounds"], n_vars=res.n_vars, n_vars_emu=ref["n_vars"],
               n_cells=res.n_cells, n_cells_emu=ref["n_cells"], nnz=res.nnz, nnz_emu=ref["nnz"], n_kept=res.n_kept,
               cov_same=bool(np.array_equal(res.kf_cov, ref["cov"])), slack_same=bool(np.array_equal(res.kf_slack, ref["slack"])),
               status=res.status, first_call_s=t_first, build_us=res.time_build_us, solve_us=res.time_solve_us)
    # timings: host-view path and device-resident path
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); eng.solve(view); ts.append(time.perf_counter() - t0)
    rec["host_path_ms_med"] = float(np.median(ts) * 1e3)
    dv = DeviceView(eng, view)
    td = []
    for _ in range(10):
        t0 = time.perf_counter(); r2 = eng.solve(dv); td.append(time.perf_counter() - t0)
    rec["dev_path_ms_med"] = float(np.median(td) * 1e3)
    rec["dev_same"] = bool(np.array_equal(r2.keep, ref["keep"]))
    st = eng.stats(); rec["kernel_ms"] = st["last_device_ms"]; rec["grid"] = st["grid_ctas"]
    rec["build_us"] = r2.time_build_us; rec["solve_us"] = r2.time_solve_us
    dv.free(); eng.close()
    print(json.dumps(rec), flush=True)
    out.append(rec)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)
print("ALL_SAME", all(r["same_bitmask"] and r["dev_same"] and r["cov_same"] and r["slack_same"] for r in out))

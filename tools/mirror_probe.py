"""Loads B c2 windows into one device mirror and runs mss_mirror_solve a few times (for ncu launch lists of the mk_* kernels)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ms_slam_b200 import msgen, engine as E
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = msgen.CONFIGS["c2"]
eng = E.Engine(N=cfg["N"], lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA)
mb = bench.MirrorBatch(eng, E, "c2", list(range(B)), cfg["n_feat"])
for _ in range(steps):
    t0 = time.perf_counter(); rc = mb.step(); dt = time.perf_counter() - t0
    st = mb.mir.stats()
    print(rc, "call ms", dt * 1e3, "build", st["last_build_ms"], "solve", st["last_solve_ms"], flush=True)

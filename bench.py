#!/usr/bin/env python
"""bench.py -- sparsification windows/sec on the KITTI-00-shaped 500 KF x 200k MP window (BASELINE.json metric).

A "step" = one mss_solve_batch over a batch of B independent synthetic c2 windows per GPU (msgen-v1 seeds, in the transport
form FlattenWindow emits: valid slots only, map points in discovery order, MSS_LAYOUT_PACKED16; the batch is larger than L2,
so no flush is needed between steps).  With N GPUs the global batch is N*B windows, window w is solved by rank w % N and the
result slots (keep bitmask + row coverage) are all-gathered with NCCL inside the call.

  value : whole-job windows/s, views and result arrays resident in HBM, CUDA events on the engine's stream, max over ranks
  e2e   : the same call with HOST views (one pinned blob per window) and host result arrays: the H2D of every view (it
          overlaps the solve: the persistent kernel waits per window for a ready flag) + D2H of the results in the timed region
  roofline     : the persistent kernel: algorithmic bytes / its CUDA-event duration, against the
                 measured HBM copy bandwidth (MEASURED_PEAKS.json); traffic = ncu dram bytes of the same launch (profiles/)
  cpu_baseline : the oracle port (HiGHS) timed on this box's host cores on a bounded sample (rank 0, N=1)
  --impl reference : times the reference's CPU algorithm (oracle port: HiGHS, GUROBI is not installable) on the same metric
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sparsification windows/sec (500 KF x 200k MP)"
UNIT = "windows/s"
WORKLOAD = "c2"                    # BASELINE.json configs[1]
C2_KF = 500
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def floor_bytes(K, H, M, F, O):
    """Solver-independent floor of one window = SURVEY 8(d) ALG_BYTES(I=0) without the build terms: the view is read once,
    the result slot is written once."""
    b_in = 4 * (K + 1) + 6 * F + 4 * M + 4 * (M + 1) + 4 * O + 4 * H
    b_out = 4 * ((M + 31) // 32) + 8 * (K + H) + 64
    return b_in + b_out


def view_floor_bytes(v):
    """floor_bytes for a view object in either layout: its arrays read once + the result slot written once"""
    return v.input_bytes() + 4 * ((v.M + 31) // 32) + 8 * (v.K + v.H) + 64


def kernel_alg_bytes(views, results, row_entries, var_visits):
    """Algorithmic bytes of one launch of the implemented kernel, a LOWER bound of what it has to move:
         view arrays read once, observations streamed a second time by the fill pass     6F + 4(K+1) + 8M + 8O + 4H (+ 4(M+1))
         per map point: zeroed counters + seen mark, W2 + W4 passes                      9M + 30M
         per keyframe-row entry: one 64-bit reduction and one CSR write in W1            12Z
         per entry read by a later row phase (device counter): entry + state gather       5 * row_entries
         per map point visited by a later variable phase (device counter)                17 * var_visits
         result slots written once
       not counted: 64-bit reductions of the PROP / GREEDY / D1 row phases, live-list and FREE-list writes."""
    total = 0
    for v, r in zip(views, results):
        # observations are streamed a second time by the fill pass (SoA: their pointers too)
        again = 4 * v.O if hasattr(v, "obs_pairs") else 4 * v.O + 4 * (v.M + 1)
        total += view_floor_bytes(v) + again + 39 * v.M + 12 * int(r.nnz)
    return total + 5 * int(row_entries) + 17 * int(var_visits)


def survey_alg_bytes(K, H, M, F, O, Z, G, I):
    """ALG_BYTES = B_build + I*B_iter + 3*B_iter + B_out with I = rounds actually executed.
    (M = map-point table, Z = incidences, G = occupied cells; formula restated in DESIGN.md section 5.)"""
    b_in = 6 * F + 9 * M + 4 * O + 4 * H
    b_build = b_in + 16 * Z + 4 * M + 4 * (K + G + H + M)
    b_iter = 12 * Z + 8 * M + 8 * (K + G + H)
    b_out = M // 8 + 8 * (K + H) + 16
    return b_build + (I + 3) * b_iter + b_out


# ------------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------------------------------
def cpu_sample(seed=12345, workload="c2", milp=False, time_limit=None):
    """One CPU sample: the reference's model (oracle/ilp_model.py restates MapSparsification.cc:58-157) of one full-size
    window, assembled and solved by HiGHS (GUROBI is not installable here: no network, no licence).  milp=False: the LP
    relaxation only -- the root node every MILP solve (GUROBI upstream) must at least pay, i.e. an optimistic estimate of
    the reference's time; milp=True: the MILP at the reference's MIPGap 0.002 (MapSparsification.cc:155-156)."""
    from ms_slam_b200 import msgen
    from oracle import ilp_model as om
    view, N = msgen.make_config(workload, seed)
    t0 = time.perf_counter()
    if milp:
        sol = om.solve_ilp(view, N, msgen.LAMBDA, msgen.GRID_LAMBDA, time_limit=time_limit)
    else:
        sol = om.solve_lp(view, N, msgen.LAMBDA, msgen.GRID_LAMBDA)
    dt = time.perf_counter() - t0
    return dt, sol


def _cpu_worker(args):
    seed, workload = args
    dt, sol = cpu_sample(seed, workload)
    return dt


def _cpu_milp_worker(args):
    seed, workload, limit = args
    dt, sol = cpu_sample(seed, workload, milp=True, time_limit=limit)
    return dict(workload=workload, seed=seed, seconds=dt, objective=float(sol.objective), status=int(sol.status),
                mip_gap=None if sol.mip_gap is None else float(sol.mip_gap), time_limit_s=limit)


def cpu_parallel_step(pool, seeds, workload="c2"):
    """One CPU step: len(seeds) independent full-size windows solved concurrently, one process (= one HiGHS thread) per
    window -- the way a multi-core host would run the reference over independent windows.  Returns wall seconds."""
    t0 = time.perf_counter()
    list(pool.map(_cpu_worker, [(s, workload) for s in seeds]))
    return time.perf_counter() - t0


def _cpu_pool():
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    procs = max(1, min(os.cpu_count() or 1, 32))
    return ProcessPoolExecutor(max_workers=procs, mp_context=mp.get_context("spawn")), procs


def cpu_baseline_leg():
    """Bounded CPU sample next to the GPU number: every host core solves one full c2 window (500 KF x 200k MP; LP relaxation
    of the reference's model, see cpu_sample) at the same time; plus ONE MILP datapoint -- what GUROBI actually does
    (MapSparsification.cc:153-157) -- on an EuRoC-shaped c3 window (100 KF x 30k MP) under a time limit."""
    pool, procs = _cpu_pool()
    with pool:
        cpu_parallel_step(pool, range(procs), "c1")                     # start the workers (imports, HiGHS start-up)
        milp_f = pool.submit(_cpu_milp_worker, (0, "c3", 100.0))        # alone first: its time is not shared with the LP step
        try:
            milp = milp_f.result(timeout=400)
        except Exception as e:      # noqa: BLE001
            milp = {"error": repr(e)}
        dt = cpu_parallel_step(pool, [12345 + i for i in range(procs)], "c2")
    return {"value": procs / dt, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{procs} c2 windows of 500 KF x 200000 MP solved concurrently, one process per core: assemble + HiGHS LP "
                      f"relaxation only ({dt:.1f} s wall; the reference's GUROBI MILP at MIPGap 0.002 costs at least its root LP)",
            "seconds": dt,
            "milp_datapoint": dict(milp, note="HiGHS MILP at the reference's MIPGap 0.002 on ONE c3 window (100 KF x 30k MP), one core; "
                                              "the same solve at c2 size does not finish within minutes")}


def reference_arm(args, rank):
    """--impl reference: the reference's CPU algorithm on the stated config (full 500 KF x 200k MP windows, not scaled):
    each step solves one window per host core concurrently."""
    if rank != 0:
        return 0
    pool, procs = _cpu_pool()
    times = []
    with pool:
        for w in range(max(args.warmup, 1)):
            cpu_parallel_step(pool, range(procs), "c1")        # warm-up: worker start, imports, HiGHS start-up
        for s in range(args.steps):
            times.append(cpu_parallel_step(pool, [777 + 1000 * s + i for i in range(procs)], "c2"))
    total = float(np.sum(times))
    value = procs * args.steps / total
    sample = (f"each step: {procs} full c2 windows (500 KF x 200000 MP) solved concurrently (one process per core), assemble + HiGHS "
              f"LP relaxation of the reference's ILP")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: 500 KF x 200000 MP KITTI-00-shaped windows, msgen-v1", "windows_per_step": procs,
                       "note": "HiGHS stand-in for GUROBI (not installable: no network/licence); LP relaxation only = optimistic "
                               "(a lower bound of the reference's MILP time)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------------
def numa_bind(local_rank):
    """Pin this process to the CPUs next to its GPU before any pinned allocation, so the pinned blobs land on the GPU's NUMA
    node (first touch) and the N ranks of a box do not pull their H2D traffic through one socket.  Best effort."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to GPU {local_rank}"
    except Exception as e:      # noqa: BLE001
        return f"not bound ({type(e).__name__})"
    return "not bound (single node)"


class Batch:
    """B windows per GPU of one workload, in both memory kinds, with the ctypes arrays of one mss_solve_batch call."""

    def __init__(self, eng, E, workload, B, world, rank, layout="packed16", order="discovery", unsorted=False, host_arm=True):
        from concurrent.futures import ThreadPoolExecutor
        from ms_slam_b200 import msgen, dist as msd
        from ms_slam_b200.window import pack_view
        self.eng, self.E, self.workload = eng, E, workload
        cfg = msgen.CONFIGS[workload]
        self.cfg, self.N = cfg, cfg["N"]
        self.nwin = B * world
        self.mine = msd.local_windows(self.nwin, rank, world)

        def make(w):
            # transport form: what the C++ FlattenWindow emits (empty slots and window-keyframe observations left out; the
            # window, the model and the result are the same -- tests/test_gpu_parity.py::test_compact_view_same_result)
            v = msgen.make_config(workload, seed=w)[0].compact()
            v = v.discovery_order() if order == "discovery" else v
            if layout == "packed16":
                return pack_view(v, tokens16=True)
            return pack_view(v, sort_slots=not unsorted) if layout == "packed" else v
        self.make = make
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:   # numpy releases the GIL in the heavy parts
            self.views = dict(zip(self.mine, pool.map(make, self.mine)))
        K, H, M = cfg["K"], cfg["H"], cfg["M"]
        self.K, self.H, self.M = K, H, M
        self.words, self.rows = (M + 31) // 32, K + H
        self.in_bytes = sum(v.input_bytes() for v in self.views.values())
        nwin = self.nwin
        self.pins = []
        # ---- `value` arm: views resident in HBM; the keep bitmask (+ row coverage) of every OWNED window comes back to a
        #      pinned host buffer inside the timed call; windows of other ranks: header only, their bitmasks
        #      stay in the all-gathered device buffer
        self.dviews = {w: E.DeviceView(eng, self.views[w]) for w in self.mine}
        self.cv = (E.mss_window_view * nwin)()
        self.cr = (E.mss_result * nwin)()
        self.keep_dev_arm = {}
        for w in range(nwin):
            if w in self.dviews:
                self.cv[w] = self.dviews[w].c_view()
                self.cv[w].result_memory = E.RESULT_HOST
                kb = self.pin_zeros(self.words, np.uint32)
                self.keep_dev_arm[w] = kb
                self.cr[w].keep_bits = kb.ctypes.data
            else:
                self.cv[w] = E.mss_window_view(K, H, M, 0, 0, E.MEM_HOST)
        # ---- `e2e` arm: pinned host blobs in, pinned host results out (owned windows)
        self.hv = (E.mss_window_view * nwin)()
        self.hr = (E.mss_result * nwin)()
        self.keep_host_arm = {}
        for w in range(nwin):
            if w in self.views and host_arm:
                v = self.views[w]
                if layout in ("packed", "packed16"):
                    self.hv[w] = E.packed_c_view(v.K, v.H, v.M, v.F, v.O, E.MEM_HOST,
                                                 *self.pin_blob([v.feat_ptr, v.slots, v.mp_nobs16, v.obs_pairs, v.okf_total]),
                                                 tokens16=layout == "packed16", nobs8=bool(v.meta.get("nobs8")))
                else:
                    self.hv[w] = E.mss_window_view(v.K, v.H, v.M, v.F, v.O, E.MEM_HOST,
                                                   *self.pin_blob([v.feat_ptr, v.feat_mp, v.feat_cell, v.mp_nobs, v.mp_obs_ptr,
                                                                   v.mp_obs_kf, v.okf_total]))
                kb = self.pin_zeros(self.words, np.uint32)
                self.keep_host_arm[w] = kb
                self.hr[w].keep_bits = kb.ctypes.data
                self.hr[w].kf_cov = self.pin_zeros(self.rows, np.int32).ctypes.data
                self.hr[w].kf_slack = self.pin_zeros(self.rows, np.int32).ctypes.data
            else:
                self.hv[w] = E.mss_window_view(K, H, M, 0, 0, E.MEM_HOST)

    def pin_zeros(self, n, dtype):
        p = self.eng.pinned((max(n, 1),), dtype)
        p.array[...] = 0
        self.pins.append(p)
        return p.array

    def pin_blob(self, arrays):
        """one pinned blob per window, arrays back to back at 16-byte boundaries (what FlattenWindow lays out): the engine
        moves such a view with a single copy"""
        offs, total = [], 0
        for a in arrays:
            offs.append(total)
            total += (a.nbytes + 15) // 16 * 16
        p = self.eng.pinned((max(total, 16),), np.uint8)
        self.pins.append(p)
        for a, o in zip(arrays, offs):
            p.array[o:o + a.nbytes] = a.view(np.uint8).reshape(-1)
        return [p.array.ctypes.data + o for o in offs]

    def free(self):
        for d in self.dviews.values():
            d.free()
        for p in self.pins:
            p.free()
        self.dviews, self.pins = {}, []


class MirrorBatch:
    """The same B windows per GPU held in ONE persistent device mirror (include/mss.h mss_mirror_*): the maps are
    loaded once, untimed (in a SLAM run they arrive as deltas while the map is built); a step then sends K keyframe handles per
    window and gets a bitmask over map-point handles back."""

    def __init__(self, eng, E, workload, windows, S):
        from concurrent.futures import ThreadPoolExecutor
        from ms_slam_b200 import msgen
        from ms_slam_b200 import mirror as MR
        self.eng, self.MR = eng, MR
        self.mir = MR.Mirror(eng, S)
        self.n = len(windows)
        cfg = msgen.CONFIGS[workload]
        span_kf = cfg["K"] + cfg["H"]

        def make(i):
            # map-point handles in discovery order: the recorder hands them out at first sight, keyframe after keyframe, i.e.
            # in the order the map was built (the view arm numbers its transport form the same way)
            v = msgen.make_config(workload, seed=windows[i])[0].compact().discovery_order()
            return MR.arrays_from_view(v, S=S, seed=windows[i], kf0=i * span_kf, mp0=0, shuffle=False)
        t0 = time.perf_counter()
        self.kfs, mp0 = [], 0
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            for i, L in enumerate(pool.map(make, range(self.n))):
                # map-point handles of window i start where window i-1 ended (disjoint maps = independent windows)
                for key in ("slot_mp", "obs_mp"):
                    L[key] = np.where(L[key] >= 0, L[key] + mp0, -1).astype(np.int32)
                L["mp0"] = mp0
                self.mir.load(L)
                self.kfs.append(np.ascontiguousarray(L["window"], np.int32))
                mp0 += L["n_mp"]
        self.load_s = time.perf_counter() - t0
        st = self.mir.stats()
        words = (st["n_map_points"] + 31) // 32
        self.cw = (MR.mss_mirror_window * self.n)()
        self.cr = (E.mss_result * self.n)()
        self.pins = []
        for i in range(self.n):
            p = eng.pinned((words + 1,), np.uint32)
            p.array[...] = 0
            self.pins.append(p)
            self.cw[i].K, self.cw[i].kf = self.kfs[i].size, self.kfs[i].ctypes.data
            self.cw[i].del_bits, self.cw[i].del_words, self.cw[i].apply = p.array.ctypes.data, words + 1, 0

    def step(self):
        return self.mir.lib.mss_mirror_solve(self.mir.handle, self.n, self.cw, self.cr)

    def deleted_count(self, i):
        w = self.cw[i]
        lo, hi = w.h_lo >> 5, (w.h_hi + 31) >> 5
        return int(np.unpackbits(self.pins[i].array[lo:hi].view(np.uint8)).sum())

    def free(self):
        self.mir.close()
        for p in self.pins:
            p.free()
        self.pins = []


def roofline_of(batch, kern_ms, st, peak):
    """roofline numbers of one launch of the persistent kernel over `batch` (this rank's windows)"""
    views, cr, mine = batch.views, batch.cr, batch.mine
    fb = sum(view_floor_bytes(v) for v in views.values())
    sb = sum(survey_alg_bytes(views[w].K, views[w].H, views[w].M, views[w].F, views[w].O, cr[w].nnz, cr[w].n_cells, cr[w].rounds)
             for w in mine)
    s0 = sum(survey_alg_bytes(views[w].K, views[w].H, views[w].M, views[w].F, views[w].O, cr[w].nnz, cr[w].n_cells, 0) for w in mine)
    ab = kernel_alg_bytes([views[w] for w in mine], [cr[w] for w in mine], st["last_row_entries"], st["last_var_visits"])
    sec = kern_ms * 1e-3
    rounds = [int(cr[w].rounds) for w in mine]
    return {"bound": "hbm", "achieved": ab / sec / 1e9, "peak": peak, "unit": "GB/s", "frac": ab / sec / 1e9 / peak,
            "traffic": None, "kernel": "mss_persistent_kernel", "kernel_ms": kern_ms, "windows_per_launch": len(mine),
            "alg_bytes_per_launch": ab,
            "alg_bytes_definition": "DESIGN.md section 4 (lower bound of what THIS kernel moves): views read once + outside "
                                    "observations streamed twice + 39 B/map point + 12 B/incidence (build) + 5 B per entry read by "
                                    f"later row phases ({int(st['last_row_entries'])} entries, device counter) + 17 B per later "
                                    f"variable visit ({int(st['last_var_visits'])}) + result slots",
            "survey_floor_I0": {"bytes_per_launch": s0, "achieved": s0 / sec / 1e9, "frac": s0 / sec / 1e9 / peak,
                                "ms_at_peak": s0 / (peak * 1e9) * 1e3,
                                "definition": "SURVEY 8(d) ALG_BYTES(I=0) = B_build + 3*B_iter + B_out: the solver-independent "
                                              "compulsory traffic of build + round/repair/verify; comparable across rounds"},
            "survey_formula": {"bytes_per_launch": sb, "achieved": sb / sec / 1e9,
                               "definition": "SURVEY 8(d) planning formula B_build + (I+3)*B_iter + B_out with I = rounds executed "
                                             f"(mean {float(np.mean(rounds)):.1f}, max {max(rounds)}); it assumes every round "
                                             "streams the whole CSR, which this work-efficient kernel does not do"},
            "io_floor": {"bytes_per_launch": fb, "achieved": fb / sec / 1e9, "frac": fb / sec / 1e9 / peak,
                         "definition": "every view read once + result slots written once"}}


def host_api_leg(N, lam, glam):
    """What the reference-facing C++ API costs around the solve: ORB_SLAM3::MapSparsification (libmss_host.so) over a
    c2-sized final flush and over a live 30-keyframe window driven through Run(): flatten_ms (pointer graph -> view, host),
    solve_ms (C-ABI call incl. copies), apply_ms (SetBadFlag fan-out + forwarding, host)."""
    from ms_slam_b200 import msgen
    from ms_slam_b200.host_mirror import World
    out = {}
    keys = ("K", "H", "M", "n_vars", "n_kept", "n_deleted", "rounds", "components", "flatten_ms", "solve_ms", "apply_ms", "status",
            "mirror", "delta_ops", "build_ms", "h2d_bytes", "d2h_bytes")
    for label, name, seed, kw, over in (("c2_flush", "c2", 0, {}, {}),
                                        ("live_windows", "live", 0, dict(window_length=30), dict(K=60, M=12000))):
        view, Nw = msgen.make_config(name, seed, **over)
        for mode, mkw in (("device_mirror", dict(mirror=True)), ("flatten_batched_handback", dict(mirror=False)),
                          ("flatten_per_point_setbadflag", dict(mirror=False, batched_handback=False))):
            try:
                w = World(view, N=Nw, lam=lam, grid_lam=glam, **kw, **mkw)
                w.start()
                if label == "live_windows":         # two successive 30-keyframe windows: the second one shows the steady state
                    w.feed(0, 30)                   # LocalMapping's hook: > 10 queued keyframes trigger a window
                    w.wait_forwarded(30, 60000)
                    w.feed(30, 30)
                    w.wait_forwarded(60, 60000)
                w.finish(120000)
                reps = [r for r in w.reports() if r["K"] > 0]
                if label == "live_windows":
                    out.setdefault(label, {})[mode] = {f"window_{i + 1}": {k: r.get(k) for k in keys} for i, r in enumerate(reps[:2])}
                else:
                    r = max(reps, key=lambda x: x["K"]) if reps else {}
                    out.setdefault(label, {})[mode] = {k: r.get(k) for k in keys}
                w.close()
            except Exception as e:      # noqa: BLE001
                out.setdefault(label, {})[mode] = {"error": repr(e)}
    out["what"] = ("ORB_SLAM3::MapSparsification (libmss_host.so) driven like System / LocalMapping drive it; ms per window on the "
                   "host: flatten_ms = pointer graph -> view (flatten modes) or drain of the recorded deltas into the device mirror "
                   "(device_mirror; the first window also uploads the whole map), solve_ms = the C-ABI call incl. copies, apply_ms = "
                   "hand-back into the map + forwarding; flatten_per_point_setbadflag is the reference's own hand-back loop")
    return out


def after_path_leg(eng, E, torch, local_rank, peak):
    """The two steps right behind the path, on the device: compaction of the per-keypoint arrays of the
    keyframes of one c2-sized window (KeyFrame::EraseBadDescriptor, 68 B per row: pure data movement, HBM roofline) and the BoW
    re-transform of the surviving descriptors against an ORBvoc-shaped vocabulary (k = 10, L = 6: 1.1 M nodes, 35 MB,
    L2-resident).  Timed with CUDA events on the engine's stream around the C-ABI call (it includes the small descriptor
    upload and the read-back of the row counts / vectors)."""
    from ms_slam_b200 import bow as BW
    from ms_slam_b200.mirror import KeyframePayload, compact_keyframes
    out = {}
    stream = torch.cuda.ExternalStream(eng.lib.mss_stream(eng.handle), device=torch.device("cuda", local_rank))
    rng = np.random.default_rng(0)
    nkf, rows, frac = 500, 2000, 0.15                                    # c2: 500 keyframes x 2000 keypoints, ~15 % of the points survive
    def payloads():
        ps = []
        for _ in range(nkf):
            ps.append(KeyframePayload(eng, rows, rng.random(rows) < frac, rng.integers(0, 256, size=(rows, 32), dtype=np.uint8),
                                      rng.integers(0, 2**31, size=(rows, 7), dtype=np.int64).astype(np.uint32),
                                      rng.random(rows, dtype=np.float32), rng.random(rows, dtype=np.float32)))
        return ps
    try:
        times, ktimes, left = [], [], None
        for it in range(4):                                              # in place: fresh arrays every time (first one = warm-up)
            ps = payloads()
            from ms_slam_b200 import mirror as MR
            MR._declare(eng.lib)
            carr = (MR.mss_kf_payload * nkf)(*[p.c_struct() for p in ps])       # (marshalling outside the timed call)
            left = np.zeros(nkf, np.int32)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = eng.lib.mss_compact_keyframes(eng.handle, nkf, carr, left.ctypes.data)
            e1.record(stream)
            e1.synchronize()
            assert rc == 0
            times.append(e0.elapsed_time(e1))
            ktimes.append(eng.stats()["last_device_ms"])
            if it < 3:
                for p in ps:
                    p.free()
        ms = float(np.median(times[1:]))
        moved = nkf * rows * (68 + 1) + int(left.sum()) * 68             # every row and its flag read once, survivors written once
        kms = float(np.median(ktimes[1:]))
        out["compaction"] = {"keyframes": nkf, "rows_per_keyframe": rows, "rows_left": int(left.sum()), "call_ms": ms, "kernel_ms": kms,
                             "algorithmic_bytes": moved, "roofline": {"bound": "hbm", "achieved": moved / kms / 1e6, "peak": peak, "unit": "GB/s",
                                                                      "frac": moved / kms / 1e6 / peak},
                             "what": "mss_compact_keyframes over the keyframes of one c2-sized window, one launch (kernel_ms: CUDA events around "
                                     "the kernel inside the library); the call also uploads the 500 array descriptors and reads the row counts back"}
        # ---- BoW re-transform of the survivors ---------------------------------------------------------------------------------
        voc = BW.synthetic_vocabulary(k=10, L=6, seed=0, ragged=False)
        V = BW.Vocabulary(eng, voc)
        counts = [int(x) for x in left]
        ptrs = [p.ptr["descriptors"] for p in ps]
        times = []
        barr = (BW.mss_bow_keyframe * nkf)()
        keep_alive, bk = [], []
        for q in range(nkf):                                                 # host arrays for every output, marshalled once
            n = max(counts[q], 1)
            o = [np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64), np.zeros(1, np.int32),
                 np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(1, np.int32)]
            keep_alive.append(o)
            barr[q] = BW.mss_bow_keyframe(counts[q], q, ptrs[q], *[a.ctypes.data for a in o])
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = eng.lib.mss_bow_transform(V.handle, nkf, barr, 4, 1 if it == 3 else 0)
            e1.record(stream)
            e1.synchronize()
            assert rc == 0
            times.append(e0.elapsed_time(e1))
            bk.append(eng.stats()["last_device_ms"])
        ms = float(np.median(times[1:]))
        nd = int(sum(counts))
        out["bow_transform"] = {"keyframes": nkf, "descriptors": nd, "vocabulary_nodes": int(voc["parent"].size), "call_ms": ms, "kernels_ms": float(np.median(bk[1:])),
                                "descriptors_per_s": nd / (ms * 1e-3), "node_descriptor_bytes_read": nd * 6 * 10 * 32,
                                "achieved_gbs_from_l2": nd * 6 * 10 * 32 / ms / 1e6,
                                "what": "mss_bow_transform: tree descent (6 levels x 10 children x 32 B per descriptor, from L2) + per-keyframe "
                                        "BowVector / FeatureVector + read-back of the vectors to the host"}
        V.close()
        for p in ps:
            p.free()
    except Exception as e:      # noqa: BLE001
        out["error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="c2 windows per GPU per step")
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mirror", action="store_true", help="skip the persistent-device-mirror arm")
    ap.add_argument("--e2e-handles", type=int, default=2, help="handles (host threads) taking alternate steps in the e2e arm")
    ap.add_argument("--no-extras", action="store_true", help="skip latency / host-API / c3 / c5 sub-entries (N=1, rank 0)")
    ap.add_argument("--order", default="discovery", choices=["generated", "discovery"],
                    help="map-point numbering of the synthetic windows: FlattenWindow's discovery order "
                         "(mnIndexForSparsification, MapSparsification.cc:91-99) or as msgen draws them (random)")
    ap.add_argument("--unsorted-slots", action="store_true", help="packed layout: keep the slots of a keyframe in slot order")
    ap.add_argument("--layout", default="packed16", choices=["packed16", "packed", "soa"],
                    help="transport layout of the views (include/mss.h mss_layout)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_arm(args, rank)

    import torch
    import torch.distributed as dist
    from ms_slam_b200 import msgen, dist as msd
    from ms_slam_b200 import engine as E

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = numa_bind(local_rank) if world > 1 else "single process"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = msgen.CONFIGS[args.workload]
    N = cfg["N"]
    eng = E.Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local_rank)
    if world > 1:
        uid = msd.broadcast_unique_id(eng, rank)
        eng.comm_init(uid, rank, world)

    B = args.batch
    batch = Batch(eng, E, args.workload, B, world, rank, args.layout, args.order, args.unsorted_slots, host_arm=not args.no_e2e)
    nwin, mine, views, K, H, M = batch.nwin, batch.mine, batch.views, batch.K, batch.H, batch.M
    stream = torch.cuda.ExternalStream(eng.lib.mss_stream(eng.handle), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(cviews, cres, steps, warmup, n=None, e=None, stream=stream):
        e, n = e or eng, n or nwin
        for _ in range(warmup):
            rc = e.solve_batch_raw(cviews, cres, n)
            assert rc == 0, e.lib.mss_last_error(e.handle)
        barrier()
        l0 = e.stats()["kernel_launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kern_ms = []
        e0.record(stream)
        for _ in range(steps):
            rc = e.solve_batch_raw(cviews, cres, n)
            kern_ms.append(e.stats()["last_device_ms"])
        e1.record(stream)
        barrier()
        assert rc == 0, e.lib.mss_last_error(e.handle)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(np.mean(kern_ms)), e.stats()["kernel_launches"] - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, kern_ms, launches = run(batch.cv, batch.cr, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    st_dev = eng.stats()
    e2e = None
    if not args.no_e2e:
        ms_host, _, _ = run(batch.hv, batch.hr, args.steps, args.warmup)
        st_host = eng.stats()
        single = {"value": nwin * args.steps / (ms_host * 1e-3), "unit": UNIT, "ms_per_step": ms_host / args.steps,
                  "what": "one handle, one thread: every call waits for the solve of the last windows to arrive and for its results"}
        # ---- the same calls from TWO handles on two host threads (a handle is single-threaded, like GRBEnv upstream): the
        #      staging copies of one batch travel while the other batch is being solved (one FIFO staging stream per device),
        #      so PCIe stays busy; every step still copies its inputs up and its results down inside the timed region ----------
        import threading
        nh = max(2, args.e2e_handles)
        extra, lanes, keep_x = [], [(eng, batch.hr)], []
        for _ in range(nh - 1):
            eng_b = E.Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local_rank)
            if world > 1:
                eng_b.comm_init(msd.broadcast_unique_id(eng_b, rank), rank, world)
            hr_b = (E.mss_result * nwin)()
            keep_b = {}
            for w in range(nwin):
                if w in batch.views:
                    keep_b[w] = batch.pin_zeros(batch.words, np.uint32)
                    hr_b[w].keep_bits = keep_b[w].ctypes.data
                    hr_b[w].kf_cov = batch.pin_zeros(batch.rows, np.int32).ctypes.data
                    hr_b[w].kf_slack = batch.pin_zeros(batch.rows, np.int32).ctypes.data
            extra.append(eng_b); lanes.append((eng_b, hr_b)); keep_x.append(keep_b)
        rcs = [0] * nh

        def worker(t, steps):
            e, res = lanes[t]
            for _ in range(steps):
                rc = e.solve_batch_raw(batch.hv, res, nwin)
                if rc != 0:
                    rcs[t] = rc
                    return
        for e, res in lanes:                                             # warm-up of every handle (arenas, staging buffers)
            for _ in range(args.warmup):
                assert e.solve_batch_raw(batch.hv, res, nwin) == 0, e.lib.mss_last_error(e.handle)
        steps_each = (args.steps + nh - 1) // nh
        barrier()
        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(t, steps_each)) for t in range(nh)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        torch.cuda.synchronize()
        ms_pipe = (time.perf_counter() - t0) * 1e3
        barrier()
        assert rcs == [0] * nh, [e.lib.mss_last_error(e.handle) for e, _ in lanes]
        if world > 1:
            t = torch.tensor([ms_pipe], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_pipe = float(t.item())
        keep_b = keep_x[0]
        eng_b = extra[0]
        for w in mine:
            assert np.array_equal(keep_b[w], batch.keep_host_arm[w]), f"window {w}: the two handles disagree"
        st_b = eng_b.stats()
        e2e = {"value": nwin * nh * steps_each / (ms_pipe * 1e-3), "unit": UNIT, "ms_per_step": ms_pipe / (nh * steps_each),
               "steps": nh * steps_each, "handles": nh,
               "h2d_bytes_per_step": int(st_host["last_h2d_bytes"]), "d2h_bytes_per_step": int(st_host["last_d2h_bytes"]),
               "timing": "host wall clock around all steps, device synchronised on both sides, max over ranks",
               "single_call": single,
               "what": "mss_solve_batch with one pinned host blob per window in, pinned host keep bits + row coverage of the "
                       f"owned windows out, every step; {nh} handles on {nh} host threads take the steps in turn, so the copies of "
                       "one step overlap the solve of the other (h2d / d2h bytes: per step, from the handle's own counters "
                       f"{int(st_b['last_h2d_bytes'])} / {int(st_b['last_d2h_bytes'])})"}
        for e in extra:
            e.close()

    # ---- the same windows from the persistent device mirror: K keyframe handles up, deleted-handle bitmask down -------------
    e2e_mirror = None
    if not args.no_e2e and not args.no_mirror:
        try:
            # (a mirror belongs to one device: with several ranks every rank keeps the maps of its own windows in its own
            # mirror, on a handle without a communicator -- the deleted-handle bitmasks are consumed by the host that owns them)
            eng_m = eng if world == 1 else E.Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local_rank)
            mb = MirrorBatch(eng_m, E, args.workload, list(mine), cfg["n_feat"])
            for _ in range(args.warmup):
                rc = mb.step()
                assert rc == 0, eng_m.lib.mss_last_error(eng_m.handle)
            barrier()
            stream_m = torch.cuda.ExternalStream(eng_m.lib.mss_stream(eng_m.handle), device=torch.device("cuda", local_rank))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            bms, sms = [], []
            e0.record(stream_m)
            for _ in range(args.steps):
                rc = mb.step()
                st_m = mb.mir.stats()
                bms.append(st_m["last_build_ms"]); sms.append(st_m["last_solve_ms"])
            e1.record(stream_m)
            barrier()
            assert rc == 0, eng_m.lib.mss_last_error(eng_m.handle)
            ms_m = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms_m], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_m = float(t.item())
            same = all(mb.cr[i].objective == batch.cr[w].objective and mb.cr[i].n_kept == batch.cr[w].n_kept and
                       mb.deleted_count(i) == batch.cr[w].n_vars - batch.cr[w].n_kept for i, w in enumerate(mine))
            assert same, "mirror arm and view arm disagree"
            e2e_mirror = {"value": nwin * args.steps / (ms_m * 1e-3), "unit": UNIT, "ms_per_step": ms_m / args.steps,
                          "h2d_bytes_per_step": int(st_m["last_h2d_bytes"]), "d2h_bytes_per_step": int(st_m["last_d2h_bytes"]),
                          "assembly_ms_per_step": float(np.mean(bms)), "solve_kernel_ms_per_step": float(np.mean(sms)),
                          "mirror_device_bytes": int(st_m["device_bytes"]), "map_load_s_untimed": mb.load_s,
                          "same_result_as_view_arm": True,
                          "what": "mss_mirror_solve: the maps of all windows live in one persistent device mirror (loaded once, "
                                  "untimed: in a SLAM run they arrive as deltas); per step K keyframe handles per window go up, the view "
                                  "is assembled on the device, the deleted-map-point bitmask comes back"}
            mb.free()
            if eng_m is not eng:
                eng_m.close()
        except Exception as e:      # noqa: BLE001
            e2e_mirror = {"error": repr(e)}

    # ---- what was timed is right: every owned window solved; both arms returned the same bitmask -----------------------
    cr = batch.cr
    for w in mine:
        assert cr[w].status == 0 and cr[w].n_kept > 0 and cr[w].rounds > 0
        if not args.no_e2e:
            if not np.array_equal(batch.keep_dev_arm[w], batch.keep_host_arm[w]):
                ref = eng.solve(batch.views[w]).keep_bits
                raise AssertionError(f"window {w}: device-view and host-view arms differ (value arm == fresh solve: "
                                     f"{np.array_equal(batch.keep_dev_arm[w], ref)}, e2e arm == fresh solve: {np.array_equal(batch.keep_host_arm[w], ref)})")
    r0 = cr[mine[0]]
    quality = {"objective_w0": r0.objective, "kept_w0": r0.n_kept, "vars_w0": r0.n_vars, "rounds_w0": r0.rounds}

    # ---- cross-rank parity: the all-gathered bitmasks of windows solved by OTHER ranks, fetched to the host in one extra call,
    #      against a local re-solve of the same windows on this GPU without a communicator (>= 8 foreign windows per rank)
    cross = None
    if world > 1:
        av = (E.mss_window_view * nwin)()
        ar = (E.mss_result * nwin)()
        keep_all = {}
        for w in range(nwin):
            av[w] = batch.cv[w]
            if w not in batch.dviews:
                av[w].result_memory = E.RESULT_HOST
            keep_all[w] = batch.pin_zeros(batch.words, np.uint32)
            ar[w].keep_bits = keep_all[w].ctypes.data
        rc = eng.solve_batch_raw(av, ar, nwin)
        assert rc == 0, eng.lib.mss_last_error(eng.handle)
        foreign = [w for w in range(nwin) if w not in batch.dviews]
        pick = foreign[rank::max(1, len(foreign) // 8)][:8] if foreign else []
        solo = E.Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local_rank)
        ok = True
        for w in pick:
            r = solo.solve(batch.make(w))
            ok = ok and np.array_equal(r.keep_bits, keep_all[w]) and r.objective == ar[w].objective and r.n_kept == ar[w].n_kept
        for w in mine:
            ok = ok and np.array_equal(keep_all[w], batch.keep_dev_arm[w])
        solo.close()
        flag = torch.tensor([1 if ok else 0, len(pick)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        cross = bool(flag[0].item())
        if not cross:
            raise SystemExit(f"bench.py: rank {rank}: an all-gathered bitmask differs from the local re-solve")
        cross_n = int(flag[1].item())

    if rank == 0:
        peak, peak_src = peaks()
        roof = roofline_of(batch, kern_ms, st_dev, peak)
        roof["peak_source"] = peak_src
        roof["note"] = ("the solve is bound by dependent latency (spread-address state gathers and 64-bit reductions per undecided "
                        "entry per round, block and group barriers), not by DRAM bandwidth; `traffic` is the ncu dram read+write "
                        "of the same launch (profiles/)")
        line = {
            "metric": METRIC, "value": nwin * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {K} KF x {M} MP KITTI-00-shaped windows (msgen-v1 seeds 0..{nwin-1}) in the transport "
                                   "form FlattenWindow emits: valid slots + outside observations only, map points numbered "
                                   + ("in discovery order (mnIndexForSparsification)" if args.order == "discovery" else "as generated (random)")
                                   + f", {args.layout} layout" + (", one-byte nObs table" if any(v.meta.get("nobs8") for v in batch.views.values()) else ""),
                       "windows_per_gpu_per_step": B, "global_windows_per_step": nwin,
                       "N": N, "lambda": msgen.LAMBDA, "grid_lambda": msgen.GRID_LAMBDA,
                       "parallelism": f"window w -> rank w % {world}; one NCCL all-gather of result slots" if world > 1 else "single GPU",
                       "value_includes": "views resident in HBM; D2H of the owned windows' keep bitmask + row coverage inside the timed call",
                       "l2": f"per-step inputs {batch.in_bytes/1e6:.0f} MB per GPU > 126 MB L2: no flush needed" if batch.in_bytes > 126e6
                             else f"per-step inputs {batch.in_bytes/1e6:.0f} MB per GPU (< L2)",
                       "kernel_ms_per_step": kern_ms, "grid_ctas": st_dev["grid_ctas"], "quality": quality, "numa": numa},
            "e2e": e2e,
            "e2e_mirror": e2e_mirror,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        if cross is not None:
            line["cross_rank_parity"] = cross
            line["cross_rank_windows_checked_per_rank"] = cross_n
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                line["roofline"]["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_extras:
            # ---- single-window latency (what the live SLAM thread sees): device-resident view, kernel time of one launch ----
            lat = {}
            for name in ("c2", "live"):
                from ms_slam_b200.window import pack_view
                vv = pack_view(msgen.make_config(name, 0)[0].compact().discovery_order(), tokens16=True)
                e1 = E.Engine(N=msgen.CONFIGS[name]["N"], lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local_rank)
                dv = E.DeviceView(e1, vv)
                ks, hs = [], []
                for i in range(25):
                    t0 = time.perf_counter()
                    e1.solve(dv)
                    hs.append((time.perf_counter() - t0) * 1e3)
                    ks.append(e1.stats()["last_device_ms"])
                lat[name] = {"kernel_us_median": float(np.median(ks[5:]) * 1e3), "call_us_median": float(np.median(hs[5:]) * 1e3),
                             "grid_ctas": e1.stats()["grid_ctas"]}
                dv.free(); e1.close()
            line["single_window_latency"] = lat
            # ---- every window certified on the device: lower bound of the reference ILP (mss_set_dual_bound) -------------------
            try:
                eng.set_dual_bound(True)
                _, k_on, _ = run(batch.cv, batch.cr, 5, 2)
                gaps = [batch.cr[w].objective / batch.cr[w].dual_bound - 1.0 for w in mine]
                same = all(np.array_equal(batch.keep_dev_arm[w], batch.keep_host_arm[w]) for w in mine) if not args.no_e2e else None
                line["certificate"] = {"kernel_ms_per_step": k_on, "cost_vs_plain": k_on / kern_ms - 1.0, "gap_max": float(max(gaps)),
                                       "gap_mean": float(np.mean(gaps)), "windows": len(gaps), "all_within_1pct": bool(max(gaps) <= 0.01),
                                       "same_selection": same,
                                       "what": "objective / dual_bound - 1 per window, dual_bound proven on the device (csrc/mss_bound.cuh): "
                                               "a certified optimality gap against the reference ILP, no CPU solver involved"}
                eng.set_dual_bound(False)
            except Exception as e:      # noqa: BLE001
                line["certificate"] = {"error": repr(e)}
            # ---- the other single-GPU configs: roofline sub-entries (c3 EuRoC-shaped; c5 4Seasons-shaped, HBM-bound) ----------
            batch.free()
            sub = {}
            for name, nb, steps in (("c3", 512, 10), ("c5", 16, 5)):
                try:
                    e1 = E.Engine(N=msgen.CONFIGS[name]["N"], lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local_rank)
                    b1 = Batch(e1, E, name, nb, 1, 0, args.layout, args.order, args.unsorted_slots, host_arm=False)
                    s1 = torch.cuda.ExternalStream(e1.lib.mss_stream(e1.handle), device=torch.device("cuda", local_rank))
                    ms1, k1, _ = run(b1.cv, b1.cr, steps, 3, n=b1.nwin, e=e1, stream=s1)
                    r1 = roofline_of(b1, k1, e1.stats(), peak)
                    sub[name] = {"windows_per_s": b1.nwin * steps / (ms1 * 1e-3), "ms_per_step": ms1 / steps, "windows_per_step": b1.nwin,
                                 "inputs_mb_per_step": b1.in_bytes / 1e6,
                                 "roofline": {k: r1[k] for k in ("achieved", "frac", "kernel_ms", "alg_bytes_per_launch", "survey_floor_I0", "io_floor")}}
                    b1.free(); e1.close()
                except Exception as e:      # noqa: BLE001
                    sub[name] = {"error": repr(e)}
            line["other_configs"] = sub
            # ---- the reference-facing C++ API around the solve -----------------------------------------------------------------
            line["host_api"] = host_api_leg(N, msgen.LAMBDA, msgen.GRID_LAMBDA)
            # ---- the steps right behind the path, on the device (f2 compaction, f4 BoW re-transform) --------------------------
            e2 = E.Engine(N=N, lam=msgen.LAMBDA, grid_lam=msgen.GRID_LAMBDA, device=local_rank)
            line["after_path"] = after_path_leg(e2, E, torch, local_rank, peak)
            e2.close()
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

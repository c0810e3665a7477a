"""world_size-2 gloo test of the multi-GPU host logic: window -> rank sharding, rank-major result slots and the single
all-gather that reassembles every window's result on every rank (the device path does the same with NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, LAM, GLAM


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ms_slam_b200 import msgen, dist as msd
    from oracle import emulate as em                      # stand-in solver for the CPU-only box (test infrastructure)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        nwin = 5                                              # ragged: rank 0 owns 3 windows, rank 1 owns 2
        views = [msgen.make_config("live", 100 + w, M=800 + 150 * w, H=4 if w % 2 else 0)[0] for w in range(nwin)]
        shapes = [(v.K, v.H, v.M) for v in views]
        layout = msd.SlotLayout(shapes, world)
        local = np.zeros(layout.words_per_rank, np.uint32)
        full_local = np.zeros(layout.total_words, np.uint32)
        for w in msd.local_windows(nwin, rank, world):
            assert msd.owner(w, world) == rank
            r = em.solve(views[w], 100, LAM, GLAM)
            hdr = dict(status=0, rounds=r["rounds"], n_max=r["n_max"], n_vars=r["n_vars"], n_cells=r["n_cells"], nnz=r["nnz"],
                       n_kept=r["n_kept"], uncovered=r["uncovered"], total_slack=int(r["slack"].sum()), sum_cost=r["sum_cost"])
            layout.pack(full_local, w, hdr, em.pack_bits(r["keep"]), r["cov"], r["slack"])
        local[:] = full_local[rank * layout.words_per_rank:(rank + 1) * layout.words_per_rank]
        gathered = msd.allgather_slots(layout, local)
        out = []
        for w in range(nwin):
            u = layout.unpack(gathered, w, LAM, GLAM)
            out.append((u["objective"], u["n_kept"], u["keep_bits"].tobytes(), u["cov"].tobytes()))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_allgather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]                                  # every rank holds every window's result
    # and it equals the single-process solve of each window
    from ms_slam_b200 import msgen
    from oracle import emulate as em
    for w, (obj, n_kept, bits, cov) in enumerate(got[0]):
        v = msgen.make_config("live", 100 + w, M=800 + 150 * w, H=4 if w % 2 else 0)[0]
        r = em.solve(v, 100, LAM, GLAM)
        assert obj == r["objective"] and n_kept == r["n_kept"]
        assert bits == em.pack_bits(r["keep"]).tobytes()
        assert cov == r["cov"].astype(np.int32).tobytes()


def test_slot_layout_arithmetic():
    from ms_slam_b200 import dist as msd
    lay = msd.SlotLayout([(10, 2, 100), (20, 0, 1000), (5, 5, 33)], nranks=2)
    assert lay.spr == 2 and lay.slot_stride == msd.slot_words(20, 0, 1000)
    offs = [lay.offset(w) for w in range(3)]
    assert offs == [0, 2 * lay.slot_stride, lay.slot_stride]
    assert msd.local_windows(7, 1, 3) == [1, 4] and msd.owner(5, 4) == 1

"""Multi-GPU sharding of independent windows.

Windows drawn from disjoint covisibility components share no map point, so they are solved independently:
window ``w`` belongs to rank ``w % nranks``.  The only data that crosses GPUs is the per-window *result slot*
(header + keep-bitmask + per-row coverage), exchanged with ONE regular all-gather: every rank owns
``spr = ceil(nwin / nranks)`` slots of ``slot_stride`` uint32 words in a rank-major buffer.

libmss.so does this on device (result slots are written by the kernel straight into the rank's send region and
all-gathered in place with NCCL, csrc/mss_engine.cu).  This module is the same layout arithmetic on the host, used by
bench.py to set the communicator up through torch.distributed and by the world_size-2 gloo tests.
"""
from __future__ import annotations

import numpy as np

HDR_WORDS = 16                 # kHdrWords in csrc/mss_kernels.cuh
SLOT_MAGIC = 0x4D535331        # "MSS1"


def owner(w: int, nranks: int) -> int:
    return w % nranks


def local_windows(nwin: int, rank: int, nranks: int):
    return list(range(rank, nwin, nranks))


def slot_words(K: int, H: int, M: int) -> int:
    return HDR_WORDS + (M + 31) // 32 + 2 * (K + H)


class SlotLayout:
    """Rank-major slot table of one batch: window w -> rank w % nranks, slot w // nranks."""

    def __init__(self, shapes, nranks: int):
        """shapes: list of (K, H, M) for every window of the batch (known to every rank)."""
        self.shapes = list(shapes)
        self.nwin = len(self.shapes)
        self.nranks = nranks
        self.spr = (self.nwin + nranks - 1) // nranks
        self.slot_stride = max([slot_words(*s) for s in self.shapes], default=HDR_WORDS)
        self.words_per_rank = self.spr * self.slot_stride
        self.total_words = self.words_per_rank * nranks

    def offset(self, w: int) -> int:
        return ((w % self.nranks) * self.spr + w // self.nranks) * self.slot_stride

    def pack(self, buf: np.ndarray, w: int, header: dict, keep_bits, cov, slack) -> None:
        K, H, M = self.shapes[w]
        o = self.offset(w)
        words = (M + 31) // 32
        hdr = np.zeros(HDR_WORDS, np.uint32)
        sc = int(header.get("sum_cost", 0))
        hdr[0] = np.uint32(header.get("status", 0) & 0xFFFFFFFF)
        hdr[1], hdr[2], hdr[3] = header.get("rounds", 0), header.get("n_max", 0), header.get("n_vars", 0)
        hdr[4], hdr[5], hdr[6] = header.get("n_cells", 0), header.get("nnz", 0), header.get("n_kept", 0)
        hdr[7], hdr[8] = header.get("uncovered", 0), header.get("total_slack", 0)
        hdr[9], hdr[10] = sc & 0xFFFFFFFF, sc >> 32
        hdr[14], hdr[15] = SLOT_MAGIC, w
        buf[o:o + HDR_WORDS] = hdr
        buf[o + HDR_WORDS:o + HDR_WORDS + words] = np.asarray(keep_bits, np.uint32)
        r0 = o + HDR_WORDS + words
        buf[r0:r0 + K + H] = np.asarray(cov).astype(np.uint32)
        buf[r0 + K + H:r0 + 2 * (K + H)] = np.asarray(slack).astype(np.uint32)

    def unpack(self, buf: np.ndarray, w: int, lam: float, grid_lam: float) -> dict:
        K, H, M = self.shapes[w]
        o = self.offset(w)
        words = (M + 31) // 32
        hdr = buf[o:o + HDR_WORDS]
        if int(hdr[14]) != SLOT_MAGIC:
            raise ValueError(f"slot of window {w} was not written")
        r0 = o + HDR_WORDS + words
        sum_cost = (int(hdr[10]) << 32) | int(hdr[9])
        return dict(status=int(np.int32(hdr[0])), rounds=int(hdr[1]), n_max=int(hdr[2]), n_vars=int(hdr[3]),
                    n_cells=int(hdr[4]), nnz=int(hdr[5]), n_kept=int(hdr[6]), uncovered=int(hdr[7]),
                    total_slack=int(hdr[8]), sum_cost=sum_cost,
                    objective=float(sum_cost) + grid_lam * int(hdr[7]) + lam * int(hdr[8]),
                    keep_bits=buf[o + HDR_WORDS:o + HDR_WORDS + words].copy(),
                    cov=buf[r0:r0 + K + H].astype(np.int32), slack=buf[r0 + K + H:r0 + 2 * (K + H)].astype(np.int32))


def allgather_slots(layout: SlotLayout, local_buf: np.ndarray, group=None) -> np.ndarray:
    """All-gather the per-rank slot regions with torch.distributed (any backend); returns the full rank-major buffer."""
    import torch
    import torch.distributed as dist
    assert local_buf.size == layout.words_per_rank
    t = torch.from_numpy(local_buf.view(np.int32).copy())
    outs = [torch.empty_like(t) for _ in range(layout.nranks)]
    dist.all_gather(outs, t, group=group)
    return torch.cat(outs).numpy().view(np.uint32)


def broadcast_unique_id(engine, rank: int, src: int = 0, group=None) -> bytes:
    """Rank `src` creates the NCCL unique id inside libmss; everybody receives the 128 bytes."""
    import torch
    import torch.distributed as dist
    if rank == src:
        uid = engine.unique_id()
        t = torch.tensor(list(uid), dtype=torch.uint8)
    else:
        t = torch.zeros(128, dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())

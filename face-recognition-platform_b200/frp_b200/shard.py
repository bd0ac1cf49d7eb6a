"""Row-sharded gallery: one process per GPU, contiguous blocks of the global
insertion sequence per rank, queries replicated, per-shard top-k merged after one
all-gather (NCCL over NVLink/NVSwitch).  The reference is single-process; this is
the scaling layer of SURVEY.md section 8(e).

Data path per step (all on the rank's stream):
    frp_match_device (local threshold + top-k, exact)          -> packed [B,k] lists
    torch.distributed.all_gather_into_tensor (B*k*16 + 8 B/rank)
    frp_merge_topk_device (stable merge by (score, global row)) -> [B,k]
The merge is deterministic and identical on every rank, so every rank holds the
global result (no broadcast back).

``local_match`` / ``merge`` are injectable so the rendezvous, partitioning and
gather plumbing can be exercised with the gloo backend on CPU in the tests.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def record_bytes(B: int, k: int) -> int:
    """Size of one rank's packed result record (include/frp_b200.h, frp_shard_record_bytes):
    idx[B,k] i64 | score[B,k] f32 | count[B] i32 | status[B] i32, padded to 16 bytes."""
    n = B * k * 8 + B * k * 4 + B * 4 + B * 4
    return (n + 15) // 16 * 16


def unpack_records(buf: np.ndarray, R: int, B: int, k: int):
    """Split R concatenated records (uint8 array) into idx[R,B,k], score[R,B,k], count[R,B], status[R,B]."""
    nb = record_bytes(B, k)
    rec = np.asarray(buf, dtype=np.uint8).reshape(R, nb)
    o1 = B * k * 8
    o2 = o1 + B * k * 4
    o3 = o2 + B * 4
    idx = np.ascontiguousarray(rec[:, :o1]).view(np.int64).reshape(R, B, k)
    score = np.ascontiguousarray(rec[:, o1:o2]).view(np.float32).reshape(R, B, k)
    count = np.ascontiguousarray(rec[:, o2:o3]).view(np.int32).reshape(R, B)
    status = np.ascontiguousarray(rec[:, o3:o3 + B * 4]).view(np.int32).reshape(R, B)
    return idx, score, count, status


def shard_bounds(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the global row sequence owned by ``rank``
    (first ``total_rows % world_size`` ranks get one extra row)."""
    base, rem = divmod(int(total_rows), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_streams(stream_ids, world_size: int, rank: int) -> list:
    """Streams (cameras) whose frames ``rank`` aligns.  Alignment has no cross-frame state, so
    the stage shards by stream with no exchange: every rank aligns the detections
    of its own cameras and feeds the embeddings into the (row-sharded) matcher.  The assignment
    is deterministic and balanced: ids in sorted order, dealt round-robin."""
    ids = sorted(stream_ids, key=lambda v: (str(type(v)), v))
    return [sid for i, sid in enumerate(ids) if i % int(world_size) == int(rank)]


class ShardedMatcher:
    """Per-rank matcher over a row shard.

    Parameters
    ----------
    gallery : frp_b200.gallery.Gallery holding this rank's rows (``set_base`` is
        applied here so results carry global row indices).
    total_rows, rank, world_size : the partition.
    group : torch.distributed process group (None = default).
    """

    def __init__(self, gallery, total_rows: int, rank: int, world_size: int, group=None,
                 local_match: Optional[Callable] = None, merge: Optional[Callable] = None,
                 exchange: str = "nccl"):
        """exchange: "nccl" = all_gather_into_tensor + merge kernel; "p2p" = ONE fused kernel
        that pushes the record into every peer's symmetric-memory buffer over NVLink, waits
        on system-scope flags and merges (frp_push_merge_device) — no collective call."""
        import torch  # plumbing only: device buffers, streams, the collective
        self.torch = torch
        self.g = gallery
        self.rank, self.world = int(rank), int(world_size)
        self.group = group
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.exchange = exchange
        self._p2p = {}
        self.lo, self.hi = shard_bounds(total_rows, world_size, rank)
        if gallery is not None:
            gallery.set_base(self.lo)
        self._local_match = local_match or self._cuda_local_match
        self._merge = merge or self._cuda_merge
        self._bufs = {}

    # -- buffers (allocated once per (B, k), reused every step) ----------------
    def _buffers(self, B: int, k: int, device):
        key = (B, k, str(device))
        if key not in self._bufs:
            t = self.torch
            # packed per-rank record: idx[B,k] i64 | score[B,k] f32 | count[B] i32 | status[B] i32
            nbytes = record_bytes(B, k)
            send = t.empty(nbytes, dtype=t.uint8, device=device)
            recv = t.empty(self.world * nbytes, dtype=t.uint8, device=device)
            out_idx = t.empty((B, k), dtype=t.int64, device=device)
            out_score = t.empty((B, k), dtype=t.float32, device=device)
            out_count = t.empty((B,), dtype=t.int32, device=device)
            any_status = t.zeros((1,), dtype=t.int32, device=device)
            self._bufs[key] = (nbytes, send, recv, out_idx, out_score, out_count, any_status)
        return self._bufs[key]

    @staticmethod
    def _views(buf, B: int, k: int, torch):
        o1 = B * k * 8
        o2 = o1 + B * k * 4
        o3 = o2 + B * 4
        idx = buf[:o1].view(torch.int64).view(B, k)
        score = buf[o1:o2].view(torch.float32).view(B, k)
        count = buf[o2:o3].view(torch.int32)
        status = buf[o3:o3 + B * 4].view(torch.int32)
        return idx, score, count, status

    # -- default CUDA implementations ------------------------------------------
    def _cuda_local_match(self, q_dev, B, tau, k, flags, idx, score, count, status, stream):
        self.g.match_device(q_dev, B, tau, k, flags, idx, score, count, status, stream)

    def _cuda_merge(self, metric, R, B, k, records, out_idx, out_score, out_count, any_status, stream):
        from ._capi import check, lib, ptr
        check(lib.frp_merge_packed_device(int(metric), R, B, k, ptr(records), ptr(out_idx), ptr(out_score),
                                          ptr(out_count), ptr(any_status), stream or None))

    # -- one step -----------------------------------------------------------------
    def match(self, q_dev, tau: float, k: int, flags: int = 0, metric: Optional[int] = None):
        """q_dev: [B, dim] fp32 tensor on this rank's device (same on every rank).
        Returns device tensors (idx[B,k] global rows, score[B,k], count[B], any_fallback[1])."""
        t = self.torch
        import torch.distributed as dist
        B = int(q_dev.shape[0])
        device = q_dev.device
        nbytes, send, recv, out_idx, out_score, out_count, any_status = self._buffers(B, k, device)
        stream = t.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0
        idx, score, count, status = self._views(send, B, k, t)
        self._local_match(q_dev, B, tau, k, flags, idx, score, count, status, stream)
        m = metric if metric is not None else self.g.metric
        if self.exchange == "p2p":
            from ._capi import check, lib, ptr
            st = self._p2p_state(B, k, device)
            st["epoch"] += 1
            check(lib.frp_push_merge_device(int(m), self.world, self.rank, B, k, ptr(send), st["peers_dev"],
                                            st["epoch"] & 0xFFFFFFFF, ptr(st["counter"]), ptr(out_idx),
                                            ptr(out_score), ptr(out_count), ptr(any_status), stream or None))
            return out_idx, out_score, out_count, any_status
        if self.world > 1:
            dist.all_gather_into_tensor(recv, send, group=self.group)
        else:
            recv = send               # a single shard: merge straight from the local record
        self._merge(m, self.world, B, k, recv, out_idx, out_score, out_count, any_status, stream)
        return out_idx, out_score, out_count, any_status

    def _p2p_state(self, B: int, k: int, device):
        """Exchange buffer of this rank for (B, k): symmetric memory mapped into every rank
        (torch.distributed._symmetric_memory = CUDA VMM handles over NVLink), allocated and
        rendezvoused once; a plain device buffer when there is a single rank."""
        key = (B, k, str(device))
        if key in self._p2p:
            return self._p2p[key]
        t = self.torch
        from ._capi import lib
        nbytes = int(lib.frp_exchange_buffer_bytes(self.world, B, k))
        if self.world > 1:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(nbytes, dtype=t.uint8, device=device)
            buf.zero_()
            t.cuda.synchronize(device)
            hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            hdl.barrier()                         # every rank's flags are zero before the first push
            peers_dev = int(hdl.buffer_ptrs_dev)
            keep = (buf, hdl)
        else:
            buf = t.zeros(nbytes, dtype=t.uint8, device=device)
            ptrs = t.tensor([buf.data_ptr()], dtype=t.int64, device=device)
            t.cuda.synchronize(device)
            peers_dev = ptrs.data_ptr()
            keep = (buf, ptrs)
        st = {"peers_dev": peers_dev, "counter": t.zeros(1, dtype=t.int32, device=device), "epoch": 0,
              "keep": keep}
        self._p2p[key] = st
        return st

    def match_certified(self, q_dev, tau: float, k: int, flags: int = 0, metric: Optional[int] = None):
        """``match`` plus the fallback ladder of the host API: if any shard could not
        certify its fp16 prefilter (status word != 0 after the merge, identical on every
        rank, so all ranks take the same branch), the step is repeated against the fp32
        master and, if needed, with the widest candidate lists.  Costs one device->host
        read of a 4-byte flag per step.  Returns (idx, score, count, level) with level
        0 = certified by the fp16 path, 1/2 = certified by a fallback, 3 = uncertified
        (more than 64 exact ties at the cut)."""
        from ._capi import MATCH_FORCE_EXACT, MATCH_FORCE_GEMM, MATCH_FORCE_SCAN, MATCH_WIDE
        out = self.match(q_dev, tau, k, flags, metric)
        if int(out[3].item()) == 0:
            return out[0], out[1], out[2], 0
        base = flags & ~(MATCH_FORCE_GEMM | MATCH_FORCE_SCAN)
        for level, extra in ((1, MATCH_FORCE_EXACT), (2, MATCH_FORCE_EXACT | MATCH_WIDE)):
            out = self.match(q_dev, tau, k, base | extra, metric)
            if int(out[3].item()) == 0:
                return out[0], out[1], out[2], level
        return out[0], out[1], out[2], 3

"""Row-sharded gallery: one process per GPU, contiguous blocks of the global
insertion sequence per rank, queries replicated, per-shard top-k merged after one
all-gather (NCCL over NVLink/NVSwitch).  The reference is single-process; this is
the scaling layer of SURVEY.md section 8(e).

Data path per step:
    frp_match_record_device (local threshold + top-k, exact; the re-rank kernel writes the
                             packed record straight into the all-gather send buffer)
    torch.distributed.all_gather_into_tensor (B*k*16 + 8 B bytes per rank)
    frp_merge_packed_device (same comparator as the re-rank: float64 score, then global row)
The merge is deterministic and identical on every rank, so every rank holds the
global result (no broadcast back).

``match`` runs the three pieces on the caller's stream.  ``submit`` / ``result`` run the
same pieces as a two-slot pipeline over two streams: the all-gather + merge of step i
overlap the prefilter of step i+1 (queries are independent), so the NCCL latency and the
inter-rank skew of one step are hidden behind the next one.

``local_match`` / ``merge`` are injectable so the rendezvous, partitioning and
gather plumbing can be exercised with the gloo backend on CPU in the tests.

``ShardedGalleryDict`` is the mutable form: names map to (shard, local row, global seq);
new faces go to the least-full shard and carry the next global sequence number, which the
kernels report instead of ``base_row + row`` (frp_gallery_set_seq), so results and tie order
are those of ONE insertion-ordered dict no matter how many shards hold the rows.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np


def record_bytes(B: int, k: int) -> int:
    """Size of one rank's packed result record (include/frp_b200.h, frp_shard_record_bytes):
    idx[B,k] i64 | score[B,k] f64 | count[B] i32 | status[B] i32, padded to 16 bytes."""
    n = B * k * 8 + B * k * 8 + B * 4 + B * 4
    return (n + 15) // 16 * 16


def unpack_records(buf: np.ndarray, R: int, B: int, k: int):
    """Split R concatenated records (uint8 array) into idx[R,B,k], score[R,B,k] (f64), count[R,B], status[R,B]."""
    nb = record_bytes(B, k)
    rec = np.asarray(buf, dtype=np.uint8).reshape(R, nb)
    o1 = B * k * 8
    o2 = o1 + B * k * 8
    o3 = o2 + B * 4
    idx = np.ascontiguousarray(rec[:, :o1]).view(np.int64).reshape(R, B, k)
    score = np.ascontiguousarray(rec[:, o1:o2]).view(np.float64).reshape(R, B, k)
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
        applied here so results carry global row indices; pass ``set_base=False`` for a
        gallery that carries explicit sequences).
    total_rows, rank, world_size : the partition.
    group : torch.distributed process group (None = default).
    """

    def __init__(self, gallery, total_rows: int, rank: int, world_size: int, group=None,
                 local_match: Optional[Callable] = None, merge: Optional[Callable] = None,
                 exchange: str = "nccl", set_base: bool = True):
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
        if gallery is not None and set_base:
            gallery.set_base(self.lo)
        self._local_match = local_match or self._cuda_local_match
        self._merge = merge or self._cuda_merge
        self._bufs = {}
        self._pipe = None

    # -- buffers (allocated once per (B, k, slot), reused every step) ----------------
    def _buffers(self, B: int, k: int, device, slot: int = 0):
        key = (B, k, str(device), slot)
        if key not in self._bufs:
            t = self.torch
            nbytes = record_bytes(B, k)
            send = t.zeros(nbytes, dtype=t.uint8, device=device)
            recv = t.zeros(self.world * nbytes, dtype=t.uint8, device=device)
            out_idx = t.empty((B, k), dtype=t.int64, device=device)
            out_score = t.empty((B, k), dtype=t.float32, device=device)
            out_count = t.empty((B,), dtype=t.int32, device=device)
            any_status = t.zeros((1,), dtype=t.int32, device=device)
            self._bufs[key] = (nbytes, send, recv, out_idx, out_score, out_count, any_status)
        return self._bufs[key]

    @staticmethod
    def _views(buf, B: int, k: int, torch):
        o1 = B * k * 8
        o2 = o1 + B * k * 8
        o3 = o2 + B * 4
        idx = buf[:o1].view(torch.int64).view(B, k)
        score = buf[o1:o2].view(torch.float64).view(B, k)
        count = buf[o2:o3].view(torch.int32)
        status = buf[o3:o3 + B * 4].view(torch.int32)
        return idx, score, count, status

    # -- default CUDA implementations ------------------------------------------
    def _cuda_local_match(self, q_dev, B, tau, k, flags, idx, score, count, status, stream):
        # the four views are one packed record starting at idx: the re-rank kernel fills it in place
        from ._capi import check, lib, ptr
        check(lib.frp_match_record_device(self.g.handle, ptr(q_dev), int(B), float(tau), int(k), int(flags),
                                          idx.data_ptr(), stream or None))

    def _cuda_merge(self, metric, R, B, k, records, out_idx, out_score, out_count, any_status, stream):
        from ._capi import check, lib, ptr
        check(lib.frp_merge_packed_device(int(metric), R, B, k, ptr(records), ptr(out_idx), ptr(out_score), None,
                                          ptr(out_count), ptr(any_status), stream or None))

    def _stream_of(self, device):
        t = self.torch
        return t.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0

    # -- one step, everything on the caller's stream -----------------------------------
    def match(self, q_dev, tau: float, k: int, flags: int = 0, metric: Optional[int] = None):
        """q_dev: [B, dim] fp32 tensor on this rank's device (same on every rank).
        Returns device tensors (idx[B,k] global rows, score[B,k], count[B], any_fallback[1])."""
        t = self.torch
        import torch.distributed as dist
        B = int(q_dev.shape[0])
        device = q_dev.device
        nbytes, send, recv, out_idx, out_score, out_count, any_status = self._buffers(B, k, device)
        stream = self._stream_of(device)
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

    # -- CUDA graph of one step ---------------------------------------------------------
    def capture(self, q_static, tau: float, k: int, flags: int = 0, metric: Optional[int] = None):
        """Capture one step (local match -> all-gather -> merge) reading ``q_static`` into a CUDA
        graph: a step is 5-6 short dependent launches around one long kernel, and at 8 ranks
        (0.25 ms of kernels per step) the gaps between them are ~15 % of the step.  Returns
        (graph, outputs): ``q_static.copy_(q); graph.replay()`` runs a step on the current
        stream; ``outputs`` = (idx, score, count, any_fallback) device tensors it refills.
        The warm-up step runs on the capture stream itself so every workspace the library
        needs for that stream exists before capture starts (no allocation inside the graph)."""
        t = self.torch
        dev = q_static.device
        side = t.cuda.Stream(device=dev)
        side.wait_stream(t.cuda.current_stream(dev))
        with t.cuda.stream(side):
            for _ in range(2):
                self.match(q_static, tau, k, flags, metric)
        side.synchronize()
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g, stream=side):
            out = self.match(q_static, tau, k, flags, metric)
        t.cuda.current_stream(dev).wait_stream(side)
        return g, out

    # -- two-slot pipeline: exchange of step i behind the prefilter of step i+1 -----------
    def _pipeline(self, device, depth: int):
        if self._pipe is None or self._pipe["device"] != device or self._pipe["depth"] != depth:
            t = self.torch
            cuda = device.type == "cuda"
            self._pipe = {
                "device": device, "depth": depth, "n": 0,
                "xstream": t.cuda.Stream(device=device, priority=-1) if cuda else None,
                "matched": [t.cuda.Event() if cuda else None for _ in range(depth)],
                "done": [t.cuda.Event() if cuda else None for _ in range(depth)],
                "inflight": [None] * depth,
            }
        return self._pipe

    def submit(self, q_dev, tau: float, k: int, flags: int = 0, metric: Optional[int] = None, depth: int = 2,
               host_out=None) -> int:
        """Queue one step; returns a ticket for ``result`` / ``wait_host``.  The local match runs
        on the caller's stream, the all-gather + merge on the exchange stream.  At most ``depth``
        steps are in flight: a slot is reused only after ``result`` has been called for its
        previous ticket (the caller owns the slot's output tensors until then).
        host_out = (idx, score, count, flag) pinned host tensors: the merged result is also copied
        to them on the exchange stream, so the device->host traffic of step i never waits for the
        prefilter of step i+1."""
        t = self.torch
        import torch.distributed as dist
        if self.exchange != "nccl":
            raise ValueError("the pipelined path uses the NCCL exchange")
        if self.world > 1 and not getattr(self, "_reserved", False) and q_dev.device.type == "cuda":
            # the all-gather + merge of step i must RUN while the prefilter of step i+1 holds the
            # SMs: keep a few out of the prefilter's grid (FRP_RESERVED_SMS overrides; 0 = none)
            import os
            from ._capi import check, lib
            check(lib.frp_set_reserved_sms(int(os.environ.get("FRP_RESERVED_SMS", "2"))))
            self._reserved = True
        B = int(q_dev.shape[0])
        device = q_dev.device
        p = self._pipeline(device, depth)
        ticket = p["n"]
        slot = ticket % depth
        if p["inflight"][slot] is not None:
            raise RuntimeError("pipeline slot %d still holds ticket %d: call result() first" % (slot, p["inflight"][slot][0]))
        nbytes, send, recv, out_idx, out_score, out_count, any_status = self._buffers(B, k, device, slot)
        cuda = device.type == "cuda"
        cur = t.cuda.current_stream(device) if cuda else None
        idx, score, count, status = self._views(send, B, k, t)
        if cuda:
            cur.wait_event(p["done"][slot])      # the slot's previous exchange has consumed `send`
        self._local_match(q_dev, B, tau, k, flags, idx, score, count, status, cur.cuda_stream if cuda else 0)
        m = metric if metric is not None else self.g.metric
        if cuda:
            p["matched"][slot].record(cur)
            xs = p["xstream"]
            xs.wait_event(p["matched"][slot])
            with t.cuda.stream(xs):
                if self.world > 1:
                    dist.all_gather_into_tensor(recv, send, group=self.group)
                else:
                    recv = send
                self._merge(m, self.world, B, k, recv, out_idx, out_score, out_count, any_status, xs.cuda_stream)
                if host_out is not None:
                    for dst, src in zip(host_out, (out_idx, out_score, out_count, any_status)):
                        dst.copy_(src, non_blocking=True)
                p["done"][slot].record(xs)
        else:
            if self.world > 1:
                dist.all_gather_into_tensor(recv, send, group=self.group)
            else:
                recv = send
            self._merge(m, self.world, B, k, recv, out_idx, out_score, out_count, any_status, 0)
            if host_out is not None:
                for dst, src in zip(host_out, (out_idx, out_score, out_count, any_status)):
                    dst.copy_(src)
        p["inflight"][slot] = (ticket, out_idx, out_score, out_count, any_status)
        p["n"] += 1
        return ticket

    def result(self, ticket: int):
        """Outputs of a submitted step; the caller's current stream is made to wait for the
        step's merge (no host synchronisation)."""
        p = self._pipe
        slot = ticket % p["depth"]
        ent = p["inflight"][slot]
        if ent is None or ent[0] != ticket:
            raise RuntimeError("ticket %d is not in flight" % ticket)
        if p["device"].type == "cuda":
            self.torch.cuda.current_stream(p["device"]).wait_event(p["done"][slot])
        p["inflight"][slot] = None
        return ent[1], ent[2], ent[3], ent[4]

    def wait_host(self, ticket: int):
        """Block the HOST until a submitted step (merge and host copies) has completed."""
        p = self._pipe
        slot = ticket % p["depth"]
        ent = p["inflight"][slot]
        if ent is None or ent[0] != ticket:
            raise RuntimeError("ticket %d is not in flight" % ticket)
        if p["device"].type == "cuda":
            p["done"][slot].synchronize()
        p["inflight"][slot] = None
        return ent[1], ent[2], ent[3], ent[4]

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
        read of a 4-byte flag per step (``CertifiedPipeline`` defers that read by one step).
        Returns (idx, score, count, level) with level 0 = certified by the fp16 path,
        1/2 = certified by a fallback, 3 = uncertified (more than 64 exact ties at the cut)."""
        out = self.match(q_dev, tau, k, flags, metric)
        if int(out[3].item()) == 0:
            return out[0], out[1], out[2], 0
        return self._ladder(q_dev, tau, k, flags, metric)

    def _ladder(self, q_dev, tau, k, flags, metric):
        from ._capi import MATCH_FORCE_EXACT, MATCH_FORCE_GEMM, MATCH_FORCE_SCAN, MATCH_WIDE
        base = flags & ~(MATCH_FORCE_GEMM | MATCH_FORCE_SCAN)
        out = None
        for level, extra in ((1, MATCH_FORCE_EXACT), (2, MATCH_FORCE_EXACT | MATCH_WIDE)):
            out = self.match(q_dev, tau, k, base | extra, metric)
            if int(out[3].item()) == 0:
                return out[0], out[1], out[2], level
        return out[0], out[1], out[2], 3


class CertifiedPipeline:
    """End-to-end stepping with host buffers over a ``ShardedMatcher``: every step copies
    its queries from pinned host memory, runs the pipelined sharded match and copies the
    merged result back to pinned host memory — all asynchronous, the copies of step i on the
    exchange stream behind its merge.  The certificate flag of step i is read on the host while
    step i+1 is already running (a deferred check instead of a blocking ``.item()`` per step);
    a step whose flag is set is repeated through the fallback ladder before its result is
    handed out, so callers only ever see certified answers.

        pipe = CertifiedPipeline(matcher, B, dim, k, device)
        for q in batches:                      # q: pinned host tensor [B, dim]
            done = pipe.push(q, tau)           # result of the PREVIOUS step, or None
        last = pipe.flush()
    Results are (idx[B,k] int64, score[B,k] f32, count[B] i32, level) in pinned host memory,
    valid until the next-but-one push."""

    def __init__(self, matcher: ShardedMatcher, B: int, dim: int, k: int, device, depth: int = 2):
        t = matcher.torch
        self.m, self.B, self.k, self.device, self.depth = matcher, B, k, device, depth
        self.slots = []
        self._cuda = t.device(device).type == "cuda"
        pin = (lambda x: x.pin_memory()) if self._cuda else (lambda x: x)      # CPU tensors: the gloo tests
        for _ in range(depth):
            self.slots.append({
                "q": t.empty((B, dim), dtype=t.float32, device=device),
                "host": (pin(t.empty((B, k), dtype=t.int64)), pin(t.empty((B, k), dtype=t.float32)),
                         pin(t.empty((B,), dtype=t.int32)), pin(t.zeros((1,), dtype=t.int32))),
                "tau": None, "flags": 0, "ticket": None,
            })
        self.n = 0
        self.pending = None

    def push(self, hq, tau: float, flags: int = 0):
        s = self.slots[self.n % self.depth]
        s["q"].copy_(hq, non_blocking=True)
        s["tau"], s["flags"] = tau, flags
        s["ticket"] = self.m.submit(s["q"], tau, self.k, flags, depth=self.depth, host_out=s["host"])
        prev, self.pending = self.pending, s
        self.n += 1
        return self._finish(prev) if prev is not None else None

    def _finish(self, s):
        t = self.m.torch
        self.m.wait_host(s["ticket"])
        idx, score, count, flag = s["host"]
        level = 0
        if int(flag[0]) != 0:
            # rare: repeat this step through the ladder (synchronous), then copy again
            oi, os_, oc, level = self.m._ladder(s["q"], s["tau"], self.k, s["flags"], None)
            idx.copy_(oi); score.copy_(os_); count.copy_(oc)
            if self._cuda:
                t.cuda.current_stream(self.device).synchronize()
        return idx, score, count, level

    def flush(self):
        if self.pending is None:
            return None
        s, self.pending = self.pending, None
        return self._finish(s)


class ShardedGalleryDict:
    """Mutable row-sharded gallery with the dict semantics of ``GalleryDict`` (SURVEY 8e:
    "appends round-robin or to the least-full shard with global seq kept per row").

    * ``d[name] = vec``: a new name gets the next global sequence number and goes to the shard
      with the fewest live rows (lowest shard index on ties); an existing name is updated in
      place (position = sequence kept, face_service.py:374).
    * ``del d[name]``: the row is tombstoned in its shard (face_service.py:521-523).
    * results carry global sequence numbers (frp_gallery_set_seq); within a shard rows are
      appended in increasing sequence, so local tie order == global tie order, and the merge
      breaks cross-shard ties by sequence: answers are those of ONE insertion-ordered dict.

    ``shards`` = number of shards; ``rank`` = None owns every shard in this process (virtual
    shards on one GPU: tests, single-GPU capacity splits), an int owns only that shard — every
    rank then applies the same mutation stream (SPMD) and only the owner touches its device
    gallery; the others just keep the routing tables in step.  ``match`` merges the shards'
    records with frp_merge_packed_device (all-gather first when the shards live on other ranks).
    """

    def __init__(self, shards: int, dim: int, metric: int, device: int = 0, rank: Optional[int] = None,
                 group=None, gallery_factory=None, merge: Optional[Callable] = None):
        from .gallery import Gallery
        self.R, self.dim, self.metric, self.device, self.rank, self.group = int(shards), int(dim), int(metric), device, rank, group
        factory = gallery_factory or (lambda dim_, metric_, device_: Gallery(dim_, metric_, device_))
        owned = range(self.R) if rank is None else [int(rank)]
        self.galleries = {r: factory(self.dim, self.metric, device) for r in owned}
        self._where: Dict[str, Tuple[int, int, int]] = {}    # name -> (shard, local row, seq); insertion order == seq order
        self._name_of_seq: Dict[int, str] = {}
        self._rows = [0] * self.R          # appended rows per shard (incl. tombstones)
        self._live = [0] * self.R
        self._next_seq = 0
        self._merge = merge
        self._matchers = {}

    # -- dict protocol --------------------------------------------------------------------
    def __len__(self):
        return len(self._where)

    def __contains__(self, name):
        return name in self._where

    def keys(self):
        return list(self._where.keys())

    def name_of(self, seq: int):
        return self._name_of_seq.get(int(seq))

    def shard_of(self, name: str) -> int:
        return self._where[name][0]

    def live_per_shard(self):
        return list(self._live)

    def __setitem__(self, name: str, value) -> None:
        vec = np.asarray(value, dtype=np.float64).reshape(-1)
        if vec.shape[0] != self.dim:
            raise ValueError("embedding length %d does not match the gallery's %d" % (vec.shape[0], self.dim))
        if name in self._where:
            r, row, _ = self._where[name]
            if r in self.galleries:
                self.galleries[r].update(row, vec)
            return
        r = min(range(self.R), key=lambda i: (self._live[i], i))         # least-full shard
        seq = self._next_seq
        self._next_seq += 1
        row = self._rows[r]
        if r in self.galleries:
            g = self.galleries[r]
            first = g.append(vec[None, :])
            assert first == row, (first, row)
            g.set_seq(row, [seq])
        self._rows[r] += 1
        self._live[r] += 1
        self._where[name] = (r, row, seq)
        self._name_of_seq[seq] = name

    def __delitem__(self, name: str) -> None:
        r, row, seq = self._where.pop(name)
        del self._name_of_seq[seq]
        self._live[r] -= 1
        if r in self.galleries:
            self.galleries[r].remove(row)

    def rename(self, old: str, new: str, value=None) -> None:
        """Dict semantics of pop + assign: the face moves to the END of the order (new sequence).
        ``value`` = the embedding to store under the new name; when omitted the fp32 master row
        (normalised for cosine galleries) is read back from the owning shard."""
        if old not in self._where:
            raise KeyError(old)
        if value is None:
            r, row, _ = self._where[old]
            value = self.galleries[r].get_row(row) if r in self.galleries else None
            if self.rank is not None:
                value = self._bcast_row(value, r)
        del self[old]
        self[new] = value

    def _bcast_row(self, vec, owner):
        import torch
        import torch.distributed as dist
        t = torch.zeros(self.dim, dtype=torch.float32) if vec is None else torch.from_numpy(np.asarray(vec, np.float32))
        dev = torch.device("cuda", self.device) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.broadcast(t, owner, group=self.group)
        return t.cpu().numpy()

    # -- match ----------------------------------------------------------------------------
    def match(self, queries, tau: float, k: int, flags: int = 0):
        """Host queries [B, dim] -> (names[B][<=k], score[B,k] f32, count[B], level).
        Every owned shard runs frp_match_record_device; the records are merged by
        frp_merge_packed_device (virtual shards: concatenated in place; one shard per rank:
        all-gathered first).  An uncertified step walks the same ladder as frp_match (fp32
        master, then the widest lists); level 3 = still uncertified."""
        from ._capi import MATCH_FORCE_EXACT, MATCH_FORCE_GEMM, MATCH_FORCE_SCAN, MATCH_WIDE
        base = flags & ~(MATCH_FORCE_GEMM | MATCH_FORCE_SCAN)
        out = None
        for level, f in enumerate((flags, base | MATCH_FORCE_EXACT, base | MATCH_FORCE_EXACT | MATCH_WIDE)):
            out = self._match_once(queries, tau, k, f)
            if out[3] == 0:
                return out[0], out[1], out[2], level
        return out[0], out[1], out[2], 3

    def _match_once(self, queries, tau, k, flags):
        import torch
        from ._capi import check, lib, ptr
        q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
        B = q.shape[0]
        dev = torch.device("cuda", self.device)
        qd = torch.from_numpy(q).to(dev)
        nb = record_bytes(B, k)
        recv = torch.zeros(self.R * nb, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        for r, g in self.galleries.items():
            slot = recv[r * nb:(r + 1) * nb] if self.rank is None else recv[:nb]
            check(lib.frp_match_record_device(g.handle, ptr(qd), B, float(tau), int(k), int(flags), slot.data_ptr(), st or None))
        if self.rank is not None and self.R > 1:
            import torch.distributed as dist
            send = recv[:nb].clone()
            dist.all_gather_into_tensor(recv, send, group=self.group)
        out_idx = torch.empty((B, k), dtype=torch.int64, device=dev)
        out_score = torch.empty((B, k), dtype=torch.float32, device=dev)
        out_count = torch.empty((B,), dtype=torch.int32, device=dev)
        anyst = torch.zeros((1,), dtype=torch.int32, device=dev)
        check(lib.frp_merge_packed_device(int(self.metric), self.R, B, k, ptr(recv), ptr(out_idx), ptr(out_score), None,
                                          ptr(out_count), ptr(anyst), st or None))
        idx, sc, cnt = out_idx.cpu().numpy(), out_score.cpu().numpy(), out_count.cpu().numpy()
        names: List[List[Optional[str]]] = [[self.name_of(int(idx[b, j])) for j in range(int(cnt[b]))] for b in range(B)]
        return names, sc, cnt, int(anyst.item())

    def close(self):
        for g in self.galleries.values():
            g.close()
        self.galleries = {}

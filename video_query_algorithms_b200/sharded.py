"""Multi-GPU scan, one process per GPU (torchrun): clip-range shards, no data-path collective except the
merge of per-rank results.

Each rank owns rows [rank * n_local, (rank + 1) * n_local) of the search set.  After its local scan a
rank's payload is `[4 + 2k] int64` = n_match, n_near, n_tie, n_topk | top-k global rows | top-k score bits;
ranks exchange payloads with ONE all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests) and every
rank merges the gathered buffer: counts are summed, the global top-k is the k best under
(score descending, global row ascending) — the ranking rule of reference ticket.py:266.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import check, lib, ptr


def pack_payload(counts, rows, scores, k):
    """Host-side twin of what `select_compact` writes on the device (tests / CPU ranks)."""
    out = np.empty(4 + 2 * k, np.int64)
    out[:4] = counts
    out[4:4 + k] = -1
    out[4 + k:] = np.float32(-np.inf).view(np.uint32)
    n = len(rows)
    out[4:4 + n] = rows
    out[4 + k:4 + k + n] = np.asarray(scores, np.float32).view(np.uint32).astype(np.int64)
    return out


def unpack_payload(payload, k):
    counts = payload[:4].copy()
    n = int(counts[3])
    rows = payload[4:4 + n].copy()
    scores = payload[4 + k:4 + k + n].astype(np.uint32).view(np.float32).copy()
    return counts, rows, scores


def merge_payloads_host(gathered, world, k):
    """gathered: int64 [world, 4 + 2k] -> merged payload (host merge through the C ABI's vq_merge_topk)."""
    g = np.ascontiguousarray(gathered, dtype=np.int64).reshape(world, 4 + 2 * k)
    rows = np.ascontiguousarray(g[:, 4:4 + k])
    scores = np.ascontiguousarray(g[:, 4 + k:].astype(np.uint32).view(np.float32))
    so, ro, n = np.empty(k, np.float32), np.empty(k, np.int64), C.c_int32()
    check(lib().vq_merge_topk(world, k, ptr(scores), ptr(rows), ptr(so), ptr(ro), C.byref(n)), "vq_merge_topk")
    counts = g[:, :4].sum(axis=0)
    counts[3] = n.value
    return pack_payload(counts, ro[:n.value], so[:n.value], k)


class _DevArray:
    """Zero-copy __cuda_array_interface__ view of library-owned device memory."""

    def __init__(self, ptr_value, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr_value, False), "version": 3}


class RankScan:
    """One rank's view: local shard handle + the exchange.  `dist` is torch.distributed (already
    initialised).  exchange = "p2p": payloads are stored straight into the peers' inboxes over NVLink and
    merged by one kernel (vq_scan_exchange_enqueue, csrc/vq_exchange.cu); exchange = "p2p-lagged": the same kernel
    pushes step i and merges step i-1 (a stream of queries: ranks never wait for the slowest rank of the current
    step; call flush() after the last step); exchange = "nccl": one all_gather_into_tensor + the merge kernel.
    Everything is enqueued on the caller's stream."""

    def __init__(self, handle, k, device_index, dist=None, torch=None, exchange="p2p"):
        self.handle, self.k, self.device_index = handle, int(k), device_index
        self.dist, self.torch = dist, torch
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n_payload = 4 + 2 * self.k
        self.exchange = exchange if self.world > 1 else "none"
        self._payload_t = None
        self._x = None
        dev = torch.device("cuda", device_index)
        if self.exchange == "nccl":
            self.gathered = torch.zeros(self.world * self.n_payload, dtype=torch.int64, device=dev)
            self.merged = torch.zeros(self.n_payload, dtype=torch.int64, device=dev)
        elif self.exchange in ("p2p", "p2p-lagged"):
            self._x = C.c_void_p()
            check(lib().vq_exchange_create(C.byref(self._x), device_index, self.world, self.rank), "vq_exchange_create")
            mine = np.zeros(64, np.uint8)
            check(lib().vq_exchange_local_handle(self._x, ptr(mine)), "vq_exchange_local_handle")
            t_mine = torch.from_numpy(mine).to(dev)
            t_all = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(t_all, t_mine)             # setup only: 64-byte IPC handles
            handles = np.ascontiguousarray(t_all.cpu().numpy())
            check(lib().vq_exchange_connect(self._x, ptr(handles)), "vq_exchange_connect")
            dist.barrier()
            p = C.c_void_p()
            check(lib().vq_exchange_merged(self._x, C.byref(p)), "vq_exchange_merged")
            self.merged = torch.as_tensor(_DevArray(p.value, self.n_payload), device=dev)

    def close(self):
        if self._x is not None:
            self.torch.cuda.synchronize()
            if self.world > 1:
                self.dist.barrier()                                  # nobody unmaps while a peer may still write
            lib().vq_exchange_destroy(self._x)
            self._x = None

    def _payload_view(self):
        """Zero-copy torch view of the payload the library wrote on the device."""
        if self._payload_t is None:
            p, n = C.c_void_p(), C.c_int32()
            check(lib().vq_scan_payload(self.handle, C.byref(p), C.byref(n)), "vq_scan_payload")
            self._payload_t = self.torch.as_tensor(_DevArray(p.value, n.value),
                                                   device=self.torch.device("cuda", self.device_index))
        return self._payload_t

    def enqueue(self, target_dev_ptr, params, stream_ptr):
        """Local scan + (world > 1) exchange + merge, all enqueued on `stream_ptr`; no host sync."""
        check(lib().vq_scan_enqueue(self.handle, C.c_void_p(target_dev_ptr), C.byref(params), C.c_void_p(stream_ptr)),
              "vq_scan_enqueue")
        if self.exchange == "p2p":
            check(lib().vq_scan_exchange_enqueue(self.handle, self._x, C.c_void_p(stream_ptr)), "vq_scan_exchange_enqueue")
        elif self.exchange == "p2p-lagged":
            check(lib().vq_scan_exchange_enqueue_lagged(self.handle, self._x, C.c_void_p(stream_ptr)),
                  "vq_scan_exchange_enqueue_lagged")
        elif self.exchange == "nccl":
            self.dist.all_gather_into_tensor(self.gathered, self._payload_view())
            check(lib().vq_merge_payloads_enqueue(self.device_index, C.c_void_p(self.gathered.data_ptr()), self.world,
                                                  self.k, C.c_void_p(self.merged.data_ptr()), C.c_void_p(stream_ptr)),
                  "vq_merge_payloads_enqueue")

    def flush(self, stream_ptr):
        """Peer-memory exchange: merge the last pushed step if it is lagged, and make the caller's stream wait for the
        exchange's own stream (enqueued; no host sync).  No-op for NCCL."""
        if self.exchange in ("p2p", "p2p-lagged"):           # also joins the exchange's own stream into `stream_ptr`
            check(lib().vq_exchange_flush_enqueue(self._x, C.c_void_p(stream_ptr)), "vq_exchange_flush_enqueue")

    def kernels_per_step(self):
        return 4 + {"none": 0, "p2p": 1, "p2p-lagged": 1, "nccl": 1}[self.exchange]

    def result(self):
        """(counts[4], global top-k rows, scores) after the stream has been synchronised."""
        if self._x is not None:
            check(lib().vq_exchange_check(self._x), "vq_exchange_check")       # a peer that never delivered -> VQError, not a hang
        t = self.merged if self.world > 1 else self._payload_view()
        return unpack_payload(t.cpu().numpy(), self.k)

    def exchange_times(self, parts=False):
        """the exchange kernels' own times (ms) since the last call (after a synchronisation); parts=True: also
        [n, 3] = pushes | wait for the peers' flags | merge"""
        if self._x is None:
            return (np.empty(0, np.float32), np.empty((0, 3), np.float32)) if parts else np.empty(0, np.float32)
        out, split, n = np.empty(256, np.float32), np.empty((256, 3), np.float32), C.c_int32()
        check(lib().vq_exchange_kernel_times(self._x, 256, ptr(out), ptr(split), C.byref(n)), "vq_exchange_kernel_times")
        return (out[:n.value], split[:n.value]) if parts else out[:n.value]


def exchange_host(payload, dist, torch, k):
    """CPU ranks (gloo): allgather numpy payloads and merge on the host."""
    world = dist.get_world_size()
    mine = torch.from_numpy(np.ascontiguousarray(payload, dtype=np.int64))
    gathered = torch.empty(world * mine.numel(), dtype=torch.int64)
    dist.all_gather_into_tensor(gathered, mine)
    return merge_payloads_host(gathered.numpy(), world, k)


# ---------------------------------------------------------------------------------------------------
# The other exchanges of SURVEY.md §8(e): variable-length ordered lists, per-query top-k of the batched
# path, labelled similarities for the weight update.  All three are small next to the scan they follow
# (KBs to a few MB), so they are plain collectives on whatever backend `dist` runs (NCCL over NVLink with
# one rank per GPU, gloo in the CPU tests); `device` = the rank's torch.device for NCCL, None for gloo.

def _t(a, torch, device):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device) if device is not None else t


def all_gather_np(a, dist, torch, device=None):
    """numpy [..] on every rank -> numpy [world, ..] on every rank (same shape on all ranks)."""
    world = dist.get_world_size()
    mine = _t(a, torch, device).reshape(-1)
    out = torch.empty(world * mine.numel(), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().reshape((world,) + tuple(np.shape(a)))


class HostMailbox:
    """All-gather of small host records between the rank processes of one box through shared memory
    (vq_hostx_*, csrc/vq_hostx.cu): no device hop, no stream synchronisation.  Setup is a collective on `dist`
    (a unique segment name from rank 0, create / attach around a barrier, unlink once everybody is attached)."""

    def __init__(self, dist, slot_bytes=1 << 16, timeout_s=60.0):
        import os
        import secrets
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.slot_bytes, self.timeout_s = int(slot_bytes), float(timeout_s)
        name = ["/vq-%d-%s" % (os.getpid(), secrets.token_hex(6))] if self.rank == 0 else [None]
        dist.broadcast_object_list(name, src=0)
        self._x = C.c_void_p()
        err = None
        for turn in (0, 1):                                  # rank 0 creates, barrier, the others attach, barrier
            if (self.rank == 0) == (turn == 0):
                try:
                    check(lib().vq_hostx_create(C.byref(self._x), name[0].encode(), self.world, self.rank,
                                                self.slot_bytes), "vq_hostx_create")
                except _ffi.VQError as e:
                    err, self._x = e, C.c_void_p()
            dist.barrier()
        flags = [None] * self.world
        dist.all_gather_object(flags, err is None)
        if self._x:
            lib().vq_hostx_unlink(self._x)
        if not all(flags):                                   # e.g. ranks in different containers: nobody uses it
            self.close()
            raise _ffi.VQError("host mailbox unavailable on ranks %s: %s"
                               % ([r for r, f in enumerate(flags) if not f], err or "see those ranks"))

    @classmethod
    def try_create(cls, dist, **kw):
        """A mailbox when every rank of `dist` runs on this host and can attach, else None (callers then use the
        collectives of `dist`).  Every rank gets the same answer."""
        import socket
        hosts = [None] * dist.get_world_size()
        dist.all_gather_object(hosts, socket.gethostname())
        if len(set(hosts)) != 1:
            return None
        try:
            return cls(dist, **kw)
        except _ffi.VQError:
            return None

    def fits(self, a):
        return a.nbytes <= self.slot_bytes

    def all_gather(self, a):
        """numpy [..] on every rank -> numpy [world, ..] on every rank (same shape and dtype on all ranks)."""
        a = np.ascontiguousarray(a)
        out = np.empty((self.world,) + a.shape, a.dtype)
        check(lib().vq_hostx_allgather(self._x, ptr(a), a.nbytes, ptr(out), self.timeout_s), "vq_hostx_allgather")
        return out

    def close(self):
        if self._x:
            lib().vq_hostx_destroy(self._x)
            self._x = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_batch(counts, topk_rows, topk_scores, dist, torch, device=None):
    """Per-rank results of the batched path (counts [Q, 2], top-k global rows / scores [Q, k]) -> the search
    set's: counts summed, per-query top-k merged with the ranking rule (vq_merge_topk_batch).  One allgather of
    Q * (2 + 2k) int64 per rank (413 KB at Q = 256, k = 100)."""
    from .store import merge_topk_batch
    counts = np.asarray(counts, np.int64)
    Q, k = topk_rows.shape
    pay = np.empty((Q, 2 + 2 * k), np.int64)
    pay[:, :2] = counts
    pay[:, 2:2 + k] = topk_rows
    pay[:, 2 + k:] = np.ascontiguousarray(topk_scores, dtype=np.float32).view(np.uint32)
    g = all_gather_np(pay, dist, torch, device)
    total = g[:, :, :2].sum(axis=0)
    if k == 0:
        return total, np.empty((Q, 0), np.int64), np.empty((Q, 0), np.float32)
    rows_o, sc_o = merge_topk_batch(g[:, :, 2:2 + k], g[:, :, 2 + k:].astype(np.uint32).view(np.float32))
    return total, rows_o, sc_o


def gather_sims(partial, dist, torch, device=None):
    """Labelled similarities for the weight update (hyperparameter.py:45-65): every rank holds float64 [L, S] with
    the rows it owns filled in and zeros elsewhere; one all_reduce(SUM) of 8*L*S bytes gives every rank the full
    table (x + 0 is exact, so the values are the owners' bit for bit)."""
    t = _t(np.asarray(partial, np.float64), torch, device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


TIE_CAP = 512         # tie-band entries per rank that ride in the summary record (the band is 2 * COMPUTE_EPS wide: ~10 per 1M clips)
_F32_NINF_BITS = int(np.array([-np.inf], np.float32).view(np.uint32)[0])


def _bits(scores):
    return np.ascontiguousarray(scores, dtype=np.float32).view(np.uint32).astype(np.int64)


def _floats(bits):
    return np.ascontiguousarray(bits).astype(np.uint32).view(np.float32)


class RankSummary:
    """What ONE all_gather of a fixed-size record per rank gives every rank: each rank's first global row and list
    lengths (so list positions map to their owners without further traffic), the merged top-k, each rank's best
    near miss and — when every rank's tie band fits TIE_CAP entries — the tie band."""

    def __init__(self, first_rows, counts, topk, near_best, ties):
        self.first_rows, self.counts, self.topk, self.near_best, self.ties = first_rows, counts, topk, near_best, ties

    @property
    def total(self):
        return self.counts.sum(axis=0)


class _Scratch:
    """Persistent numpy buffers with their addresses taken once: the summary record is packed and merged on every rank's
    per-query path, where `ndarray.ctypes` (about 4 us per pointer) would cost more than the work itself."""

    def __init__(self, **shapes):
        self.a, self.p = {}, {}
        for name, (shape, dtype) in shapes.items():
            self.a[name] = np.empty(shape, dtype)
            self.p[name] = self.a[name].ctypes.data


_PACK, _MERGE = {}, {}


def summary_record(first_row, counts3, topk_rows, topk_scores, k, near_best=None, ties=None):
    """int64 [5 + 2k + 3 + 2*TIE_CAP]: first row | n_match n_near n_tie n_topk | top-k rows | top-k score bits |
    best near miss (position in this rank's near-miss list, global row, score bits; row -1 = none) | tie rows | tie bits.
    Packed by the library (vq_summary_pack)."""
    k = int(k)
    sc = _PACK.get((k, TIE_CAP))
    if sc is None:
        sc = _PACK[(k, TIE_CAP)] = _Scratch(c3=(3, np.int64), t_rows=(max(k, 1), np.int64), t_sc=(max(k, 1), np.float32),
                                           nb=(3, np.int64), nb_sc=(1, np.float32), tie_r=(TIE_CAP, np.int64), tie_s=(TIE_CAP, np.float32))
    A, P = sc.a, sc.p
    A["c3"][:] = counts3
    n_top = len(topk_rows)
    A["t_rows"][:n_top] = topk_rows
    A["t_sc"][:n_top] = topk_scores
    nb_p = None
    if near_best is not None and near_best[1] >= 0:
        A["nb_sc"][0] = near_best[2]
        A["nb"][0], A["nb"][1], A["nb"][2] = near_best[0], near_best[1], int(A["nb_sc"].view(np.uint32)[0])
        nb_p = P["nb"]
    n_ties = int(A["c3"][2])
    if ties is not None:
        n_ties = len(ties[0])
        if n_ties <= TIE_CAP:
            A["tie_r"][:n_ties] = ties[0]
            A["tie_s"][:n_ties] = ties[1]
    elif n_ties:
        n_ties = TIE_CAP + 1                                 # no list at hand: marked as "rides with the lists"
    rec = np.empty(5 + 2 * k + 3 + 2 * TIE_CAP, np.int64)
    check(lib().vq_summary_pack(int(first_row), P["c3"], k, n_top, P["t_rows"], P["t_sc"], nb_p, TIE_CAP, n_ties,
                                P["tie_r"], P["tie_s"], rec.ctypes.data), "vq_summary_pack")
    return rec


def exchange_summary(rec, k, dist, torch, device=None, mailbox=None):
    """all_gather of summary_record -> RankSummary (identical on every rank); through the host mailbox when there
    is one and the record fits its slots, else a collective of `dist`.  Merged by the library (vq_summary_merge)."""
    g = mailbox.all_gather(rec) if mailbox is not None and mailbox.fits(rec) else all_gather_np(rec, dist, torch, device)
    g = np.ascontiguousarray(g, dtype=np.int64)
    world, k = g.shape[0], int(k)
    sc = _MERGE.get((world, k, TIE_CAP))
    if sc is None:
        sc = _MERGE[(world, k, TIE_CAP)] = _Scratch(
            first=(world, np.int64), counts=((world, 3), np.int64), t_rows=(max(k, 1), np.int64), t_sc=(max(k, 1), np.float32),
            best=(3, np.int64), tie_r=(world * TIE_CAP, np.int64), tie_s=(world * TIE_CAP, np.float32))
    A, P = sc.a, sc.p
    n_top, n_ties = C.c_int32(), C.c_int64()
    check(lib().vq_summary_merge(g.ctypes.data, world, k, TIE_CAP, P["first"], P["counts"], P["t_rows"], P["t_sc"],
                                 C.byref(n_top), P["best"], P["tie_r"], P["tie_s"], C.byref(n_ties)), "vq_summary_merge")
    best = A["best"]
    near_best = None if best[1] < 0 else (int(best[0]), int(best[1]), float(best[2:3].astype(np.uint32).view(np.float32)[0]))
    ties = None if n_ties.value < 0 else (A["tie_r"][:n_ties.value].copy(), A["tie_s"][:n_ties.value].copy())
    return RankSummary(A["first"].copy(), A["counts"].copy(), (A["t_rows"][:n_top.value].copy(), A["t_sc"][:n_top.value].copy()),
                       near_best, ties)


_HOST_OUT = {}


def _host_out(torch, device, n):
    """(int64 [>= n], fp32 [>= n]) host tensors the gathered lists land in — pinned for a CUDA device, kept per device
    and grown geometrically (pinned allocations cost milliseconds)."""
    key = str(device)
    have = _HOST_OUT.get(key)
    if have is None or have[0].numel() < n:
        cap = max(1 << 16, 1 << int(max(n, 1) - 1).bit_length())
        pin = device.type == "cuda"
        have = (torch.empty(cap, dtype=torch.int64, pin_memory=pin), torch.empty(cap, dtype=torch.float32, pin_memory=pin))
        _HOST_OUT[key] = have
    return have


def gather_lists_torch(lists, which, summary, dist, torch, copy=True):
    """Several ordered lists of this rank -> the search set's, with ONE all_gather and no host staging on the way in:
    `lists` = [(rows, scores)] as torch tensors on the rank's device — int32 LOCAL rows and fp32 scores, exactly what
    the scan left in HBM — for the columns `which` of the summary's counts.  On the device: rows become int64 global
    rows, the lists go back to back into one buffer of two planes (12 bytes per entry, padded to the longest rank's
    total), one all_gather (NVLink), then each list's segments are concatenated in rank order = database order, which
    the reference's seeded sampling walks (ticket.py:326-341).  One device-to-host copy per list into pinned memory, one
    stream synchronisation.  copy=False returns views of that pinned memory, valid until the next call."""
    rank, world = dist.get_rank(), dist.get_world_size()
    counts = summary.counts[:, which]
    cap = int(counts.sum(axis=1).max(initial=0))
    cap += cap & 1                                            # int64 views of the gathered rows need even offsets
    if cap == 0:
        return [(np.empty(0, np.int64), np.empty(0, np.float32)) for _ in which]
    dev = lists[0][0].device
    buf = torch.zeros(3 * cap, dtype=torch.int32, device=dev)
    b_rows, b_scores = buf[:2 * cap].view(torch.int64), buf[2 * cap:].view(torch.float32)
    o = 0
    for (rows, scores), n in zip(lists, counts[rank]):
        n = int(n)
        if n:
            b_rows[o:o + n] = rows[:n].to(torch.int64) + int(summary.first_rows[rank])
            b_scores[o:o + n] = scores[:n]
        o += n
    out = torch.empty(world * 3 * cap, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, buf)
    g = out.view(world, 3 * cap)
    starts = np.concatenate([np.zeros((world, 1), np.int64), np.cumsum(counts, axis=1)], axis=1)
    h_rows, h_scores = _host_out(torch, dev, int(counts.sum()))
    spans, o = [], 0
    for j in range(len(which)):
        n_j = int(counts[:, j].sum())
        if n_j:
            seg = [(int(starts[r, j]), int(counts[r, j])) for r in range(world)]
            h_rows[o:o + n_j].copy_(torch.cat([g[r, :2 * cap].view(torch.int64)[a:a + n] for r, (a, n) in enumerate(seg)]),
                                    non_blocking=True)
            h_scores[o:o + n_j].copy_(torch.cat([g[r, 2 * cap:].view(torch.float32)[a:a + n] for r, (a, n) in enumerate(seg)]),
                                      non_blocking=True)
        spans.append((o, n_j))
        o += n_j
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()
    np_rows, np_scores = h_rows.numpy(), h_scores.numpy()
    res = [(np_rows[a:a + n], np_scores[a:a + n]) for a, n in spans]
    return [(r.copy(), s_.copy()) for r, s_ in res] if copy else res


def gather_lists_root(lists, which, summary, dist, torch, root=0, copy=True):
    """The same lists delivered to ONE rank only (the rank that talks to the API; nobody else consumes whole lists):
    every other rank sends its entries — exactly as many as it has, no padding — straight from device memory to `root`
    (NCCL send / recv over NVLink; gloo in the CPU tests); `root` receives all segments into one device buffer in rank
    order = database order, makes the rows global there and copies each list once into pinned host memory.  Returns the
    lists on `root` (copy=False: views of the pinned memory, valid until the next call) and None elsewhere.  Against the
    all-gather this moves 1/world of the bytes over the fabric and 1/world of them over PCIe, on one rank instead of all."""
    rank, world = dist.get_rank(), dist.get_world_size()
    counts = summary.counts[:, which]                          # [world, len(which)]
    per_rank = counts.sum(axis=1)
    n_mine = int(per_rank[rank])
    dev = lists[0][0].device
    # my segment: int32 local rows of all lists back to back, then their fp32 scores (bit pattern) — 8 bytes per entry
    mine = torch.empty(2 * n_mine, dtype=torch.int32, device=dev)
    o = 0
    for (rows, scores), n in zip(lists, counts[rank]):
        n = int(n)
        if n:
            mine[o:o + n] = rows[:n]
            mine[n_mine + o:n_mine + o + n] = scores[:n].view(torch.int32)
        o += n
    if rank != root:
        if n_mine:
            dist.send(mine, root)
        return None
    total = int(per_rank.sum())
    if total == 0:
        return [(np.empty(0, np.int64), np.empty(0, np.float32)) for _ in which]
    segs = [mine if r == rank else torch.empty(2 * int(per_rank[r]), dtype=torch.int32, device=dev) for r in range(world)]
    ops = [dist.P2POp(dist.irecv, segs[r], r) for r in range(world) if r != rank and int(per_rank[r])]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    h_rows, h_scores = _host_out(torch, dev, total)
    spans, o = [], 0
    starts = np.concatenate([np.zeros((world, 1), np.int64), np.cumsum(counts, axis=1)], axis=1)
    for j in range(len(which)):
        n_j = int(counts[:, j].sum())
        if n_j:
            r_parts, s_parts = [], []
            for r in range(world):
                a, n, nr = int(starts[r, j]), int(counts[r, j]), int(per_rank[r])
                if n:
                    r_parts.append(segs[r][a:a + n].to(torch.int64) + int(summary.first_rows[r]))
                    s_parts.append(segs[r][nr + a:nr + a + n].view(torch.float32))
            h_rows[o:o + n_j].copy_(torch.cat(r_parts), non_blocking=True)
            h_scores[o:o + n_j].copy_(torch.cat(s_parts), non_blocking=True)
        spans.append((o, n_j))
        o += n_j
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()
    np_rows, np_scores = h_rows.numpy(), h_scores.numpy()
    res = [(np_rows[a:a + n], np_scores[a:a + n]) for a, n in spans]
    return [(r.copy(), s_.copy()) for r, s_ in res] if copy else res


class ShardedLists:
    """The search set's ordered match / near-miss / tie lists left where the scan put them: every rank holds its own
    segment (global rows + fp32 scores in its pinned host mirror); the segments in rank order ARE the lists in database
    order, and the per-rank counts every rank received with the summary record map a list position to its owner without
    further traffic.  `local(which)` = this rank's segment; `span(which, rank)` = the positions it covers."""

    _COL = {"matches": 0, "near_misses": 1, "ties": 2}

    def __init__(self, summary, rank, local_lists):
        self.counts = summary.counts                           # [world, 3]
        self.offsets = np.concatenate([np.zeros((1, 3), np.int64), np.cumsum(summary.counts, axis=0)])[:-1]
        self.rank, self._local = rank, local_lists

    def local(self, which):
        return self._local[self._COL[which]]

    def span(self, which, rank=None):
        r, c = self.rank if rank is None else rank, self._COL[which]
        return int(self.offsets[r, c]), int(self.offsets[r, c] + self.counts[r, c])

    def total(self, which):
        return int(self.counts[:, self._COL[which]].sum())

    def owner(self, which, position):
        c = self._COL[which]
        return int(np.searchsorted(self.offsets[:, c] + self.counts[:, c], position, side="right"))


def gather_positions_multi(requests, summary, local_gather, dist, torch, device=None, mailbox=None):
    """requests: [(list column in the summary's counts, positions in the search set's list)], the same on every
    rank.  Each rank fetches the entries its own lists hold with `local_gather(column, local positions) -> (global
    rows, fp32 scores)`; ONE exchange of 16 bytes per position — an all-gather through the host mailbox summed here, or
    an all_reduce of `dist` — gives every rank all of them (x + 0 on integers: exact)."""
    rank = dist.get_rank()
    reqs = [(c, np.asarray(p, np.int64).reshape(-1)) for c, p in requests]
    out = np.zeros((sum(len(p) for _, p in reqs), 2), np.int64)
    o = 0
    for c, pos in reqs:
        n_list = summary.counts[:, c]
        if len(pos) and (pos.min() < 0 or pos.max() >= int(n_list.sum())):
            raise _ffi.VQError("gather: position outside the list of %d entries" % int(n_list.sum()))
        base = int(n_list[:rank].sum())
        sel = np.flatnonzero((pos >= base) & (pos < base + int(n_list[rank])))
        if len(sel):
            r, sc = local_gather(c, np.ascontiguousarray(pos[sel] - base))
            out[o + sel, 0] = r
            out[o + sel, 1] = _bits(sc)
        o += len(pos)
    if len(out) and mailbox is not None and mailbox.fits(out):
        out = mailbox.all_gather(out).sum(axis=0)
    elif len(out):
        t = _t(out, torch, device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out = t.cpu().numpy()
    res, o = [], 0
    for _, pos in reqs:
        res.append((out[o:o + len(pos), 0].copy(), _floats(out[o:o + len(pos), 1])))
        o += len(pos)
    return res


class RankStore:
    """One rank's shard of a search set as a FeatureStore (`first_global_row` = the start of its clip range) plus
    the collectives that turn per-rank results into the search set's.  Every rank calls every method (they are
    collectives) and every rank gets the same result.  A single-query call costs two exchanges: one fixed-size
    summary record per rank (RankSummary), then either the packed lists (scan) or the sampled entries (gather).
    The small ones go through the host mailbox when all ranks share a box, the lists through `dist`."""

    _COL = {"matches": 0, "near_misses": 1, "ties": 2}

    def __init__(self, store, dist, torch, device=None, host_mailbox=True):
        self.store, self.dist, self.torch, self.device = store, dist, torch, device
        self.world = dist.get_world_size()
        self.lo = store.first_global_row
        self.hi = store.first_global_row + store.n_rows
        self.summary = None
        # ranks of one box exchange their small per-query records through shared memory (collective setup)
        self.mailbox = HostMailbox.try_create(dist) if host_mailbox else None

    def close(self):
        if self.mailbox is not None:
            self.mailbox.close()
            self.mailbox = None

    def _summarise(self, res, k, near_best=None):
        rows, scores = self.store.topk() if k else (np.empty(0, np.int64), np.empty(0, np.float32))
        ties = self.store.ties(copy=False)
        rec = summary_record(self.lo, [res.n_match, res.n_near, res.n_tie], rows, scores, k, near_best, ties)
        self.summary = exchange_summary(rec, k, self.dist, self.torch, self.device, self.mailbox)
        return ties

    def _device_lists(self, columns):
        """The last scan's ordered lists as torch views of the library's device memory (no copy)."""
        torch, out = self.torch, []
        dl = self.store.device_lists()
        for c in columns:
            rows_p, scores_p, n = dl[c]
            if n == 0:
                out.append((torch.empty(0, dtype=torch.int32, device=self.device),
                            torch.empty(0, dtype=torch.float32, device=self.device)))
            else:
                out.append((torch.as_tensor(_DevArray(rows_p, n, "<i4"), device=self.device),
                            torch.as_tensor(_DevArray(scores_p, n, "<f4"), device=self.device)))
        return out

    def scan(self, target_features, weights, threshold, lower_limit, eps, topk=0, copy=True, lists="sharded", root=0):
        """Full single-query result: counts [match, near, tie] and the merged top-k on every rank, plus the ordered
        global match / near-miss / tie lists:
          lists="sharded" (default): each rank publishes ITS segment into its pinned host mirror exactly like a single-GPU
            scan and nothing else moves — the result is a ShardedLists (segments in rank order = the lists in database
            order; every rank knows every segment's length).  No consumer of the path needs the whole lists on every rank.
          lists="root": the whole lists on rank `root` only (the one that talks to the API; None elsewhere), sent device to
            device without padding (gather_lists_root).
          lists="all": replicated on every rank with one padded all-gather (gather_lists_torch) — world x the bytes.
        One exchange of a fixed-size summary record per rank (host mailbox) in every mode."""
        if lists == "sharded":
            res = self.store.scan(target_features, weights, threshold, lower_limit, eps, topk=topk, lists=True)
            self._summarise(res, int(topk))
            sm = self.summary
            local = [self.store.matches(copy=copy), self.store.near_misses(copy=copy), self.store.ties(copy=copy)]
            return sm.total, ShardedLists(sm, self.dist.get_rank(), local), sm.topk
        res = self.store.scan(target_features, weights, threshold, lower_limit, eps, topk=topk, lists=False)
        self._summarise(res, int(topk))
        sm = self.summary
        which = [0, 1] if sm.ties is not None else [0, 1, 2]
        if lists == "root":
            out = gather_lists_root(self._device_lists(which), which, sm, self.dist, self.torch, root=root, copy=copy)
            if out is not None and sm.ties is not None:
                out.append(sm.ties)
            return sm.total, out, sm.topk
        out = gather_lists_torch(self._device_lists(which), which, sm, self.dist, self.torch, copy=copy)
        if sm.ties is not None:
            out.append(sm.ties)
        return sm.total, out, sm.topk

    def scan_select(self, target_features, weights, threshold, lower_limit, eps, topk=0):
        """The review round's variant (ticket.py:311-356 samples a few dozen clips): every rank's match / near-miss
        lists stay on its device; what crosses ranks is one summary record per rank (counts, top-k, tie band, best
        near miss).  Returns (counts [match, near, tie], (top-k rows, scores), tie list, best near miss
        (position in the search set's near-miss list, global row, score) or None); fetch sampled list entries with
        gather() / gather_many()."""
        res = self.store.scan(target_features, weights, threshold, lower_limit, eps, topk=topk, lists=False)
        self._summarise(res, int(topk), self.store.near_best())
        sm = self.summary
        tie_list = sm.ties if sm.ties is not None else \
            gather_lists_torch(self._device_lists([2]), [2], sm, self.dist, self.torch)[0]
        return sm.total, sm.topk, tie_list, sm.near_best

    def gather_many(self, requests):
        """[(which, positions)] -> [(global rows, fp32 scores)]: entries of the search set's ordered match / near-miss /
        tie lists of the last scan at the given positions (the same on every rank), one collective for all requests."""
        names = {v: n for n, v in self._COL.items()}
        return gather_positions_multi([(self._COL[w], p) for w, p in requests], self.summary,
                                      lambda c, local: self.store.gather(names[c], local),
                                      self.dist, self.torch, self.device, self.mailbox)

    def gather(self, which, positions):
        return self.gather_many([(which, positions)])[0]

    def scan_batch(self, targets, weights, threshold, lower_limit, topk=0):
        c, r, s, ms = self.store.scan_batch(targets, weights, threshold, lower_limit, topk=topk)
        if topk == 0:
            r, s = np.empty((len(c), 0), np.int64), np.empty((len(c), 0), np.float32)
        return gather_batch(c, r, s, self.dist, self.torch, self.device) + (ms,)

    def labelled_sims(self, target_features, global_rows):
        rows = np.asarray(global_rows, np.int64)
        out = np.zeros((len(rows), len(self.store.streams)), np.float64)
        own = np.flatnonzero((rows >= self.lo) & (rows < self.hi))
        if len(own):
            out[own] = self.store.labelled_sims(target_features, rows[own])
        return gather_sims(out, self.dist, self.torch, self.device)

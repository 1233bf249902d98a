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

    def __init__(self, ptr_value, n_int64):
        self.__cuda_array_interface__ = {"shape": (n_int64,), "typestr": "<i8", "data": (ptr_value, False), "version": 3}


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
        """Lagged exchange: merge the last pushed step (enqueued; no host sync).  No-op for the other modes."""
        if self.exchange == "p2p-lagged":
            check(lib().vq_exchange_flush_enqueue(self._x, C.c_void_p(stream_ptr)), "vq_exchange_flush_enqueue")

    def kernels_per_step(self):
        return 4 + {"none": 0, "p2p": 1, "p2p-lagged": 1, "nccl": 1}[self.exchange]

    def result(self):
        """(counts[4], global top-k rows, scores) after the stream has been synchronised."""
        t = self.merged if self.world > 1 else self._payload_view()
        return unpack_payload(t.cpu().numpy(), self.k)


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


def gather_lists(rows, scores, dist, torch, device=None):
    """One rank's ordered match (or near-miss / tie) list -> the search set's list on every rank.
    Ranks own ascending clip ranges, so concatenating in rank order keeps database order — the order of the
    reference's `scores` dict that its seeded sampling walks (ticket.py:326-341).  Lengths differ per rank:
    counts are gathered first, then the lists padded to the longest one."""
    rows = np.asarray(rows, np.int64)
    scores = np.asarray(scores, np.float32)
    counts = all_gather_np(np.array([len(rows)], np.int64), dist, torch, device)[:, 0]
    cap = int(counts.max())
    if cap == 0:
        return np.empty(0, np.int64), np.empty(0, np.float32)
    pad = np.zeros((cap, 2), np.int64)
    pad[:len(rows), 0] = rows
    pad[:len(rows), 1] = scores.view(np.uint32)
    g = all_gather_np(pad, dist, torch, device)
    out_r = np.concatenate([g[r, :counts[r], 0] for r in range(len(counts))])
    out_s = np.concatenate([g[r, :counts[r], 1] for r in range(len(counts))]).astype(np.uint32).view(np.float32)
    return out_r, out_s


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


class RankStore:
    """One rank's shard of a search set as a FeatureStore (`first_global_row` = the start of its clip range) plus
    the collectives that turn per-rank results into the search set's.  Every rank calls every method (they are
    collectives) and every rank gets the same result."""

    def __init__(self, store, dist, torch, device=None):
        self.store, self.dist, self.torch, self.device = store, dist, torch, device
        self.world = dist.get_world_size()
        self.lo = store.first_global_row
        self.hi = store.first_global_row + store.n_rows

    def scan(self, target_features, weights, threshold, lower_limit, eps, topk=0):
        """Full single-query result with host buffers: counts [match, near, tie], ordered global match /
        near-miss / tie lists and the merged top-k."""
        res = self.store.scan(target_features, weights, threshold, lower_limit, eps, topk=topk)
        k = int(topk)
        rows, scores = self.store.topk() if k else (np.empty(0, np.int64), np.empty(0, np.float32))
        payload = pack_payload([res.n_match, res.n_near, res.n_tie, len(rows)], rows, scores, k)
        merged = merge_payloads_host(all_gather_np(payload, self.dist, self.torch, self.device), self.world, k)
        counts, t_rows, t_scores = unpack_payload(merged, k)
        lists = [gather_lists(*fn(copy=False), self.dist, self.torch, self.device)
                 for fn in (self.store.matches, self.store.near_misses, self.store.ties)]
        return counts[:3], lists, (t_rows, t_scores)

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

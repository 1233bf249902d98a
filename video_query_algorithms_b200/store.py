"""HBM-resident feature store and its scan API (host side of include/vq.h).

Replaces the per-job HTTP pull of `Ticket._get_candidate_features` (reference
src/models/ticket.py:358-382): a search set's feature rows are laid out once, clip-major
`[clip][stream][split][dim]` fp32, sharded by contiguous clip range over the GPUs of the box, and
stay resident across broker ticks.  Row order is the order of first appearance of each clip id in
the `search-sets/features` response — the insertion order of the reference's `scores` dict, which
its seeded `random.sample` depends on (ticket.py:326-333).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi
from ._ffi import ScanCounts, ScanParams, VQError, check, lib, ptr


@dataclass
class ScanResult:
    n_match: int
    n_near: int
    n_tie: int
    n_topk: int
    scan_ms: float                 # device time of the scan kernel, max over shards


class Shard:
    """One vq_store on one device holding global rows [first, first + n)."""

    def __init__(self, device, n_rows, n_streams, n_splits, dim, first_global_row=0):
        self.handle = C.c_void_p()
        check(lib().vq_store_create(C.byref(self.handle), device, n_rows, n_streams, n_splits, dim,
                                    first_global_row), "vq_store_create")
        self.device, self.n_rows, self.first = device, n_rows, first_global_row
        self.n_streams, self.n_splits, self.dim = n_streams, n_splits, dim

    def close(self):
        if self.handle:
            lib().vq_store_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_ptr(self):
        p = C.c_void_p()
        check(lib().vq_store_device_ptr(self.handle, C.byref(p)), "vq_store_device_ptr")
        return p.value


def make_params(weights, threshold, lower_limit, eps, topk=0, want_sims=False):
    p = ScanParams()
    for i, w in enumerate(weights):
        p.weights[i] = float(w)
    p.threshold, p.lower_limit, p.eps = float(threshold), float(lower_limit), float(eps)
    p.topk, p.want_sims = int(topk), int(bool(want_sims))
    return p


class RecordIndex:
    """What one pass over a `search-sets/features`-style response says about the rows it fills (host code, no device
    call, no vector touched yet): clip ids in row order, the split numbers, the feature length, which (row, stream,
    split) slots are present and — per slot — the record that fills it."""

    def __init__(self, order, splits, dim, n_streams, rec, row, si, pi, first=None):
        self.order, self.splits, self.dim, self.n_streams = order, splits, int(dim), int(n_streams)
        self.n_rows = len(order)
        self.rec, self.row, self.si, self.pi = rec, row, si, pi            # one entry per filled slot, sorted by row
        self.row_floats = self.n_streams * len(splits) * self.dim
        self.present = np.zeros((self.n_rows, self.n_streams, len(splits)), bool)
        self.present[row, si, pi] = True
        # position of each slot's FIRST record in the response (a dict keeps a key where it was first inserted, whatever
        # overwrites its value later): what the reference's per-job dict order is made of (FeatureStore.dict_order)
        self.first = rec if first is None else first

    def slot_positions(self):
        """int64 [n_rows, S, P]: response position of each slot's first record (-1: absent), or None when the order of
        the rows is the same whatever split a job's target starts with — every slot present and every (stream, split)
        listing the clips in row order — which is the case for everything load_db.py writes in one go."""
        pos = np.full(self.present.shape, -1, np.int64)
        pos[self.row, self.si, self.pi] = self.first
        if self.present.all() and (self.n_rows < 2 or bool(np.all(pos[1:] > pos[:-1]))):
            return None
        return pos

    def fill(self, feature_rows, r0, r1, out, threads=0):
        """float32(feature vectors) of rows [r0, r1) into `out` (float32 [>= (r1 - r0) * row_floats], C-contiguous); slots
        no record fills are zeroed.  The unboxing runs in libvq_pyhost (several threads, csrc/vq_pyhost.c)."""
        lo, hi = np.searchsorted(self.row, [r0, r1])
        n_slots = (r1 - r0) * self.n_streams * len(self.splits)
        flat = out.reshape(-1)[:(r1 - r0) * self.row_floats]
        if hi - lo != n_slots:
            flat[:] = 0
        if hi == lo:
            return
        P = len(self.splits)
        dest = np.ascontiguousarray((((self.row[lo:hi] - r0) * self.n_streams + self.si[lo:hi]) * P + self.pi[lo:hi])
                                    * self.dim, dtype=np.int64)
        rec = np.ascontiguousarray(self.rec[lo:hi], dtype=np.int64)
        if threads <= 0:
            import os
            threads = min(len(os.sched_getaffinity(0)), 16)
        _ffi.check_py(_ffi.pyhost().vq_py_fill_chunk(feature_rows, _ffi.ptr(rec), _ffi.ptr(dest), len(rec), self.dim,
                                                     _ffi.ptr(flat), threads), "vq_py_fill_chunk")


def index_feature_rows(feature_rows, streams, feature_name, held=None):
    """One pass over `search-sets/features`-style records -> RecordIndex (and the records as a list).
    Applies the reference's filters (ticket.py:374-381: stream in `streams`, name == `feature_name`) and its dict
    semantics: clips in order of first appearance (= the insertion order of the reference's `scores` dict, which its
    seeded sampling walks), a later record of the same (clip, stream, split) overwrites an earlier one, slots no record
    fills are absent.  `held(clip ids) -> bool array` drops clips a store already holds (append).
    Row order on ragged data: the reference's `scores` dict is filled while it walks the FIRST stream split by split
    (ticket.py:146-160), so a clip enters at its first record of (first stream, lowest split it has there) — for
    complete data simply the order of first appearance, otherwise clips that lack the first split come after all
    clips that have it.  Clips without any first-stream record (the reference raises KeyError for them) go last."""
    if not isinstance(feature_rows, list):
        feature_rows = list(feature_rows)
    streams = tuple(streams)
    n = len(feature_rows)
    clip, stream = np.empty(n, np.int64), np.empty(n, np.int32)
    split, length = np.empty(n, np.int32), np.empty(n, np.int32)
    _ffi.check_py(_ffi.pyhost().vq_py_index_records(feature_rows, streams, feature_name, _ffi.ptr(clip), _ffi.ptr(stream),
                                                    _ffi.ptr(split), _ffi.ptr(length)), "vq_py_index_records")
    rec = np.flatnonzero(stream >= 0)                          # kept records, in response order
    clip, si, sp, length = clip[rec], stream[rec].astype(np.int64), split[rec].astype(np.int64), length[rec]
    u, first_idx, inv = np.unique(clip, return_index=True, return_inverse=True)
    by_first = np.argsort(first_idx, kind="stable")
    rank = np.empty(len(u), np.int64)
    rank[by_first] = np.arange(len(u))
    row = rank[inv]                                            # row of each kept record: first appearance of its clip
    order = u[by_first]
    n_rows = len(order)
    # ragged first stream: re-seat the rows by (lowest first-stream split, position of that record)
    s0 = np.flatnonzero(si == 0)
    key0 = np.full(n_rows, np.inf)
    key1 = np.arange(n_rows, dtype=np.float64)
    if len(s0):
        o = np.lexsort((s0, sp[s0], row[s0]))
        r_sorted = row[s0][o]
        firsts = np.concatenate([[True], r_sorted[1:] != r_sorted[:-1]])
        er = r_sorted[firsts]
        key0[er] = sp[s0][o][firsts]
        key1[er] = s0[o][firsts]
    if n_rows > 1 and (np.any(key0[1:] < key0[:-1]) or np.any((key0[1:] == key0[:-1]) & (key1[1:] < key1[:-1]))):
        perm = np.lexsort((key1, key0))
        seat = np.empty(n_rows, np.int64)
        seat[perm] = np.arange(n_rows)
        order, row = order[perm], seat[row]
    if held is not None and n_rows:
        keep = ~np.asarray(held(order.tolist()), dtype=bool)
        remap = np.where(keep, np.cumsum(keep) - 1, -1)
        order = order[keep]
        row = remap[row]
        live = row >= 0
        rec, row, si, sp, length = rec[live], row[live], si[live], sp[live], length[live]
        n_rows = len(order)
    splits = sorted(set(sp.tolist()))
    dim = int(length[0]) if len(length) else 0
    pi = np.searchsorted(np.asarray(splits, np.int64), sp) if len(sp) else sp
    # one record per slot — the last one in response order — sorted by row
    slot = (row * len(streams) + si) * max(len(splits), 1) + pi
    first = rec
    if len(slot):
        uniq, first_idx = np.unique(slot, return_index=True)          # first record of each slot: its place in the reference's dicts
        _, last_rev = np.unique(slot[::-1], return_index=True)       # last record of each slot: its value
        keep = len(slot) - 1 - last_rev                               # (both in ascending slot order)
        o = np.argsort(row[keep], kind="stable")
        first = rec[first_idx][o]
        keep = keep[o]
        rec, row, si, pi = rec[keep], row[keep], si[keep], pi[keep]
    return RecordIndex(order, splits, dim, len(streams), rec, row, si, pi, first), feature_rows


def pack_feature_rows(feature_rows, streams, feature_name, held=None):
    """`search-sets/features`-style records -> the store's row layout in host memory (see index_feature_rows for the
    filters and the row order).  Returns (clip ids in row order, sorted split numbers, X float32 [n, S, P, dim],
    present bool [n, S, P]).  FeatureStore.from_feature_rows does the same through pinned chunks without ever holding X."""
    idx, feature_rows = index_feature_rows(feature_rows, streams, feature_name, held=held)
    order = idx.order.tolist()
    if not order:
        return order, idx.splits, np.zeros((0, len(streams), len(idx.splits), 0), np.float32), np.zeros((0, len(streams), len(idx.splits)), bool)
    X = np.empty((idx.n_rows, len(streams), len(idx.splits), idx.dim), np.float32)
    idx.fill(feature_rows, 0, idx.n_rows, X)
    return order, idx.splits, X, idx.present


class FeatureStore:
    """A search set's features in HBM, as one shard per device."""

    def __init__(self, n_rows, streams, splits, dim=1024, devices=None, clip_ids=None,
                 first_global_row=0):
        n_dev = C.c_int()
        check(lib().vq_device_count(C.byref(n_dev)), "vq_device_count")
        if n_dev.value < 1:
            raise VQError("no CUDA device visible: the scoring path has no CPU fallback")
        if devices is None:
            devices = list(range(n_dev.value)) if n_rows >= 65536 * n_dev.value else [0]
        self.streams = tuple(streams)
        self.splits = [int(p) for p in splits]
        self.dim, self.n_rows, self.first_global_row = int(dim), int(n_rows), int(first_global_row)
        self.row_shape = (len(self.streams), len(self.splits), self.dim)
        per = -(-self.n_rows // len(devices)) if self.n_rows else 0
        self.shards = []
        for i, dev in enumerate(devices):
            lo = min(i * per, self.n_rows)
            hi = min(lo + per, self.n_rows)
            if hi > lo or i == 0:
                self.shards.append(Shard(dev, hi - lo, len(self.streams), len(self.splits), self.dim,
                                         self.first_global_row + lo))
        self.clip_ids = None
        self._row_of = None
        self._row_memo = {}
        self.present = None          # bool [N, S, P] when some clip lacks some split, else None
        self.slot_pos = None         # int64 [N, S, P] response positions when the dict order depends on the target (dict_order)
        if clip_ids is not None:
            self.set_clip_ids(clip_ids)
        self.last = None

    # ------------------------------------------------------------------ construction
    def set_clip_ids(self, clip_ids):
        self.clip_ids = np.asarray(clip_ids, dtype=np.int64)
        assert self.clip_ids.shape == (self.n_rows,)
        self._row_of = None
        self._row_memo = {}

    def _index(self):
        """(sorted clip ids, their local rows): the id -> row table as two arrays, looked up by binary search.  A search
        set of a million clips makes a Python dict a 100 ms, 100 MB affair; an id table that is already ascending (the
        usual case: ids are issued in insertion order) needs no work at all."""
        if self._row_of is None:
            ids = self.clip_ids
            if ids.size < 2 or bool(np.all(ids[1:] > ids[:-1])):
                self._row_of = (ids, None)
            else:
                order = np.argsort(ids, kind="stable")
                self._row_of = (ids[order], order)
        return self._row_of

    def _lookup(self, clip_ids):
        """local rows of the given ids (-1 where the store does not hold the id); vectorised."""
        sorted_ids, order = self._index()
        q = np.asarray(clip_ids, dtype=np.int64).reshape(-1)
        pos = np.searchsorted(sorted_ids, q)
        pos_c = np.minimum(pos, max(len(sorted_ids) - 1, 0))
        hit = (len(sorted_ids) > 0) & (sorted_ids[pos_c] == q) if len(sorted_ids) else np.zeros(len(q), bool)
        rows = pos_c if order is None else order[pos_c]
        return np.where(hit, rows, -1)

    def _row_or_minus1(self, clip_id):
        """scalar lookup, memoised (labelled clips are looked up again and again across rounds)"""
        memo = self.__dict__.setdefault("_row_memo", {})
        c = int(clip_id)
        r = memo.get(c)
        if r is None:
            r = int(self._lookup([c])[0])
            if len(memo) < 1_000_000:
                memo[c] = r
        return r

    def row_of(self, clip_id):
        r = self._row_or_minus1(clip_id)
        if r < 0:
            raise KeyError(clip_id)
        return r

    def has_clip(self, clip_id):
        return clip_id is not None and self._row_or_minus1(clip_id) >= 0

    def rows_of(self, clip_ids):
        rows = self._lookup(clip_ids)
        if (rows < 0).any():
            raise KeyError(int(np.asarray(clip_ids).reshape(-1)[int(np.argmax(rows < 0))]))
        return (self.first_global_row + rows).astype(np.int64)

    def _shard_of(self, global_row):
        for sh in self.shards:
            if sh.first <= global_row < sh.first + sh.n_rows:
                return sh
        raise VQError("row %d is not in the store" % global_row)

    def _check_range(self, first_row, n_rows, who):
        if first_row < 0 or n_rows < 0 or first_row + n_rows > self.n_rows:
            raise VQError("%s: rows [%d, %d) outside the store of %d rows" % (who, first_row, first_row + n_rows,
                                                                              self.n_rows))

    def upload(self, first_row, rows):
        """rows: float32 [n, S, P, dim] (or [n, S*P*dim]) for local rows first_row.. of the store."""
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, int(np.prod(self.row_shape)))
        self._check_range(first_row, rows.shape[0], "upload")
        g0 = self.first_global_row + first_row
        for sh in self.shards:
            lo, hi = max(g0, sh.first), min(g0 + rows.shape[0], sh.first + sh.n_rows)
            if hi > lo:
                part = rows[lo - g0:hi - g0]
                check(lib().vq_store_upload(sh.handle, lo - sh.first, hi - lo, ptr(part)), "vq_store_upload")

    def download(self, first_row, n_rows):
        self._check_range(first_row, n_rows, "download")
        out = np.empty((n_rows,) + self.row_shape, np.float32)
        g0 = self.first_global_row + first_row
        for sh in self.shards:
            lo, hi = max(g0, sh.first), min(g0 + n_rows, sh.first + sh.n_rows)
            if hi > lo:
                part = np.empty((hi - lo,) + self.row_shape, np.float32)
                check(lib().vq_store_download(sh.handle, lo - sh.first, hi - lo, ptr(part)), "vq_store_download")
                out[lo - g0:hi - g0] = part
        return out

    def set_present(self, present):
        """present: bool [N, S, P]; slots that are False must have been uploaded as zeros."""
        present = np.asarray(present, dtype=bool)
        self._present_version = getattr(self, "_present_version", 0) + 1
        self._eff_key = None                                    # the device table is rewritten below: no target-specific one is loaded
        if present.all():
            self.present = None
            for sh in self.shards:
                check(lib().vq_store_set_split_weights(sh.handle, None), "vq_store_set_split_weights")
            return
        self.present = present
        self._apply_split_weights(present.sum(axis=2))
        self._eff_key = (np.ones(self.row_shape[:2], bool).tobytes(), self._present_version)

    def _apply_split_weights(self, counts):
        if (counts == 0).any():
            raise VQError("a clip has no feature row at all for one stream; the reference raises KeyError "
                          "for such a search set (ticket.py:177)")
        inv = (1.0 / counts).astype(np.float32)
        for sh in self.shards:
            lo = sh.first - self.first_global_row
            part = np.ascontiguousarray(inv[lo:lo + sh.n_rows])
            check(lib().vq_store_set_split_weights(sh.handle, ptr(part)), "vq_store_set_split_weights")

    # ------------------------------------------------------------------ incremental growth
    def append(self, rows, clip_ids=None, present=None):
        """Append clips after the last row (load_db.py adds clips to a search set, reference load_db.py:10-28).
        rows: float32 [n, S, P, dim]; clip_ids: their ids (required when the store has an id table);
        present: bool [n, S, P] when some new clip lacks some split.  The last shard grows (geometric
        reallocation inside the library), everything resident stays resident."""
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, int(np.prod(self.row_shape)))
        n_new = rows.shape[0]
        if n_new == 0:
            return
        if self.clip_ids is not None:
            if clip_ids is None or len(clip_ids) != n_new:
                raise VQError("append: the store has a clip-id table; pass the %d new clip ids" % n_new)
            new_ids = np.asarray(clip_ids, dtype=np.int64)
            dup = [int(c) for c in new_ids[self._lookup(new_ids) >= 0]]
            if dup or len(set(new_ids.tolist())) != n_new:
                raise VQError("append: clip ids already in the store or repeated: %s" % (dup[:5] or "repeated ids"))
        sh = self.shards[-1]
        check(lib().vq_store_append(sh.handle, n_new, ptr(rows)), "vq_store_append")
        sh.n_rows += n_new
        self.n_rows += n_new
        if self.clip_ids is not None:
            self.clip_ids = np.concatenate([self.clip_ids, new_ids])
            self._row_of = None
            self._row_memo = {}
        self._eff_key = None                                    # the library drops the per-row split weights on an append
        self._present_version = getattr(self, "_present_version", 0) + 1
        if self.present is not None or present is not None:
            old = self.present if self.present is not None else np.ones((self.n_rows - n_new,) + self.row_shape[:2], bool)
            new = np.ones((n_new,) + self.row_shape[:2], bool) if present is None else np.asarray(present, dtype=bool)
            self.set_present(np.concatenate([old, new]))
        self.last = None

    def append_feature_rows(self, feature_rows, feature_name):
        """Append the clips of a `search-sets/features`-style response that the store does not hold yet (same
        filters and first-appearance order as from_feature_rows); returns the number of clips added."""
        idx, feature_rows = index_feature_rows(feature_rows, self.streams, feature_name, held=lambda ids: self._lookup(ids) >= 0)
        return self._append_indexed(idx, feature_rows)

    def _append_indexed(self, idx, feature_rows):
        if not idx.n_rows:
            return 0
        extra = [p for p in idx.splits if p not in self.splits]
        if extra:
            raise VQError("append_feature_rows: split %d is not one of the store's splits %s" % (extra[0], self.splits))
        if idx.dim != self.dim:
            raise VQError("append_feature_rows: feature length %d, the store holds %d" % (idx.dim, self.dim))
        X = np.empty((idx.n_rows, len(self.streams), len(idx.splits), idx.dim), np.float32)
        idx.fill(feature_rows, 0, idx.n_rows, X)
        present = idx.present
        if idx.splits != self.splits:                               # the new clips lack some split: widen to the store's slots
            at = [self.splits.index(p) for p in idx.splits]
            Xw = np.zeros((idx.n_rows,) + self.row_shape, np.float32)
            pw = np.zeros((idx.n_rows,) + self.row_shape[:2], bool)
            Xw[:, :, at], pw[:, :, at] = X, present
            X, present = Xw, pw
        old_pos, n_old = getattr(self, "slot_pos", None), self.n_rows
        self.append(X, clip_ids=idx.order, present=present)
        new_pos = idx.slot_positions()
        if old_pos is not None or new_pos is not None:
            # appended clips come after everything held, in every (stream, split) listing
            if old_pos is None:
                old_pos = np.tile(np.arange(n_old, dtype=np.int64)[:, None, None], (1,) + self.row_shape[:2])
            if new_pos is None:
                new_pos = np.tile(np.arange(idx.n_rows, dtype=np.int64)[:, None, None], (1, len(self.streams), len(idx.splits)))
            wide = np.full((idx.n_rows,) + self.row_shape[:2], -1, np.int64)
            wide[:, :, [self.splits.index(p) for p in idx.splits]] = np.where(new_pos >= 0, new_pos + int(old_pos.max()) + 1, -1)
            self.slot_pos = np.concatenate([old_pos, wide])
        return idx.n_rows

    def sync_feature_rows(self, feature_rows, feature_name):
        """Bring the resident store up to date with a fresh `search-sets/features` response (load_db.py adds clips to a
        search set between ticks, reference load_db.py:10-28; the reference itself re-reads everything per job).
        Returns the number of clips appended (0: nothing new), or None when the response is not the store's rows plus
        new ones at the end — clips removed, reordered, other splits — and the caller must rebuild from the response."""
        idx, feature_rows = index_feature_rows(feature_rows, self.streams, feature_name)
        n_old = self.n_rows
        if idx.n_rows < n_old or not np.array_equal(idx.order[:n_old], self.clip_ids) or \
                any(p not in self.splits for p in idx.splits) or (idx.n_rows and idx.dim != self.dim):
            return None
        at = [self.splits.index(p) for p in idx.splits]
        old_present = np.ones((n_old,) + self.row_shape[:2], bool) if self.present is None else self.present
        new_present = np.zeros((n_old,) + self.row_shape[:2], bool)
        new_present[:, :, at] = idx.present[:n_old]
        if not np.array_equal(old_present, new_present):
            return None                                             # an old clip gained or lost a feature row
        if idx.n_rows == n_old:
            return 0
        tail = idx.row >= n_old
        sub = RecordIndex(idx.order[n_old:], idx.splits, idx.dim, idx.n_streams, idx.rec[tail], idx.row[tail] - n_old,
                          idx.si[tail], idx.pi[tail])
        return self._append_indexed(sub, feature_rows)

    def fill_synthetic(self, seed, means=None):
        m = None if means is None else np.asarray(means, dtype=np.float32)
        for sh in self.shards:
            check(lib().vq_store_fill_synthetic(sh.handle, int(seed), ptr(m)), "vq_store_fill_synthetic")

    CHUNK_BYTES = 64 << 20           # pinned staging per buffer of the ingest pipeline (two buffers)

    @classmethod
    def from_feature_rows(cls, feature_rows, streams, feature_name, devices=None):
        """Build from the `search-sets/features` API response (list of feature dicts), applying the
        reference's filters (ticket.py:374-381: stream in streams, name == feature_name).  The rows never exist as one
        host array: they are unboxed chunk by chunk into pinned staging while the previous chunk is on its way to HBM."""
        streams = tuple(streams)
        idx, feature_rows = index_feature_rows(feature_rows, streams, feature_name)
        if not idx.n_rows:
            raise VQError("search set has no '%s' features for streams %s" % (feature_name, streams))
        st = cls(idx.n_rows, streams, idx.splits, idx.dim, devices=devices, clip_ids=idx.order)
        st._ingest(idx, feature_rows)
        st.set_present(idx.present)
        st.slot_pos = idx.slot_positions()
        return st

    def _ingest(self, idx, feature_rows):
        """Two pinned buffers: fill one (libvq_pyhost, several threads) while the other is in flight (one
        cudaMemcpyAsync per chunk and shard); a buffer is refilled only after its copy has completed."""
        rows_per = max(1, min(idx.n_rows, self.CHUNK_BYTES // (idx.row_floats * 4)))
        bufs = [self._alloc_staging(rows_per * idx.row_floats) for _ in range(2 if idx.n_rows > rows_per else 1)]
        try:
            for i, r0 in enumerate(range(0, idx.n_rows, rows_per)):
                r1 = min(r0 + rows_per, idx.n_rows)
                buf = bufs[i % len(bufs)]
                if i >= len(bufs):
                    self._sync_uploads()                            # the copy out of this buffer (chunk i - 2) must be done;
                idx.fill(feature_rows, r0, r1, buf)                 # chunk i - 1 is at most still in flight while this fills
                self._upload_async(r0, buf[:(r1 - r0) * idx.row_floats])
            self._sync_uploads()
        finally:
            for b in bufs:
                self._free_staging(b)

    def upload_pipelined(self, first_row, rows):
        """upload() for large host arrays: the rows go through two pinned staging buffers, the copy into one overlapping
        the transfer out of the other (a pageable source is otherwise staged by the driver in small synchronous pieces)."""
        rf = int(np.prod(self.row_shape))
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, rf)
        n = rows.shape[0]
        self._check_range(first_row, n, "upload_pipelined")
        rows_per = max(1, self.CHUNK_BYTES // (rf * 4))
        if n <= rows_per:
            return self.upload(first_row, rows)
        bufs = [self._alloc_staging(rows_per * rf) for _ in range(2)]
        try:
            for i, r0 in enumerate(range(0, n, rows_per)):
                r1 = min(r0 + rows_per, n)
                buf = bufs[i % 2]
                if i >= 2:
                    self._sync_uploads()
                buf[:(r1 - r0) * rf] = rows[r0:r1].reshape(-1)
                self._upload_async(first_row + r0, buf[:(r1 - r0) * rf])
            self._sync_uploads()
        finally:
            for b in bufs:
                self._free_staging(b)

    # staging / asynchronous upload primitives (test doubles replace these four).  Page-locking memory costs milliseconds
    # per buffer, so the last two staging buffers are kept for the next build or append (at most 2 x CHUNK_BYTES).
    _STAGING_POOL = []               # [(address, n_floats)] of idle pinned buffers, process-wide

    def _alloc_staging(self, n_floats):
        n_floats = int(n_floats)
        pool = FeatureStore._STAGING_POOL
        for i, (addr, cap) in enumerate(pool):
            if cap >= n_floats:
                pool.pop(i)
                break
        else:
            p = C.c_void_p()
            check(lib().vq_pinned_alloc(C.byref(p), n_floats * 4), "vq_pinned_alloc")
            addr, cap = p.value, n_floats
        whole = np.ctypeslib.as_array(C.cast(C.c_void_p(addr), C.POINTER(C.c_float)), shape=(cap,))
        self.__dict__.setdefault("_staging_caps", {})[addr] = cap
        return whole[:n_floats]

    def _free_staging(self, a):
        addr = a.ctypes.data
        cap = self.__dict__.get("_staging_caps", {}).pop(addr, None)
        if cap is None:
            return
        pool = FeatureStore._STAGING_POOL
        pool.append((addr, cap))
        pool.sort(key=lambda e: -e[1])
        while len(pool) > 2:                                      # keep the two largest
            lib().vq_pinned_free(C.c_void_p(pool.pop()[0]))

    def _upload_async(self, first_row, flat):
        n = len(flat) // int(np.prod(self.row_shape))
        g0 = self.first_global_row + first_row
        rf = int(np.prod(self.row_shape))
        for sh in self.shards:
            lo, hi = max(g0, sh.first), min(g0 + n, sh.first + sh.n_rows)
            if hi > lo:
                part = flat[(lo - g0) * rf:(hi - g0) * rf]
                check(lib().vq_store_upload_async(sh.handle, lo - sh.first, hi - lo, ptr(part)), "vq_store_upload_async")

    def _sync_uploads(self):
        for sh in self.shards:
            check(lib().vq_store_sync(sh.handle), "vq_store_sync")

    def close(self):
        for sh in self.shards:
            sh.close()
        self.shards = []

    # ------------------------------------------------------------------ target layout
    def pack_target(self, target_features, dtype=np.float32):
        """{stream: {split: vector}} -> [S, P, dim] in the store's slot order; splits the target
        lacks are zero (they then contribute nothing, ticket.py:149-152)."""
        T = np.zeros(self.row_shape, dtype)
        have = np.zeros(self.row_shape[:2], bool)
        for si, s in enumerate(self.streams):
            for pi, p in enumerate(self.splits):
                v = target_features.get(s, {}).get(p)
                if v is not None:
                    T[si, pi] = np.asarray(v, dtype=np.float64)
                    have[si, pi] = True
        return T, have

    def dict_order(self, target_features):
        """Position of every row in the reference's `scores` dict for a job with this target, or None when that is the row
        order itself.  The reference fills the dict while it walks the target's streams and, per stream, the target's splits
        IN THE TARGET'S ORDER, listing each split's clips in response order (ticket.py:146-160): a clip enters at the first
        such listing that has it.  The store's rows follow the ascending-split walk, which is what every bootstrapped target
        and every reference clip whose records arrive in split order produce; on a ragged search set a reference clip whose
        `video-clips/features` response lists, say, split 2 before split 1 seats the clips differently, and the seeded
        sampling (ticket.py:333,341) walks THAT order."""
        pos = getattr(self, "slot_pos", None)
        if pos is None or not self.n_rows:
            return None
        n = self.n_rows
        key_stream = np.full(n, len(self.streams), np.int64)
        key_split = np.zeros(n, np.int64)
        key_pos = np.arange(n, dtype=np.int64)
        for si in reversed(range(len(self.streams))):
            by_split = target_features.get(self.streams[si], {})
            walk = [self.splits.index(p) for p in by_split if p in self.splits]          # the target's own split order
            for j in reversed(range(len(walk))):
                has = pos[:, si, walk[j]] >= 0
                key_stream[has], key_split[has], key_pos[has] = si, j, pos[has, si, walk[j]]
        order = np.lexsort((key_pos, key_split, key_stream))
        if np.array_equal(order, np.arange(n)):
            return None
        place = np.empty(n, np.int64)
        place[order] = np.arange(n)
        return place

    def _sync_split_weights_for_target(self, have):
        """The reference averages over splits that BOTH the target and the clip have (ticket.py:146-160), so the per-row
        1/n_splits table on the device depends on the target of the job at hand.  The store is shared by all jobs of a
        search set: what an earlier job uploaded is replaced — or dropped, when this target and every clip have every
        split — before anything is scanned.  `_eff_key` names what the device holds (None = the default 1/n_splits)."""
        if have.all() and self.present is None:
            if getattr(self, "_eff_key", None) is not None:
                for sh in self.shards:
                    check(lib().vq_store_set_split_weights(sh.handle, None), "vq_store_set_split_weights")
                self._eff_key = None
            return
        key = (have.tobytes(), getattr(self, "_present_version", 0))
        if getattr(self, "_eff_key", None) != key:
            present = np.ones((self.n_rows,) + self.row_shape[:2], bool) if self.present is None else self.present
            self._apply_split_weights((present & have[None]).sum(axis=2))
            self._eff_key = key

    # ------------------------------------------------------------------ scan
    def scan(self, target_features, weights, threshold, lower_limit, eps, topk=0, want_sims=False, lists=True, packed=None):
        """One fused pass: similarities, scores, match / near-miss / tie lists, top-k.
        lists=False is the selection round's variant: the match and near-miss lists stay on the device (only counts,
        top-k, tie band and the best near miss come back); fetch sampled entries with gather().
        packed = pack_target(target_features, float32) from an earlier call (a job scans one target several times, and
        turning 2 x 1024 boxed floats into an array costs as much host time as 5 % of a 1M-clip scan)."""
        T, have = packed if packed is not None else self.pack_target(target_features, np.float32)
        self._sync_split_weights_for_target(have)
        w = [weights[s] for s in self.streams] if isinstance(weights, dict) else list(weights)
        p = make_params(w, threshold, lower_limit, eps, topk, want_sims)
        Tc = np.ascontiguousarray(T)
        # enqueue on every shard first (each has its own stream), then wait: shards run concurrently
        self._near_best = [None] * len(self.shards)
        self._merged_topk = None
        if len(self.shards) == 1:
            c = ScanCounts()
            self._scan_shard(0, Tc, p, c, lists)
            counts = [c]
        else:
            counts = self._scan_multi(Tc, p, lists)
        self._lists_on_host = bool(lists)
        self.last = ScanResult(sum(c.n_match for c in counts), sum(c.n_near for c in counts),
                               sum(c.n_tie for c in counts), min(int(topk), sum(c.n_topk for c in counts)),
                               max(c.scan_ms for c in counts))
        self._last_counts = counts
        self._last_topk = topk
        return self.last

    def _scan_shard(self, i, Tc, p, c, lists):
        if lists:
            check(lib().vq_scan(self.shards[i].handle, ptr(Tc), C.byref(p), C.byref(c)), "vq_scan")
        else:
            pos, row, sc = C.c_int64(), C.c_int64(), C.c_float()
            check(lib().vq_scan_select(self.shards[i].handle, ptr(Tc), C.byref(p), C.byref(c), C.byref(pos), C.byref(row),
                                       C.byref(sc)), "vq_scan_select")
            self._near_best[i] = (pos.value, row.value, sc.value)

    def _scan_multi(self, Tc, p, lists=True):
        """All shards with ONE library call from this thread (vq_scan_multi): the work is enqueued on every shard's
        stream before any of them is waited for, and the shards' top-k lists come back merged."""
        n = len(self.shards)
        handles = (C.c_void_p * n)(*[sh.handle for sh in self.shards])
        counts = (ScanCounts * n)()
        near = np.empty((n, 3), np.int64)
        k = max(int(p.topk), 1)
        rows, scores, n_top = np.empty(k, np.int64), np.empty(k, np.float32), C.c_int32()
        check(lib().vq_scan_multi(handles, n, ptr(Tc), C.byref(p), int(bool(lists)), counts, ptr(near), k, ptr(rows), ptr(scores),
                                  C.byref(n_top)), "vq_scan_multi")
        if not lists:
            self._near_best = [(int(a), int(b), float(np.array([c], np.int64).astype(np.uint32).view(np.float32)[0]))
                               for a, b, c in near]
        self._merged_topk = (rows[:n_top.value], scores[:n_top.value])
        return list(counts)

    def device_lists(self):
        """Device addresses of the last scan's ordered lists on a single-shard store: [(rows pointer (uint32 LOCAL rows),
        scores pointer (fp32), entries)] for matches, near misses, tie band — for consumers that move the lists between
        devices without a host round trip (sharded.RankStore).  Valid until the next scan on this store."""
        if len(self.shards) != 1:
            raise VQError("device_lists: single-shard stores only (one store per rank)")
        v = _ffi.ScanDeviceView()
        check(lib().vq_scan_view(self.shards[0].handle, C.byref(v)), "vq_scan_view")
        c = self._last_counts[0]
        return [(v.match_rows_dev, v.match_scores_dev, c.n_match), (v.near_rows_dev, v.near_scores_dev, c.n_near),
                (v.tie_rows_dev, v.tie_scores_dev, c.n_tie)]

    def _host_view(self, shard, which):
        """Read-only numpy views of the library's pinned host mirror (valid until the next scan)."""
        rp, sp, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        check(lib().vq_scan_host_list(shard.handle, which, C.byref(rp), C.byref(sp), C.byref(n)), "vq_scan_host_list")
        if n.value == 0:
            return np.empty(0, np.int64), np.empty(0, np.float32)
        r = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int64)), shape=(n.value,))
        s = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_float)), shape=(n.value,))
        r.flags.writeable = False
        s.flags.writeable = False
        return r, s

    def near_best(self):
        """(position in the near-miss list, global row, score) of the best near miss of a lists=False scan — highest
        score, first in database order (the clip ticket.py:335-340 holds out of the sampling); None if there is none."""
        best, base = None, 0
        for nb, c in zip(self._near_best, self._last_counts):
            if nb is not None and nb[1] >= 0 and (best is None or nb[2] > best[2]):     # ties: the earlier shard wins
                best = (base + nb[0], nb[1], nb[2])
            base += c.n_near
        return best

    def gather_many(self, requests):
        """[(which, positions)] -> [(global rows, fp32 scores)] with ONE library call for all requests and all shards
        (a review round fetches its sampled matches and near misses together, ticket.py:333,341)."""
        names = {"matches": 0, "near_misses": 1, "ties": 2}
        pos = [np.asarray(p, dtype=np.int64).reshape(-1) for _, p in requests]
        n = sum(len(p) for p in pos)
        rows, scores = np.empty(n, np.int64), np.empty(n, np.float32)
        if n:
            which = np.concatenate([np.full(len(p), names[w], np.int32) for (w, _), p in zip(requests, pos)])
            allp = np.ascontiguousarray(np.concatenate(pos))
            handles = (C.c_void_p * len(self.shards))(*[sh.handle for sh in self.shards])
            check(lib().vq_gather_list_multi(handles, len(self.shards), n, ptr(which), ptr(allp), ptr(rows), ptr(scores)),
                  "vq_gather_list_multi")
        out, o = [], 0
        for p in pos:
            out.append((rows[o:o + len(p)], scores[o:o + len(p)]))
            o += len(p)
        return out

    def gather(self, which, positions):
        """Entries of the ordered match / near-miss / tie list of the last scan at the given list positions:
        (global rows, fp32 scores).  One small round trip instead of the whole list."""
        if len(self.shards) > 1:
            return self.gather_many([(which, positions)])[0]
        idx = {"matches": 0, "near_misses": 1, "ties": 2}[which]
        attr = ("n_match", "n_near", "n_tie")[idx]
        pos = np.asarray(positions, dtype=np.int64).reshape(-1)
        rows, scores = np.empty(len(pos), np.int64), np.empty(len(pos), np.float32)
        base = 0
        for sh, c in zip(self.shards, self._last_counts):
            n = getattr(c, attr)
            sel = np.flatnonzero((pos >= base) & (pos < base + n))
            if len(sel):
                local = np.ascontiguousarray(pos[sel] - base)
                r, s = np.empty(len(sel), np.int64), np.empty(len(sel), np.float32)
                check(lib().vq_gather_list(sh.handle, idx, len(sel), ptr(local), ptr(r), ptr(s)), "vq_gather_list")
                rows[sel], scores[sel] = r, s
            base += n
        if len(pos) and (pos.min() < 0 or pos.max() >= base):
            raise VQError("gather: position outside the %s list of %d entries" % (which, base))
        return rows, scores

    def _fetch_list(self, fn, attr, which, copy):
        if len(self.shards) > 1 and which < 2 and self._lists_on_host:
            # vq_scan_multi published every shard's segment into one host mirror: the search set's list, no concatenation
            rp, sp, n = C.c_void_p(), C.c_void_p(), C.c_int64()
            check(lib().vq_scan_multi_host_list(self.shards[0].handle, which, C.byref(rp), C.byref(sp), C.byref(n)),
                  "vq_scan_multi_host_list")
            if n.value == 0:
                return np.empty(0, np.int64), np.empty(0, np.float32)
            r = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int64)), shape=(n.value,))
            s_ = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_float)), shape=(n.value,))
            if copy:
                return r.copy(), s_.copy()
            r.flags.writeable = False
            s_.flags.writeable = False
            return r, s_
        rows, scores = [], []
        copy = copy or not (self._lists_on_host or which == 2)      # after a lists=False scan only the tie band is mirrored
        for sh, c in zip(self.shards, self._last_counts):
            if not copy:
                r, s = self._host_view(sh, which)
            else:
                n = getattr(c, attr)
                r = np.empty(n, np.int64)
                s = np.empty(n, np.float32)
                check(fn(sh.handle, n, ptr(r), ptr(s)), attr)
            rows.append(r)
            scores.append(s)
        if len(rows) == 1:
            return rows[0], scores[0]
        return np.concatenate(rows), np.concatenate(scores)

    def matches(self, copy=True):
        """(global rows ascending, fp32 scores) of {score >= threshold}.  copy=False returns read-only views of
        the pinned host mirror the scan published into: no copy, valid until the next scan on this store."""
        return self._fetch_list(lib().vq_fetch_matches, "n_match", 0, copy)

    def near_misses(self, copy=True):
        return self._fetch_list(lib().vq_fetch_near, "n_near", 1, copy)

    def ties(self, copy=True):
        return self._fetch_list(lib().vq_fetch_ties, "n_tie", 2, copy)

    def ranked(self, which="matches"):
        """The whole match (or near-miss) list of the last scan ranked on the device: score descending, database
        order among equal scores (report order, ticket.py:266).  Shards are ranked separately and merged here."""
        idx = {"matches": 0, "near_misses": 1}[which]
        attr = ("n_match", "n_near")[idx]
        rows, scores = [], []
        for sh, c in zip(self.shards, self._last_counts):
            n = getattr(c, attr)
            r, s = np.empty(n, np.int64), np.empty(n, np.float32)
            check(lib().vq_fetch_ranked(sh.handle, idx, n, ptr(r), ptr(s)), "vq_fetch_ranked")
            rows.append(r)
            scores.append(s)
        if len(rows) == 1:
            return rows[0], scores[0]
        r, s = np.concatenate(rows), np.concatenate(scores)
        order = np.lexsort((r, -s.astype(np.float64)))
        return r[order], s[order]

    def rank_list(self, which, place):
        """The match (or near-miss) list of the last scan sorted on the device by (score descending, `place` ascending):
        `place` [n] is each list entry's rank in the caller's own order (the tie-break; distinct values).  Returns
        (the places in sorted order, fp32 scores).  Shards are sorted separately (K7) and merged here."""
        idx = {"matches": 0, "near_misses": 1}[which]
        attr = ("n_match", "n_near")[idx]
        place = np.ascontiguousarray(place, dtype=np.uint32)
        outs, base = [], 0
        for sh, c in zip(self.shards, self._last_counts):
            n = getattr(c, attr)
            o, sc = np.empty(n, np.uint32), np.empty(n, np.float32)
            check(lib().vq_rank_list(sh.handle, idx, n, ptr(np.ascontiguousarray(place[base:base + n])), ptr(o), ptr(sc)),
                  "vq_rank_list")
            outs.append((o, sc))
            base += n
        if base != len(place):
            raise VQError("rank_list: %d places for a %s list of %d entries" % (len(place), which, base))
        if len(outs) == 1:
            return outs[0]
        o, sc = np.concatenate([a for a, _ in outs]), np.concatenate([b for _, b in outs])
        order = np.lexsort((o, -sc.astype(np.float64)))
        return o[order], sc[order]

    def topk(self):
        k = self._last_topk
        if k == 0:
            return np.empty(0, np.int64), np.empty(0, np.float32)
        n_l = len(self.shards)
        if n_l > 1 and getattr(self, "_merged_topk", None) is not None:
            return self._merged_topk                               # merged inside vq_scan_multi
        if n_l == 1:
            n = self._last_counts[0].n_topk
            r, s = np.empty(n, np.int64), np.empty(n, np.float32)
            check(lib().vq_fetch_topk(self.shards[0].handle, n, ptr(r), ptr(s)), "vq_fetch_topk")
            return r, s
        sc = np.full((n_l, k), -np.inf, np.float32)
        rw = np.full((n_l, k), -1, np.int64)
        for i, (sh, c) in enumerate(zip(self.shards, self._last_counts)):
            r = np.empty(c.n_topk, np.int64)
            s = np.empty(c.n_topk, np.float32)
            check(lib().vq_fetch_topk(sh.handle, c.n_topk, ptr(r), ptr(s)), "vq_fetch_topk")
            sc[i, :c.n_topk], rw[i, :c.n_topk] = s, r
        so, ro, n = np.empty(k, np.float32), np.empty(k, np.int64), C.c_int32()
        check(lib().vq_merge_topk(n_l, k, ptr(sc), ptr(rw), ptr(so), ptr(ro), C.byref(n)), "vq_merge_topk")
        return ro[:n.value], so[:n.value]

    def scores(self):
        out = np.empty(self.n_rows, np.float32)
        for sh in self.shards:
            lo = sh.first - self.first_global_row
            part = np.empty(sh.n_rows, np.float32)
            check(lib().vq_fetch_scores(sh.handle, 0, sh.n_rows, ptr(part)), "vq_fetch_scores")
            out[lo:lo + sh.n_rows] = part
        return out

    def scores_at(self, global_rows):
        """fp32 scores of the last scan at the given global rows: one round trip per shard that holds any of them."""
        rows = np.asarray(global_rows, dtype=np.int64).reshape(-1)
        out = np.empty(len(rows), np.float32)
        lo_g, hi_g = self.first_global_row, self.first_global_row + self.n_rows
        if len(rows) and (rows.min() < lo_g or rows.max() >= hi_g):
            raise VQError("scores_at: rows outside the store [%d, %d)" % (lo_g, hi_g))
        for sh in self.shards:
            sel = np.flatnonzero((rows >= sh.first) & (rows < sh.first + sh.n_rows))
            if len(sel):
                local = np.ascontiguousarray(rows[sel] - sh.first)
                part = np.empty(len(sel), np.float32)
                check(lib().vq_fetch_scores_at(sh.handle, len(sel), ptr(local), ptr(part)), "vq_fetch_scores_at")
                out[sel] = part
        return out

    def sims(self):
        out = np.empty((self.n_rows, len(self.streams)), np.float32)
        for sh in self.shards:
            lo = sh.first - self.first_global_row
            part = np.empty((sh.n_rows, len(self.streams)), np.float32)
            check(lib().vq_fetch_sims(sh.handle, 0, sh.n_rows, ptr(part)), "vq_fetch_sims")
            out[lo:lo + sh.n_rows] = part
        return out

    # ------------------------------------------------------------------ batched queries (tcgen05)
    TIE_CAP = 4096               # tie-band entries kept per query and shard by the batched kernel

    def scan_batch(self, targets, weights, threshold, lower_limit, topk=0, debug_scores=False, eps=0.0):
        """Score Q targets against the whole store in one pass per shard on the tensor cores.
        targets: list of {stream: {split: vector}} or array [Q, S, P, dim].
        Returns (counts [Q, 2] = matches, near misses; topk_rows [Q, k]; topk_scores [Q, k]; kernel ms)
        or, with debug_scores, the fp32 score matrix [Q, n_rows] (single-shard stores only).
        eps > 0 also reports every query's tie band — rows within eps of the threshold or of the near-miss limit —
        in self.batch_ties = (counts [Q], [(global rows ascending, fp32 scores) per query])."""
        if isinstance(targets, np.ndarray):
            T = np.ascontiguousarray(targets, dtype=np.float32).reshape((-1,) + self.row_shape)
            have = np.ones(self.row_shape[:2], bool)
        else:
            packed = [self.pack_target(t, np.float32) for t in targets]
            T = np.stack([t for t, _ in packed])
            have = packed[0][1]
            if any(not np.array_equal(h, have) for _, h in packed[1:]):
                raise VQError("scan_batch: the targets of one batch must have the same (stream, split) slots — the mean over "
                              "splits (ticket.py:155-157) is one per-row table for the whole pass")
        self._sync_split_weights_for_target(have)              # never inherit what an earlier single-query job left
        Q = T.shape[0]
        w = [weights[s] for s in self.streams] if isinstance(weights, dict) else list(weights)
        p = make_params(w, threshold, lower_limit, float(eps), topk)
        self.batch_ties = None
        if debug_scores:
            if len(self.shards) != 1:
                raise VQError("scan_batch(debug_scores=True) needs a single-shard store")
            out = np.empty((Q, self.n_rows), np.float32)
            check(lib().vq_scan_batch_scores(self.shards[0].handle, ptr(T), Q, C.byref(p), ptr(out)),
                  "vq_scan_batch_scores")
            return out
        counts = np.zeros((Q, 2), np.int64)
        k = max(topk, 1)
        n_sh = len(self.shards)
        c_l = [np.zeros((Q, 2), np.int64) for _ in range(n_sh)]
        rows_l = [np.full((Q, k), -1, np.int64) for _ in range(n_sh)]
        sc_l = [np.full((Q, k), -np.inf, np.float32) for _ in range(n_sh)]
        ms_l = [C.c_float() for _ in range(n_sh)]
        want_ties = eps > 0
        tc_l = [np.zeros(Q, np.int64) for _ in range(n_sh)] if want_ties else None
        tr_l = [np.empty((Q, self.TIE_CAP), np.int64) for _ in range(n_sh)] if want_ties else None
        ts_l = [np.empty((Q, self.TIE_CAP), np.float32) for _ in range(n_sh)] if want_ties else None
        errs = []

        def run(i):                                   # one call per shard; ctypes releases the GIL, devices run concurrently
            try:
                if want_ties:
                    check(lib().vq_scan_batch_ties(self.shards[i].handle, ptr(T), Q, C.byref(p), ptr(c_l[i]), ptr(rows_l[i]),
                                                   ptr(sc_l[i]), ptr(tc_l[i]), self.TIE_CAP, ptr(tr_l[i]), ptr(ts_l[i]),
                                                   C.byref(ms_l[i])), "vq_scan_batch_ties")
                else:
                    check(lib().vq_scan_batch(self.shards[i].handle, ptr(T), Q, C.byref(p), ptr(c_l[i]), ptr(rows_l[i]),
                                              ptr(sc_l[i]), C.byref(ms_l[i])), "vq_scan_batch")
            except Exception as e:                    # surfaced below
                errs.append(e)

        if n_sh == 1:
            run(0)
        else:
            import threading
            ts = [threading.Thread(target=run, args=(i,)) for i in range(n_sh)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        if errs:
            raise errs[0]
        for c in c_l:
            counts += c
        if want_ties:
            lists = []
            for q in range(Q):
                parts = [(tr_l[i][q, :min(int(tc_l[i][q]), self.TIE_CAP)], ts_l[i][q, :min(int(tc_l[i][q]), self.TIE_CAP)])
                         for i in range(n_sh)]
                lists.append((np.concatenate([a for a, _ in parts]), np.concatenate([b for _, b in parts])))
            self.batch_ties = (sum(tc_l), lists)
        ms = max(m.value for m in ms_l)
        if topk == 0:
            return counts, np.empty((Q, 0), np.int64), np.empty((Q, 0), np.float32), ms
        if len(self.shards) == 1:
            return counts, rows_l[0], sc_l[0], ms
        rows_o, sc_o = merge_topk_batch(np.stack(rows_l), np.stack(sc_l))
        return counts, rows_o, sc_o, ms

    # ------------------------------------------------------------------ labelled subset (fp64)
    def labelled_sims(self, target_features, global_rows):
        """float64 [n, S] similarities of the given rows (any shard) against an fp64 target."""
        T, have = self.pack_target(target_features, np.float64)
        self._sync_split_weights_for_target(have)
        rows = np.asarray(global_rows, dtype=np.int64)
        out = np.empty((len(rows), len(self.streams)), np.float64)
        Tc = np.ascontiguousarray(T)
        lo_g, hi_g = self.first_global_row, self.first_global_row + self.n_rows
        if len(rows) and (rows.min() < lo_g or rows.max() >= hi_g):
            raise VQError("labelled_sims: rows outside the store [%d, %d)" % (lo_g, hi_g))
        for sh in self.shards:
            sel = np.flatnonzero((rows >= sh.first) & (rows < sh.first + sh.n_rows))
            if len(sel):
                r = np.ascontiguousarray(rows[sel])
                o = np.empty((len(sel), len(self.streams)), np.float64)
                check(lib().vq_labelled_sims(sh.handle, ptr(Tc), ptr(r), len(sel), ptr(o)), "vq_labelled_sims")
                out[sel] = o
        return out

    def bootstrap_target(self, valid_rows, invalid_rows, mu, slots=None):
        """float64 [S, P, dim]: new target from labelled rows (target_clip.py:161-261), solved on
        the GPU that holds the rows.  Labelled rows spread over several shards are first gathered
        onto the first shard's device through a scratch store.  slots: bool [S, P], the (stream, split) problems to
        solve (default all); the others come back zero.  A singular system raises VQError."""
        valid = np.ascontiguousarray(valid_rows, dtype=np.int64)
        invalid = np.ascontiguousarray(invalid_rows if invalid_rows is not None else [], dtype=np.int64)
        allr = np.concatenate([valid, invalid])
        lo_g, hi_g = self.first_global_row, self.first_global_row + self.n_rows
        if len(valid) == 0 or allr.min() < lo_g or allr.max() >= hi_g:
            raise VQError("bootstrap_target: need >= 1 valid row and all rows inside the store [%d, %d)" % (lo_g, hi_g))
        sh = self._shard_of(int(allr[0]))
        if not all(sh.first <= r < sh.first + sh.n_rows for r in allr):
            return self._bootstrap_gathered(valid, invalid, mu, slots)
        out = np.empty(self.row_shape, np.float64)
        mask = None if slots is None else np.ascontiguousarray(np.asarray(slots, bool).reshape(self.row_shape[:2]), dtype=np.uint8)
        check(lib().vq_bootstrap_target(sh.handle, ptr(valid), len(valid), ptr(invalid) if len(invalid) else None,
                                        len(invalid), float(mu), ptr(mask), ptr(out)), "vq_bootstrap_target")
        return out

    def _bootstrap_gathered(self, valid, invalid, mu, slots=None):
        allr = np.concatenate([valid, invalid])
        uniq = np.unique(allr)
        feats = np.concatenate([self.download(int(r) - self.first_global_row, 1) for r in uniq])
        tmp = FeatureStore(len(uniq), self.streams, self.splits, self.dim, devices=[self.shards[0].device])
        try:
            tmp.upload(0, feats)
            pos = {int(r): i for i, r in enumerate(uniq)}
            v = np.array([pos[int(r)] for r in valid], np.int64)
            iv = np.array([pos[int(r)] for r in invalid], np.int64)
            return tmp.bootstrap_target(v, iv, mu, slots)
        finally:
            tmp.close()


def merge_topk_batch(rows, scores):
    """rows int64 / scores fp32 [n_lists, Q, k] (padding -1 / -inf) -> the k best per query under
    (score descending, global row ascending), the ranking rule of reference ticket.py:266."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n_lists, Q, k = rows.shape
    rows_o, sc_o, n = np.empty((Q, k), np.int64), np.empty((Q, k), np.float32), np.empty(Q, np.int32)
    check(lib().vq_merge_topk_batch(n_lists, Q, k, ptr(scores), ptr(rows), ptr(sc_o), ptr(rows_o), ptr(n)),
          "vq_merge_topk_batch")
    return rows_o, sc_o


def loss_grid(sims, labels, weight_grid, threshold_grid, ballast, replicates=None, device=0):
    """losses [R, n_w, n_th] (float64) on the GPU — hyperparameter.py:56-65 for R index sets.
    replicates: list of index arrays into the L labelled clips; None = one replicate of all L."""
    sims = np.ascontiguousarray(sims, dtype=np.float64)
    L = sims.shape[0]
    if sims.shape[1] != 2:
        raise VQError("optimize_weights is defined for exactly two streams (hyperparameter.py:58)")
    lab = np.ascontiguousarray(np.asarray(labels).astype(bool).astype(np.uint8))
    if replicates is None:
        replicates = [np.arange(L, dtype=np.int32)]
    off = np.zeros(len(replicates) + 1, np.int32)
    off[1:] = np.cumsum([len(r) for r in replicates])
    idx = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in replicates]))
    wg = np.ascontiguousarray(weight_grid, dtype=np.float64)
    tg = np.ascontiguousarray(threshold_grid, dtype=np.float64)
    out = np.empty((len(replicates), len(wg), len(tg)), np.float64)
    check(lib().vq_loss_grid(device, ptr(sims), ptr(lab), L, ptr(wg), len(wg), ptr(tg), len(tg), float(ballast),
                             ptr(off), ptr(idx), len(replicates), ptr(out)), "vq_loss_grid")
    return out


# One store per (api url, search set, streams, feature name); outlives broker ticks.
_REGISTRY = {}


def get_store(key, builder):
    st = _REGISTRY.get(key)
    if st is None:
        st = builder()
        _REGISTRY[key] = st
    return st


def register_store(key, store):
    _REGISTRY[key] = store


def evict_least_recently_used(keep=None):
    """Close the resident store that has gone unused the longest (never `keep`); True if one was closed.  A broker that
    serves many search sets fills HBM with their stores; when building another one runs out of device memory the caller
    evicts and retries (Ticket._build_store) instead of failing every job of the new search set from then on."""
    victims = [(getattr(st, "last_used", 0.0), k) for k, st in _REGISTRY.items() if k != keep]
    if not victims:
        return False
    invalidate(min(victims, key=lambda v: v[0])[1])
    return True


def invalidate(key=None):
    """Drop a cached store (or all): call when load_db.py has added clips to the search set."""
    for k in [key] if key is not None else list(_REGISTRY):
        st = _REGISTRY.pop(k, None)
        if st is not None:
            st.close()

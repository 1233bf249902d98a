"""Vectorised draws from Python's own Mersenne Twister stream.

The reference resamples with `random.choices(range(n), k=k)` (target_clip.py:302-305), i.e. k times
`floor(random() * float(n))` in a Python loop — 0.5 s for 1000 bootstrap replicates of 5000 labels.  numpy's legacy
`RandomState` is the same MT19937 with the same 53-bit double construction, so the identical index sequence can be
produced in one vector call: copy the generator state over, draw, copy the advanced state back.  The caller's
generator ends in exactly the state `random.choices` would have left it in, so every later draw is unchanged."""
from __future__ import annotations

import numpy as np


def choices_range(rng, n, k, repeats=1):
    """`repeats` consecutive calls of `rng.choices(range(n), k=k)` as one int64 array [repeats, k] ([k] when
    repeats == 1); `rng` is the `random` module or a `random.Random`.  One state transfer for all of them."""
    if k <= 0 or repeats <= 0:
        return np.empty((repeats, 0) if repeats != 1 else 0, np.int64)
    version, internal, gauss_next = rng.getstate()
    if version != 3 or len(internal) != 625:                      # not CPython's MT state: take the plain path
        out = np.asarray([rng.choices(range(n), k=k) for _ in range(repeats)], dtype=np.int64)
        return out[0] if repeats == 1 else out
    rs = np.random.RandomState()
    rs.set_state(("MT19937", np.asarray(internal[:624], dtype=np.uint32), int(internal[624]), 0, 0.0))
    idx = np.floor(rs.random_sample(repeats * k) * float(n)).astype(np.int64)
    _, keys, pos, _, _ = rs.get_state()
    rng.setstate((version, tuple(keys.tolist()) + (int(pos),), gauss_next))
    return idx if repeats == 1 else idx.reshape(repeats, k)


NATIVE_SAMPLE_MIN = 64       # below this many picks Python's own loop is as fast as the state transfer


def sample_range(rng, n, k):
    """`rng.sample(range(n), k)` as an int64 array, leaving `rng` in the state the call would have left it in.
    The review rounds draw a few dozen list positions; the finalize round has no size limit (compute_matches.py:78-84)
    and the reference's `random.sample(M.items(), len(M))` is then a seeded permutation of every match — one
    interpreter loop iteration per clip.  Large draws run in the library on the generator's state (vq_mt_sample_range,
    csrc/vq_rng.cu); small ones, and generators that are not CPython's MT19937, take `rng.sample` itself."""
    n, k = int(n), int(k)
    if k < NATIVE_SAMPLE_MIN:
        return np.asarray(rng.sample(range(n), k), dtype=np.int64)
    version, internal, gauss_next = rng.getstate()
    if version != 3 or len(internal) != 625:
        return np.asarray(rng.sample(range(n), k), dtype=np.int64)
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")
    import ctypes as C
    from ._ffi import check, lib, ptr
    state = np.asarray(internal[:624], dtype=np.uint32)
    pos = C.c_int32(int(internal[624]))
    out = np.empty(k, np.int64)
    check(lib().vq_mt_sample_range(ptr(state), C.byref(pos), n, k, ptr(out)), "vq_mt_sample_range")
    rng.setstate((version, tuple(state.tolist()) + (pos.value,), gauss_next))
    return out


def _cpython_set_table_size(n_unique):
    """Table size of a CPython set after inserting n_unique distinct keys one by one (setobject.c: resize when
    fill * 5 >= mask * 3, to the first power of two above used * 4, or used * 2 beyond 50000 keys)."""
    size = 8
    while True:
        threshold = -(-(size - 1) * 3 // 5)                       # smallest fill with fill * 5 >= mask * 3
        if n_unique < threshold:
            return size
        used = threshold
        new = 8
        while new <= used * (2 if used > 50000 else 4):
            new <<= 1
        size = new


def set_order(draws, n):
    """`list(set(draws))` for draws in range(n) (the reference drops duplicates this way, target_clip.py:306) without
    building the set when its iteration order is known: small non-negative ints hash to themselves, so once the table
    is larger than every key each key sits in its own slot and iteration is ascending.  Otherwise the real set."""
    draws = np.asarray(draws)
    if n <= 16 * max(draws.size, 1):                                # presence mask: O(n), no sort, no hashing
        present = np.zeros(n, dtype=bool)
        present[draws] = True
        u = np.flatnonzero(present)
    else:
        u = np.unique(draws)
    if n <= _cpython_set_table_size(len(u)):
        return u
    return np.fromiter(set(np.asarray(draws).tolist()), dtype=np.int64)

"""VQSYN-1: counter-based synthetic feature database (CPU side).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py may import this
package.  The CUDA generator (`csrc/vq_store.cu`, `vq_store_fill_synthetic`) implements the same
function; any row can be regenerated here bit-identically, which is how shards of a
100M-clip database are spot-checked without ever holding it on the host.

The reference ships no generator; the shape follows its real features (`data/features/**`:
all values >= 0, rgb mean ~2.5, flow ~0.9) so that scores spread over [0.55, 1] instead of
collapsing to one value.

Definition (all integer maths mod 2^64, all float maths single-rounded fp32, no FMA):
    mix(z)      = splitmix64 finaliser
    u01(k)      = float32(mix(k * 0x9E3779B97F4A7C15 + seed) >> 40) * 2^-24         in [0, 1)
    base[s][d]  = (xb * xb) * m3_s,  m3_s = fp32(3) * fp32(mean_s),   xb = u01(((2^48 + s) * L) + d)
    alpha[c]    = a8,  a = u01(2^56 + c), a2 = a*a, a4 = a2*a2, a8 = a4*a4
    noise       = (x * x) * m3_s,             x  = u01(((c * S + s) * L) + d)
    feat[c,s,d] = alpha*base + (1 - alpha)*noise        (two products, one sum, each rounded)
with c the GLOBAL clip row, S streams, L = n_splits * 1024 floats per stream.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
MIX1 = np.uint64(0xBF58476D1CE4E5B9)
MIX2 = np.uint64(0x94D049BB133111EB)
BASE_TAG = np.uint64(1) << np.uint64(48)
ALPHA_TAG = np.uint64(1) << np.uint64(56)
DEFAULT_MEANS = (2.5, 0.9, 1.7, 1.3)   # per stream; reference data: rgb ~2.5, flow ~0.9
DEFAULT_SEED = 20261018


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * MIX1
    z = (z ^ (z >> np.uint64(27))) * MIX2
    return z ^ (z >> np.uint64(31))


def u01(k, seed):
    with np.errstate(over="ignore"):
        z = _mix(np.asarray(k, dtype=np.uint64) * GOLDEN + np.uint64(seed))
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def alpha(rows, seed):
    a = u01(ALPHA_TAG + np.asarray(rows, dtype=np.uint64), seed)
    a2 = a * a
    a4 = a2 * a2
    return a4 * a4


def base_vector(seed, stream, n_streams, stream_len, means=DEFAULT_MEANS):
    d = np.arange(stream_len, dtype=np.uint64)
    with np.errstate(over="ignore"):
        kb = (BASE_TAG + np.uint64(stream)) * np.uint64(stream_len) + d
    xb = u01(kb, seed)
    return (xb * xb) * (np.float32(3.0) * np.float32(means[stream]))


def rows(seed, row_ids, n_streams=2, stream_len=1024, means=DEFAULT_MEANS):
    """Features of the given global clip rows: float32 [len(row_ids), n_streams, stream_len]."""
    r = np.asarray(row_ids, dtype=np.uint64).reshape(-1)
    out = np.empty((r.shape[0], n_streams, stream_len), np.float32)
    d = np.arange(stream_len, dtype=np.uint64)
    a = alpha(r, seed)[:, None]
    one_minus = np.float32(1.0) - a
    S = np.uint64(n_streams)
    L = np.uint64(stream_len)
    for s in range(n_streams):
        b = base_vector(seed, s, n_streams, stream_len, means)
        with np.errstate(over="ignore"):
            k = (r[:, None] * S + np.uint64(s)) * L + d[None, :]
        x = u01(k, seed)
        noise = (x * x) * (np.float32(3.0) * np.float32(means[s]))
        out[:, s, :] = a * b[None, :] + one_minus * noise
    return out


def database(seed, n_clips, n_streams=2, stream_len=1024, first_row=0, chunk=8192,
             means=DEFAULT_MEANS):
    """Whole shard [n_clips, n_streams, stream_len] float32, generated in chunks."""
    out = np.empty((n_clips, n_streams, stream_len), np.float32)
    for lo in range(0, n_clips, chunk):
        hi = min(n_clips, lo + chunk)
        out[lo:hi] = rows(seed, np.arange(first_row + lo, first_row + hi), n_streams,
                          stream_len, means)
    return out


def pick_reference_row(seed, n_clips, want_alpha=0.9):
    """Deterministic choice of the reference clip: the row whose alpha is closest to 0.9."""
    a = alpha(np.arange(n_clips), seed)
    return int(np.argmin(np.abs(a - np.float32(want_alpha))))

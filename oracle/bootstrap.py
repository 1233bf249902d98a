"""CPU oracle for target bootstrapping.  TEST INFRASTRUCTURE.

Restates `src/models/target_clip.py` of the reference on arrays.  Pinned by the golden
"revise" / "finalize" rounds recorded from the reference itself (`tests/golden/`).

feats: float64 [n, S, P, D] — the labelled clips' features in the order the reference's
`features_for_matches` returns them (order of the matches list, `target_clip.py:129-134`).
"""
from __future__ import annotations

import numpy as np


def random_fraction(n, fraction, replacement, rng):
    """Indices kept by `TargetClip._random_fraction` — `target_clip.py:297-309`.
    `list(set(...))` drops duplicates and orders small ints ascending by hash."""
    t = max(round(n * fraction), 1)
    if replacement is False:
        samples = rng.sample(range(n), t)
    else:
        samples = rng.choices(range(n), k=t)
    return list(set(samples))


def solve_valid(X):
    """Min-norm w with w . x_i = 1 for all rows x_i — `target_clip.py:194-198`.
    X: [n, D] rows.  (The reference holds X as D x n columns.)"""
    Xc = np.asarray(X, dtype=np.float64).T
    M = np.matmul(Xc.T, Xc)
    M_inv = np.linalg.inv(M)
    mu = np.sum(M_inv, axis=1).reshape([-1, 1])
    return np.dot(Xc, mu).T[0]


def solve_valid_invalid(X, Y, mu_h):
    """`target_clip.py:248-260`.  X: [n, D] valid rows, Y: [m, D] invalid rows."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    tr = np.trace(np.matmul(Y, Y.T))
    scale = mu_h / tr
    M = np.eye(Y.shape[1]) + scale * np.matmul(Y.T, Y)
    M_inv = np.linalg.inv(M)
    B = np.matmul(X, np.matmul(M_inv, X.T))
    B_inv = np.linalg.inv(B)
    w_1 = np.matmul(np.matmul(M_inv, X.T), B_inv)
    w_2 = M_inv - np.matmul(np.matmul(w_1, X), M_inv)
    w_3 = np.sum(np.matmul(w_2, scale * Y.T), axis=1).reshape([-1, 1])
    return (w_3 + np.sum(w_1, axis=1).reshape([-1, 1])).T[0]


def dynamic_target_adjustment(valid, invalid, b_fraction, replacement, mu_h, rng):
    """`target_clip.py:84-103` + `:161-261`.  valid/invalid: [n, S, P, D] / [m, S, P, D] (m may be 0).
    RNG order: valid resample first, then invalid (`:227-230`); the valid-only branch draws
    only when b_fraction != 1 or replacement (`:181-182`)."""
    valid = np.asarray(valid, dtype=np.float64)
    n, S, P, D = valid.shape
    out = np.empty((S, P, D), np.float64)
    if invalid is not None and len(invalid) > 0:
        invalid = np.asarray(invalid, dtype=np.float64)
        iv = random_fraction(n, b_fraction, replacement, rng)
        ii = random_fraction(len(invalid), b_fraction, replacement, rng)
        for s in range(S):
            for p in range(P):
                out[s, p] = solve_valid_invalid(valid[iv, s, p], invalid[ii, s, p], mu_h)
        return out
    if b_fraction != 1 or replacement is True:
        iv = random_fraction(n, b_fraction, replacement, rng)
    else:
        iv = list(range(n))
    for s in range(S):
        for p in range(P):
            out[s, p] = solve_valid(valid[iv, s, p])
    return out


def target_by_bagging(valid, invalid, nbags, mu_h, rng):
    """Mean of nbags resampled-with-replacement targets — `target_clip.py:145-159`."""
    bags = [dynamic_target_adjustment(valid, invalid, 1, True, mu_h, rng) for _ in range(nbags)]
    return np.average(bags, axis=0)


def avg_new_old(new, old, f_memory):
    """`target_clip.py:75-82`."""
    if old is None:
        return new
    return np.multiply(f_memory, new) + np.multiply(1 - f_memory, old)


def get_target_features(ref, valid, invalid, previous, bootstrap, has_latest_result,
                        bootstrap_type, f_bootstrap, f_memory, nbags, mu_h, rng):
    """Case analysis of `TargetClip.get_target_features` — `target_clip.py:26-73`."""
    from .scoring import scale_target
    if not bootstrap or not has_latest_result:
        return scale_target(ref)
    if valid is None or len(valid) == 0:
        return scale_target(ref)
    if bootstrap_type == "simple":
        return dynamic_target_adjustment(valid, invalid, f_bootstrap, False, mu_h, rng)
    if bootstrap_type == "partial_update":
        new = dynamic_target_adjustment(valid, invalid, f_bootstrap, False, mu_h, rng)
        return avg_new_old(new, previous, f_memory)
    if bootstrap_type == "bagging":
        return target_by_bagging(valid, invalid, nbags, mu_h, rng)
    raise Exception("Error: bootstrap_type should be one of 'simple', 'partial_update', or 'bagging'")

"""CPU oracle for the match-scoring path: vectorised float64 restatement.

TEST INFRASTRUCTURE — not a product path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import anything under oracle/.

Parity status: the reference's own tests pin nothing for this path (its five *.ut.py files
are `assertTrue(True)`).  This restatement is therefore pinned against OUTPUTS
OF THE REFERENCE ITSELF: `tests/golden/make_golden.py` runs the unmodified reference
(`/root/reference/src/models/*.py`, driven through its public `compute_matches`) on the
reference's fixture features and on VQSYN-1 data and stores what it produced under
`tests/golden/`; `tests/test_oracle_golden.py` checks every function below against those files.

Every function cites the reference lines it restates.  Arrays replace the reference's nested
dicts: X[n, s, p, d] is clip row n (database order), stream s, split slot p, feature dim d.
All arithmetic is float64, as in the reference (Python floats / numpy float64 throughout).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- target (A1)
def scale_target(ref):
    """t[s][p] = f / (f . f)   — `src/models/target_clip.py:311-313`, applied per (stream, split)
    by `scaled_ref_clip_features` `:137-143`.  ref: [S, P, D] -> [S, P, D] float64."""
    ref = np.asarray(ref, dtype=np.float64)
    out = np.empty_like(ref)
    for s in range(ref.shape[0]):
        for p in range(ref.shape[1]):
            f = ref[s, p]
            out[s, p] = f / np.dot(f, f)
    return out


# --------------------------------------------------------------------------- similarities (A3)
def similarities(X, T, present=None):
    """sim[n, s] = mean over the splits present of T[s, p] . X[n, s, p]
    — `src/models/ticket.py:146-160` (dot `:151`, split mean `:156-157`).

    X: [N, S, P, D] (any float dtype; promoted to float64), T: [S, P, D] float64,
    present: optional bool [N, S, P] (False = that clip has no feature row for the split;
    the reference simply never sees such a row, `ticket.py:374-381`).
    Returns (sims [N, S] float64, n_splits [N, S] int)."""
    N, S, P, D = X.shape
    T = np.asarray(T, dtype=np.float64)
    sims = np.zeros((N, S), np.float64)
    cnt = np.zeros((N, S), np.int64)
    for s in range(S):
        for p in range(P):
            d = X[:, s, p, :].astype(np.float64, copy=False) @ T[s, p]
            if present is None:
                sims[:, s] += d
                cnt[:, s] += 1
            else:
                m = present[:, s, p]
                sims[:, s] += np.where(m, d, 0.0)
                cnt[:, s] += m
    with np.errstate(invalid="ignore", divide="ignore"):
        sims = sims / cnt
    return sims, cnt


# --------------------------------------------------------------------------- scores (A4)
def scores(sims, weights):
    """score = 1 - sqrt( sum_s (w_s (1 - sim_s))^2 / sum_s w_s^2 ), streams in `weights` order
    — `src/models/ticket.py:173-180`.  Same operation order as the reference loop, so given
    identical sims the result is bit-identical."""
    sims = np.asarray(sims, dtype=np.float64)
    ssum = np.zeros(sims.shape[0], np.float64)
    denom = 0.0
    for s, w in enumerate(weights):
        w = float(w)
        ssum = ssum + (w * (1 - sims[:, s])) ** 2
        denom += w ** 2
    return 1 - np.sqrt(ssum / denom)


# --------------------------------------------------------------------------- selection (A5-A7)
def lower_limit(threshold, near_miss):
    """`src/models/ticket.py:325`."""
    return threshold - near_miss * (1 - threshold)


def classify(score, threshold, near_miss):
    """Rows (database order) of the match set {score >= th} and the near-miss set
    {lower <= score < th} — `src/models/ticket.py:326-327`."""
    lo = lower_limit(threshold, near_miss)
    m = np.flatnonzero(score >= threshold)
    nm = np.flatnonzero((score >= lo) & (score < threshold))
    return m, nm


def tie_band(score, threshold, near_miss, eps):
    """Rows whose set membership an fp32 scorer may legitimately flip: within eps
    (COMPUTE_EPS, `src/models/hyperparameter.py:5`, `Dockerfile:15`) of either boundary."""
    lo = lower_limit(threshold, near_miss)
    return np.flatnonzero((np.abs(score - threshold) < eps) | (np.abs(score - lo) < eps))


def select_clips_to_review(score, clip_ids, rng, threshold=0.8, max_number_matches=20,
                           near_miss=0.5, ref_clip_id=None, user_matches=None):
    """`src/models/ticket.py:311-356`.  score/clip_ids are in database (dict-insertion) order;
    rng is Python's `random` module or a `random.Random` (the reference uses the module,
    seeded once per tick at `src/broker.py:83-84`).  Returns the ordered {clip_id: score}."""
    clip_ids = [int(c) for c in clip_ids]
    m_rows, nm_rows = classify(score, threshold, near_miss)
    match_candidates = [(clip_ids[i], float(score[i])) for i in m_rows]
    near = [(clip_ids[i], float(score[i])) for i in nm_rows]
    mscores = int(min(max_number_matches / 2, len(match_candidates)))
    m_near = int(min(max_number_matches - mscores, len(near)))
    match_scores = rng.sample(match_candidates, mscores)
    near_max = {}
    if m_near > 0:
        m_near -= 1
        j = max(range(len(near)), key=lambda i: near[i][1])     # first max in dict order
        near_max = {near[j][0]: near[j][1]}
        near = near[:j] + near[j + 1:]
    near_scores = rng.sample(near, m_near)
    out = dict(match_scores + near_scores)
    out.update(near_max)
    by_id = None
    prev = {}
    if ref_clip_id is not None and ref_clip_id in set(clip_ids):
        by_id = {c: i for i, c in enumerate(clip_ids)}
        prev[ref_clip_id] = float(score[by_id[ref_clip_id]])
    if user_matches:
        if by_id is None:
            by_id = {c: i for i, c in enumerate(clip_ids)}
        for clip, value in user_matches.items():
            if value is True:
                prev[int(clip)] = float(score[by_id[int(clip)]])
    out.update(prev)
    return out


def lowest_scoring_user_match(score, clip_ids, user_matches):
    """`src/models/ticket.py:301-309`: (min(1, scores of user-True clips), LAST user-True clip
    in database order — the reference overwrites min_clip on every True clip, `:308`)."""
    min_score, min_clip = 1, None
    for i, c in enumerate(clip_ids):
        if user_matches.get(str(int(c))) is True:
            min_score = min(min_score, float(score[i]))
            min_clip = int(c)
    return min_score, min_clip


def finalize_near_miss(threshold, low_score, eps):
    """`src/models/compute_matches.py:83-85`."""
    return max(threshold - low_score, 0) / max(1 - threshold, eps)


def topk_stable(score, k):
    """Ranking rule of the final report, `src/models/ticket.py:266`: stable sort by score,
    descending, so equal scores keep database order.  Returns rows of the k best."""
    order = np.argsort(-np.asarray(score), kind="stable")
    return order[:k]


# --------------------------------------------------------------------------- weights (A8, A9)
def weight_grid():
    """`src/models/hyperparameter.py:20` — 40 points."""
    return np.arange(0.5, 2.5, 0.05)


def threshold_grid():
    """`src/models/hyperparameter.py:21` — 31 points (float arange overshoots 1.1)."""
    return np.arange(0.5, 1.1, 0.02)


def loss_grid(sims_lab, y, wgrid=None, thgrid=None, ballast=0.0):
    """losses[iw, ith] — `src/models/hyperparameter.py:56-65`.

    sims_lab: [L, 2] float64 similarities of the labelled clips (order of `match_status`),
    y: [L] labels (bool/0/1).  loss = (0.5 th + sum_i (H(s_i - th) - y_i)(s_i - th)(1 + y_i b)) / L
    with H(0) = 1 (`np.heaviside(x, 1)` `:63`) and s_i scored with weights (1.0, w) `:58`."""
    wgrid = weight_grid() if wgrid is None else wgrid
    thgrid = threshold_grid() if thgrid is None else thgrid
    y = np.asarray(y, dtype=np.float64)
    L = len(y)
    out = 100 * np.ones((len(wgrid), len(thgrid)))
    for iw, w in enumerate(wgrid):
        s = scores(sims_lab, (1.0, w))
        for ith, th in enumerate(thgrid):
            loss = 0.5 * th
            terms = (np.heaviside(s - th, 1) - y) * (s - th) * (1 + y * ballast)
            for t in terms:            # sequential accumulation, like the reference loop
                loss += t
            out[iw, ith] = loss / L
    return out


def loss_grid_fast(sims_lab, y, wgrid=None, thgrid=None, ballast=0.0):
    """Same as loss_grid with a pairwise numpy sum (differs by ~1e-16 relative); used where
    L x replicates makes the sequential loop too slow."""
    wgrid = weight_grid() if wgrid is None else wgrid
    thgrid = threshold_grid() if thgrid is None else thgrid
    y = np.asarray(y, dtype=np.float64)
    out = np.empty((len(wgrid), len(thgrid)))
    for iw, w in enumerate(wgrid):
        s = scores(sims_lab, (1.0, w))
        d = s[:, None] - thgrid[None, :]
        terms = (np.heaviside(d, 1) - y[:, None]) * d * (1 + y[:, None] * ballast)
        out[iw] = (0.5 * thgrid + terms.sum(axis=0)) / len(y)
    return out


def quad_fit(x, y):
    """Separable parabola through 5 grid losses — `src/models/hyperparameter.py:85-114`."""
    (xa, xb, xc), (ta, tb, tc) = x
    dw = (y[4] - y[0]) * xb ** 2 + (y[2] - y[4]) * xa ** 2 - (y[2] - y[0]) * xc ** 2
    w0 = 0.5 * dw / ((y[4] - y[0]) * xb + (y[2] - y[4]) * xa - (y[2] - y[0]) * xc)
    a0 = (y[2] - y[0]) / ((xb - w0) ** 2 - (xa - w0) ** 2)
    dt = (y[3] - y[1]) * tb ** 2 + (y[2] - y[3]) * ta ** 2 - (y[2] - y[1]) * tc ** 2
    th0 = 0.5 * dt / ((y[3] - y[1]) * tb + (y[2] - y[3]) * ta - (y[2] - y[1]) * tc)
    b0 = (y[2] - y[1]) / ((tb - th0) ** 2 - (ta - th0) ** 2)
    c0 = y[2] - a0 * (xb - w0) ** 2 - b0 * (tb - th0) ** 2
    w0 = max(min(w0, xc), xa)
    th0 = max(min(th0, tc), ta)
    fit = [a0 * (xa - w0) ** 2 + b0 * (tb - th0) ** 2 + c0,
           a0 * (xb - w0) ** 2 + b0 * (ta - th0) ** 2 + c0,
           a0 * (xb - w0) ** 2 + b0 * (tb - th0) ** 2 + c0,
           a0 * (xb - w0) ** 2 + b0 * (tc - th0) ** 2 + c0,
           a0 * (xc - w0) ** 2 + b0 * (tb - th0) ** 2 + c0]
    if sum(abs(y[i] - fit[i]) for i in range(5)) > 10 ** -6:
        w0, th0 = xb, tb
    return w0, th0


def optimum_from_losses(losses, wgrid, thgrid, eps):
    """argmin (first in C order), border rule, fine tune, `threshold = opt - eps`
    — `src/models/hyperparameter.py:66-76`.  Returns (w_flow, threshold, (iw0, ith0))."""
    iw0, ith0 = np.unravel_index(np.argmin(losses, axis=None), losses.shape)
    if iw0 == 0 or ith0 == 0 or iw0 == len(wgrid) - 1 or ith0 == len(thgrid) - 1:
        w_opt, th_opt = wgrid[iw0], thgrid[ith0]
    else:
        xr = [(wgrid[iw0 - 1], wgrid[iw0], wgrid[iw0 + 1]),
              (thgrid[ith0 - 1], thgrid[ith0], thgrid[ith0 + 1])]
        yd = [losses[iw0 - 1, ith0], losses[iw0, ith0 - 1], losses[iw0, ith0],
              losses[iw0, ith0 + 1], losses[iw0 + 1, ith0]]
        w_opt, th_opt = quad_fit(xr, yd)
    return float(w_opt), float(th_opt - eps), (int(iw0), int(ith0))


def match_status(matches):
    """`src/models/hyperparameter.py:45-50`: ordered {clip: label}; user_match wins over
    is_match; later duplicates overwrite earlier ones but keep the first position."""
    out = {}
    for m in matches:
        out[m["video_clip"]] = m["user_match"] if m["user_match"] is not None else m["is_match"]
    return out


def optimize_weights(sims, clip_ids, matches, eps, ballast=0.0):
    """`src/models/hyperparameter.py:29-76` on array inputs.  Returns (w_flow, threshold, losses)."""
    status = match_status(matches)
    by_id = {int(c): i for i, c in enumerate(clip_ids)}
    rows = [by_id[int(c)] for c in status]
    y = np.array([bool(v) for v in status.values()], dtype=np.float64)
    losses = loss_grid(np.asarray(sims)[rows], y, ballast=ballast)
    w, th, _ = optimum_from_losses(losses, weight_grid(), threshold_grid(), eps)
    return w, th, losses

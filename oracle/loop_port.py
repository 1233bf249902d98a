"""Loop-faithful CPU port of the reference's scoring step.  TEST INFRASTRUCTURE.

This is the reference "as written": nested dicts of Python lists, one `np.dot(list, list)` per
(clip, stream, split), Python accumulation — the data structures and operation order of
`src/models/ticket.py:120-180,311-356`.  It exists for two reasons:
  * `bench.py --impl reference` / `cpu_baseline` time it on the GPU box's host cores, where
    `/root/reference` is absent (the reference is pure Python: there is nothing to compile
    into `oracle/_ref`, so the timed CPU arm is this port, kind "port");
  * tests pin it to the golden outputs recorded from the real reference, which shows that the
    timed thing computes what the reference computes.
It is single-threaded because the reference is (GIL-bound Python loops).
"""
from __future__ import annotations

import numpy as np


def make_candidates(X, clip_ids, streams, splits):
    """Nested dict the reference builds from the API response, `ticket.py:367-382`:
    {stream: {split: {clip: [floats]}}}.  X: [N, S, P, D].  (Payload construction — excluded
    from timings, like the HTTP fetch it stands for.)"""
    cand = {}
    for s, stream in enumerate(streams):
        cand[stream] = {}
        for p, split in enumerate(splits):
            block = X[:, s, p, :].astype(np.float64).tolist()
            cand[stream][split] = {int(c): block[i] for i, c in enumerate(clip_ids)}
    return cand


def make_target(T, streams, splits):
    """{stream: {split: [floats]}} as `scaled_ref_clip_features` stores it (`target_clip.py:142`)."""
    return {stream: {split: np.asarray(T[s, p], dtype=np.float64).tolist()
                     for p, split in enumerate(splits)} for s, stream in enumerate(streams)}


def compute_similarities(target_features, candidates):
    """`ticket.py:145-163`."""
    averaged = {}
    for stream, per_split in target_features.items():
        found = {}
        for split, tvec in per_split.items():
            for clip, cvec in candidates[stream][split].items():
                found[clip] = found.get(clip, []) + [np.dot(tvec, cvec)]
        for clip, vals in found.items():
            n = len(vals)
            entry = averaged.get(clip, {})
            entry.update({stream: [sum(vals) / n, n]})
            averaged[clip] = entry
    return averaged


def compute_scores(sims, weights):
    """`ticket.py:172-180`."""
    out = {}
    for clip, per_stream in sims.items():
        num = 0
        den = 0
        for stream, w in weights.items():
            num += (w * (1 - per_stream[stream][0])) ** 2
            den += w ** 2
        out[clip] = 1 - np.sqrt(num / den)
    return out


def candidate_sets(scores, threshold, near_miss):
    """`ticket.py:325-327`."""
    low = threshold - near_miss * (1 - threshold)
    hits = {k: v for k, v in scores.items() if v >= threshold}
    near = {k: v for k, v in scores.items() if low <= v < threshold}
    return hits, near


def scoring_step(target_features, candidates, weights, threshold, near_miss):
    """One pass of the hot path over one batch: similarities -> scores -> candidate sets."""
    sims = compute_similarities(target_features, candidates)
    sc = compute_scores(sims, weights)
    hits, near = candidate_sets(sc, threshold, near_miss)
    return sims, sc, hits, near

"""A revise round's labelled-path work on a config-3 shard: L labelled clips of a 12.5M-clip store
(102.4 GB resident) — Hyperparameter.optimize_weights (vq_labelled_sims + vq_loss_grid, store-owned scratch, no cudaMalloc
per call), the scan + review selection that follows, and a target bootstrap from 300 confirmed + 200 rejected clips.
Prints one JSON object.

  python tests/probes/revise_round_probe.py [n_clips] [n_labelled]
"""
import json
import os
import random
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("COMPUTE_EPS", ".000003")
os.environ.setdefault("RANDOM_SEED", "73459912436")
import video_query_algorithms_b200 as vq  # noqa: E402

S = ("rgb", "warped_optical_flow")
SEED, REF = 20261018, 18120


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    st = vq.FeatureStore(n, S, [1], 1024, devices=[0], clip_ids=np.arange(n))
    st.fill_synthetic(SEED)
    f = st.download(REF, 1)[0].astype(np.float64)
    tdict = {s_: {1: vq.TargetClip._scale_feature(f[i, 0]).tolist()} for i, s_ in enumerate(S)}
    w = {"rgb": 1.0, "warped_optical_flow": 1.5}
    st.scan(tdict, w, 0.8, 0.73, 3e-6)
    m_rows, m_sc = st.matches()
    n_rows, n_sc = st.near_misses()
    rng = np.random.default_rng(7)
    pick = np.sort(np.concatenate([rng.choice(m_rows, L // 2, replace=False), rng.choice(n_rows, L - L // 2, replace=False)]))
    scores = st.scores_at(pick)
    matches = [{"video_clip": int(c), "user_match": bool(v >= 0.82), "is_match": bool(v >= 0.8)} for c, v in zip(pick, scores)]
    job = {"query_id": 1, "video_id": 1, "ref_clip": 0, "ref_clip_id": REF, "search_set": 1, "number_of_matches_to_review": 20,
           "dynamic_target_adjustment": True, "matches": matches, "user_matches": {str(m["video_clip"]): m["user_match"] for m in matches[:200]}}
    t = vq.Ticket(job, "http://fake/", client=object(), schema=object(), store=st)
    t.target = types.SimpleNamespace(target_features=tdict, splits={1})
    hp = vq.Hyperparameter(w, ballast=0.1)
    t._hp = hp
    out = {"clips": n, "labelled": L}
    out["optimize_weights_ms"], _ = timed(lambda: hp.optimize_weights(t), 20)
    out["optimize_weights_ms"] *= 1e3
    out["weights"], out["threshold"] = [hp.weights[s_] for s_ in S], hp.threshold
    t.compute_scores(hp.weights)
    random.seed(a=os.environ["RANDOM_SEED"])
    out["scan_and_review_selection_ms"], _ = timed(lambda: t.select_clips_to_review(hp.threshold, 20, 0.35), 10)
    out["scan_and_review_selection_ms"] *= 1e3
    out["selected"] = len(t.matches)
    valid = np.array([m["video_clip"] for m in matches if m["user_match"]][:300], np.int64)
    invalid = np.array([m["video_clip"] for m in matches if not m["user_match"]][:200], np.int64)
    out["bootstrap_target_300_valid_200_invalid_ms"], T = timed(lambda: st.bootstrap_target(valid, invalid, 0.3), 10)
    out["bootstrap_target_300_valid_200_invalid_ms"] *= 1e3
    out["target_finite"] = bool(np.isfinite(T).all())
    out["lowest_user_match_ms"], _ = timed(lambda: t.lowest_scoring_user_match(), 20)
    out["lowest_user_match_ms"] *= 1e3
    print(json.dumps(out))
    st.close()


if __name__ == "__main__":
    main()

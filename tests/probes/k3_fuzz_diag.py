import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import numpy as np
import video_query_algorithms_b200 as vq
from oracle import scoring as sc
for seed in range(40):
    rng = np.random.default_rng(9100 + seed)
    S, P = int(rng.integers(1, 5)), int(rng.integers(1, 3))
    dim = int(rng.choice([32, 64, 256, 512, 1024]))
    n = int(rng.choice([1, 100, 127, 128, 129, 5000, 148 * 128 + 5, 40000]))
    Q = int(rng.choice([1, 5, 63, 64, 65, 128, 129, 255, 256, 257, 300]))
    if n * Q > 6_000_000:
        Q = max(1, 6_000_000 // n)
    streams = tuple("s%d" % i for i in range(S)); splits = list(range(1, P + 1))
    X = (rng.random((n, S, P, dim), dtype=np.float32) * (0.5 + rng.random((n, 1, 1, 1), dtype=np.float32))).astype(np.float32)
    present = rng.random((n, S, P)) > (0.2 if P > 1 and seed % 2 else 0.0)
    present[:, :, 0] = True
    X = X * present[..., None]
    st = vq.FeatureStore(n, streams, splits, dim, devices=[0]); st.upload(0, X); st.set_present(present)
    X64 = X.astype(np.float64)
    w = [float(v) for v in rng.uniform(0.5, 2.5, S)]
    refs = rng.integers(0, n, Q)
    T32 = np.stack([sc.scale_target(np.where(present[r][..., None], X64[r], 1.0)) for r in refs]).astype(np.float32)
    got = st.scan_batch(T32, w, 0.6, 0.35, debug_scores=True)
    worst = (0, 0, 0)
    for q in range(min(Q, 12)):
        sims64, _ = sc.similarities(X64, T32[q].astype(np.float64), None if present.all() else present)
        s64 = sc.scores(sims64, w)
        err = np.abs(got[q].astype(np.float64) - s64) / np.maximum(np.abs(s64), 0.25)
        # error relative to similarity magnitude
        i = int(err.argmax())
        if err[i] > worst[0]:
            worst = (float(err[i]), q, i, float(s64[i]), sims64[i].round(3).tolist(), float(np.abs(sims64).max()))
    # single-query K1 for the same target for comparison
    td = {s: {p_: T32[worst[1]][si, pi].astype(np.float64) for pi, p_ in enumerate(splits)} for si, s in enumerate(streams)}
    st.scan(td, w, 0.6, 0.35, 3e-6)
    k1 = st.scores()
    sims64, _ = sc.similarities(X64, T32[worst[1]].astype(np.float64), None if present.all() else present)
    s64 = sc.scores(sims64, w)
    e1 = float((np.abs(k1.astype(np.float64) - s64) / np.maximum(np.abs(s64), 0.25)).max())
    print(seed, "S%d P%d dim%d n%d Q%d missing=%s" % (S, P, dim, n, Q, not present.all()), "K3 worst %.2e" % worst[0], worst[1:], "K1 %.2e" % e1, flush=True)
    st.close()

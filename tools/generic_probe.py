"""Bandwidth of the generic (shared-memory target) K1 variant on 3-split rows."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq
S = ("rgb", "warped_optical_flow")
for splits, n in (([1, 2, 3], 300000), ([1], 900000)):
    st = vq.FeatureStore(n, S, splits, 1024, devices=[0])
    st.fill_synthetic(3)
    T = np.random.default_rng(0).random((2, len(splits), 1024))
    td = {s: {p: T[i, j] for j, p in enumerate(splits)} for i, s in enumerate(S)}
    best = 1e9
    for _ in range(5):
        r = st.scan(td, (1.0, 1.5), 0.8, 0.7, 3e-6, topk=100)
        best = min(best, r.scan_ms)
    gb = n * 2 * len(splits) * 4096 / 1e9
    print("splits %s rows %d: K1 %.3f ms  %.0f GB/s" % (splits, n, best, gb / best * 1e3))
    st.close()

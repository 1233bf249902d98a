#!/usr/bin/env python
"""K1 throughput on rows of 1, 2 and 3 splits per stream (2 streams x P x 1024 fp32), each with the kernel the library picks
and with the generic kernel forced (VQ_SCAN_GENERIC=1): python tools/scan_shapes_probe.py [GB per shard]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
S = ("rgb", "warped_optical_flow")
out = {}
for P in (1, 2, 3):
    n = int(gb * 1e9 / (2 * P * 1024 * 4))
    st = vq.FeatureStore(n, S, list(range(1, P + 1)), 1024, devices=[0])
    st.fill_synthetic(7)
    f = st.download(5, 1)[0].astype(np.float64)
    t = {s: {p + 1: vq.TargetClip._scale_feature(f[i, p]) for p in range(P)} for i, s in enumerate(S)}
    res = {}
    for mode in ("default", "generic"):
        if mode == "generic":
            os.environ["VQ_SCAN_GENERIC"] = "1"
        else:
            os.environ.pop("VQ_SCAN_GENERIC", None)
        ms = []
        for i in range(25):
            r = st.scan(t, (1.0, 1.5), 0.8, 0.73, 3e-6, topk=100, lists=False)
            if i >= 5:
                ms.append(r.scan_ms)
        sc = st.scores()
        res[mode] = {"k1_ms": float(np.mean(ms)), "tb_s": n * 2 * P * 4096 / (np.mean(ms) * 1e-3) / 1e12,
                     "n_match": r.n_match, "score_sum": float(sc.astype(np.float64).sum())}
    res["max_score_diff"] = None
    out["P=%d, %d clips" % (P, n)] = res
    st.close()
os.environ.pop("VQ_SCAN_GENERIC", None)
print(json.dumps(out, indent=1))

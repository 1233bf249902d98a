"""Small end-to-end exercise of every kernel (for compute-sanitizer runs)."""
import os, sys, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq
from oracle import synth, scoring as sc
S = ("rgb", "warped_optical_flow")
n = 4500
st = vq.FeatureStore(n, S, [1], 1024, devices=[0])
st.fill_synthetic(5)
X = synth.database(5, n).astype(np.float64)[:, :, None, :]
T = sc.scale_target(X[11])
td = {s: {1: T[i, 0]} for i, s in enumerate(S)}
r = st.scan(td, (1.0, 1.5), 0.8, 0.73, 3e-6, topk=64, want_sims=True)
print("scan", r, len(st.matches()[0]), len(st.matches(copy=False)[0]), len(st.topk()[0]), st.sims().shape)
print("ranked", st.ranked("matches")[0][:3], st.ranked("near_misses")[1][:3])
st.append(synth.database(6, 3000)[:, :, None, :])
r = st.scan(td, (1.0, 1.5), 0.8, 0.73, 3e-6, topk=64)
print("after append", st.n_rows, r, st.ranked("near_misses")[0].shape)
Tq = np.stack([T, sc.scale_target(X[12]), sc.scale_target(X[13])]).astype(np.float32)
c, rows, scores, ms = st.scan_batch(Tq, (1.0, 1.5), 0.8, 0.73, topk=20)
c2 = st.scan_batch(np.concatenate([Tq] * 60), (1.0, 1.5), 0.8, 0.73, topk=0)[0]
print("batch 180 queries, no top-k", c2[:3].tolist(), c2[177:].tolist())
print("batch", c.tolist(), rows[:, 0])
print("labelled", st.labelled_sims(td, np.arange(0, n, 500)).shape)
print("bootstrap", st.bootstrap_target(np.array([3, 9, 10, 40]), np.array([0, 20]), 0.3).shape)
random.seed(1)
print("loss", vq.loss_grid(np.random.default_rng(0).random((50, 2)), np.arange(50) % 2 == 0, sc.weight_grid(), sc.threshold_grid(), 0.1,
                            replicates=vq.resample_labelled(50, 4, random)).shape)
st3 = vq.FeatureStore(300, S, [1, 2, 3], 1024, devices=[0])
st3.upload(0, np.random.default_rng(1).random((300, 2, 3, 1024)).astype(np.float32))
T3 = np.random.default_rng(2).random((2, 3, 1024))
print("3 splits", st3.scan({s: {p: T3[i, j] for j, p in enumerate((1, 2, 3))} for i, s in enumerate(S)}, (1.0, 1.5), 0.5, 0.4, 3e-6, topk=5))
st.close(); st3.close()
print("done")

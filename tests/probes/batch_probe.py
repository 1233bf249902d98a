"""Development probe for K3: throughput and error of the batched tensor-core scan."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq
from oracle import synth, scoring as sc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S = ("rgb", "warped_optical_flow")
st = vq.FeatureStore(n, S, [1], 1024, devices=[0])
st.fill_synthetic(synth.DEFAULT_SEED)
rows = np.arange(Q) * 37 + 18120
X = synth.rows(synth.DEFAULT_SEED, rows).astype(np.float64)[:, :, None, :]
T = np.stack([sc.scale_target(x) for x in X]).astype(np.float32)
for rep in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter()
    counts, r, s, ms = st.scan_batch(T, (1.0, 1.5), 0.8, 0.73, topk=100)
    wall = (time.perf_counter() - t0) * 1e3
    flops = 2.0 * Q * n * 2048
    print("n=%d Q=%d  call %.3f ms  kernel %.3f ms  algorithmic %.1f TFLOP/s  executed bf16 (3 MMAs) %.1f TFLOP/s  clips*queries/s %.3e  HBM %.0f GB/s"
          % (n, Q, wall, ms, flops / ms / 1e9, 3 * flops / ms / 1e9, n * Q / ms * 1e3, n * 8192 / ms / 1e6))
# error vs fp64 on a sample of rows for query 0
chk = np.arange(0, min(n, 200000), 997)
Xc = synth.rows(synth.DEFAULT_SEED, chk).astype(np.float64)[:, :, None, :]
sims, _ = sc.similarities(Xc, T[0].astype(np.float64))
s64 = sc.scores(sims, (1.0, 1.5))
if n <= 1_000_000:
    got = st.scan_batch(T[:1], (1.0, 1.5), 0.8, 0.73, debug_scores=True)[0][chk]
    err = (got - s64) / np.maximum(np.abs(s64), 1e-3)
    print("score rel err: mean %.3e  max|.| %.3e" % (err.mean(), np.abs(err).max()))
print("top-1 rows", r[:3, 0], "counts", counts[:3].tolist())

"""Where the end-to-end time of one query goes (resident 1M-clip store, host buffers in and out)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S = ("rgb", "warped_optical_flow")
st = vq.FeatureStore(n, S, [1], 1024, devices=[0])
st.fill_synthetic(20261018)
f = st.download(18120, 1)[0].astype(np.float64)
t = np.stack([vq.TargetClip._scale_feature(f[s, 0]) for s in range(2)])
td = {s: {1: t[i]} for i, s in enumerate(S)}
lower = 0.8 - 0.35 * 0.2
acc = np.zeros(5)
N = 200
for it in range(N + 10):
    t0 = time.perf_counter()
    res = st.scan(td, (1.0, 1.5), 0.8, lower, 3e-6, topk=100)
    t1 = time.perf_counter()
    k = st.topk()
    t2 = time.perf_counter()
    m = st.matches(copy=False)
    t3 = time.perf_counter()
    nm = st.near_misses(copy=False)
    t4 = time.perf_counter()
    if it >= 10:
        acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, res.scan_ms * 1e-3]
acc /= N
print("per query: scan() %.1f us (K1 on device %.1f us)  topk() %.1f us  matches() %.1f us [%d]  near_misses() %.1f us [%d]  total %.1f us"
      % (acc[0] * 1e6, acc[4] * 1e6, acc[1] * 1e6, acc[2] * 1e6, len(m[0]), acc[3] * 1e6, len(nm[0]), acc[:4].sum() * 1e6))

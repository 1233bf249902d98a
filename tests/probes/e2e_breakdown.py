"""Where the host-buffer call's time goes on one GPU (1M clips): device phases vs the C call vs the Python call."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq
from video_query_algorithms_b200 import _ffi
from video_query_algorithms_b200.store import make_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S = ("rgb", "warped_optical_flow")
st = vq.FeatureStore(n, S, [1], 1024, devices=[0])
st.fill_synthetic(20261018)
f = st.download(18120, 1)[0].astype(np.float64)
t = np.stack([vq.TargetClip._scale_feature(f[s, 0]) for s in range(2)])
tdict = {s: {1: t[i]} for i, s in enumerate(S)}
lib = _ffi.lib()
h = st.shards[0].handle
T32 = np.ascontiguousarray(t.astype(np.float32).reshape(-1))
p = make_params((1.0, 1.5), 0.8, 0.73, 3e-6, topk=100)
c = _ffi.ScanCounts()


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def c_call():
    _ffi.check(lib.vq_scan(h, _ffi.ptr(T32), C.byref(p), C.byref(c)), "vq_scan")


def c_call_nolists():
    _ffi.check(lib.vq_scan_select(h, _ffi.ptr(T32), C.byref(p), C.byref(c), None, None, None), "vq_scan_select")


def py_scan():
    st.scan(tdict, (1.0, 1.5), 0.8, 0.73, 3e-6, topk=100)


def py_full():
    st.scan(tdict, (1.0, 1.5), 0.8, 0.73, 3e-6, topk=100)
    st.topk()
    st.matches(copy=False)
    st.near_misses(copy=False)


out = {"clips": n, "us_c_vq_scan": timeit(c_call), "us_c_vq_scan_select": timeit(c_call_nolists),
       "us_py_scan": timeit(py_scan), "us_py_scan_topk_lists": timeit(py_full)}
k1, sel = np.empty(1024, np.float32), np.empty(1024, np.float32)
cnt = C.c_int32()
lib.vq_scan_phase_times(h, 1024, _ffi.ptr(k1), _ffi.ptr(sel), C.byref(cnt))
out["us_k1"] = float(np.mean(k1[:cnt.value])) * 1e3
out["us_select"] = float(np.mean(sel[:cnt.value])) * 1e3
print(out)

"""Fixed cost of the exchange kernel: world = 1 (self inbox), scan vs scan + exchange, CUDA events."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import torch
import video_query_algorithms_b200 as vq
from video_query_algorithms_b200 import _ffi
from video_query_algorithms_b200.store import make_params
lib = _ffi.lib()
n = 200_000
st = vq.FeatureStore(n, ("rgb", "warped_optical_flow"), [1], 1024, devices=[0])
st.fill_synthetic(1)
h = st.shards[0].handle
t = torch.rand(2048, device="cuda:0")
p = make_params((1.0, 1.5), 0.8, 0.73, 3e-6, topk=100)
x = C.c_void_p()
_ffi.check(lib.vq_exchange_create(C.byref(x), 0, 1, 0))
stream = torch.cuda.Stream()
def run(mode, iters=300):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record(stream)
        for _ in range(iters):
            _ffi.check(lib.vq_scan_enqueue(h, C.c_void_p(t.data_ptr()), C.byref(p), C.c_void_p(stream.cuda_stream)))
            if mode == "p2p":
                _ffi.check(lib.vq_scan_exchange_enqueue(h, x, C.c_void_p(stream.cuda_stream)))
            elif mode == "lagged":
                _ffi.check(lib.vq_scan_exchange_enqueue_lagged(h, x, C.c_void_p(stream.cuda_stream)))
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
base = run("none")
print("scan only %.1f us/step; + exchange in-step %+.1f us; + exchange lagged %+.1f us" % (base, run("p2p") - base, run("lagged") - base))
import time
torch.cuda.synchronize()
for label, nn in (("enqueue only, GPU idle at start", 50),):
    t0 = time.perf_counter()
    for _ in range(nn):
        _ffi.check(lib.vq_scan_enqueue(h, C.c_void_p(t.data_ptr()), C.byref(p), C.c_void_p(stream.cuda_stream)))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%s: %.1f us of host time per vq_scan_enqueue" % (label, (t1 - t0) / nn * 1e6))

"""K1 variants on the main shape (2 x 1024 floats per clip): steady-state time of back-to-back scans."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq
from video_query_algorithms_b200 import _ffi
from video_query_algorithms_b200.store import make_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
st = vq.FeatureStore(n, ("rgb", "warped_optical_flow"), [1], 1024, devices=[0])
st.fill_synthetic(20261018)
f = st.download(18120, 1)[0].astype(np.float64)
t = np.stack([vq.TargetClip._scale_feature(f[s, 0]) for s in range(2)])
td = {"rgb": {1: t[0]}, "warped_optical_flow": {1: t[1]}}
lib = _ffi.lib()
h = st.shards[0].handle
st.scan(td, (1.0, 1.5), 0.8, 0.73, 3e-6, topk=100)       # target now on the device (store's own buffer)
import torch
p = make_params((1.0, 1.5), 0.8, 0.73, 3e-6, topk=100)
tdev = torch.tensor(t.reshape(-1), dtype=torch.float32, device="cuda:0")
stream = torch.cuda.Stream()
for rep in range(2):
    for _ in range(20):
        _ffi.check(lib.vq_scan_enqueue(h, C.c_void_p(tdev.data_ptr()), C.byref(p), C.c_void_p(stream.cuda_stream)))
    torch.cuda.synchronize()
    tmp = np.empty(1024, np.float32); cnt = C.c_int32()
    lib.vq_scan_kernel_times(h, 1024, _ffi.ptr(tmp), C.byref(cnt))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(200):
        _ffi.check(lib.vq_scan_enqueue(h, C.c_void_p(tdev.data_ptr()), C.byref(p), C.c_void_p(stream.cuda_stream)))
    e1.record(stream)
    torch.cuda.synchronize()
    lib.vq_scan_kernel_times(h, 1024, _ffi.ptr(tmp), C.byref(cnt))
    k1 = float(np.mean(tmp[:cnt.value]))
print("K1=%s blocks=%s: step %.4f ms  K1 %.4f ms  %.0f GB/s" % (os.environ.get("VQ_SCAN_K1", "reg"), os.environ.get("VQ_SCAN_BLOCKS", "-"),
      e0.elapsed_time(e1) / 200, k1, n * 8192 / k1 / 1e6))

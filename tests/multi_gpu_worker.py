"""Worker for tests/test_gpu_multi.py (launched by torchrun, one rank per GPU): sharded scan with both
exchange implementations must equal one single-GPU scan of the whole search set."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("COMPUTE_EPS", ".000003")


def main():
    import torch
    import torch.distributed as dist
    import video_query_algorithms_b200 as vq
    from video_query_algorithms_b200 import _ffi
    from video_query_algorithms_b200.sharded import RankScan
    from video_query_algorithms_b200.store import make_params
    from oracle import scoring as sc
    from oracle import synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_local, k, seed, S = 50_000, 100, 99, ("rgb", "warped_optical_flow")
    st = vq.FeatureStore(n_local, S, [1], 1024, devices=[local], first_global_row=rank * n_local)
    st.fill_synthetic(seed)
    ref = synth.pick_reference_row(seed, n_local)
    T = sc.scale_target(synth.rows(seed, [ref]).astype(np.float64)[0][:, None, :])
    target = torch.from_numpy(T.astype(np.float32).reshape(-1)).to(dev)
    params = make_params((1.0, 1.5), 0.8, 0.73, 3e-6, topk=k)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    results = {}
    for mode in ("p2p", "p2p-lagged", "nccl"):
        rs = RankScan(st.shards[0].handle, k, local, dist, torch, exchange=mode)
        for i in range(7):                                   # several steps: exercises all four inbox slots;
            params = make_params((1.0, 1.5), 0.75 + 0.01 * i, 0.73, 3e-6, topk=k)   # every step has its own threshold, so a
            rs.enqueue(target.data_ptr(), params, stream.cuda_stream)               # merge of the wrong step shows
        rs.flush(stream.cuda_stream)                         # lagged mode: merge the last step
        torch.cuda.synchronize()
        dist.barrier()
        results[mode] = rs.result()
        rs.close()
    ok = True
    if rank == 0:
        full = vq.FeatureStore(n_local * world, S, [1], 1024, devices=[local])
        full.fill_synthetic(seed)
        res = full.scan({s: {1: T[i, 0]} for i, s in enumerate(S)}, (1.0, 1.5), 0.75 + 0.01 * 6, 0.73, 3e-6, topk=k)
        rows, scores = full.topk()
        for mode, (counts, g_rows, g_scores) in results.items():
            same = (counts[0] == res.n_match and counts[1] == res.n_near and counts[2] == res.n_tie and
                    np.array_equal(g_rows, rows) and np.array_equal(g_scores, scores))
            print("exchange %s: counts %s  equal to single-GPU scan: %s" % (mode, counts.tolist(), same))
            ok = ok and same
        full.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    st.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

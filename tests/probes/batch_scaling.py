"""BASELINE config 4 (256 queries x 10M clips, tcgen05) on N GPUs, one rank per GPU under torchrun:
clip-range shards, per-rank batched scan, per-query top-k merged across ranks (sharded.RankStore.scan_batch).

  torchrun --nproc-per-node N tests/probes/batch_scaling.py --clips-total 10000000      strong scaling
  torchrun --nproc-per-node N tests/probes/batch_scaling.py --clips-per-gpu 10000000    weak scaling

Times: `kernel_ms` = the slowest rank's device time of the batched kernels (CUDA events inside the library);
`call_ms` = barrier -> RankStore.scan_batch (host buffers in and out, allgather + merge included) -> max over ranks.
Rank 0 prints one JSON line and spot-checks query 0's merged top-10 against float64 on regenerated rows."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("COMPUTE_EPS", ".000003")
S = ("rgb", "warped_optical_flow")
SEED = 20261018
REF = 18120


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips-total", type=int, default=0)
    ap.add_argument("--clips-per-gpu", type=int, default=0)
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import video_query_algorithms_b200 as vq
    from video_query_algorithms_b200.sharded import RankStore
    from oracle import scoring as sc          # checker only
    from oracle import synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_local = a.clips_per_gpu if a.clips_per_gpu else -(-a.clips_total // world)
    first = rank * n_local
    n_total = n_local * world if a.clips_per_gpu else a.clips_total
    n_mine = max(0, min(n_local, n_total - first))
    st = vq.FeatureStore(n_mine, S, [1], 1024, devices=[local], first_global_row=first)
    st.fill_synthetic(SEED)
    rs = RankStore(st, dist, torch, dev)
    q_rows = REF + np.arange(a.queries) * 37
    X = synth.rows(SEED, q_rows).astype(np.float64)[:, :, None, :]
    T = np.stack([sc.scale_target(x) for x in X]).astype(np.float32)
    best_k, best_c = None, None
    for _ in range(a.reps):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        counts, rows, scores, ms = rs.scan_batch(T, (1.0, 1.5), 0.8, 0.73, topk=a.topk)
        call = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ms, call], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_ms, c_ms = t.tolist()
        best_k = k_ms if best_k is None else min(best_k, k_ms)
        best_c = c_ms if best_c is None else min(best_c, c_ms)
    if rank == 0:
        top = rows[0][:10]
        Xc = synth.rows(SEED, top).astype(np.float64)[:, :, None, :]
        sims, _ = sc.similarities(Xc, T[0].astype(np.float64))
        s64 = sc.scores(sims, (1.0, 1.5))
        ok = bool(np.all(np.abs(scores[0][:10] - s64) < 1e-5) and rows[0][0] == REF)
        flops = 2.0 * a.queries * n_total * 2048
        print(json.dumps({
            "config": "configs[3]: batched %d-query scoring vs %d clips on %d GPU(s), %d clips per GPU" % (a.queries, n_total, world, n_local),
            "n_gpus": world, "scaling": "weak" if a.clips_per_gpu else "strong", "kernel_ms_max_over_ranks": best_k,
            "call_ms_host_buffers_allgather_merge": best_c, "clips_x_queries_per_s": n_total * a.queries / best_k * 1e3,
            "algorithmic_tflops": flops / best_k / 1e9, "executed_tflops_bf16x2": 3 * flops / best_k / 1e9,
            "query0_counts": counts[0].tolist(), "query0_top10_matches_float64": ok}), flush=True)
    rs.close()
    st.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

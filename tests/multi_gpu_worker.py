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
    # the other exchanges of SURVEY §8(e) through RankStore: ordered lists, batched per-query top-k, labelled sims
    from video_query_algorithms_b200.sharded import RankStore
    rstore = RankStore(st, dist, torch, dev)
    tf = {s: {1: T[i, 0]} for i, s in enumerate(S)}
    r_counts, r_lists, (r_trows, r_tsc) = rstore.scan(tf, (1.0, 1.5), 0.81, 0.73, 3e-6, topk=k, lists="all")
    # the default: every rank keeps its own segment of the lists (nothing replicated), and rank 1 alone gets them whole
    d_counts, d_sh, (d_trows, d_tsc) = rstore.scan(tf, (1.0, 1.5), 0.81, 0.73, 3e-6, topk=k)
    seg_ok = True
    for c, name in enumerate(("matches", "near_misses", "ties")):
        a, b = d_sh.span(name)
        rows_l, sc_l = d_sh.local(name)
        seg_ok = seg_ok and np.array_equal(rows_l, r_lists[c][0][a:b]) and np.array_equal(sc_l, r_lists[c][1][a:b])
        seg_ok = seg_ok and d_sh.total(name) == len(r_lists[c][0]) and (b == a or d_sh.owner(name, a) == rank)
    seg_ok = seg_ok and list(d_counts) == list(r_counts) and np.array_equal(d_trows, r_trows) and np.array_equal(d_tsc, r_tsc)
    root = world - 1
    o_counts, o_lists, _ = rstore.scan(tf, (1.0, 1.5), 0.81, 0.73, 3e-6, topk=k, lists="root", root=root)
    if rank == root:
        seg_ok = seg_ok and all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(o_lists, r_lists))
    else:
        seg_ok = seg_ok and o_lists is None
    seg_flag = torch.tensor([1 if seg_ok else 0], device=dev)
    dist.all_reduce(seg_flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("rank store sharded lists and root-only gather agree with the replicated lists: %s" % bool(seg_flag.item()))
    # review round: lists stay on the ranks; counts / top-k / tie band / best near miss / sampled positions cross
    s_counts, (s_trows, s_tsc), s_ties, s_best = rstore.scan_select(tf, (1.0, 1.5), 0.81, 0.73, 3e-6, topk=k)
    rng = np.random.default_rng(5)                           # same positions on every rank
    pos_m = rng.integers(0, int(s_counts[0]), 25)
    pos_n = np.concatenate([rng.integers(0, int(s_counts[1]), 25), [0, int(s_counts[1]) - 1]])
    s_m, s_n, s_e = rstore.gather_many([("matches", pos_m), ("near_misses", pos_n), ("ties", [])])
    s_m1 = rstore.gather("matches", pos_m[:3])
    assert np.array_equal(s_m1[0], s_m[0][:3]) and np.array_equal(s_m1[1], s_m[1][:3])
    nq = 24
    Tq = np.stack([sc.scale_target(synth.rows(seed, [100 + 977 * q]).astype(np.float64)[0][:, None, :]) for q in range(nq)])
    Tq = Tq.astype(np.float32)
    b_counts, b_rows, b_sc, _ = rstore.scan_batch(Tq, (1.0, 1.5), 0.81, 0.73, topk=17)
    z_counts, z_rows, _, _ = rstore.scan_batch(Tq, (1.0, 1.5), 0.81, 0.73, topk=0)
    labelled = np.arange(7, n_local * world, 997, dtype=np.int64)
    l_sims = rstore.labelled_sims(tf, labelled)
    ok = bool(seg_flag.item())
    if rank == 0:
        full = vq.FeatureStore(n_local * world, S, [1], 1024, devices=[local])
        full.fill_synthetic(seed)
        res = full.scan(tf, (1.0, 1.5), 0.81, 0.73, 3e-6, topk=k)
        f_lists = [full.matches(), full.near_misses(), full.ties()]
        f_trows, f_tsc = full.topk()
        same = (list(r_counts) == [res.n_match, res.n_near, res.n_tie] and np.array_equal(r_trows, f_trows) and
                np.array_equal(r_tsc, f_tsc) and
                all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(r_lists, f_lists)))
        print("rank store scan (lists of %d / %d / %d entries): equal to single-GPU scan: %s"
              % (len(r_lists[0][0]), len(r_lists[1][0]), len(r_lists[2][0]), same))
        ok = ok and same and len(r_lists[0][0]) > 0
        nm_rows, nm_sc = f_lists[1]
        j = int(np.argmax(nm_sc))                            # first maximum in database order (ticket.py:335-340)
        same = (list(s_counts) == [res.n_match, res.n_near, res.n_tie] and np.array_equal(s_trows, f_trows) and
                np.array_equal(s_tsc, f_tsc) and np.array_equal(s_ties[0], f_lists[2][0]) and
                np.array_equal(s_ties[1], f_lists[2][1]) and s_best == (j, int(nm_rows[j]), float(nm_sc[j])) and
                np.array_equal(s_m[0], f_lists[0][0][pos_m]) and np.array_equal(s_m[1], f_lists[0][1][pos_m]) and
                np.array_equal(s_n[0], nm_rows[pos_n]) and np.array_equal(s_n[1], nm_sc[pos_n]) and len(s_e[0]) == 0)
        print("rank store selection scan (best near miss at list position %d, %d + %d sampled positions): equal to "
              "single-GPU scan: %s" % (s_best[0], len(pos_m), len(pos_n), same))
        ok = ok and same
        fc, fr, fs, _ = full.scan_batch(Tq, (1.0, 1.5), 0.81, 0.73, topk=17)
        same = (np.array_equal(fc, b_counts) and np.array_equal(fr, b_rows) and np.array_equal(fs, b_sc) and
                np.array_equal(z_counts, fc) and z_rows.shape == (nq, 0))
        print("rank store batched scan (%d queries, top-17): equal to single-GPU batched scan: %s" % (nq, same))
        ok = ok and same
        same = np.array_equal(full.labelled_sims(tf, labelled), l_sims)
        print("rank store labelled sims (%d rows over %d ranks): equal: %s" % (len(labelled), world, same))
        ok = ok and same
        res = full.scan({s: {1: T[i, 0]} for i, s in enumerate(S)}, (1.0, 1.5), 0.75 + 0.01 * 6, 0.73, 3e-6, topk=k)
        rows, scores = full.topk()
        for mode, (counts, g_rows, g_scores) in results.items():
            same = (counts[0] == res.n_match and counts[1] == res.n_near and counts[2] == res.n_tie and
                    np.array_equal(g_rows, rows) and np.array_equal(g_scores, scores))
            print("exchange %s: counts %s  equal to single-GPU scan: %s" % (mode, counts.tolist(), same))
            ok = ok and same
        full.close()
    # the same selection scan with the small exchanges on NCCL instead of the host mailbox
    rstore2 = RankStore(st, dist, torch, dev, host_mailbox=False)
    n_counts, n_top, n_ties, n_best = rstore2.scan_select(tf, (1.0, 1.5), 0.81, 0.73, 3e-6, topk=k)
    n_m, n_n = rstore2.gather_many([("matches", pos_m), ("near_misses", pos_n)])
    same = (rstore.mailbox is not None and np.array_equal(n_counts, s_counts) and np.array_equal(n_top[0], s_trows) and
            n_best == s_best and np.array_equal(n_m[0], s_m[0]) and np.array_equal(n_n[1], s_n[1]) and
            np.array_equal(n_ties[0], s_ties[0]))
    if rank == 0:
        print("host mailbox and NCCL small exchanges agree: %s" % same)
    ok = ok and same
    rstore.close()
    # a peer that never delivers must end the exchange kernel with an error, not wedge the GPU: the last rank skips a step
    os.environ["VQ_EXCHANGE_TIMEOUT_S"] = "0.5"
    rs_t = RankScan(st.shards[0].handle, k, local, dist, torch, exchange="p2p")
    del os.environ["VQ_EXCHANGE_TIMEOUT_S"]
    timed_out = None
    if rank != world - 1:
        rs_t.enqueue(target.data_ptr(), params, stream.cuda_stream)
        rs_t.flush(stream.cuda_stream)
        torch.cuda.synchronize()
        try:
            rs_t.result()
            timed_out = False
        except vq.VQError as e:
            timed_out = "gave up" in str(e)
    t_flag = torch.tensor([1 if timed_out in (True, None) else 0], device=dev)
    dist.all_reduce(t_flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("exchange deadline: a missing peer ends the kernel with an error: %s" % bool(t_flag.item()))
    ok = ok and bool(t_flag.item())
    rs_t.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    st.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

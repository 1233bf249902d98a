"""World-size-2 test of the multi-rank path on CPU (gloo): per-rank payloads -> one allgather -> merge.
The local scan itself needs a GPU, so each rank computes its shard's results with the oracle and packs them
exactly as the device does; what is tested is the exchange format, the global row numbering and the merge
rule (counts summed, top-k by score descending then global row ascending)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, n_total, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import scoring as sc
    from oracle import synth
    from video_query_algorithms_b200 import sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = n_total // world
    first = rank * n_local
    X = synth.database(5, n_local, first_row=first).astype(np.float64)[:, :, None, :]
    ref = synth.rows(5, [17]).astype(np.float64)[0][:, None, :]
    T = sc.scale_target(ref)
    sims, _ = sc.similarities(X, T)
    score = sc.scores(sims, (1.0, 1.5)).astype(np.float32)
    score[::7] = score[3]                                   # force ties across ranks
    m, nm = sc.classify(score.astype(np.float64), 0.8, 0.35)
    top = sc.topk_stable(score, k)
    payload = sharded.pack_payload([len(m), len(nm), 0, len(top)], first + top, score[top], k)
    merged = sharded.exchange_host(payload, dist, torch, k)
    np.save(os.path.join(out_dir, "merged_%d.npy" % rank), merged)
    np.save(os.path.join(out_dir, "score_%d.npy" % rank), score)
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [10, 64])
def test_two_rank_exchange_and_merge(tmp_path, k):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    from oracle import scoring as sc
    from video_query_algorithms_b200 import sharded
    world, n_total, port = 2, 600, _free_port()
    mp.spawn(_rank_main, args=(world, port, n_total, k, str(tmp_path)), nprocs=world, join=True)
    score = np.concatenate([np.load(tmp_path / ("score_%d.npy" % r)) for r in range(world)])
    merged = [np.load(tmp_path / ("merged_%d.npy" % r)) for r in range(world)]
    assert np.array_equal(merged[0], merged[1])             # every rank holds the same merged result
    counts, rows, scores = sharded.unpack_payload(merged[0], k)
    m, nm = sc.classify(score.astype(np.float64), 0.8, 0.35)
    assert counts[0] == len(m) and counts[1] == len(nm) and counts[3] == k
    assert np.array_equal(rows, sc.topk_stable(score, k))   # ties -> lower global row first
    assert np.array_equal(scores, score[rows])


def test_payload_roundtrip_and_padding():
    sys.path.insert(0, ROOT)
    from video_query_algorithms_b200 import sharded
    p = sharded.pack_payload([5, 2, 1, 3], [40, 7, 9], [0.9, 0.9, -0.25], 8)
    counts, rows, scores = sharded.unpack_payload(p, 8)
    assert list(counts) == [5, 2, 1, 3] and list(rows) == [40, 7, 9]
    assert np.allclose(scores, [0.9, 0.9, -0.25]) and p[4 + 3] == -1
    g = np.stack([p, sharded.pack_payload([1, 1, 0, 2], [3, 50], [0.9, 0.1], 8)])
    counts, rows, scores = sharded.unpack_payload(sharded.merge_payloads_host(g, 2, 8), 8)
    assert list(counts[:3]) == [6, 3, 1] and list(rows) == [3, 7, 40, 50, 9]


def _rank_collectives(rank, world, port, n_total, k, nq, out_dir):
    """Each rank plays its shard with the oracle and runs the §8(e) collectives on gloo: ordered lists of
    different lengths (one rank's may be empty), per-query top-k of a batch, labelled similarities."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import scoring as sc
    from oracle import synth
    from video_query_algorithms_b200 import sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = n_total // world
    first = rank * n_local
    X = synth.database(5, n_local, first_row=first).astype(np.float64)[:, :, None, :]
    refs = synth.rows(5, list(range(3, 3 + nq))).astype(np.float64)[:, :, None, :]
    counts = np.zeros((nq, 2), np.int64)
    t_rows = np.full((nq, k), -1, np.int64)
    t_sc = np.full((nq, k), -np.inf, np.float32)
    for q in range(nq):
        sims, _ = sc.similarities(X, sc.scale_target(refs[q]))
        score = sc.scores(sims, (1.0, 1.5)).astype(np.float32)
        score[::5] = np.float32(0.5)                         # ties inside and across ranks
        m, nm = sc.classify(score.astype(np.float64), 0.8, 0.35)
        counts[q] = len(m), len(nm)
        top = sc.topk_stable(score, min(k, n_local))
        t_rows[q, :len(top)] = first + top
        t_sc[q, :len(top)] = score[top]
        if q == 0:
            score0, sims0 = score, sims
    # lists: rank 1 has an impossible threshold -> empty list on that rank
    th = 0.8 if rank == 0 else 2.0
    m, _ = sc.classify(score0.astype(np.float64), th, 0.35)
    g_rows, g_sc = sharded.gather_lists(first + m, score0[m], dist, torch)
    e_rows, e_sc = sharded.gather_lists(np.empty(0, np.int64), np.empty(0, np.float32), dist, torch)
    assert len(e_rows) == 0 and len(e_sc) == 0
    b_counts, b_rows, b_sc = sharded.gather_batch(counts, t_rows, t_sc, dist, torch)
    z_counts, z_rows, _ = sharded.gather_batch(counts, t_rows[:, :0], t_sc[:, :0], dist, torch)
    assert np.array_equal(z_counts, b_counts) and z_rows.shape == (nq, 0)
    labelled = np.arange(1, n_total, 37, dtype=np.int64)     # labelled clips spread over both ranks
    part = np.zeros((len(labelled), 2))
    own = np.flatnonzero((labelled >= first) & (labelled < first + n_local))
    part[own] = sims0[labelled[own] - first]
    full = sharded.gather_sims(part, dist, torch)
    np.savez(os.path.join(out_dir, "coll_%d.npz" % rank), g_rows=g_rows, g_sc=g_sc, b_counts=b_counts, b_rows=b_rows,
             b_sc=b_sc, sims_full=full, labelled=labelled, score0=score0, sims0=sims0, counts=counts)
    dist.destroy_process_group()


def test_two_rank_lists_batch_topk_and_labelled_sims(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    from oracle import scoring as sc
    from oracle import synth
    world, n_total, k, nq, port = 2, 400, 12, 5, _free_port()
    mp.spawn(_rank_collectives, args=(world, port, n_total, k, nq, str(tmp_path)), nprocs=world, join=True)
    out = [np.load(tmp_path / ("coll_%d.npz" % r)) for r in range(world)]
    for key in ("g_rows", "g_sc", "b_counts", "b_rows", "b_sc", "sims_full"):
        assert np.array_equal(out[0][key], out[1][key]), key            # every rank holds the same result
    # lists: rank 0's matches (rank 1 contributed none), in database order
    score0 = out[0]["score0"]
    m, _ = sc.classify(score0.astype(np.float64), 0.8, 0.35)
    assert np.array_equal(out[0]["g_rows"], m) and np.array_equal(out[0]["g_sc"], score0[m])
    # batch: against the whole search set scored in one piece
    X = synth.database(5, n_total).astype(np.float64)[:, :, None, :]
    refs = synth.rows(5, list(range(3, 3 + nq))).astype(np.float64)[:, :, None, :]
    n_local = n_total // world
    for q in range(nq):
        sims, _ = sc.similarities(X, sc.scale_target(refs[q]))
        score = sc.scores(sims, (1.0, 1.5)).astype(np.float32)
        for r in range(world):
            score[r * n_local:(r + 1) * n_local:5] = np.float32(0.5)
        mm, nm = sc.classify(score.astype(np.float64), 0.8, 0.35)
        assert list(out[0]["b_counts"][q]) == [len(mm), len(nm)]
        top = sc.topk_stable(score, k)
        assert np.array_equal(out[0]["b_rows"][q], top) and np.array_equal(out[0]["b_sc"][q], score[top])
    # labelled similarities: the owners' values bit for bit
    sims_all = np.concatenate([out[r]["sims0"] for r in range(world)])
    assert np.array_equal(out[0]["sims_full"], sims_all[out[0]["labelled"]])

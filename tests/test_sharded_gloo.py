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
    # single-query summary + packed lists: rank 1 has an impossible threshold -> empty match list on that rank; the
    # tie band is faked from rows whose score is exactly 0.5 (more than TIE_CAP entries on each rank -> list gather)
    th = 0.8 if rank == 0 else 2.0
    m, _ = sc.classify(score0.astype(np.float64), th, 0.35)
    _, nm = sc.classify(score0.astype(np.float64), 0.51, 0.35)   # ties at 0.5 top the near-miss band: the first one wins
    ties = np.flatnonzero(score0 == np.float32(0.5))
    sharded.TIE_CAP = 16                                     # (512 in production; every rank uses the same value)
    assert len(ties) > sharded.TIE_CAP
    lb = None
    if len(nm):
        j = int(np.argmax(score0[nm]))                       # first maximum in list order
        lb = (j, first + int(nm[j]), float(score0[nm[j]]))
    top0 = sc.topk_stable(score0, k)
    rec = sharded.summary_record(first, [len(m), len(nm), len(ties)], first + top0, score0[top0], k, lb,
                                 (first + ties, score0[ties]))
    sm = sharded.exchange_summary(rec, k, dist, torch)
    assert sm.ties is None
    # the same exchange through the shared-memory mailbox (both ranks run on this host) gives the same summary
    mb = sharded.HostMailbox.try_create(dist)
    assert mb is not None
    sm_mb = sharded.exchange_summary(rec, k, dist, torch, mailbox=mb)
    assert np.array_equal(sm_mb.counts, sm.counts) and np.array_equal(sm_mb.first_rows, sm.first_rows)
    assert np.array_equal(sm_mb.topk[0], sm.topk[0]) and np.array_equal(sm_mb.topk[1], sm.topk[1])
    assert sm_mb.near_best == sm.near_best and sm_mb.ties is None
    tiny = sharded.HostMailbox(dist, slot_bytes=64)          # records that do not fit the slots use `dist`
    assert not tiny.fits(rec) and sharded.exchange_summary(rec, k, dist, torch, mailbox=tiny).near_best == sm.near_best
    tiny.close()
    def dev(idx):                                            # what a scan leaves on the device: int32 LOCAL rows, fp32 scores
        return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(score0[idx])

    mine = [dev(m), dev(nm), dev(ties)]
    (g_rows, g_sc), (n_rows, n_sc), (tie_rows, tie_sc) = sharded.gather_lists_torch(mine, [0, 1, 2], sm, dist, torch)
    views = sharded.gather_lists_torch(mine[:2], [0, 1], sm, dist, torch, copy=False)
    assert np.array_equal(views[0][0], g_rows) and np.array_equal(views[1][1], n_sc)
    # the same lists delivered to ONE rank (unpadded point-to-point segments), and left sharded
    for root in (0, 1):
        at_root = sharded.gather_lists_root(mine, [0, 1, 2], sm, dist, torch, root=root)
        if rank == root:
            for (a_r, a_s), (b_r, b_s) in zip(at_root, [(g_rows, g_sc), (n_rows, n_sc), (tie_rows, tie_sc)]):
                assert np.array_equal(a_r, b_r) and np.array_equal(a_s, b_s)
        else:
            assert at_root is None
    sh = sharded.ShardedLists(sm, rank, [(first + m, score0[m]), (first + nm, score0[nm]), (first + ties, score0[ties])])
    for name, (w_rows, w_sc) in (("matches", (g_rows, g_sc)), ("near_misses", (n_rows, n_sc)), ("ties", (tie_rows, tie_sc))):
        a, b = sh.span(name)
        assert np.array_equal(sh.local(name)[0], w_rows[a:b]) and np.array_equal(sh.local(name)[1], w_sc[a:b])
        assert sh.total(name) == len(w_rows) and all(sh.owner(name, p) == rank for p in range(a, b))
        assert sh.span(name, 0)[1] == sh.span(name, 1)[0] and sh.span(name, 0)[0] == 0
    # a short tie band rides in the summary record itself; k = 0 and no near miss at all
    rec2 = sharded.summary_record(first, [0, 0, 2], [], [], 0, None, (first + ties[:2], score0[ties[:2]]))
    sm2 = sharded.exchange_summary(rec2, 0, dist, torch)
    assert sm2.near_best is None and len(sm2.topk[0]) == 0 and list(sm2.total) == [0, 0, 2 * world]
    assert sharded.gather_lists_torch([dev(np.empty(0, np.int64))], [0], sm2, dist, torch)[0][0].shape == (0,)
    # sampled positions of two lists with one collective: last, first, first of rank 1, repeats
    n_near = sm.counts[:, 1]
    pos = np.array([int(n_near.sum()) - 1, 0, int(n_near[0]), 3, 3], np.int64)
    pos_m = np.array([len(g_rows) - 1, 0], np.int64)
    local = {0: (first + m, score0[m]), 1: (first + nm, score0[nm])}
    (p_rows, p_sc), (pm_rows, pm_sc), (e_rows, _) = sharded.gather_positions_multi(
        [(1, pos), (0, pos_m), (2, [])], sm, lambda c, loc: (local[c][0][loc], local[c][1][loc]), dist, torch)
    assert len(e_rows) == 0
    (q_rows, q_sc), (qm_rows, qm_sc), _ = sharded.gather_positions_multi(
        [(1, pos), (0, pos_m), (2, [])], sm, lambda c, loc: (local[c][0][loc], local[c][1][loc]), dist, torch, mailbox=mb)
    assert np.array_equal(q_rows, p_rows) and np.array_equal(q_sc, p_sc) and np.array_equal(qm_rows, pm_rows)
    mb.close()
    try:
        sharded.gather_positions_multi([(1, [int(n_near.sum())])], sm, None, dist, torch)
        raise AssertionError("position beyond the list accepted")
    except sharded._ffi.VQError:
        pass
    b_counts, b_rows, b_sc = sharded.gather_batch(counts, t_rows, t_sc, dist, torch)
    z_counts, z_rows, _ = sharded.gather_batch(counts, t_rows[:, :0], t_sc[:, :0], dist, torch)
    assert np.array_equal(z_counts, b_counts) and z_rows.shape == (nq, 0)
    labelled = np.arange(1, n_total, 37, dtype=np.int64)     # labelled clips spread over both ranks
    part = np.zeros((len(labelled), 2))
    own = np.flatnonzero((labelled >= first) & (labelled < first + n_local))
    part[own] = sims0[labelled[own] - first]
    full = sharded.gather_sims(part, dist, torch)
    np.savez(os.path.join(out_dir, "coll_%d.npz" % rank), g_rows=g_rows, g_sc=g_sc, p_rows=p_rows, p_sc=p_sc, pos=pos,
             best=np.array(sm.near_best[:2], np.int64), best_sc=np.float32(sm.near_best[2]), n_rows=n_rows, n_sc=n_sc,
             t_rows=tie_rows, t_sc=tie_sc, pm_rows=pm_rows, pm_sc=pm_sc, s_top=sm.topk[0], s_top_sc=sm.topk[1],
             s_total=sm.total, sm2_ties=sm2.ties[0], b_counts=b_counts, b_rows=b_rows,
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
    for key in ("g_rows", "g_sc", "b_counts", "b_rows", "b_sc", "sims_full", "p_rows", "p_sc", "best", "best_sc",
                "n_rows", "n_sc", "t_rows", "t_sc", "pm_rows", "pm_sc", "s_top", "s_top_sc", "s_total", "sm2_ties"):
        assert np.array_equal(out[0][key], out[1][key]), key            # every rank holds the same result
    # lists: rank 0's matches (rank 1 contributed none), in database order
    score0 = out[0]["score0"]
    m, _ = sc.classify(score0.astype(np.float64), 0.8, 0.35)
    assert np.array_equal(out[0]["g_rows"], m) and np.array_equal(out[0]["g_sc"], score0[m])
    # near-miss / tie lists, merged top-k, sampled positions and the best near miss against the search set in one piece
    score_all = np.concatenate([out[r]["score0"] for r in range(world)])
    _, nm = sc.classify(score_all.astype(np.float64), 0.51, 0.35)
    ties = np.flatnonzero(score_all == np.float32(0.5))
    o = out[0]
    assert np.array_equal(o["n_rows"], nm) and np.array_equal(o["n_sc"], score_all[nm])
    assert np.array_equal(o["t_rows"], ties) and np.array_equal(o["t_sc"], score_all[ties])
    assert list(o["s_total"]) == [len(m), len(nm), len(ties)]
    top = sc.topk_stable(score_all, k)
    assert np.array_equal(o["s_top"], top) and np.array_equal(o["s_top_sc"], score_all[top])
    assert np.array_equal(o["p_rows"], nm[o["pos"]]) and np.array_equal(o["p_sc"], score_all[nm[o["pos"]]])
    assert list(o["pm_rows"]) == [m[-1], m[0]] and list(o["pm_sc"]) == [score_all[m[-1]], score_all[m[0]]]
    j = int(np.argmax(score_all[nm]))
    assert list(o["best"]) == [j, nm[j]] and o["best_sc"] == score_all[nm[j]]
    assert score_all[nm[j]] == np.float32(0.5) and nm[j] == 0             # the tie on both ranks: first in database order
    n_local = n_total // world
    assert list(o["sm2_ties"]) == [0, 5, n_local, n_local + 5]
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

"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): clip-range shards + exchange + merge
must reproduce a single-GPU scan of the whole search set bit for bit, for both exchange implementations."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_scan_equals_single_gpu_scan():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert "exchange p2p:" in r.stdout and "exchange p2p-lagged:" in r.stdout and "exchange nccl:" in r.stdout
    assert "rank store scan" in r.stdout and "rank store batched scan" in r.stdout and "rank store labelled sims" in r.stdout
    assert "rank store selection scan" in r.stdout and "host mailbox and NCCL small exchanges agree: True" in r.stdout
    assert "rank store sharded lists and root-only gather agree with the replicated lists: True" in r.stdout
    assert "exchange deadline: a missing peer ends the kernel with an error: True" in r.stdout


def _devices():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return list(range(min(n, 4)))


def test_single_process_store_over_real_devices_equals_one_device():
    """The broker's arrangement (one process, reference src/broker.py:62-92) with the search set sharded over several REAL
    devices: FeatureStore.scan is one library call (vq_scan_multi) that enqueues on every shard's stream from this thread
    and waits once.  Everything a single-device store returns must come back bit for bit: scores, ordered lists, tie
    band, merged top-k, best near miss, sampled gathers, device-ranked lists, labelled fp64 similarities."""
    import numpy as np
    import video_query_algorithms_b200 as vq
    from oracle import scoring as sc
    from oracle import synth
    devs = _devices()
    n, seed, S = 300_007, 31, ("rgb", "warped_optical_flow")
    one = vq.FeatureStore(n, S, [1], 1024, devices=[0])
    many = vq.FeatureStore(n, S, [1], 1024, devices=devs)
    assert len(many.shards) == len(devs) and {sh.device for sh in many.shards} == set(devs)
    one.fill_synthetic(seed)
    many.fill_synthetic(seed)
    ref = synth.pick_reference_row(seed, n)
    T = sc.scale_target(synth.rows(seed, [ref]).astype(np.float64)[0][:, None, :])
    tf = {s: {1: T[i, 0]} for i, s in enumerate(S)}
    lo = sc.lower_limit(0.8, 0.35)
    for lists in (True, False):
        a = one.scan(tf, (1.0, 1.5), 0.8, lo, 3e-6, topk=100, lists=lists)
        b = many.scan(tf, (1.0, 1.5), 0.8, lo, 3e-6, topk=100, lists=lists)
        assert (a.n_match, a.n_near, a.n_tie, a.n_topk) == (b.n_match, b.n_near, b.n_tie, b.n_topk) and a.n_match > 1000
        assert np.array_equal(one.scores(), many.scores())
        for f in ("topk", "ties") + (("matches", "near_misses") if lists else ()):
            ra, sa = getattr(one, f)()
            rb, sb = getattr(many, f)()
            assert np.array_equal(ra, rb) and np.array_equal(sa, sb), f
        if lists:
            for f in ("matches", "near_misses"):
                ra, sa = one.ranked(f)
                rb, sb = many.ranked(f)
                assert np.array_equal(ra, rb) and np.array_equal(sa, sb)
                place = np.random.default_rng(1).permutation(len(ra)).astype(np.uint32)
                pa, qa = one.rank_list(f, place)
                pb, qb = many.rank_list(f, place)
                assert np.array_equal(pa, pb) and np.array_equal(qa, qb)
        else:
            assert one.near_best() == many.near_best()
            pos = np.random.default_rng(2).integers(0, a.n_match, 64)
            ga, gb = one.gather("matches", pos), many.gather("matches", pos)
            assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1])
    rows = np.arange(5, n, 4999, dtype=np.int64)
    assert np.array_equal(one.labelled_sims(tf, rows), many.labelled_sims(tf, rows))
    assert np.array_equal(one.scores_at(rows), many.scores_at(rows))
    one.close()
    many.close()


def test_compute_matches_on_a_store_spanning_real_devices(tmp_path, monkeypatch):
    """compute_matches end to end (new -> revise rounds of golden scenario D, 10k clips; new -> revise -> finalize of the
    fixture scenario A) with the ticket's store sharded over the box's devices: states, weights, threshold, scores, the
    selected clips in order and the final report equal the one-device run's."""
    import random
    import numpy as np
    import video_query_algorithms_b200 as vq
    from fake_api import FakeRepository
    from scenarios import Scenario
    devs = _devices()
    for name in ("D_synth10k", "A_brooklyn_bagging"):
        runs = []
        for j, devices in enumerate(([0], devs)):
            work = tmp_path / ("%s_run%d" % (name, j)) / "work"      # the report goes to ../final_reports/ of the cwd
            work.mkdir(parents=True)
            monkeypatch.chdir(work)
            scn = Scenario(name)
            api, qid = scn.build_api()
            vq.invalidate()
            rule, tickets, out = scn.label_rule(), [], []

            def factory(job, url):
                t = vq.Ticket(job, url, client=api.client(), devices=devices)
                t.topk = 10
                tickets.append(t)
                return t

            for i, r in enumerate(scn.rounds):
                if i > 0:
                    api.label_latest_round(qid, rule)
                api.request(qid, r["kind"])
                hp = vq.Hyperparameter(**scn.hp())
                random.seed(a=scn.seed)
                vq.compute_matches(FakeRepository(api), hp, ticket_factory=factory)
                t = tickets[-1]
                assert len(t.feature_store().shards) == len(devices)
                out.append((api.queries[qid]["process_state"], dict(hp.weights), hp.threshold, t.scores.array().copy(),
                            list(t.matches.items()), list(t.ranked[0]), list(t.tie_band)))
            out.append(list(api.uploaded_reports))
            runs.append(out)
        a, b = runs
        assert len(a) == len(b)
        for ra, rb in zip(a[:-1], b[:-1]):
            assert ra[0] == rb[0] and ra[1] == rb[1] and ra[2] == rb[2] and np.array_equal(ra[3], rb[3])
            assert ra[4] == rb[4] and ra[5] == rb[5] and ra[6] == rb[6]
        strip = lambda rep: [ln for ln in rep.splitlines() if not ln.startswith("Query:")]
        assert [strip(x) for x in a[-1]] == [strip(x) for x in b[-1]]
    vq.invalidate()

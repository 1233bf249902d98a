"""GPU parity tests: the CUDA path (through the C ABI, include/vq.h) against the CPU oracle and the
golden vectors recorded from the reference.  Run on the B200 box: pytest -m gpu.

Tolerances (BASELINE.json north_star): scores within 1e-5 relative of the float64 reference
(checked as |d| <= 1e-5 * max(|ref|, 0.05): a score is 1 - distance, so near zero its fp32 absolute
error of ~1e-7 - one ulp of the similarity - is not a meaningful relative error); match / near-miss
sets and top-k rows bit-exact except for rows within COMPUTE_EPS of a boundary (those must appear
in the reported tie band); updated weights / threshold within 1e-5.
"""
import os
import random

import numpy as np
import pytest

from oracle import bootstrap as ob
from oracle import scoring as sc
from oracle import synth
from scenarios import SCENARIOS, Scenario

pytestmark = pytest.mark.gpu

EPS = 3e-6
STREAMS = ("rgb", "warped_optical_flow")


@pytest.fixture(scope="module")
def vq():
    import video_query_algorithms_b200 as vq
    return vq


def tdict(T, splits=(1,)):
    return {s: {p: T[si, pi] for pi, p in enumerate(splits)} for si, s in enumerate(STREAMS)}


def assert_scores_close(got, want, tol=1e-5, floor=0.05):
    err = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), floor)
    assert err.max() <= tol, "max rel err %.3e at row %d" % (err.max(), err.argmax())


def assert_sets_match(got_rows, want_rows, score64, boundaries, eps=EPS):
    """Bit-exact except rows whose float64 score is within eps of a boundary."""
    diff = np.setxor1d(got_rows, want_rows)
    for r in diff:
        assert min(abs(score64[r] - b) for b in boundaries) < eps, \
            "row %d (score %.9f) differs outside the tie band" % (r, score64[r])
    return len(diff)


# ---------------------------------------------------------------------------- the C ABI from C
def test_plain_c_consumer_of_the_abi_gets_what_the_python_host_gets(vq, tmp_path):
    """tests/c_abi/scan_consumer.c — C99, only include/vq.h and libvq_b200.so, no Python in the process — builds a synthetic
    shard, derives the target from one of its rows, scans with host buffers and prints counts, top-k and list checksums;
    the same query through the Python host (ctypes) must give the same numbers bit for bit.  This is the boundary a cgo /
    JNI / N-API binding would sit on."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "scan_consumer")
    lib_dir = os.path.join(root, "video_query_algorithms_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c_abi", "scan_consumer.c"), "-L", lib_dir, "-lvq_b200",
                    "-Wl,-rpath," + lib_dir, "-o", exe], check=True, capture_output=True, text=True)
    n, seed, ref = 50000, 20261018, 18120
    out = subprocess.run([exe, str(n), str(seed), str(ref)], check=True, capture_output=True, text=True, timeout=300).stdout
    got = json.loads(out.strip().splitlines()[-1])
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.fill_synthetic(seed)
    f = st.download(ref, 1)[0].astype(np.float64)
    T = np.stack([vq.TargetClip._scale_feature(f[s, 0]) for s in range(2)])[:, None, :]
    res = st.scan(tdict(T), (1.0, 1.5), 0.8, 0.8 - 0.35 * (1 - 0.8), EPS, topk=100)
    (m_rows, m_sc), (n_rows, n_sc), (k_rows, k_sc) = st.matches(), st.near_misses(), st.topk()
    assert (got["n_match"], got["n_near"], got["n_tie"], got["n_topk"]) == (res.n_match, res.n_near, res.n_tie, res.n_topk)
    assert res.n_match > 100 and res.n_near > 100
    assert got["match_rows_sum"] == int(m_rows.sum()) and got["near_rows_sum"] == int(n_rows.sum())
    assert got["match_bits_sum"] == int(m_sc.view(np.uint32).astype(np.uint64).sum())
    assert got["near_bits_sum"] == int(n_sc.view(np.uint32).astype(np.uint64).sum())
    assert got["topk_rows"] == k_rows.tolist() and np.float32(got["top1_score"]) == k_sc[0]
    st.close()


# ---------------------------------------------------------------------------- generator
def test_synthetic_generator_is_bit_identical_to_cpu_twin(vq):
    st = vq.FeatureStore(5000, STREAMS, [1], 1024, devices=[0], first_global_row=123456789)
    st.fill_synthetic(synth.DEFAULT_SEED)
    for lo, n in ((0, 3), (2500, 64), (4999, 1)):
        got = st.download(lo, n)[:, :, 0, :]
        want = synth.rows(synth.DEFAULT_SEED, 123456789 + lo + np.arange(n))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    st.close()


# ---------------------------------------------------------------------------- scan parity
@pytest.mark.parametrize("n", [10000, 4097, 33, 1])
def test_scan_matches_oracle_on_synthetic(vq, n):
    seed = synth.DEFAULT_SEED
    X = synth.database(seed, n)                                   # [n, 2, 1024] float32
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.upload(0, X[:, :, None, :])
    X64 = X.astype(np.float64)[:, :, None, :]
    ref = synth.pick_reference_row(seed, n)
    T = sc.scale_target(X64[ref])
    sims64, _ = sc.similarities(X64, T)
    th, near = 0.8, 0.35
    lo = sc.lower_limit(th, near)
    for w in ((1.0, 1.5), (1.0, 0.5)):
        s64 = sc.scores(sims64, w)
        k = min(100, n)
        res = st.scan(tdict(T), w, th, lo, EPS, topk=k, want_sims=True)
        got = st.scores()
        assert_scores_close(got, s64)
        assert np.abs(st.sims() - sims64).max() <= 1e-5 * np.abs(sims64).max()
        m64, nm64 = sc.classify(s64, th, near)
        m_rows, m_sc = st.matches()
        n_rows, n_sc = st.near_misses()
        assert res.n_match == len(m_rows) and res.n_near == len(n_rows)
        assert np.all(np.diff(m_rows) > 0) and np.all(np.diff(n_rows) > 0)       # database order
        assert np.array_equal(m_sc, got[m_rows]) and np.array_equal(n_sc, got[n_rows])
        assert_sets_match(m_rows, m64, s64, (th,))
        assert_sets_match(n_rows, nm64, s64, (th, lo))
        # the device's own sets are exactly {fp32 score >= th} in float64 comparison
        assert np.array_equal(m_rows, np.flatnonzero(got.astype(np.float64) >= th))
        assert np.array_equal(n_rows, np.flatnonzero((got.astype(np.float64) >= lo) & (got.astype(np.float64) < th)))
        t_rows, _ = st.ties()
        g64 = got.astype(np.float64)
        assert np.array_equal(t_rows, np.flatnonzero((np.abs(g64 - th) < EPS) | (np.abs(g64 - lo) < EPS)))
        # ranking: stable descending (ticket.py:266)
        r_rows, r_sc = st.topk()
        assert len(r_rows) == k
        assert np.array_equal(r_rows, sc.topk_stable(got, k))               # exact on the device's scores
        want = sc.topk_stable(s64, k)
        for a, b in zip(r_rows, want):                                       # vs float64: same up to fp32 ties
            assert a == b or abs(s64[a] - s64[b]) < EPS
    st.close()


def test_topk_with_massive_ties_keeps_database_order(vq):
    """Thousands of identical rows share one score: the candidate set overflows the fast path
    (radix-select branch) and the ranking must fall back to database order (stable sort, ticket.py:266)."""
    base = synth.database(3, 40)
    X = np.repeat(base[7:8], 6000, axis=0)
    X[100], X[2000], X[5999] = base[1], base[2], base[3]
    st = vq.FeatureStore(len(X), STREAMS, [1], 1024, devices=[0])
    st.upload(0, X[:, :, None, :])
    T = sc.scale_target(base[7].astype(np.float64)[:, None, :])
    st.scan(tdict(T), (1.0, 1.5), 0.8, 0.73, EPS, topk=1000)
    got = st.scores()
    rows, scores = st.topk()
    assert np.array_equal(rows, sc.topk_stable(got, 1000))
    assert np.array_equal(rows[:3], [0, 1, 2]) and 100 not in rows[:99]
    st.close()


@pytest.mark.parametrize("n", [3000, 70001])
def test_full_ranking_of_the_match_list(vq, n):
    """Finalize-round ranking on the device (vq_fetch_ranked): the whole match / near-miss list by score descending,
    database order among equal scores — with heavy ties (scores quantised by construction) and list lengths that
    are not powers of two, below and above one 2048-key sort tile."""
    rng = np.random.default_rng(n)
    base = synth.database(21, 64)                              # 64 distinct clips, each repeated many times -> ties
    pick = rng.integers(0, 64, n)
    X = base[pick][:, :, None, :]
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.upload(0, X)
    T = sc.scale_target(base[7].astype(np.float64)[:, None, :])
    res = st.scan(tdict(T), (1.0, 1.5), 0.62, 0.55, EPS, topk=10)
    assert max(res.n_match, res.n_near) > 4 * 2048 or n < 5000     # several sort tiles and global steps
    got = st.scores()
    for which, (rows, sc_) in (("matches", st.matches()), ("near_misses", st.near_misses())):
        r, s_ = st.ranked(which)
        order = np.lexsort((rows, -sc_.astype(np.float64)))    # score descending, row ascending
        assert np.array_equal(r, rows[order]) and np.array_equal(s_, sc_[order])
        assert np.array_equal(s_, got[r])
    m2 = st.matches(copy=False)                                # the database-order mirror is untouched by the ranking
    assert np.array_equal(m2[0], np.flatnonzero(got.astype(np.float64) >= 0.62))
    st.close()


def test_nan_rows_maximum_topk_and_topk_beyond_the_store(vq):
    """Edge cases: a clip whose features hold a NaN has a NaN score and belongs to no list, exactly like the
    reference (every comparison with NaN is False, ticket.py:326-327) and does not disturb its neighbours;
    top-k at the ABI maximum (1024) and larger than the store."""
    n = 2600
    X = synth.database(31, n)
    bad = [0, 77, 1299, n - 1]
    X[bad, 1, 5] = np.nan
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.upload(0, X[:, :, None, :])
    X64 = X.astype(np.float64)[:, :, None, :]
    T = sc.scale_target(X64[10])
    th, lo = 0.7, 0.6
    res = st.scan(tdict(T), (1.0, 1.5), th, lo, EPS, topk=1024)
    got = st.scores()
    sims64, _ = sc.similarities(X64, T)
    s64 = sc.scores(sims64, (1.0, 1.5))
    assert np.isnan(got[bad]).all() and np.isnan(s64[bad]).all()
    ok = np.setdiff1d(np.arange(n), bad)
    assert_scores_close(got[ok], s64[ok])
    assert not np.intersect1d(st.matches()[0], bad).size and not np.intersect1d(st.near_misses()[0], bad).size
    assert res.n_match == np.count_nonzero(got[ok].astype(np.float64) >= th)
    rows, scs = st.topk()
    assert len(rows) == 1024 and not np.intersect1d(rows, bad).size
    want = ok[sc.topk_stable(got[ok], 1024)]
    assert np.array_equal(rows, want)
    gb = st.scan_batch(T[None].astype(np.float32), (1.0, 1.5), th, lo, debug_scores=True)[0]
    assert np.isnan(gb[bad]).all()
    cb = st.scan_batch(T[None].astype(np.float32), (1.0, 1.5), th, lo, topk=5)[0]
    assert cb[0, 0] == np.count_nonzero(gb[ok].astype(np.float64) >= th)
    small = vq.FeatureStore(7, STREAMS, [1], 1024, devices=[0])
    small.upload(0, X[100:107, :, None, :])
    r = small.scan(tdict(T), (1.0, 1.5), th, lo, EPS, topk=100)      # k larger than the store: all rows, ranked
    assert r.n_topk == 7 and np.array_equal(small.topk()[0], sc.topk_stable(small.scores(), 7))
    with pytest.raises(vq.VQError):
        small.scan(tdict(T), (1.0, 1.5), th, lo, EPS, topk=1025)
    st.close()
    small.close()


def test_selection_scan_gathers_what_the_full_lists_hold(vq):
    """vq_scan_select / vq_gather_list (the review round's path: lists stay on the device) against the full-list
    scan: same counts, tie band and top-k; gathered entries equal the list entries at those positions; the best near
    miss is the first maximum of the near-miss list in database order — also across two shards."""
    n = 30011
    X = synth.database(17, n)
    X[5000] = X[77]                                               # a duplicated clip: equal scores, two positions
    for devices in ([0], [0, 0]):
        st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=devices)
        st.upload(0, X[:, :, None, :])
        ref = int(np.argmin(np.abs(synth.alpha(np.arange(n), 17) - 0.9)))      # a clip with many matches and near misses
        T = sc.scale_target(X[ref].astype(np.float64)[:, None, :])
        full = st.scan(tdict(T), (1.0, 1.5), 0.8, 0.73, EPS, topk=50)
        assert full.n_match > 100 and full.n_near > 100
        m, nm, ti, tk = st.matches(), st.near_misses(), st.ties(), st.topk()
        lite = st.scan(tdict(T), (1.0, 1.5), 0.8, 0.73, EPS, topk=50, lists=False)
        assert (full.n_match, full.n_near, full.n_tie, full.n_topk) == (lite.n_match, lite.n_near, lite.n_tie, lite.n_topk)
        assert np.array_equal(st.ties(copy=False)[0], ti[0]) and np.array_equal(st.topk()[0], tk[0])
        rng = np.random.default_rng(0)
        for name, (rows, scs) in (("matches", m), ("near_misses", nm)):
            pos = rng.choice(len(rows), size=min(40, len(rows)), replace=False)
            r, s = st.gather(name, pos)
            assert np.array_equal(r, rows[pos]) and np.array_equal(s, scs[pos])
            with pytest.raises(vq.VQError):
                st.gather(name, [len(rows)])
        jbest = int(np.argmax(nm[1]))
        assert st.near_best() == (jbest, int(nm[0][jbest]), float(nm[1][jbest]))
        assert np.array_equal(st.matches()[0], m[0])              # whole lists remain fetchable (device copy path)
        # a near-miss band that contains the duplicated pair as its maximum: the first one in database order wins
        sc77 = float(st.scores()[77])
        th_up = float(np.nextafter(np.float32(sc77), np.float32(2.0)))          # the band's maximum is exactly sc77
        st.scan(tdict(T), (1.0, 1.5), th_up, sc77 - 0.05, EPS, lists=False)
        pos_b, row_b, s_b = st.near_best()
        assert row_b == 77 and s_b == sc77
        st.close()


def test_lists_larger_than_the_host_mirror_grow_it(vq):
    """The pinned host mirror starts at max(n/8, 65536) entries per list; a scan whose lists are longer sets the
    overflow flag, the mirror grows and the lists are published again: whole-store match list, then a whole-store
    near-miss list, then a normal scan, all equal to what the scores say."""
    n = 200_000
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0], first_global_row=7)
    st.fill_synthetic(4)
    ref = synth.pick_reference_row(4, n)
    T = sc.scale_target(synth.rows(4, [ref]).astype(np.float64)[0][:, None, :])
    for th, lo in ((-5.0, -6.0), (5.0, -5.0), (0.8, 0.73)):
        res = st.scan(tdict(T), (1.0, 1.5), th, lo, EPS, topk=10)
        got = st.scores().astype(np.float64)
        want_m = np.flatnonzero(got >= th) + 7
        want_n = np.flatnonzero((got >= lo) & (got < th)) + 7
        for copy in (False, True):
            m, nm = st.matches(copy=copy), st.near_misses(copy=copy)
            assert res.n_match == len(want_m) and np.array_equal(m[0], want_m) and np.array_equal(m[1], got[want_m - 7].astype(np.float32))
            assert res.n_near == len(want_n) and np.array_equal(nm[0], want_n)
    assert res.n_match < 65536 < n                                  # the first two scans overflowed the initial mirror
    st.close()


def test_scan_handles_empty_store(vq):
    st = vq.FeatureStore(0, STREAMS, [1], 1024, devices=[0])
    T = np.ones((2, 1, 1024))
    res = st.scan(tdict(T), (1.0, 1.5), 0.8, 0.7, EPS, topk=10)
    assert (res.n_match, res.n_near, res.n_tie) == (0, 0, 0)
    assert len(st.topk()[0]) == 0
    st.close()


def test_scan_three_splits_with_missing_slots(vq):
    """Fixture-shaped rows (3 splits); some clips lack a split: mean over the splits present."""
    rng = np.random.default_rng(5)
    n = 700
    X = rng.random((n, 2, 3, 1024)).astype(np.float32)
    present = np.ones((n, 2, 3), bool)
    present[rng.integers(0, n, 60), rng.integers(0, 2, 60), rng.integers(0, 3, 60)] = False
    present[:, :, 0] = True                                    # never lose every split
    Xz = X * present[..., None]
    st = vq.FeatureStore(n, STREAMS, [1, 2, 3], 1024, devices=[0])
    st.upload(0, Xz)
    st.set_present(present)
    T = sc.scale_target(X[7].astype(np.float64))
    sims64, cnt = sc.similarities(X.astype(np.float64), T, present)
    s64 = sc.scores(sims64, (1.0, 1.5))
    st.scan(tdict(T, (1, 2, 3)), (1.0, 1.5), 0.8, 0.7, EPS, topk=50, want_sims=True)
    assert_scores_close(st.scores(), s64)
    lab = st.labelled_sims(tdict(T, (1, 2, 3)), np.arange(0, n, 7))
    assert np.abs(lab - sims64[::7]).max() < 1e-12 * np.abs(sims64).max() + 1e-13
    st.close()


@pytest.mark.parametrize("streams,splits,dim", [(("rgb",), [1], 1024), (("a", "b", "c"), [1, 2], 512),
                                                 (("a", "b", "c", "d"), [1], 256), (("rgb", "flow"), [1], 2048)])
def test_generic_shapes_single_and_batched(vq, streams, splits, dim):
    """Any stream count <= 4, several splits, other feature sizes: the shared-memory-target variant of K1 and
    the stream-sequential K3 against the oracle (compute_scores is generic in the streams, ticket.py:176)."""
    rng = np.random.default_rng(len(streams) * 100 + dim)
    n, S, P = 1500, len(streams), len(splits)
    X = (rng.random((n, S, P, dim)) * rng.random((n, 1, 1, 1)) * 3).astype(np.float32)
    w = [1.0, 1.5, 0.7, 2.0][:S]
    st = vq.FeatureStore(n, streams, splits, dim, devices=[0])
    st.upload(0, X)
    X64 = X.astype(np.float64)
    refs = [5, 77, 300]
    T = np.stack([sc.scale_target(X64[r]) for r in refs])                 # [3, S, P, dim]
    td = lambda t: {s: {p: t[si, pi] for pi, p in enumerate(splits)} for si, s in enumerate(streams)}
    for qi, r in enumerate(refs):
        sims64, _ = sc.similarities(X64, T[qi])
        s64 = sc.scores(sims64, w)
        th = float(np.quantile(s64, 0.9))
        lo = th - 0.1
        res = st.scan(td(T[qi]), w, th, lo, EPS, topk=20, want_sims=True)
        got = st.scores()
        assert_scores_close(got, s64)
        assert_sets_match(st.matches()[0], np.flatnonzero(s64 >= th), s64, (th,))
        assert np.array_equal(st.topk()[0], sc.topk_stable(got, 20))
    th = 0.5
    T32 = T.astype(np.float32)
    gotb = st.scan_batch(T32, w, th, th - 0.1, debug_scores=True)
    counts, rows, scores, _ = st.scan_batch(T32, w, th, th - 0.1, topk=20)
    for qi in range(3):
        sims64, _ = sc.similarities(X64, T32[qi].astype(np.float64))
        s64 = sc.scores(sims64, w)
        # the tensor-core path's error (-7e-7, accumulation truncation) is relative to the SIMILARITY, which in
        # this random data reaches 3 while scores go down to 0: compare against max(|score|, 0.25)
        assert_scores_close(gotb[qi], s64, floor=0.25)
        assert counts[qi, 0] == np.count_nonzero(gotb[qi].astype(np.float64) >= th)
        assert np.array_equal(rows[qi], sc.topk_stable(gotb[qi], 20))
    st.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("VQ_FUZZ_SEEDS", "8"))))
def test_scan_on_randomly_drawn_shapes_matches_oracle(vq, seed):
    """Randomly drawn store shapes — 1-4 streams, 1-3 splits, feature lengths that are any multiple of 4 (so every
    partial-chunk path of the generic scan kernel: a stream shorter than one warp pass, a ragged last chunk), row counts
    from 1 to a few thousand (fewer rows than a block, more than a selection chunk), random missing (clip, stream, split)
    slots, random weights, threshold and top-k — against the float64 oracle: scores 1e-5, sets bit-exact outside the tie
    band, top-k equal to the stable ranking of the device's scores, labelled fp64 similarities to 1e-12."""
    rng = np.random.default_rng(4200 + seed)
    S, P = int(rng.integers(1, 5)), int(rng.integers(1, 4))
    dim = int(rng.choice([4, 12, 100, 132, 516, 1000, 1024, 1540, 2052]))
    n = int(rng.choice([1, 3, 31, 33, 129, 1000, 4097, 6000]))
    streams = tuple("s%d" % i for i in range(S))
    splits = sorted(int(p_) for p_ in rng.choice(np.arange(1, 8), P, replace=False))
    X = (rng.random((n, S, P, dim)) * (0.2 + rng.random((n, 1, 1, 1)) * 2)).astype(np.float32)
    present = rng.random((n, S, P)) > (0.15 if P > 1 and seed % 2 else 0.0)
    present[:, :, 0] = True                                    # never lose every split of a stream
    X = X * present[..., None]
    st = vq.FeatureStore(n, streams, splits, dim, devices=[0])
    st.upload(0, X)
    st.set_present(present)
    X64 = X.astype(np.float64)
    w = [float(v) for v in rng.uniform(0.5, 2.5, S)]
    td = lambda t: {s: {p_: t[si, pi] for pi, p_ in enumerate(splits)} for si, s in enumerate(streams)}
    for r in {0, n // 2, n - 1}:
        T = sc.scale_target(np.where(present[r][..., None], X64[r], 1.0))      # a full target even when the clip lacks a split
        sims64, _ = sc.similarities(X64, T, None if present.all() else present)
        s64 = sc.scores(sims64, w)
        th = float(np.quantile(s64, 0.8)) if n > 4 else float(s64.min())
        lo = th - 0.2 * (1 - th)
        k = int(rng.choice([0, 1, 17, 100, 1024]))
        res = st.scan(td(T), w, th, lo, EPS, topk=k, want_sims=True)
        got = st.scores()
        assert_scores_close(got, s64)
        assert_sets_match(st.matches()[0], np.flatnonzero(s64 >= th), s64, (th,))
        assert_sets_match(st.near_misses()[0], np.flatnonzero((s64 >= lo) & (s64 < th)), s64, (th, lo))
        g64 = got.astype(np.float64)
        assert res.n_match == np.count_nonzero(g64 >= th) and res.n_near == np.count_nonzero((g64 >= lo) & (g64 < th))
        assert np.array_equal(st.topk()[0], sc.topk_stable(got, min(k, n)))
        rows = np.unique(rng.integers(0, n, 40))
        lab = st.labelled_sims(td(T), rows)
        assert np.abs(lab - sims64[rows]).max() <= 1e-12 * max(np.abs(sims64).max(), 1.0)
    st.close()


def test_two_shards_merge_like_one(vq):
    """Clip-range sharding inside one process (both shards on device 0 here)."""
    n = 9001
    X = synth.database(11, n)[:, :, None, :]
    one = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    two = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0, 0])
    assert len(two.shards) == 2
    one.upload(0, X)
    two.upload(0, X)
    T = sc.scale_target(X[5].astype(np.float64))
    a = one.scan(tdict(T), (1.0, 1.5), 0.8, 0.73, EPS, topk=64)
    b = two.scan(tdict(T), (1.0, 1.5), 0.8, 0.73, EPS, topk=64)
    assert (a.n_match, a.n_near, a.n_tie) == (b.n_match, b.n_near, b.n_tie)
    assert np.array_equal(one.scores(), two.scores())
    for f in ("matches", "near_misses", "topk"):
        ra, sa = getattr(one, f)()
        rb, sb = getattr(two, f)()
        assert np.array_equal(ra, rb) and np.array_equal(sa, sb)
    # batched path: the shards are scanned concurrently (one thread each) and the per-query top-k merged on the host
    Tq = np.stack([sc.scale_target(X[r].astype(np.float64)) for r in (5, 4600, 8999)]).astype(np.float32)
    ca, ra, sa, _ = one.scan_batch(Tq, (1.0, 1.5), 0.8, 0.73, topk=25)
    cb, rb, sb, _ = two.scan_batch(Tq, (1.0, 1.5), 0.8, 0.73, topk=25)
    assert np.array_equal(ca, cb) and np.array_equal(ra, rb) and np.array_equal(sa, sb)
    one.close()
    two.close()


def test_append_grows_the_store_like_a_rebuild(vq):
    """Incremental growth (load_db.py adds clips): a store filled by three appends (two of them re-allocating)
    scans, ranks and batch-scans exactly like a store built in one upload; split weights follow."""
    rng = np.random.default_rng(12)
    n0, adds, S, P, dim = 700, (1, 900, 5000), 2, 2, 256
    n = n0 + sum(adds)
    X = (rng.random((n, S, P, dim)) * rng.random((n, 1, 1, 1)) * 2).astype(np.float32)
    present = rng.random((n, S, P)) > 0.15
    present[:, :, 0] = True                                   # every clip keeps at least one split per stream
    X[~present] = 0
    ids = np.arange(n) * 3 + 11
    full = vq.FeatureStore(n, ("a", "b"), [1, 2], dim, devices=[0], clip_ids=ids)
    full.upload(0, X)
    full.set_present(present)
    grown = vq.FeatureStore(n0, ("a", "b"), [1, 2], dim, devices=[0], clip_ids=ids[:n0])
    grown.upload(0, X[:n0])
    grown.set_present(present[:n0])
    at = n0
    for k in adds:
        grown.append(X[at:at + k], clip_ids=ids[at:at + k], present=present[at:at + k])
        at += k
    assert grown.n_rows == n and grown.row_of(ids[-1]) == n - 1
    with pytest.raises(vq.VQError):
        grown.append(X[:1], clip_ids=ids[:1])                  # id already present
    ref = int(np.flatnonzero(present.all(axis=(1, 2)))[0])          # a reference clip with every split
    T = sc.scale_target(X[ref].astype(np.float64))
    td = {s: {p: T[si, pi] for pi, p in enumerate([1, 2])} for si, s in enumerate(("a", "b"))}
    ra = full.scan(td, (1.0, 1.5), 0.6, 0.5, EPS, topk=40)
    rb = grown.scan(td, (1.0, 1.5), 0.6, 0.5, EPS, topk=40)
    assert (ra.n_match, ra.n_near, ra.n_tie) == (rb.n_match, rb.n_near, rb.n_tie)
    assert np.array_equal(full.scores(), grown.scores())
    for f in ("matches", "near_misses", "topk"):
        a, b = getattr(full, f)(), getattr(grown, f)()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(full.download(0, n), grown.download(0, n))
    T32 = T[None].astype(np.float32)
    assert np.array_equal(full.scan_batch(T32, (1.0, 1.5), 0.6, 0.5, debug_scores=True),
                          grown.scan_batch(T32, (1.0, 1.5), 0.6, 0.5, debug_scores=True))
    full.close()
    grown.close()
    # the same through the API-response path: a store built from the first clips + append_feature_rows(all rows)
    def feats(clips):
        return [{"dnn_stream_id": s, "dnn_stream_split": p, "name": "global_pool", "video_clip_id": int(ids[c]),
                 "feature_vector": X[c, si, pi].tolist()}
                for c in clips for si, s in enumerate(("a", "b")) for pi, p in enumerate([1, 2]) if present[c, si, pi]]
    some, more = list(range(40)), list(range(25, 90))
    one = vq.FeatureStore.from_feature_rows(feats(list(range(90))), ("a", "b"), "global_pool", devices=[0])
    two = vq.FeatureStore.from_feature_rows(feats(some), ("a", "b"), "global_pool", devices=[0])
    assert two.append_feature_rows(feats(more), "global_pool") == 50 and two.n_rows == 90
    assert np.array_equal(one.clip_ids, two.clip_ids) and np.array_equal(one.download(0, 90), two.download(0, 90))
    one.scan(td, (1.0, 1.5), 0.6, 0.5, EPS, topk=10)
    two.scan(td, (1.0, 1.5), 0.6, 0.5, EPS, topk=10)
    assert np.array_equal(one.scores(), two.scores())
    one.close()
    two.close()


def test_exchange_kernel_in_step_lagged_and_flush_on_one_rank(vq):
    """The peer-memory exchange kernel with world = 1 (its own inbox is the only peer), so that the in-step / lagged /
    flush logic and the four inbox slots are exercised on a single-GPU box too: in-step merges step i, lagged
    leaves step i-1 in the merged buffer until the flush.  (The multi-rank run is tests/test_gpu_multi.py.)"""
    import ctypes as C
    import torch
    from video_query_algorithms_b200 import _ffi
    from video_query_algorithms_b200.store import make_params
    from video_query_algorithms_b200.sharded import unpack_payload
    lib = _ffi.lib()
    n, k = 20000, 32
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0], first_global_row=1000)
    st.fill_synthetic(99)
    ref = synth.pick_reference_row(99, n)
    T = sc.scale_target(synth.rows(99, [ref]).astype(np.float64)[0][:, None, :])
    target = torch.from_numpy(T.astype(np.float32).reshape(-1)).to("cuda:0")
    h = st.shards[0].handle
    stream = torch.cuda.Stream()
    x = C.c_void_p()
    _ffi.check(lib.vq_exchange_create(C.byref(x), 0, 1, 0), "vq_exchange_create")
    mp = C.c_void_p()
    _ffi.check(lib.vq_exchange_merged(x, C.byref(mp)), "vq_exchange_merged")

    from video_query_algorithms_b200.sharded import _DevArray
    merged_t = torch.as_tensor(_DevArray(mp.value, 4 + 2 * k), device=torch.device("cuda", 0))   # zero-copy view, as RankScan does

    def merged():
        torch.cuda.synchronize()
        return unpack_payload(merged_t.cpu().numpy(), k)

    def expected(th):
        res = st.scan(tdict(T), (1.0, 1.5), th, 0.73, EPS, topk=k)
        return (res.n_match, res.n_near, res.n_tie), st.topk()

    ths = [0.76 + 0.01 * i for i in range(6)]                  # six steps: every inbox slot is reused at least once
    want = [expected(th) for th in ths]
    for lagged in (False, True):
        for i, th in enumerate(ths):
            p = make_params((1.0, 1.5), th, 0.73, EPS, topk=k)
            _ffi.check(lib.vq_scan_enqueue(h, C.c_void_p(target.data_ptr()), C.byref(p), C.c_void_p(stream.cuda_stream)), "enqueue")
            fn = lib.vq_scan_exchange_enqueue_lagged if lagged else lib.vq_scan_exchange_enqueue
            _ffi.check(fn(h, x, C.c_void_p(stream.cuda_stream)), "exchange")
            counts, rows, scores = merged()
            j = i - 1 if lagged else i                           # lagged: the previous step's result
            if j >= 0 and not (lagged and i == 0):
                assert tuple(counts[:3]) == want[j][0] and np.array_equal(rows, want[j][1][0]) and np.array_equal(scores, want[j][1][1])
        _ffi.check(lib.vq_exchange_flush_enqueue(x, C.c_void_p(stream.cuda_stream)), "flush")
        counts, rows, scores = merged()
        assert tuple(counts[:3]) == want[-1][0] and np.array_equal(rows, want[-1][1][0]) and np.array_equal(scores, want[-1][1][1])
    _ffi.check(lib.vq_exchange_destroy(x), "vq_exchange_destroy")
    st.close()


@pytest.mark.parametrize("world,k", [(8, 100), (8, 1000), (3, 7), (16, 64), (20, 5)])
def test_exchange_kernel_between_ranks_emulated_on_one_device(vq, world, k):
    """The multi-block exchange kernel between `world` ranks that all live on device 0 (vq_exchange_connect_local: one
    process, inboxes addressed directly), so that a single-GPU box runs what tests/test_gpu_multi.py runs over NVLink:
    every rank pushes into every inbox, waits for every flag and merges; each rank's merged result must equal one scan
    of the whole search set — counts summed, global top-k under (score desc, global row asc).  The search set has many
    exact ties inside and across ranks (rows drawn with repetition), ranks shorter than k, a one-row and an empty rank;
    k = 1000 at world 8 takes the merge that reads the inbox directly (candidates do not fit in shared memory); world 20
    has more ranks than blocks.  In-step and lagged mode, six steps each (every inbox slot reused), then the flush."""
    import ctypes as C
    import torch
    from video_query_algorithms_b200 import _ffi
    from video_query_algorithms_b200.store import make_params
    from video_query_algorithms_b200.sharded import _DevArray, unpack_payload
    lib = _ffi.lib()
    rng = np.random.default_rng(world * 1000 + k)
    base = synth.rows(7, np.arange(1500))                                     # [1500, 2, 1024] fp32
    sizes = [700, 1, 0, 300, 50, 700, 650, 12] + [int(v) for v in rng.integers(1, 400, 32)]
    sizes = sizes[:world]
    picks = [rng.integers(0, len(base), n_r) for n_r in sizes]
    firsts = np.concatenate([[0], np.cumsum(sizes)])
    total = int(firsts[-1])
    full = vq.FeatureStore(total, STREAMS, [1], 1024, devices=[0])
    full.upload(0, base[np.concatenate(picks)][:, :, None, :])
    stores = []
    for r in range(world):
        st = vq.FeatureStore(sizes[r], STREAMS, [1], 1024, devices=[0], first_global_row=int(firsts[r]))
        if sizes[r]:
            st.upload(0, base[picks[r]][:, :, None, :])
        stores.append(st)
    T = sc.scale_target(base[3].astype(np.float64)[:, None, :])
    target = torch.from_numpy(T.astype(np.float32).reshape(-1)).to("cuda:0")
    streams = [torch.cuda.Stream() for _ in range(world)]
    xs = (C.c_void_p * world)()
    for r in range(world):
        x = C.c_void_p()
        _ffi.check(lib.vq_exchange_create(C.byref(x), 0, world, r), "vq_exchange_create")
        xs[r] = x
    _ffi.check(lib.vq_exchange_connect_local(xs, world), "vq_exchange_connect_local")
    views = []
    for r in range(world):
        mp = C.c_void_p()
        _ffi.check(lib.vq_exchange_merged(xs[r], C.byref(mp)), "vq_exchange_merged")
        views.append(torch.as_tensor(_DevArray(mp.value, 4 + 2 * k), device=torch.device("cuda", 0)))

    def expected(th):
        res = full.scan(tdict(T), (1.0, 1.5), th, 0.6, EPS, topk=k)
        rows, scores = full.topk()
        return (res.n_match, res.n_near, res.n_tie), rows.copy(), scores.copy()

    ths = [0.70 + 0.02 * i for i in range(6)]
    want = [expected(th) for th in ths]
    assert len(np.unique(want[0][2])) < len(want[0][2])                      # the top-k holds exact ties

    def check_all(j, what):
        torch.cuda.synchronize()
        for r in range(world):
            _ffi.check(lib.vq_exchange_check(xs[r]), "vq_exchange_check")
            counts, rows, scores = unpack_payload(views[r].cpu().numpy(), k)
            assert tuple(int(c) for c in counts[:3]) == want[j][0], (what, r, counts)
            assert np.array_equal(rows, want[j][1]) and np.array_equal(scores, want[j][2]), (what, r)

    for lagged in (False, True):
        fn = lib.vq_scan_exchange_enqueue_lagged if lagged else lib.vq_scan_exchange_enqueue
        for i, th in enumerate(ths):
            p = make_params((1.0, 1.5), th, 0.6, EPS, topk=k)
            for r in range(world):                                            # nothing here waits on the host: all ranks' kernels are in flight together
                sp = C.c_void_p(streams[r].cuda_stream)
                _ffi.check(lib.vq_scan_enqueue(stores[r].shards[0].handle, C.c_void_p(target.data_ptr()), C.byref(p), sp), "enqueue")
                _ffi.check(fn(stores[r].shards[0].handle, xs[r], sp), "exchange")
            if not lagged:
                check_all(i, "in-step %d" % i)
            elif i in (2, 5):
                check_all(i - 1, "lagged %d" % i)
        for r in range(world):
            _ffi.check(lib.vq_exchange_flush_enqueue(xs[r], C.c_void_p(streams[r].cuda_stream)), "flush")
        check_all(len(ths) - 1, "flush")
    times = np.empty(256, np.float32)
    parts = np.empty((256, 3), np.float32)
    n_t = C.c_int32()
    _ffi.check(lib.vq_exchange_kernel_times(xs[0], 256, _ffi.ptr(times), _ffi.ptr(parts), C.byref(n_t)), "vq_exchange_kernel_times")
    assert n_t.value >= 12 and (times[:n_t.value] > 0).all() and (parts[:n_t.value] >= 0).all()
    print("exchange kernel, world %d on one device, k %d: median %.1f us" % (world, k, 1e3 * float(np.median(times[:n_t.value]))))
    for r in range(world):
        _ffi.check(lib.vq_exchange_destroy(xs[r]), "vq_exchange_destroy")
    for st in stores:
        st.close()
    full.close()


# ---------------------------------------------------------------------------- batched queries (tcgen05)
@pytest.mark.parametrize("n,nq", [(5003, 70), (300, 3), (20000, 300)])
def test_batched_tensor_core_scan_matches_oracle(vq, n, nq):
    """K3: Q targets in one pass (bf16x2 split on tcgen05, 256 queries per pass; 300 queries = two passes, 3 and 70
    queries = MMAs narrower than 256).  Scores within 1e-5 of float64; per-query counts and top-k equal to the
    oracle's up to rows within COMPUTE_EPS of a boundary / of each other."""
    seed = 77
    X = synth.database(seed, n)
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.upload(0, X[:, :, None, :])
    X64 = X.astype(np.float64)[:, :, None, :]
    rng = np.random.default_rng(3)
    refs = rng.choice(n, nq, replace=nq > n)
    alphas = synth.alpha(np.arange(n), seed)
    refs[0] = int(np.argmin(np.abs(alphas - 0.9)))
    T = np.stack([sc.scale_target(X64[r]) for r in refs])            # [Q, 2, 1, 1024]
    w, th, near = (1.0, 1.5), 0.8, 0.35
    lo = sc.lower_limit(th, near)
    k = min(50, n)
    got = st.scan_batch(T.astype(np.float32), w, th, lo, debug_scores=True)
    counts, rows, scores, ms = st.scan_batch(T.astype(np.float32), w, th, lo, topk=k)
    assert got.shape == (nq, n) and counts.shape == (nq, 2) and rows.shape == (nq, k)
    T32 = T.astype(np.float32).astype(np.float64)
    for q in range(nq):
        sims64, _ = sc.similarities(X64, T32[q])
        s64 = sc.scores(sims64, w)
        assert_scores_close(got[q], s64)
        m64, nm64 = sc.classify(s64, th, near)
        g = got[q].astype(np.float64)
        assert counts[q, 0] == np.count_nonzero(g >= th) and counts[q, 1] == np.count_nonzero((g >= lo) & (g < th))
        assert_sets_match(np.flatnonzero(g >= th), m64, s64, (th,))
        assert np.array_equal(rows[q], sc.topk_stable(got[q], k))
        assert np.array_equal(scores[q], got[q][rows[q]])
        for a, b in zip(rows[q], sc.topk_stable(s64, k)):
            assert a == b or abs(s64[a] - s64[b]) < EPS
    # the tie band per query (rows within eps of the threshold or of the near-miss limit, compared in double like K2):
    # a wide band so that it is populated; same counts / top-k as the run without it
    wide = 2e-3
    c2, r2, s2, _ = st.scan_batch(T.astype(np.float32), w, th, lo, topk=k, eps=wide)
    assert np.array_equal(c2, counts) and np.array_equal(r2, rows) and np.array_equal(s2, scores)
    t_counts, t_lists = st.batch_ties
    n_band = 0
    for q in range(nq):
        g = got[q].astype(np.float64)
        want = np.flatnonzero((np.abs(g - th) < wide) | (np.abs(g - lo) < wide))
        assert t_counts[q] == len(want) and np.array_equal(t_lists[q][0], want) and np.array_equal(t_lists[q][1], got[q][want])
        n_band += len(want)
    assert n_band > 0
    st.scan_batch(T.astype(np.float32), w, th, lo, topk=k)
    assert st.batch_ties is None
    st.close()


def test_batched_scan_counts_only_and_ragged_tail(vq):
    """K3 without a top-k (no candidate lists at all) on a shard whose size is not a multiple of the 128-clip
    tile, spanning the short seeding launch and a full one: counts equal the single-query scan's counts."""
    n, seed = 148 * 128 + 4 * 128 + 77, 5
    X = synth.database(seed, n)
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.upload(0, X[:, :, None, :])
    X64 = X.astype(np.float64)[:, :, None, :]
    refs = [3, n // 2, n - 1]
    T = np.stack([sc.scale_target(X64[r]) for r in refs])
    w, th, lo = (1.0, 1.5), 0.8, sc.lower_limit(0.8, 0.35)
    counts, rows, scores, _ = st.scan_batch(T.astype(np.float32), w, th, lo, topk=0)
    assert rows.shape == (3, 0) and scores.shape == (3, 0)
    got = st.scan_batch(T.astype(np.float32), w, th, lo, debug_scores=True)
    for q in range(3):
        g = got[q].astype(np.float64)
        assert counts[q, 0] == np.count_nonzero(g >= th) and counts[q, 1] == np.count_nonzero((g >= lo) & (g < th))
        sims64, _ = sc.similarities(X64, T[q].astype(np.float32).astype(np.float64))
        assert_scores_close(got[q], sc.scores(sims64, w))
    st.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("VQ_FUZZ_SEEDS", "6"))))
def test_batched_scan_on_randomly_drawn_shapes_matches_oracle(vq, seed):
    """K3 on randomly drawn shapes: 1-4 streams, 1-2 splits, stream lengths 32..2048 (multiples of the 32-wide K block),
    1..300 queries (narrow MMAs, exactly 256, more than one pass), shard sizes around the 128-clip tile and the launch
    boundaries, any top-k, missing slots — scores against float64 (1e-5 of max(|score|, |similarity|, 0.25): its error is
    relative to the similarity), per-query counts equal to a double comparison of its own scores, top-k equal to the stable ranking
    of its own scores, tie band equal to the rows of its own scores within eps."""
    rng = np.random.default_rng(9100 + seed)
    S, P = int(rng.integers(1, 5)), int(rng.integers(1, 3))
    dim = int(rng.choice([32, 64, 256, 512, 1024]))
    n = int(rng.choice([1, 100, 127, 128, 129, 5000, 148 * 128 + 5, 40000]))
    Q = int(rng.choice([1, 5, 63, 64, 65, 128, 129, 255, 256, 257, 300]))
    if n * Q > 6_000_000:
        Q = max(1, 6_000_000 // n)
    streams = tuple("s%d" % i for i in range(S))
    splits = list(range(1, P + 1))
    X = (rng.random((n, S, P, dim), dtype=np.float32) * (0.5 + rng.random((n, 1, 1, 1), dtype=np.float32))).astype(np.float32)
    present = rng.random((n, S, P)) > (0.2 if P > 1 and seed % 2 else 0.0)
    present[:, :, 0] = True
    X = X * present[..., None]
    st = vq.FeatureStore(n, streams, splits, dim, devices=[0])
    st.upload(0, X)
    st.set_present(present)
    X64 = X.astype(np.float64)
    w = [float(v) for v in rng.uniform(0.5, 2.5, S)]
    refs = rng.integers(0, n, Q)
    T32 = np.stack([sc.scale_target(np.where(present[r][..., None], X64[r], 1.0)) for r in refs]).astype(np.float32)
    th = 0.6
    lo = th - 0.25
    k = int(rng.choice([0, 1, 20, 100]))
    wide = 1e-3
    got = st.scan_batch(T32, w, th, lo, debug_scores=True)
    counts, rows, scores, _ = st.scan_batch(T32, w, th, lo, topk=k, eps=wide)
    t_counts, t_lists = st.batch_ties
    assert got.shape == (Q, n) and rows.shape == (Q, k)
    kk = min(k, n)
    for q in range(Q):
        if q < 12 or q % 37 == 0:                                  # float64 check on a subset of the queries (host time)
            sims64, _ = sc.similarities(X64, T32[q].astype(np.float64), None if present.all() else present)
            s64 = sc.scores(sims64, w)
            # the operand split's error is relative to the SIMILARITY (3 * 2^-18 per product, random signs), which in this
            # data reaches 2-3 where the score passes through 0: measured against max(|score|, max |sim|, 0.25)
            scale = np.maximum(np.maximum(np.abs(s64), np.abs(sims64).max(axis=1)), 0.25)
            err = np.abs(got[q].astype(np.float64) - s64) / scale
            assert err.max() <= 1e-5, "query %d: max err %.3e at row %d" % (q, err.max(), err.argmax())
        g = got[q].astype(np.float64)
        assert counts[q, 0] == np.count_nonzero(g >= th) and counts[q, 1] == np.count_nonzero((g >= lo) & (g < th))
        assert np.array_equal(rows[q, :kk], sc.topk_stable(got[q], kk)) and np.array_equal(scores[q, :kk], got[q][rows[q, :kk]])
        assert (rows[q, kk:] == -1).all()
        band = np.flatnonzero((np.abs(g - th) < wide) | (np.abs(g - lo) < wide))
        assert t_counts[q] == len(band)
        if len(band) <= st.TIE_CAP:
            assert np.array_equal(t_lists[q][0], band) and np.array_equal(t_lists[q][1], got[q][band])
    st.close()


# ---------------------------------------------------------------------------- labelled subset (fp64)
def test_loss_grid_and_replicates_match_oracle(vq):
    rng = np.random.default_rng(1)
    L = 300
    sims = 0.6 + 0.5 * rng.random((L, 2))
    y = rng.random(L) < 0.4
    wg, tg = sc.weight_grid(), sc.threshold_grid()
    for ballast in (0.0, 0.3):
        got = vq.loss_grid(sims, y, wg, tg, ballast)[0]
        want = sc.loss_grid(sims, y, wg, tg, ballast)
        assert np.abs(got - want).max() < 1e-13
    random.seed("73459912436")
    reps = vq.resample_labelled(L, 25, random)
    got = vq.loss_grid(sims, y, wg, tg, 0.1, replicates=reps)
    for r, idx in enumerate(reps):
        want = sc.loss_grid_fast(sims[idx], y[idx], wg, tg, 0.1)
        assert np.abs(got[r] - want).max() < 1e-13


def test_bootstrap_target_matches_oracle(vq):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "fixture_brooklyn.npz"))
    X = z["X"]                                                # [87, 2, 3, 1024] float64
    st = vq.FeatureStore(X.shape[0], STREAMS, [1, 2, 3], 1024, devices=[0])
    st.upload(0, X)
    X32 = X.astype(np.float32).astype(np.float64)
    valid, invalid = np.array([3, 9, 10, 40, 41, 70]), np.array([0, 20, 55, 56])
    for mu, inv in ((0.0, invalid[:0]), (0.0, invalid), (0.3, invalid)):
        got = st.bootstrap_target(valid, inv, mu)
        for s in range(2):
            for p in range(3):
                want = (ob.solve_valid_invalid(X32[valid, s, p], X32[inv, s, p], mu) if len(inv)
                        else ob.solve_valid(X32[valid, s, p]))
                assert np.abs(got[s, p] - want).max() <= 1e-8 * np.abs(want).max()
    st.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("VQ_FUZZ_SEEDS", "6"))))
def test_labelled_path_on_randomly_drawn_problems_matches_oracle(vq, seed):
    """The fp64 kernels of the weight update and the target bootstrap on randomly drawn problems: K5's loss grid for any
    number of labelled clips (1 .. 3000), label mixes incl. all-True / all-False, grid sizes other than 40 x 31, ballast,
    scores sitting exactly on a threshold (H(0) = 1), replicate index sets with repeats and of different lengths (1e-13
    against the oracle's loop); K6's solve for 1-4 streams, 1-3 splits, feature lengths 64-1024, 1 .. 60 valid and
    0 .. 30 invalid rows, mu = 0 and > 0, duplicate-free rows drawn from well-conditioned data, slot masks (1e-8 of the
    oracle's numpy solve, hyperparameter.py:56-65, target_clip.py:194-198,248-260)."""
    rng = np.random.default_rng(5300 + seed)
    # ---- K5
    L = int(rng.choice([1, 2, 7, 100, 1000, 3000]))
    sims = 0.4 + 0.8 * rng.random((L, 2))
    mode = seed % 4
    y = np.ones(L, bool) if mode == 1 else np.zeros(L, bool) if mode == 2 else rng.random(L) < rng.uniform(0.1, 0.9)
    wg = sc.weight_grid() if seed % 2 else np.linspace(0.3, 3.0, int(rng.integers(1, 60)))
    tg = sc.threshold_grid() if seed % 3 else np.linspace(0.2, 1.2, int(rng.integers(1, 50)))
    if L > 2:                                                      # a clip whose score equals a grid threshold exactly
        sims[0] = (1.0 - (1.0 - tg[len(tg) // 2]), 1.0 - (1.0 - tg[len(tg) // 2]))
    ballast = float(rng.choice([0.0, 0.1, 0.75]))
    got = vq.loss_grid(sims, y, wg, tg, ballast)[0]
    want = sc.loss_grid(sims, y, wg, tg, ballast) if L <= 1000 else sc.loss_grid_fast(sims, y, wg, tg, ballast)
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-13 * max(1.0, np.abs(want).max())
    reps = [rng.integers(0, L, int(rng.integers(1, L + 2))) for _ in range(int(rng.integers(1, 12)))]
    got = vq.loss_grid(sims, y, wg, tg, ballast, replicates=reps)
    for r, idx in enumerate(reps):
        want = sc.loss_grid_fast(sims[idx], y[idx], wg, tg, ballast)
        assert np.abs(got[r] - want).max() < 1e-13 * max(1.0, np.abs(want).max())
    # ---- K6
    S, P = int(rng.integers(1, 5)), int(rng.integers(1, 4))
    dim = int(rng.choice([64, 256, 1024]))
    n = 400
    X = (rng.random((n, S, P, dim)) + 0.1).astype(np.float32)
    st = vq.FeatureStore(n, tuple("s%d" % i for i in range(S)), list(range(1, P + 1)), dim, devices=[0])
    st.upload(0, X)
    X64 = X.astype(np.float64)
    for _ in range(3):
        nv, ni = int(rng.integers(1, min(60, dim // 2))), int(rng.integers(0, 30))
        rows = rng.choice(n, nv + ni, replace=False)
        valid, invalid = rows[:nv], rows[nv:]
        mu = float(rng.choice([0.0, 0.25, 1.0]))
        slots = None if seed % 2 else rng.random((S, P)) < 0.6
        got = st.bootstrap_target(valid, invalid if ni else None, mu, slots)
        for s_ in range(S):
            for p_ in range(P):
                if slots is not None and not slots[s_, p_]:
                    assert not got[s_, p_].any()
                    continue
                want = (ob.solve_valid_invalid(X64[valid, s_, p_], X64[invalid, s_, p_], mu) if ni
                        else ob.solve_valid(X64[valid, s_, p_]))
                assert np.abs(got[s_, p_] - want).max() <= 1e-8 * np.abs(want).max(), (S, P, dim, nv, ni, mu)
    st.close()


# ---------------------------------------------------------------------------- end to end vs reference goldens
@pytest.mark.parametrize("name", SCENARIOS)
def test_compute_matches_reproduces_reference_rounds(vq, name, tmp_path, monkeypatch):
    from fake_api import FakeRepository
    scn = Scenario(name)
    api, qid = scn.build_api()
    (tmp_path / "work").mkdir()
    monkeypatch.chdir(tmp_path / "work")                     # final report goes to ../final_reports/
    vq.invalidate()
    rule = scn.label_rule()
    tickets = []

    def factory(job, url):
        t = vq.Ticket(job, url, client=api.client(), devices=[0])
        t.topk = 10
        tickets.append(t)
        return t

    for i, r in enumerate(scn.rounds):
        if i > 0:
            api.label_latest_round(qid, rule)
        api.request(qid, r["kind"])
        hp = vq.Hyperparameter(**scn.hp())
        random.seed(a=scn.seed)                              # broker.py:83-84
        vq.compute_matches(FakeRepository(api), hp, ticket_factory=factory)
        t = tickets[-1]
        assert api.queries[qid]["process_state"] == r["process_state"]
        assert [hp.weights[s] for s in scn.streams] == pytest.approx(r["weights"], rel=1e-5)
        assert hp.threshold == pytest.approx(r["threshold"], rel=1e-5)
        want_scores = scn.arr(i, "scores")
        assert np.array_equal(t.feature_store().clip_ids, scn.arr(i, "clip_order"))
        assert_scores_close(t.scores.array(), want_scores)
        T = np.array([[t.target.target_features[s][p] for p in r["splits"]] for s in scn.streams])
        assert np.abs(T - scn.arr(i, "target")).max() <= 1e-6 * np.abs(scn.arr(i, "target")).max()
        if "losses" in scn.arrays.files and "r%d_losses" % i in scn.arrays.files:
            assert np.abs(hp.losses - scn.arr(i, "losses")).max() < 1e-6
        # selected clips: same ids in the same order unless a boundary tie was reported
        got_ids = list(t.matches)
        want_ids = [k for k, _ in r["selected"]]
        if not t.tie_band:
            assert got_ids == want_ids
            got_sc = np.array([t.matches[k] for k in got_ids])
            assert_scores_close(got_sc, np.array([v for _, v in r["selected"]]))
        assert len(t.ranked[0]) == min(10, len(want_scores))
        assert list(t.ranked[0]) == [int(scn.clip_ids[j]) for j in sc.topk_stable(t.scores.array(), len(t.ranked[0]))]
    # the store is built once per search set and reused by later ticks (the reference re-downloads every job)
    assert api.calls.count(("search-sets", "features")) == 1
    if scn.rounds[-1]["kind"] == "finalize":
        assert len(api.uploaded_reports) == 1
        body = api.uploaded_reports[0].splitlines()
        ref = scn.meta["final_report"].splitlines()
        assert len(body) == len(ref)
        got_clips = [ln.split(",")[4] for ln in body[-len(scn.rounds[-1]["selected"]):]]
        ref_clips = [ln.split(",")[4] for ln in ref[-len(scn.rounds[-1]["selected"]):]]
        # ranking is by score, stable and descending (ticket.py:266); clips whose float64 scores are
        # within COMPUTE_EPS of each other (e.g. several bootstrapped matches at score ~1) may swap
        ref_score = {str(k): v for k, v in scn.rounds[-1]["selected"]}
        assert sorted(got_clips) == sorted(ref_clips) or tickets[-1].tie_band
        for a, b in zip(got_clips, ref_clips):
            if a in ref_score and b in ref_score:
                assert a == b or abs(ref_score[a] - ref_score[b]) < EPS
    vq.invalidate()


def test_job_error_states_follow_the_reference(vq, tmp_path, monkeypatch):
    """compute_matches.py:47-52,92-94: fatal query errors -> state 5 + note; recoverable -> note and go on;
    an empty selection -> state 5 'No matches were found'."""
    from fake_api import FakeRepository
    scn = Scenario("A_brooklyn_bagging")
    monkeypatch.chdir(tmp_path)
    vq.invalidate()
    factory = lambda api: (lambda job, url: vq.Ticket(job, url, client=api.client(), devices=[0]))
    # (1) reference time outside the video: no ref clip
    api, qid = scn.build_api()
    api.queries[qid]["ref_clip_id"] = None
    api.request(qid, "new")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=factory(api))
    assert api.queries[qid]["process_state"] == 5 and "Fatal Error" in api.queries[qid]["notes"]
    # (2) revise with dynamic target adjustment but no confirmed match: note, fall back to the scaled ref clip
    vq.invalidate()
    api, qid = scn.build_api()
    api.request(qid, "new")
    random.seed(a=scn.seed)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=factory(api))
    api.label_latest_round(qid, lambda m: False)
    api.request(qid, "revise")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=factory(api))
    assert api.queries[qid]["process_state"] == 4
    assert "Changing dynamic target adjustment to False" in api.queries[qid]["notes"]
    # (3) nothing selectable: ref clip outside the search set and an unreachable threshold
    vq.invalidate()
    api, qid = scn.build_api()
    ss = api.queries[qid]["search_set_to_query"]
    api.search_sets[ss]["clip_ids"] = [c for c in api.search_sets[ss]["clip_ids"] if c != api.queries[qid]["ref_clip_id"]]
    api.request(qid, "new")
    hp = dict(scn.hp(), default_threshold=5.0, near_miss_default=0.0)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**hp), ticket_factory=factory(api))
    assert api.queries[qid]["process_state"] == 5 and "No matches were found" in api.queries[qid]["notes"]
    vq.invalidate()


def test_missing_library_or_bad_arguments_fail_loudly(vq):
    from video_query_algorithms_b200 import _ffi
    with pytest.raises(vq.VQError):
        vq.FeatureStore(10, STREAMS, [1], 1022, devices=[0])            # dim not a multiple of 4
    st = vq.FeatureStore(10, STREAMS, [1], 1024, devices=[0])
    with pytest.raises(vq.VQError):
        st.scan(tdict(np.ones((2, 1, 1024))), (0.0, 0.0), 0.8, 0.7, EPS)   # all-zero weights
    with pytest.raises(vq.VQError):
        st.labelled_sims(tdict(np.ones((2, 1, 1024))), np.array([10]))    # row outside the store
    with pytest.raises(vq.VQError):
        st.bootstrap_target(np.array([3, 99]), None, 0.0)                 # row outside the store
    st.close()


def test_calls_from_a_new_thread_every_tick(vq):
    """The broker runs every tick on a fresh thread (reference src/broker.py:90-92, threading.Timer): a store built
    on one thread must serve scans, list fetches, the fp64 labelled path and batched queries from any other."""
    import threading
    n, seed = 30_000, 11
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.fill_synthetic(seed)
    T = sc.scale_target(synth.rows(seed, [123]).astype(np.float64)[0][:, None, :])
    lo = sc.lower_limit(0.8, 0.35)
    lab = np.arange(5, n, 701, dtype=np.int64)

    def tick(out):
        try:
            r = st.scan(tdict(T), (1.0, 1.5), 0.8, lo, EPS, topk=50)
            out["scan"] = (r.n_match, r.n_near, r.n_tie, st.matches()[0].copy(), st.near_misses()[0].copy(), st.topk()[0].copy(),
                           st.scores().copy())
            out["lab"] = st.labelled_sims(tdict(T), lab)
            c, rows, scs, _ = st.scan_batch(np.stack([T, T]).astype(np.float32), (1.0, 1.5), 0.8, lo, topk=5)
            out["batch"] = (c.copy(), rows.copy(), scs.copy())
            out["grid"] = vq.loss_grid(out["lab"], (out["lab"][:, 0] > 0.9).astype(np.uint8), np.arange(0.5, 2.5, 0.05),
                                       np.arange(0.5, 1.1, 0.02), 0.1)
        except Exception as e:                     # surfaced on the main thread
            out["error"] = e

    results = []
    for _ in range(3):                              # three ticks, three threads, one store
        out = {}
        th = threading.Thread(target=tick, args=(out,))
        th.start()
        th.join()
        assert "error" not in out, out.get("error")
        results.append(out)
    main = {}
    tick(main)
    assert "error" not in main, main.get("error")
    for out in results:
        assert out["scan"][:3] == main["scan"][:3]
        for a, b in zip(out["scan"][3:], main["scan"][3:]):
            assert np.array_equal(a, b)
        assert np.array_equal(out["lab"], main["lab"]) and np.array_equal(out["grid"], main["grid"])
        for a, b in zip(out["batch"], main["batch"]):
            assert np.array_equal(a, b)
    assert main["scan"][0] > 0 and main["batch"][0][0, 0] == main["batch"][0][1, 0]
    st.close()


# ---------------------------------------------------------------------------- full size (BASELINE configs 2 and 3)
@pytest.mark.parametrize("n, first", [(1_000_000, 0), (12_500_000, 87_500_000)],
                         ids=["config2-1M", "config3-last-shard-of-100M"])
def test_full_size_scan_properties(vq, n, first):
    """1M clips (8.19 GB, config 2) and the last 12.5M-clip shard of the 100M-clip DB (102.4 GB, global rows
    87.5M..100M, config 3) generated on the device: determinism, list invariants, and a float64 CPU check of
    every top-k / tie row plus a seeded sample of the matches, near misses and the rest (rows regenerated on
    the CPU from the counter-based generator)."""
    import torch
    if n * 8192 * 1.05 > torch.cuda.get_device_properties(0).total_memory:
        pytest.skip("device memory too small for %d clips" % n)
    seed = synth.DEFAULT_SEED
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0], first_global_row=first)
    st.fill_synthetic(seed)
    ref = 18120                                   # alpha ~ 0.9: ~9 % of the clips match
    T = sc.scale_target(synth.rows(seed, [ref]).astype(np.float64)[0][:, None, :])
    th, lo, w = 0.8, sc.lower_limit(0.8, 0.35), (1.0, 1.5)
    r1 = st.scan(tdict(T), w, th, lo, EPS, topk=100)
    s1 = st.scores()
    m1, n1, k1 = st.matches(), st.near_misses(), st.topk()
    r2 = st.scan(tdict(T), w, th, lo, EPS, topk=100)
    assert np.array_equal(s1, st.scores()) and (r1.n_match, r1.n_near) == (r2.n_match, r2.n_near)
    assert np.array_equal(m1[0], st.matches()[0]) and np.array_equal(k1[0], st.topk()[0])
    g = s1.astype(np.float64)
    assert np.array_equal(m1[0], first + np.flatnonzero(g >= th))
    assert np.array_equal(n1[0], first + np.flatnonzero((g >= lo) & (g < th)))
    assert np.array_equal(m1[1], s1[m1[0] - first]) and np.array_equal(n1[1], s1[n1[0] - first])
    top = sc.topk_stable(s1, 100)
    assert np.array_equal(k1[0], first + top) and np.array_equal(k1[1], s1[top])
    assert r1.n_match == len(m1[0]) > n // 50 and r1.n_near == len(n1[0]) > n // 100
    if first <= ref < first + n:
        assert s1[ref - first] == pytest.approx(1.0, abs=1e-6) and k1[0][0] == ref
    rng = np.random.default_rng(0)
    t_rows = st.ties()[0]
    check = np.unique(np.concatenate([k1[0], t_rows, rng.choice(m1[0], 300), rng.choice(n1[0], 300),
                                      first + rng.integers(0, n, 1500), [first, first + n - 1]]))
    X = synth.rows(seed, check).astype(np.float64)[:, :, None, :]
    sims64, _ = sc.similarities(X, T)
    s64 = sc.scores(sims64, w)
    assert_scores_close(s1[check - first], s64)
    in_m = np.isin(check, m1[0])
    in_n = np.isin(check, n1[0])
    for j, r in enumerate(check):
        ok_m = (s64[j] >= th) == in_m[j]
        ok_n = ((s64[j] >= lo) and (s64[j] < th)) == in_n[j]
        if not (ok_m and ok_n):
            assert r in t_rows, "row %d flipped outside the reported tie band" % r
    st.close()


# ---------------------------------------------------------------------------- round-2 additions
def test_partial_update_target_on_the_gpu_matches_reference(vq):
    """Scenario E: bootstrap_type 'partial_update' (target_clip.py:75-82: blend of the new target with the stored one)
    through the product's TargetClip with the real store behind it — the solve runs in vq_bootstrap_target (K6).
    Recorded by calling the reference's TargetClip directly (through compute_matches the reference crashes on it)."""
    import json
    import types
    from scenarios import GOLDEN, rng_digest
    with open(os.path.join(GOLDEN, "scn_E_partial_update.json")) as f:
        js = json.load(f)
    z = np.load(os.path.join(GOLDEN, "scn_E_partial_update.npz"))
    fx = np.load(os.path.join(GOLDEN, "fixture_shrp2.npz"))
    X, splits = fx["X"], [int(p) for p in fx["splits"]]
    ids = np.arange(js["first_clip_id"], js["first_clip_id"] + len(X))
    st = vq.FeatureStore(len(X), STREAMS, splits, 1024, devices=[0], clip_ids=ids)
    st.upload(0, X)
    prev = {s: {p: z["previous"][si, pi].tolist() for pi, p in enumerate(splits)} for si, s in enumerate(STREAMS)}

    class Client:
        def action(self, schema, keys, params=None, **kw):
            assert tuple(keys) == ("matches", "list")
            return {"results": js["matches"], "pagination": {"nextPage": None}}

    ticket = types.SimpleNamespace(client=Client(), schema=None, dynamic_target_adjustment=True, ref_clip_id=int(ids[0]),
                                   latest_query_result={"id": 1, "round": 1, "bootstrapped_target": prev},
                                   features_from_store=True, attach_store=lambda hp: st, feature_store=lambda optional=False: st)
    hp = dict(js["hp"])
    hp["streams"] = tuple(hp["streams"])
    random.seed(a=js["seed"])
    tc = vq.TargetClip(ticket, vq.Hyperparameter(**hp))
    tc.get_target_features()
    got = np.array([[tc.target_features[s][p] for p in splits] for s in STREAMS])
    assert np.abs(got - z["target"]).max() <= 1e-6 * np.abs(z["target"]).max()
    assert rng_digest() == js["rng_after"]
    assert json.dumps(tc.target_features)
    st.close()


def _ragged_api(vq, rng, n_clips=400, splits=(1, 2, 3), dim=256, ref_split_order=None):
    """A fake API whose search set is ragged: every clip has split 1 of both streams... except some that lack it in the
    first stream (they must move behind the others, ticket.py:146-160), and many lack other slots."""
    from fake_api import FakeAPI
    api = FakeAPI(page_size=13)
    base = rng.random((2, len(splits), dim))
    X = np.zeros((n_clips, 2, len(splits), dim))
    alpha = rng.random(n_clips) ** 0.5
    for c in range(n_clips):
        X[c] = alpha[c] * base + (1 - alpha[c]) * rng.random((2, len(splits), dim)) * 0.7
    present = rng.random((n_clips, 2, len(splits))) > 0.25
    present[:, 1, 0] = True                                    # every clip keeps one split of the second stream
    present[:, 0, 1] |= ~present[:, 0, 0]                      # ... and at least one of the first
    present[7] = True                                          # the reference clip is complete
    vid = api.add_video("ragged")
    cids = [api.add_clip(vid, c) for c in range(n_clips)]
    order = rng.permutation(n_clips * 2 * len(splits))         # records in arbitrary order
    for o in order:
        c, rest = divmod(int(o), 2 * len(splits))
        si, pi = divmod(rest, len(splits))
        if present[c, si, pi]:
            api.add_feature(cids[c], STREAMS[si], splits[pi], X[c, si, pi].tolist())
    # the order the API lists the reference clip's own records in (= the order a new round's target walks the splits in)
    rank = {p: i for i, p in enumerate(ref_split_order or splits)}
    api.features_by_clip[cids[7]].sort(key=lambda r: (STREAMS.index(r["dnn_stream_id"]), rank[r["dnn_stream_split"]]))
    ss = api.add_search_set("ragged", cids)
    qid = api.add_query("qr", vid, cids[7], ss, max_matches=16, dynamic_target_adjustment=True)
    return api, qid, cids, X, present


@pytest.mark.parametrize("ref_split_order", [None, (2, 3, 1)], ids=["ascending", "target-walks-2-3-1"])
def test_compute_matches_on_a_ragged_search_set_follows_the_oracle(vq, tmp_path, monkeypatch, ref_split_order):
    """A ragged search set (missing (clip, stream, split) slots, clips lacking the first split of the first stream, records
    in arbitrary order) through compute_matches on the GPU — new, revise and finalize rounds — against the float64 oracle
    on the same records: row order (the reference's dict order, ticket.py:146-160), per-clip split means, scores, weights,
    the selected clips in order.  Second case: the API lists the reference clip's records as split 2, 3, 1, so the new
    round's target walks the splits in that order and the reference's `scores` dict — which its seeded sampling draws
    from — is NOT in the store's row order (Ticket._place)."""
    from fake_api import FakeRepository
    rng = np.random.default_rng(77)
    api, qid, cids, X, present = _ragged_api(vq, rng, ref_split_order=ref_split_order)
    (tmp_path / "work").mkdir()
    monkeypatch.chdir(tmp_path / "work")
    vq.invalidate()
    tickets = []

    def factory(job, url):
        t = vq.Ticket(job, url, client=api.client(), devices=[0])
        tickets.append(t)
        return t

    hp_kw = dict(default_weights={"rgb": 1.0, "warped_optical_flow": 1.5}, default_threshold=0.7, near_miss_default=0.4,
                 bootstrap_type="simple", mu=0.2, ballast=0.1, f_bootstrap=0.7)
    for i, kind in enumerate(("new", "revise", "finalize")):
        if i > 0:
            api.label_latest_round(qid, lambda m: bool(m["score"] >= 0.78))
        api.request(qid, kind)
        hp = vq.Hyperparameter(**hp_kw)
        random.seed(a="ragged-%d" % i)
        state = random.getstate()
        vq.compute_matches(FakeRepository(api), hp, ticket_factory=factory)
        t = tickets[-1]
        st = t.feature_store()
        assert api.queries[qid]["process_state"] in (4, 7), api.queries[qid].get("notes")
        # the oracle on the same records: rows in the product's order must be the reference's dict order
        rows = api.action(("search-sets", "features"), {"id": api.queries[qid]["search_set_to_query"]})
        first_seen, lowest = {}, {}
        for j, r_ in enumerate(rows):
            if r_["dnn_stream_id"] == STREAMS[0]:
                key = (int(r_["dnn_stream_split"]), j)
                if key < lowest.get(r_["video_clip_id"], (99, 0)):
                    lowest[r_["video_clip_id"]] = key
        want_order = sorted(lowest, key=lowest.get)
        assert st.clip_ids.tolist() == want_order
        row_of = {c: k for k, c in enumerate(cids)}
        perm = np.array([row_of[c] for c in want_order])
        Xo, Po = X[perm].astype(np.float32).astype(np.float64), present[perm]
        Xo = Xo * Po[..., None]
        T = np.array([[t.target.target_features[s].get(p, np.zeros(X.shape[3])) for p in st.splits] for s in STREAMS], np.float64)
        have = np.array([[p in t.target.target_features[s] for p in st.splits] for s in STREAMS])
        sims64, _ = sc.similarities(Xo, T, Po & have[None])
        w = [hp.weights[s] for s in STREAMS]
        s64 = sc.scores(sims64, w)
        assert_scores_close(t.scores.array(), s64)
        if kind == "new":
            random.setstate(state)
            want_T = sc.scale_target(X[7])
            assert np.abs(T - want_T).max() <= 1e-6 * np.abs(want_T).max()
            th, near = hp.threshold, hp.near_miss_default
            m64, nm64 = sc.classify(s64, th, near)
            # the reference's dict for THIS target: walk the target's splits in the target's order, first stream first
            walk = list(t.target.target_features[STREAMS[0]])
            assert walk == list(ref_split_order or (1, 2, 3))
            dict_order, seen = [], set()
            for s_ in STREAMS:
                for p in t.target.target_features[s_]:
                    for r_ in rows:
                        c = r_["video_clip_id"]
                        if r_["dnn_stream_id"] == s_ and r_["dnn_stream_split"] == p and c not in seen:
                            seen.add(c)
                            dict_order.append(c)
            assert list(t.scores) == dict_order and (ref_split_order is None) == (t._place is None)
            if not t.tie_band:
                in_m = set(int(want_order[j]) for j in m64)
                m_dict = [c for c in dict_order if c in in_m]
                picked = random.sample(range(len(m_dict)), int(min(16 / 2, len(m_dict))))
                assert list(t.matches)[:len(picked)] == [m_dict[j] for j in picked]
    assert api.calls.count(("search-sets", "features")) == 1 + 3      # built once; the test itself read it three times
    vq.invalidate()


def test_partial_target_then_full_target_on_the_shared_store(vq):
    """The resident store is shared by every job of a search set: the per-row split-weight table a job with a PARTIAL
    target uploads (the reference averages over splits that both the target and the clip have, ticket.py:146-160) must
    not leak into the next job's scan, its labelled similarities or a batched scan."""
    rng = np.random.default_rng(3)
    n, S, P, dim = 3000, 2, 3, 256
    X = (rng.random((n, S, P, dim)) * rng.random((n, 1, 1, 1))).astype(np.float32)
    st = vq.FeatureStore(n, STREAMS, [1, 2, 3], dim, devices=[0])
    st.upload(0, X)
    X64 = X.astype(np.float64)
    T = sc.scale_target(X64[5])
    full = {s: {p: T[si, pi] for pi, p in enumerate([1, 2, 3])} for si, s in enumerate(STREAMS)}
    part = {s: {p: T[si, pi] for pi, p in enumerate([1, 2, 3]) if not (si == 1 and p == 3)} for si, s in enumerate(STREAMS)}
    have = np.ones((S, P), bool)
    have[1, 2] = False
    ones = np.ones((n, S, P), bool)
    w = (1.0, 1.5)
    want_full = sc.scores(sc.similarities(X64, T, ones)[0], w)
    want_part = sc.scores(sc.similarities(X64, T * have[..., None], ones & have[None])[0], w)
    rows = np.arange(0, n, 97, dtype=np.int64)
    for target, want in ((full, want_full), (part, want_part), (full, want_full), (part, want_part), (part, want_part),
                         (full, want_full)):
        st.scan(target, w, 0.7, 0.6, EPS, topk=10)
        assert_scores_close(st.scores(), want)
        T_here = T * have[..., None] if target is part else T
        l64 = sc.similarities(X64[rows], T_here, (ones & have[None])[rows] if target is part else ones[rows])[0]
        assert np.abs(st.labelled_sims(target, rows) - l64).max() <= 1e-9 * np.abs(l64).max()
    st.scan(part, w, 0.7, 0.6, EPS)
    got = st.scan_batch([full, full], w, 0.7, 0.6, debug_scores=True)       # a batch after a partial single-query job
    assert_scores_close(got[0], want_full, floor=0.25)
    with pytest.raises(vq.VQError):
        st.scan_batch([full, part], w, 0.7, 0.6, topk=3)                  # one table per pass: mixed slot sets are refused
    # an append between two jobs with the same partial target: the table is applied again to the grown shard
    st.scan(part, w, 0.7, 0.6, EPS)
    extra = (rng.random((50, S, P, dim)) * 0.5).astype(np.float32)
    st.append(extra)
    st.scan(part, w, 0.7, 0.6, EPS)
    Xg = np.concatenate([X64, extra.astype(np.float64)])
    want = sc.scores(sc.similarities(Xg, T * have[..., None], np.ones((n + 50, S, P), bool) & have[None])[0], w)
    assert_scores_close(st.scores(), want)
    st.close()


def test_singular_labelled_set_is_an_error_not_a_nan_target(vq):
    """A labelled clip listed twice makes the Gram matrix singular: numpy.linalg.inv raises in the reference
    (target_clip.py:194); here the solve kernel flags the zero pivot and the call fails loudly instead of returning NaNs."""
    X = synth.database(8, 60)[:, :, None, :]
    st = vq.FeatureStore(60, STREAMS, [1], 1024, devices=[0])
    st.upload(0, X)
    with pytest.raises(vq.VQError, match="singular"):
        st.bootstrap_target(np.array([3, 9, 3]), None, 0.0)
    ok = st.bootstrap_target(np.array([3, 9, 12]), np.array([20, 21]), 0.3)
    assert np.isfinite(ok).all()
    st.close()


@pytest.mark.parametrize("n", [1_000_000, 12_500_000], ids=["config2-1M", "config3-shard-12.5M"])
def test_finalize_selection_of_a_large_search_set_is_array_work(vq, tmp_path, monkeypatch, n):
    """Finalize on a 1M-clip store with ~90k matches: select_clips_to_review (max = inf: a seeded permutation of every
    match and near miss, ticket.py:333,341) and the report order (stable descending sort of the selection, ticket.py:266,
    ranked on the device with vq_rank_list) against their plain-Python restatement on the device's scores, and the host
    time of both (the per-clip HTTP calls of persistence are the API's contract and are not part of this)."""
    import time
    import types
    import torch
    if n * 8192 * 1.05 > torch.cuda.get_device_properties(0).total_memory:
        pytest.skip("device memory too small for %d clips" % n)
    seed = synth.DEFAULT_SEED
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0], clip_ids=np.arange(n) * 2 + 10)
    st.fill_synthetic(seed)
    T = sc.scale_target(synth.rows(seed, [18120]).astype(np.float64)[0][:, None, :])
    user = {str(10 + 2 * r): (r % 3 != 0) for r in range(1000, 600_000, 997)}        # ~600 labelled clips, a third rejected
    job = {"query_id": 1, "video_id": 1, "ref_clip": 0, "ref_clip_id": 10 + 2 * 18120, "search_set": 1,
           "number_of_matches_to_review": 20, "dynamic_target_adjustment": False, "user_matches": user}
    t = vq.Ticket(job, "http://fake/", client=object(), schema=object(), store=st)
    t.target = types.SimpleNamespace(target_features=tdict(T), splits={1})
    t._hp = vq.Hyperparameter({"rgb": 1.0, "warped_optical_flow": 1.5})
    t._weights = {"rgb": 1.0, "warped_optical_flow": 1.5}
    th, near = 0.8, 0.35
    random.seed(a="finalize")
    t.select_clips_to_review(th, float("inf"), near)                 # warm-up (mirror growth, staging allocations)
    random.seed(a="finalize")
    t0 = time.perf_counter()
    t.select_clips_to_review(th, float("inf"), near)
    t_sel = time.perf_counter() - t0
    t0 = time.perf_counter()
    r_ids, r_sc = t.ranked_selection_arrays()                        # what create_final_report walks
    t_rank = time.perf_counter() - t0
    ranked = t.ranked_selection()                                    # the same as a list of pairs (one tuple per clip)
    assert ranked == list(zip(r_ids.tolist(), r_sc.tolist()))
    state_after = None
    # the reference's selection, restated on the device's own scores
    scores = st.scores()
    ids = st.clip_ids
    g = scores.astype(np.float64)
    lo = th - near * (1 - th)
    M = {int(ids[r]): float(scores[r]) for r in np.flatnonzero(g >= th)}
    NM = {int(ids[r]): float(scores[r]) for r in np.flatnonzero((g >= lo) & (g < th))}
    assert len(M) >= 50_000
    random.seed(a="finalize")
    picked = random.sample(list(M.items()), len(M))
    best = max(NM, key=lambda k: NM[k])
    best_item = {best: NM.pop(best)}
    picked_n = random.sample(list(NM.items()), len(NM))
    want = dict(picked + picked_n)
    want.update(best_item)
    forced = {job["ref_clip_id"]: float(scores[18120])}
    forced.update({int(c): float(scores[(int(c) - 10) // 2]) for c, v in user.items() if v is True})
    want.update(forced)
    assert list(t.matches.items()) == list(want.items())
    assert ranked == sorted(want.items(), key=lambda kv: kv[1], reverse=True)
    print("finalize on %d clips: %d selected; selection %.1f ms (scan included), report order %.1f ms"
          % (n, len(want), 1e3 * t_sel, 1e3 * t_rank))
    if n == 1_000_000:
        assert t_sel + t_rank < 0.15, (t_sel, t_rank)  # the Python-object floor: the {clip: score} dict of ~150k entries
    st.close()


def test_a_full_device_reports_out_of_memory_evicts_and_goes_on(vq):
    """Two shards of 55 % of the device memory each do not fit together: creating the second one fails with the
    library's "out of memory" error (no crash, no sticky CUDA error), `evict_least_recently_used` closes the resident
    one, the retry succeeds and the new shard scans correctly — what Ticket._build_store does when a broker's search sets
    outgrow the HBM (the CPU twin is tests/test_rounds_cpu.py)."""
    import torch
    from video_query_algorithms_b200 import store as ps
    total = torch.cuda.get_device_properties(0).total_memory
    n = int(0.55 * total / 8192)
    ps.invalidate()
    a = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    a.last_used = 1.0
    ps.register_store(("test", "a"), a)
    small = vq.FeatureStore(5000, STREAMS, [1], 1024, devices=[0])          # a bystander that must keep working
    small.fill_synthetic(3)
    small.last_used = 2.0
    ps.register_store(("test", "small"), small)
    with pytest.raises(vq.VQError, match="out of memory"):
        vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    T = sc.scale_target(synth.rows(3, [7]).astype(np.float64)[0][:, None, :])
    before = small.scan(tdict(T), (1.0, 1.5), 0.8, 0.73, EPS, topk=10)      # the failed allocation left no error behind
    assert ps.evict_least_recently_used(keep=("test", "b")) and ("test", "a") not in ps._REGISTRY
    b = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    ps.register_store(("test", "b"), b)
    after = small.scan(tdict(T), (1.0, 1.5), 0.8, 0.73, EPS, topk=10)
    assert (before.n_match, before.n_near) == (after.n_match, after.n_near) and before.n_match >= 1
    assert ps.evict_least_recently_used(keep=("test", "b")) and list(ps._REGISTRY) == [("test", "b")]
    assert not ps.evict_least_recently_used(keep=("test", "b"))
    ps.invalidate()


def test_store_follows_a_growing_search_set_on_the_gpu(vq, tmp_path, monkeypatch):
    """load_db.py adds clips between ticks (reference load_db.py:10-28; the reference re-reads the search set per job): the
    resident store notices (the search-set record changed) and appends the new clips — results equal a store built fresh."""
    from fake_api import FakeRepository
    scn = Scenario("B_both_simple_mu")
    api, qid = scn.build_api()
    monkeypatch.chdir(tmp_path)
    vq.invalidate()
    ss = api.queries[qid]["search_set_to_query"]
    all_ids = list(api.search_sets[ss]["clip_ids"])
    made = []
    fac = lambda job, url: made.append(vq.Ticket(job, url, client=api.client(), devices=[0])) or made[-1]
    api.search_sets[ss]["clip_ids"] = all_ids[:120]
    api.request(qid, "new")
    random.seed(a=scn.seed)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=fac)
    st = made[-1].feature_store()
    assert st.n_rows == 120
    api.search_sets[ss]["clip_ids"] = all_ids
    api.request(qid, "new")
    random.seed(a=scn.seed)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=fac)
    assert made[-1].feature_store() is st and st.clip_ids.tolist() == all_ids
    grown_scores, grown_sel = made[-1].scores.array().copy(), list(made[-1].matches.items())
    assert_scores_close(grown_scores, scn.arr(0, "scores"))
    vq.invalidate()
    api.request(qid, "new")
    random.seed(a=scn.seed)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=fac)
    assert made[-1].feature_store() is not st
    assert np.array_equal(made[-1].scores.array(), grown_scores) and list(made[-1].matches.items()) == grown_sel
    vq.invalidate()


def test_config4_batched_scan_at_size_matches_float64_on_sampled_rows(vq):
    """BASELINE configs[3] at a size the tensor-core pipeline runs in steady state (2M clips x 256 queries, several launches
    of 16 tiles per CTA): per-query counts against the single-query scan, every top-k row and a seeded sample of rows
    against float64 on rows regenerated on the CPU."""
    n, Q, seed = 2_000_000, 256, 20261018
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0])
    st.fill_synthetic(seed)
    qrows = 18120 + 37 * np.arange(Q)
    Xq = synth.rows(seed, qrows).astype(np.float64)[:, :, None, :]
    T = np.stack([sc.scale_target(x) for x in Xq]).astype(np.float32)
    w, th, lo = (1.0, 1.5), 0.8, sc.lower_limit(0.8, 0.35)
    counts, rows, scores, ms = st.scan_batch(T, w, th, lo, topk=100)
    assert rows.shape == (Q, 100) and (rows >= 0).all()
    for q in (0, 100, 255):
        res = st.scan(tdict(T[q].astype(np.float64)), w, th, lo, EPS, topk=100)
        single = st.scores()
        # fp32 CUDA-core scores vs tensor-core scores differ by ~1e-6: counts agree up to the rows that close to a boundary
        g = single.astype(np.float64)
        slack_m = np.count_nonzero(np.abs(g - th) < 1e-5)
        slack_n = slack_m + np.count_nonzero(np.abs(g - lo) < 1e-5)
        assert abs(int(counts[q, 0]) - res.n_match) <= slack_m and abs(int(counts[q, 1]) - res.n_near) <= slack_n
        rng = np.random.default_rng(q)
        chk = np.unique(np.concatenate([rows[q], rng.integers(0, n, 400)]))
        Xc = synth.rows(seed, chk).astype(np.float64)[:, :, None, :]
        s64 = sc.scores(sc.similarities(Xc, T[q].astype(np.float64))[0], w)
        at = np.searchsorted(chk, rows[q])
        assert_scores_close(scores[q], s64[at])
        assert np.all(np.diff(scores[q]) <= 0)
        # the k-th best of the batch is at least as good as every sampled row outside the top-k (up to the error budget)
        outside = np.setdiff1d(np.arange(len(chk)), at)
        assert s64[outside].max() <= float(scores[q][-1]) + 1e-5
    assert rows[0][0] == 18120 and scores[0][0] == pytest.approx(1.0, abs=2e-6)
    st.close()


def test_config5_replicate_weight_update_at_size_matches_oracle(vq):
    """BASELINE configs[4] at its size: 1000 seeded replicates over 5000 labelled clips (labelled similarities in fp64 on
    the GPU, one loss-grid launch for all replicates) — replicate index sets equal to the reference's draws
    (random.choices + set, target_clip.py:297-309), loss grids equal to the oracle's on a sample of replicates."""
    L, R, n, seed = 5000, 1000, 200_000, 20261018
    st = vq.FeatureStore(n, STREAMS, [1], 1024, devices=[0], clip_ids=np.arange(n))
    st.fill_synthetic(seed)
    T = sc.scale_target(synth.rows(seed, [18120]).astype(np.float64)[0][:, None, :])
    rows = np.sort(np.random.default_rng(7).choice(n, L, replace=False)).astype(np.int64)
    sims = st.labelled_sims(tdict(T), rows)
    X64 = synth.rows(seed, rows[:200]).astype(np.float64)[:, :, None, :]
    want_sims, _ = sc.similarities(X64, T)
    assert np.abs(sims[:200] - want_sims).max() <= 1e-12 * np.abs(want_sims).max()
    labels = sc.scores(sims, (1.0, 1.5)) >= 0.62
    random.seed("73459912436")
    state = random.getstate()
    reps = vq.resample_labelled(L, R, random)
    after = random.getstate()
    random.setstate(state)
    for r in (0, 1, R - 1):                                     # the reference's own draws for the first replicates ...
        want = list(set(random.choices(range(L), k=L)))
        if r < 2:
            assert reps[r].tolist() == want
    wg, tg = sc.weight_grid(), sc.threshold_grid()
    got = vq.loss_grid(sims, labels, wg, tg, 0.1, replicates=reps)
    assert got.shape == (R, 40, 31)
    for r in (0, 17, 500, R - 1):
        want = sc.loss_grid_fast(sims[reps[r]], labels[reps[r]], wg, tg, 0.1)
        assert np.abs(got[r] - want).max() < 1e-13
    random.setstate(after)
    st.close()

"""CPU-only checks of the boundary and of the host-side logic (no compute calls into CUDA)."""
import ctypes
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import bootstrap as ob
from oracle import scoring as sc
from scenarios import ORACLE_ONLY_SCENARIOS, SCENARIOS, Scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    from video_query_algorithms_b200 import _ffi
    return _ffi


def header_functions():
    src = open(os.path.join(ROOT, "include", "vq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    names = header_functions()
    assert len(names) >= 25
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "libvq_b200.so does not export " + n
    assert sorted(built_lib.PROTOTYPES) == names            # the ctypes table covers the header exactly
    assert built_lib.lib().vq_abi_version() == 2


def test_header_is_plain_c(tmp_path):
    """include/vq.h is the boundary a non-Python host binds: it must compile as C99 on its own (no C++, no CUDA, no torch
    types), and the minimal use shown in INTEGRATION.md must type-check against it."""
    src = tmp_path / "use_vq.c"
    src.write_text('#include "vq.h"\n'
                   'int use(vq_store *s, const float *target, long long *rows, float *scores) {\n'
                   '    vq_scan_params p = {{1.0, 1.5}, 0.8, 0.73, 3e-6, 100, 0};\n'
                   '    vq_scan_counts c;\n'
                   '    if (vq_scan(s, target, &p, &c)) return 1;\n'
                   '    return vq_fetch_matches(s, c.n_match, (int64_t *)rows, scores);\n'
                   '}\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_plain_c_consumer_compiles_and_fails_loudly_without_a_gpu(built_lib, tmp_path):
    """tests/c_abi/scan_consumer.c (C99, -Wall -Wextra -Werror) links against the library with nothing but include/vq.h; in
    this container (no GPU) it must end with a non-zero status and the library's error text, not a crash or a result."""
    import torch
    exe = str(tmp_path / "scan_consumer")
    lib_dir = os.path.dirname(built_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "scan_consumer.c"), "-L", lib_dir, "-lvq_b200",
                    "-Wl,-rpath," + lib_dir, "-o", exe], check=True, capture_output=True, text=True)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is tests/test_gpu_parity.py's")
    r = subprocess.run([exe, "1000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and r.stdout == "" and ("failed" in r.stderr or "no CUDA device" in r.stderr), (r.returncode, r.stderr)


def test_no_gpu_means_loud_failure_not_fallback(built_lib):
    import video_query_algorithms_b200 as vq
    n = ctypes.c_int()
    rc = built_lib.lib().vq_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(vq.VQError):
        vq.FeatureStore(8, ("rgb", "warped_optical_flow"), [1], 1024)


def test_product_package_never_imports_the_oracle_or_reference():
    pkg = os.path.join(ROOT, "video_query_algorithms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f
    # the oracle is test infrastructure: besides tests/, only smoke() and bench.py's CPU legs use it
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            text = open(os.path.join(ROOT, "tools", f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_sass_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


# ---------------------------------------------------------------------------- host logic
def test_quad_fit_and_optimum_equal_oracle_on_golden_losses():
    from video_query_algorithms_b200 import Hyperparameter
    hp = Hyperparameter({"rgb": 1.0, "warped_optical_flow": 1.5})
    n = 0
    for name in SCENARIOS + ORACLE_ONLY_SCENARIOS:           # incl. a grid-border optimum (G: every label False)
        scn = Scenario(name)
        for i, r in enumerate(scn.rounds):
            key = "r%d_losses" % i
            if key not in scn.arrays.files:
                continue
            losses = scn.arrays[key]
            w, th = hp.optimum(losses)
            assert [1.0, w] == pytest.approx(r["weights"], rel=1e-12)
            assert th - scn.eps == pytest.approx(r["threshold"], rel=1e-12)
            n += 1
    assert n >= 12
    rng = np.random.default_rng(3)
    for _ in range(200):
        x = [sorted(rng.random(3) + 0.5), sorted(rng.random(3) + 0.5)]
        y = list(rng.random(5))
        assert Hyperparameter._quad_fit(x, y) == pytest.approx(sc.quad_fit(x, y), rel=1e-12, abs=1e-15)


def test_match_status_prefers_user_label():
    from video_query_algorithms_b200 import Hyperparameter
    ms = [{"video_clip": 5, "user_match": None, "is_match": True},
          {"video_clip": 7, "user_match": False, "is_match": True},
          {"video_clip": 5, "user_match": True, "is_match": False}]
    assert Hyperparameter.match_status(ms) == sc.match_status(ms) == {5: True, 7: False}


def test_random_fraction_consumes_rng_like_oracle():
    from video_query_algorithms_b200 import TargetClip, resample_labelled
    for frac, repl in ((0.5, False), (1, True), (0.3, True), (1, False)):
        random.seed("73459912436")
        a = TargetClip._random_fraction(np.arange(100, 141), frac, repl)
        sa = random.getstate()
        random.seed("73459912436")
        b = ob.random_fraction(41, frac, repl, random)
        assert list(a) == [100 + j for j in b] and random.getstate() == sa
    random.seed(5)
    reps = resample_labelled(50, 3, random)
    random.seed(5)
    assert [sorted(r) for r in reps] == [sorted(ob.random_fraction(50, 1, True, random)) for _ in range(3)]


def test_vectorised_choices_reproduce_pythons_stream_exactly():
    """_rng.choices_range draws `random.choices(range(n), k=k)` through numpy's MT19937 from the caller's own
    generator state: same indices, same generator state afterwards (so every later draw is unchanged), for the
    `random` module and for Random instances, with a pending gauss value preserved."""
    from video_query_algorithms_b200._rng import choices_range, set_order, _cpython_set_table_size
    for n, k, r in ((5000, 5000, 3), (1, 1, 1), (7, 3, 5), (41, 41, 1), (100000, 10, 2), (3, 1000, 1)):
        random.seed("73459912436")
        a = [random.choices(range(n), k=k) for _ in range(r)]
        sa, nxt = random.getstate(), random.random()
        random.seed("73459912436")
        b = choices_range(random, n, k, repeats=r)
        assert np.array_equal(np.asarray(a).reshape(b.shape), b) and random.getstate() == sa and random.random() == nxt
    r1, r2 = random.Random(5), random.Random(5)
    r1.gauss(0, 1), r2.gauss(0, 1)                                   # leaves gauss_next set
    assert r1.choices(range(50), k=50) == choices_range(r2, 50, 50).tolist() and r1.getstate() == r2.getstate()
    # list(set(draws)) without building the set when its order is known; the real set otherwise
    rng = np.random.default_rng(0)
    fallbacks = 0
    for n in (1, 2, 5, 8, 9, 31, 33, 100, 1000, 5000, 20000, 70000, 300000):
        for k in (1, 3, n // 3 + 1, n, 2 * n):
            d = rng.integers(0, n, size=k)
            fallbacks += n > _cpython_set_table_size(len(set(d.tolist())))
            assert set_order(d, n).tolist() == list(set(d.tolist())), (n, k)
    assert fallbacks > 5                                             # both branches were exercised
    from video_query_algorithms_b200 import resample_labelled
    random.seed(7)
    reps = resample_labelled(300, 40, random)
    s1 = random.getstate()
    random.seed(7)
    ref = [list(set(random.choices(range(300), k=300))) for _ in range(40)]
    assert all(list(a) == b for a, b in zip(reps, ref)) and s1 == random.getstate()


def test_sample_of_index_range_equals_sample_of_items():
    """select_clips_to_review samples index ranges; the reference samples dict items (ticket.py:333).
    Python's random.sample consumes the generator identically for equal (n, k)."""
    items = [(i * 7, i / 10) for i in range(500)]
    for k in (0, 1, 5, 6, 21, 22, 100, 500):
        random.seed(99)
        a = random.sample(items, k)
        sa = random.getstate()
        random.seed(99)
        b = [items[j] for j in random.sample(range(len(items)), k)]
        assert a == b and random.getstate() == sa


def test_batched_topk_merge_ranked_lists_unranked_lists_and_padding(built_lib):
    """vq_merge_topk_batch (per-query merge of per-shard / per-rank top-k lists): ranked input takes the k-way merge,
    unranked input the sort; both must give (score descending, global row ascending) with -1 / -inf padding."""
    from video_query_algorithms_b200.store import merge_topk_batch
    rng = np.random.default_rng(0)
    L, Q, k = 5, 33, 40
    sc_ = rng.random((L, Q, k)).astype(np.float32)
    sc_[:, :, ::3] = np.float32(0.5)                                      # ties inside and across lists
    rows = rng.permutation(L * Q * k).reshape(L, Q, k).astype(np.int64)
    for l in range(L):
        for q in range(Q):
            o = np.lexsort((rows[l, q], -sc_[l, q]))
            sc_[l, q], rows[l, q] = sc_[l, q][o], rows[l, q][o]
    rows[3, :, 25:], sc_[3, :, 25:] = -1, -np.inf                         # a short list
    rows[4], sc_[4] = -1, -np.inf                                         # an empty shard

    def want(q, rr, ss):
        a_s, a_r = ss[:, q].reshape(-1), rr[:, q].reshape(-1)
        m = a_r >= 0
        o = np.lexsort((a_r[m], -a_s[m]))[:k]
        return a_r[m][o], a_s[m][o]

    r, s_ = merge_topk_batch(rows, sc_)
    for q in range(Q):
        wr, ws = want(q, rows, sc_)
        assert np.array_equal(r[q], wr) and np.array_equal(s_[q], ws)
    shuffled_r, shuffled_s = rows.copy(), sc_.copy()
    for l in range(3):
        p_ = rng.permutation(k)
        shuffled_r[l], shuffled_s[l] = rows[l][:, p_], sc_[l][:, p_]
    r2, s2 = merge_topk_batch(shuffled_r, shuffled_s)
    assert np.array_equal(r2, r) and np.array_equal(s2, s_)
    few_r, few_s = rows[3:, :, :], sc_[3:, :, :]                          # fewer than k valid entries in total
    r3, s3 = merge_topk_batch(few_r, few_s)
    assert np.array_equal(r3[:, :25], rows[3, :, :25]) and np.all(r3[:, 25:] == -1) and np.all(np.isneginf(s3[:, 25:]))


# ---------------------------------------------------------------------------- ingest (host parsing)
def _write_csv(path, video, stream, clips, feats):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("video =%s, video url =../x/, CNN stream =%s, feature blob =global_pool, caffe model =/m.caffemodel\n"
                % (video, stream))
        for c, row in zip(clips, feats):
            f.write(",".join([str(c)] + [repr(float(x)) for x in row]) + "\n")


def test_feature_tree_parser_reads_the_load_db_layout(tmp_path):
    from video_query_algorithms_b200 import ingest
    rng = np.random.default_rng(0)
    want = {}
    for split in (1, 2):
        for stream in ("rgb", "warped_optical_flow"):
            feats = rng.random((5, 8))
            want[(stream, split)] = feats
            _write_csv(str(tmp_path / "vidA" / ("UCF101_split%d" % split) / (stream + "_global_pool_features.csv")),
                       "vidA", stream, [1, 2, 3, 4, 5], feats)
    tree = ingest.read_feature_tree(str(tmp_path))
    assert list(tree) == ["vidA"] and list(tree["vidA"]["clip_numbers"]) == [1, 2, 3, 4, 5]
    for (stream, split), feats in want.items():
        assert np.array_equal(tree["vidA"]["features"][stream][split], feats)     # repr() round-trips float64


def _python_csv(path):
    """The reference's own loop (api_load_records.py:45-58): csv.reader + int() / float() per cell."""
    import csv
    with open(path, "r") as f:
        reader = csv.reader(f)
        header = next(reader)
        clips, rows = [], []
        for row in reader:
            if row:
                clips.append(int(row[0]))
                rows.append([float(x) for x in row[1:]])
    return header, np.array(clips, np.int64), np.array(rows, np.float64).reshape(len(clips), -1)


def test_native_csv_reader_equals_csv_reader_plus_float(tmp_path):
    """vq_csv_read against Python's csv + float(): hard-to-round decimals, exponents, signs, blanks, CRLF,
    missing final newline, values outside the double range; any thread count gives the same arrays."""
    import time
    from video_query_algorithms_b200 import ingest
    rng = np.random.default_rng(3)
    n, dim = 257, 1024
    vals = rng.random((n, dim)) * 10.0 ** rng.integers(-12, 6, (n, dim))
    cells = [[repr(float(x)) for x in row] for row in vals]
    tricky = ["0.1", "1e-05", "+2.5", "-0.0", "5e-324", "1.7976931348623157e308", "2.2250738585072014e-308",
              "9007199254740993", "0.30000000000000004", "1E3", " 4.25", "7 ", "1e999", "-1e999", "1e-999", "123456789012345678901234567890",
              "8.41692461872845e-05", "0.000001", "3.", ".5"]
    cells[0][:len(tricky)] = tricky
    path = tmp_path / "rgb_global_pool_features.csv"
    with open(path, "w", newline="") as f:
        f.write("video =vidA, video url =../x/, CNN stream =rgb, feature blob =global_pool, caffe model =/m/a=b.caffemodel\r\n")
        for i, row in enumerate(cells):
            f.write(",".join([str(i + 1)] + row) + ("\r\n" if i % 2 else "\n"))
            if i == 100:
                f.write("\n")                                            # an empty line is skipped by both
        f.write(",".join(["9999"] + cells[1]))                            # last row without a newline
    t0 = time.perf_counter()
    header, clips, want = _python_csv(path)
    t_py = time.perf_counter() - t0
    t0 = time.perf_counter()
    rec = ingest.read_feature_csv(str(path))
    t_native = time.perf_counter() - t0
    assert rec["video"] == "vidA" and rec["stream"] == "rgb" and rec["feature_name"] == "global_pool"
    assert rec["weights_uri"] == header[4].split("=")[-1] == "b.caffemodel"
    assert np.array_equal(rec["clip_numbers"], clips) and clips[-1] == 9999 and len(clips) == n + 1
    assert rec["features"].shape == want.shape == (n + 1, dim)
    assert np.array_equal(rec["features"].view(np.uint64), want.view(np.uint64))      # bit for bit, signed zeros included
    one = ingest.read_feature_csv(str(path), n_threads=1)
    assert np.array_equal(one["features"].view(np.uint64), want.view(np.uint64))
    assert t_native < t_py, (t_native, t_py)


def test_native_csv_reader_rejects_malformed_files(tmp_path):
    import video_query_algorithms_b200 as vq
    from video_query_algorithms_b200 import ingest
    head = "video =v, video url =u, CNN stream =rgb, feature blob =global_pool, caffe model =m\n"
    for name, body in (("ragged", "1,0.5,0.25\n2,0.5\n"), ("text", "1,0.5,abc\n"), ("clip", "x,0.5,0.25\n"),
                       ("junk", "1,0.5,0.25x\n")):
        p = tmp_path / (name + ".csv")
        p.write_text(head + body)
        with pytest.raises(vq.VQError):
            ingest.read_feature_csv(str(p))
    with pytest.raises(vq.VQError):
        ingest.read_feature_csv(str(tmp_path / "missing.csv"))
    p = tmp_path / "empty.csv"
    p.write_text(head)
    rec = ingest.read_feature_csv(str(p))
    assert rec["features"].shape[0] == 0 and len(rec["clip_numbers"]) == 0


def test_feature_tree_parser_on_reference_fixture_if_present():
    src = "/root/reference/data/features/stock-video-clips_features"
    if not os.path.isdir(src):
        pytest.skip("reference data not present (GPU box)")
    from video_query_algorithms_b200 import ingest
    tree = ingest.read_feature_tree(src)
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixture_brooklyn.npz"))
    v = tree["DowntownBrooklynDrive_480p"]
    assert np.array_equal(v["clip_numbers"], z["clip_numbers"])
    for si, s in enumerate(("rgb", "warped_optical_flow")):
        for pi, p in enumerate(z["splits"]):
            assert np.array_equal(v["features"][s][int(p)], z["X"][:, si, pi])


def test_host_mailbox_allgather_between_threads_and_its_failure_modes(built_lib):
    """vq_hostx_* (csrc/vq_hostx.cu), the shared-memory all-gather of the rank-level path: three 'ranks' (threads with
    their own handles on one segment; ctypes drops the GIL) make 3000 calls of changing sizes and check every peer's
    record each time; a peer that never arrives is a timeout error, not a hang; oversized records, a second creation
    of the same name and attaching to a missing segment are refused."""
    import threading
    lib, VQError = built_lib.lib(), built_lib.VQError
    name = ("/vq-test-%d" % os.getpid()).encode()
    world, slot = 3, 4096
    hs = [ctypes.c_void_p() for _ in range(world)]
    for r in range(world):
        built_lib.check(lib.vq_hostx_create(ctypes.byref(hs[r]), name, world, r, slot), "vq_hostx_create")
    dup = ctypes.c_void_p()
    assert lib.vq_hostx_create(ctypes.byref(dup), name, world, 0, slot) != 0          # the name exists
    assert b"shm_open" in lib.vq_last_error()
    assert lib.vq_hostx_create(ctypes.byref(dup), name, world, 1, slot * 2) != 0      # ranks disagree on the slot size
    built_lib.check(lib.vq_hostx_unlink(hs[0]), "vq_hostx_unlink")
    assert not os.path.exists("/dev/shm" + name.decode())
    assert lib.vq_hostx_create(ctypes.byref(dup), name, world, 1, slot) != 0          # nothing to attach to any more
    errs = []

    def rank_main(r):
        try:
            rng = np.random.default_rng(1234)                                          # same sizes on every rank
            for c in range(3000):
                n = int(rng.integers(0, slot // 8 + 1))
                mine = np.arange(n, dtype=np.int64) * (r + 1) + c
                out = np.empty((world, n), np.int64)
                built_lib.check(lib.vq_hostx_allgather(hs[r], built_lib.ptr(mine), mine.nbytes, built_lib.ptr(out), 30.0),
                                "vq_hostx_allgather")
                for q in range(world):
                    if not np.array_equal(out[q], np.arange(n, dtype=np.int64) * (q + 1) + c):
                        raise AssertionError("call %d: rank %d read a wrong record of rank %d" % (c, r, q))
        except Exception as e:
            errs.append(e)

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    big = np.zeros(slot // 8 + 16, np.int64)
    out = np.empty((world, len(big)), np.int64)
    assert lib.vq_hostx_allgather(hs[0], built_lib.ptr(big), big.nbytes, built_lib.ptr(out), 1.0) == -1
    one = np.zeros(1, np.int64)
    assert lib.vq_hostx_allgather(hs[0], built_lib.ptr(one), 8, built_lib.ptr(out), 0.2) == -3   # the peers never call
    assert b"waited" in lib.vq_last_error()
    with pytest.raises(VQError):
        built_lib.check(-3, "vq_hostx_allgather")
    for h in hs:
        lib.vq_hostx_destroy(h)


def _mailbox_process(rank, world, name, slot, n_calls, q):
    try:
        sys.path.insert(0, ROOT)
        from video_query_algorithms_b200 import _ffi
        lib = _ffi.lib()
        h = ctypes.c_void_p()
        _ffi.check(lib.vq_hostx_create(ctypes.byref(h), name, world, rank, slot), "vq_hostx_create")
        rng = np.random.default_rng(99)
        bad = 0
        for c in range(n_calls):
            n = int(rng.integers(1, slot // 8 + 1))
            mine = np.full(n, (c << 8) | rank, np.int64)
            out = np.empty((world, n), np.int64)
            _ffi.check(lib.vq_hostx_allgather(h, _ffi.ptr(mine), mine.nbytes, _ffi.ptr(out), 60.0), "vq_hostx_allgather")
            bad += int(not all(np.all(out[r] == ((c << 8) | r)) for r in range(world)))
            if rank == c % world and c % 97 == 0:
                sum(range(20000))                            # one rank lags now and then: the others wait in the call
        lib.vq_hostx_destroy(h)
        q.put((rank, bad))
    except Exception as e:                                   # pragma: no cover
        q.put((rank, repr(e)))


def test_host_mailbox_between_processes(built_lib):
    """The same all-gather between 4 separate processes (the real arrangement: one rank process per GPU), 2000 calls of
    changing sizes; every record of every call must be the owner's, whole (no torn or stale slot)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    world, slot, n_calls = 4, 2048, 2000
    name = ("/vq-testp-%d" % os.getpid()).encode()
    q = ctx.Queue()                                          # no ordering between the ranks: an early rank waits for
    ps = [ctx.Process(target=_mailbox_process, args=(r, world, name, slot, n_calls, q)) for r in range(world)]
    for p_ in ps:                                            # rank 0's segment (up to 2 s) instead of failing; rank 0 is
        p_.start()                                           # started first, so a loaded machine does not stretch that wait
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p_ in ps:
        p_.join(timeout=30)
    assert res == [(r, 0) for r in range(world)], res
    assert not os.path.exists("/dev/shm" + name.decode())    # rank 0's destroy removed the name


# ---------------------------------------------------------------------------- Ticket.select_clips_to_review (host logic)
class _ArrayStore:
    """Test double for FeatureStore's result interface, serving a recorded score array: what the scan kernels would
    return for these scores (ordered lists, tie band, best near miss).  It lets the product's selection logic — sampling
    order, RNG consumption, forced clips — run against the reference's recorded rounds without a GPU."""

    def __init__(self, clip_ids, scores):
        from video_query_algorithms_b200.store import ScanResult
        self._result = ScanResult
        self.clip_ids, self.first_global_row = np.asarray(clip_ids, np.int64), 0
        self.n_rows, self.streams = len(clip_ids), ("rgb", "warped_optical_flow")
        self._s = np.asarray(scores, np.float64)
        self._row = {int(c): i for i, c in enumerate(clip_ids)}

    def has_clip(self, c):
        return c is not None and int(c) in self._row

    def row_of(self, c):
        return self._row[int(c)]

    def scan(self, target, weights, threshold, lower_limit, eps, topk=0, want_sims=False, lists=True, packed=None):
        s = self._s
        self._l = {"matches": np.flatnonzero(s >= threshold), "near_misses": np.flatnonzero((s >= lower_limit) & (s < threshold)),
                   "ties": np.flatnonzero((np.abs(s - threshold) < eps) | (np.abs(s - lower_limit) < eps))}
        self.lists = lists
        return self._result(len(self._l["matches"]), len(self._l["near_misses"]), len(self._l["ties"]), 0, 0.0)

    def _list(self, which, need_lists):
        assert self.lists or not need_lists, "whole lists read after a lists=False scan"
        return self._l[which].astype(np.int64), self._s[self._l[which]]

    def matches(self, copy=True):
        return self._list("matches", True)

    def near_misses(self, copy=True):
        return self._list("near_misses", True)

    def ties(self, copy=True):
        return self._list("ties", False)

    def gather(self, which, positions):
        assert not self.lists, "gather is the lists=False round's call"
        pos = np.asarray(positions, np.int64)
        return self._l[which][pos].astype(np.int64), self._s[self._l[which][pos]]

    def gather_many(self, requests):
        return [self.gather(w, p) for w, p in requests]

    def near_best(self):
        n = self._l["near_misses"]
        if not len(n):
            return None
        j = int(np.argmax(self._s[n]))
        return j, int(n[j]), float(self._s[n[j]])

    def topk(self):
        return np.empty(0, np.int64), np.empty(0, np.float32)

    def rows_of(self, clips):
        return np.array([self._row[int(c)] for c in clips], np.int64)

    def scores_at(self, rows):
        return self._s[np.asarray(rows, np.int64)]

    def rank_list(self, which, place):
        sc = self._s[self._l[which]]
        order = np.lexsort((np.asarray(place), -sc))
        return np.asarray(place)[order], sc[order]


@pytest.mark.parametrize("name", SCENARIOS + ORACLE_ONLY_SCENARIOS)
def test_ticket_selection_reproduces_reference_rounds_on_recorded_scores(name):
    """Product `Ticket.select_clips_to_review` on the reference's recorded scores: same clips in the same order with
    the same scores, and Python's generator left in the same state, for every round of every golden scenario (review
    rounds take the lists=False + gather path, finalize rounds the whole-list path)."""
    import types
    from scenarios import rng_digest
    from video_query_algorithms_b200 import Ticket
    scn = Scenario(name)
    hp = scn.hp()
    for i, r in enumerate(scn.rounds):
        random.seed(a=scn.seed)
        prev = r.get("match_status_input")
        if r["kind"] != "new" and prev:                      # advance the generator through the target bootstrap's draws
            dyn = scn.meta["dynamic_target_adjustment"] and any(m["user_match"] is True for m in prev)
            rows = lambda want: scn.X[[scn.row_of[m["video_clip"]] for m in prev if m["user_match"] is want]]
            ob.get_target_features(scn.X[scn.row_of[r["ref_clip_id"]]], rows(True), rows(False), None, dyn, True,
                                   hp["bootstrap_type"], hp["f_bootstrap"], hp["f_memory"], hp["nbags"], hp["mu"], random)
        assert rng_digest() == r["rng_before_select"]
        store = _ArrayStore(scn.clip_ids, scn.arr(i, "scores"))
        job = {"query_id": 1, "video_id": 1, "ref_clip": 0, "ref_clip_id": r["ref_clip_id"], "search_set": 1,
               "number_of_matches_to_review": scn.meta["max_matches"],
               "dynamic_target_adjustment": scn.meta["dynamic_target_adjustment"], "user_matches": r["user_matches"]}
        t = Ticket(job, "http://fake/", client=object(), schema=object(), store=store)
        t.target = types.SimpleNamespace(target_features={})
        t._weights = dict(zip(scn.streams, r["weights"]))
        th, mx, near = r["select_args"]
        t.select_clips_to_review(th, mx, near)
        assert store.lists == (mx == float("inf"))
        assert list(t.matches.items()) == [(k, v) for k, v in r["selected"]]
        assert rng_digest() == r["rng_after_select"]
        if mx == float("inf"):                               # report order: the reference's stable descending sort of the dict
            want = sorted(t.matches.items(), key=lambda kv: kv[1], reverse=True)
            assert t.ranked_selection() == want


def test_report_order_with_heavy_ties_and_forced_clips_outside_the_lists():
    """Finalize selection + report order on scores with many exact ties (quantised) and with forced clips (the reference
    clip, confirmed clips) below the near-miss band: Ticket.ranked_selection equals the reference's stable descending
    sort of the selection dict (ticket.py:266) — equal scores keep the order in which the selection inserted them, an
    extra clip comes after every listed clip of its score, extras of one score keep their own order."""
    import types
    from video_query_algorithms_b200 import Ticket
    for seed in range(6):
        rng = np.random.default_rng(seed)
        n = 3000
        scores = np.round(rng.uniform(0.4, 1.0, n), 2)              # ~60 distinct values: every score is shared by ~50 clips
        ids = rng.permutation(n) + 100
        low = np.flatnonzero(scores < 0.6)
        confirmed = rng.choice(low, 25, replace=False)               # below the band (threshold .8, near .35 -> lower .73)
        inband = rng.choice(np.flatnonzero(scores >= 0.75), 10, replace=False)
        user = {str(int(ids[r])): True for r in np.concatenate([confirmed[:12], inband, confirmed[12:]])}
        user.update({str(int(ids[r])): False for r in rng.choice(n, 20, replace=False) if str(int(ids[r])) not in user})
        job = {"query_id": 1, "video_id": 1, "ref_clip": 0, "ref_clip_id": int(ids[low[0]]), "search_set": 1,
               "number_of_matches_to_review": 20, "dynamic_target_adjustment": False, "user_matches": user}
        t = Ticket(job, "http://fake/", client=object(), schema=object(), store=_ArrayStore(ids, scores))
        t.target = types.SimpleNamespace(target_features={})
        t._weights = {"rgb": 1.0, "warped_optical_flow": 1.5}
        random.seed(a=seed)
        t.select_clips_to_review(0.8, float("inf"), 0.35)
        assert len(t.matches) > t._selection["n_listed"]             # some forced clips sit in neither list
        want = sorted(t.matches.items(), key=lambda kv: kv[1], reverse=True)
        assert t.ranked_selection() == want
        r_ids, r_sc = t.ranked_selection_arrays()
        assert list(zip(r_ids.tolist(), r_sc.tolist())) == want


# ---------------------------------------------------------------------------- TargetClip (host logic)
class _SolveStore:
    """Test double for the store calls TargetClip makes: rows by clip id, and `bootstrap_target` answered by the
    oracle's float64 solve (the stand-in for kernel K6), so that the product's own case analysis, resampling order,
    bagging average and output format can be checked against the reference's recorded targets without a GPU."""

    def __init__(self, scn):
        self.scn, self.first_global_row, self.present = scn, 0, None
        self.streams, self.splits = scn.streams, list(scn.splits)

    def row_of(self, clip):
        return self.scn.row_of[int(clip)]

    def has_clip(self, clip):
        return clip is not None and int(clip) in self.scn.row_of

    def bootstrap_target(self, valid_rows, invalid_rows, mu, slots=None):
        X = self.scn.X
        out = np.empty(X.shape[1:], np.float64)
        for s in range(X.shape[1]):
            for p in range(X.shape[2]):
                out[s, p] = (ob.solve_valid_invalid(X[valid_rows, s, p], X[invalid_rows, s, p], mu) if len(invalid_rows)
                             else ob.solve_valid(X[valid_rows, s, p]))
        return out


class _TargetClient:
    """The two API actions TargetClip uses: the labelled matches of the previous round, paged, and the reference
    clip's feature rows in the `video-clips/features` record layout."""

    def __init__(self, scn, prev_matches, page_size=7):
        self.scn, self.prev, self.page_size = scn, prev_matches or [], page_size

    def action(self, schema, keys, params=None, **kw):
        if keys == ["matches", "list"]:
            a = (params["page"] - 1) * self.page_size
            more = a + self.page_size < len(self.prev)
            return {"results": self.prev[a:a + self.page_size], "pagination": {"nextPage": params["page"] + 1 if more else None}}
        assert keys == ["video-clips", "features"]
        row = self.scn.X[self.scn.row_of[params["id"]]]
        return [{"dnn_stream_id": s, "dnn_stream_split": p, "name": "global_pool", "feature_vector": row[si, pi].tolist(),
                 "video_clip_id": params["id"]} for si, s in enumerate(self.scn.streams) for pi, p in enumerate(self.scn.splits)]


@pytest.mark.parametrize("name", SCENARIOS + ORACLE_ONLY_SCENARIOS)
def test_target_clip_reproduces_reference_targets(name):
    """Product `TargetClip.get_target_features` for every recorded round: the reference's target vectors (1e-9; the
    solve itself is the oracle's here) and Python's generator in the reference's state afterwards."""
    import types
    from scenarios import rng_digest
    from video_query_algorithms_b200 import Hyperparameter, TargetClip
    scn = Scenario(name)
    kw = scn.hp()
    for i, r in enumerate(scn.rounds):
        prev = r.get("match_status_input")
        store = _SolveStore(scn)
        dyn = scn.meta["dynamic_target_adjustment"]
        if r["kind"] != "new" and dyn and prev is not None and not any(m["user_match"] is True for m in prev):
            dyn = False                                      # what Ticket.catch_errors does (ticket.py:98-107)
        ticket = types.SimpleNamespace(
            client=_TargetClient(scn, prev), schema=None, dynamic_target_adjustment=dyn, ref_clip_id=r["ref_clip_id"],
            latest_query_result=None if r["kind"] == "new" else {"id": 1, "round": i, "bootstrapped_target": None},
            features_from_store=False, attach_store=lambda hp: store, feature_store=lambda optional=False: store)
        random.seed(a=scn.seed)
        tc = TargetClip(ticket, Hyperparameter(**kw))
        tc.get_target_features()
        got = np.array([[tc.target_features[s][p] for p in r["splits"]] for s in scn.streams])
        want = scn.arr(i, "target")
        assert got.shape == want.shape and np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-9
        assert rng_digest() == r["rng_after_target"]
        assert isinstance(tc.target_features[scn.streams[0]][r["splits"][0]], list)      # JSON-serialisable (ticket.py:296)
        json_ok = __import__("json").dumps(tc.target_features)
        assert json_ok


@pytest.mark.parametrize("name", SCENARIOS + ORACLE_ONLY_SCENARIOS)
def test_optimize_weights_reproduces_reference_on_recorded_similarities(name, monkeypatch):
    """Product `Hyperparameter.optimize_weights` for every recorded revise / finalize round, with the two device calls
    answered on the CPU (labelled similarities from the recording, loss grid by the oracle): label extraction and its
    order, grid handling, optimum and threshold buffer must give the reference's weights and threshold."""
    import types
    from video_query_algorithms_b200 import Hyperparameter
    from video_query_algorithms_b200 import store as product_store
    scn = Scenario(name)
    n = 0
    for i, r in enumerate(scn.rounds):
        prev = r.get("match_status_input")
        if not prev or "r%d_losses" % i not in scn.arrays.files:
            continue
        sims_all = scn.arr(i, "sims")

        def loss_grid(sims, labels, wg, tg, ballast, replicates=None, device=0):
            return sc.loss_grid(sims, np.asarray(labels, bool), wg, tg, ballast)[None]

        monkeypatch.setattr(product_store, "loss_grid", loss_grid)
        fs = types.SimpleNamespace(shards=[types.SimpleNamespace(device=0)],
                                   rows_of=lambda clips: np.array([scn.row_of[int(c)] for c in clips]),
                                   labelled_sims=lambda target, rows: sims_all[rows])
        ticket = types.SimpleNamespace(matches=prev, target=types.SimpleNamespace(target_features={}),
                                       feature_store=lambda optional=False: fs)
        hp = Hyperparameter(**scn.hp())
        hp.optimize_weights(ticket)
        assert [hp.weights[s] for s in scn.streams] == pytest.approx(r["weights"], rel=1e-10)
        assert hp.threshold == pytest.approx(r["threshold"], rel=1e-10)
        assert np.max(np.abs(hp.losses - scn.arr(i, "losses"))) < 1e-12
        n += 1
    assert n >= 1 or name == "never"


# ---------------------------------------------------------------------------- feature records -> store rows (A2)
def _records(rng, n_clips, splits, dim, streams=("rgb", "warped_optical_flow")):
    """A `search-sets/features`-style response with everything the filters and the dict semantics must cope with:
    shuffled records, duplicates of a (clip, stream, split) slot, clips lacking splits, foreign streams and names."""
    recs = []
    for c in range(100, 100 + n_clips):
        for s in streams + ("audio",):
            for p in splits:
                if rng.random() < 0.15 and p != splits[0]:                     # (every clip keeps its first split)
                    continue                                                   # this clip lacks this slot
                for _ in range(2 if rng.random() < 0.1 else 1):                # a duplicate record: the later one wins
                    recs.append({"dnn_stream_id": s, "dnn_stream_split": p, "name": "global_pool" if rng.random() < 0.9 else "fc",
                                 "feature_vector": rng.random(dim).tolist(), "video_clip_id": c})
    order = rng.permutation(len(recs))
    return [recs[i] for i in order]


def test_pack_feature_rows_follows_the_reference_dict_semantics():
    """store.pack_feature_rows against (a) a direct restatement of the reference's nested-dict build (ticket.py:367-381)
    and, when the reference tree is present, (b) the reference's own `Ticket._get_candidate_features`."""
    from video_query_algorithms_b200.store import pack_feature_rows
    streams, rng = ("rgb", "warped_optical_flow"), np.random.default_rng(11)
    ref_ticket = None
    if os.path.isdir("/root/reference/src"):
        import test_flow_cpu
        test_flow_cpu.load_reference_driver()                                 # puts the reference on sys.path (stub coreapi)
        import models.ticket as rticket
        ref_ticket = object.__new__(rticket.Ticket)
        ref_ticket.search_set = 1
    for trial in range(6):
        splits = [1, 2, 3] if trial % 2 else [1]
        recs = _records(rng, 40, splits, 16)
        order, got_splits, X, present = pack_feature_rows(recs, streams, "global_pool")
        want = {s: {p: {} for p in splits} for s in streams}                  # the reference's structure
        for tf in recs:
            if tf["dnn_stream_id"] in streams and tf["name"] == "global_pool" and tf["dnn_stream_split"] in splits:
                want[tf["dnn_stream_id"]][tf["dnn_stream_split"]][tf["video_clip_id"]] = tf["feature_vector"]
        if ref_ticket is not None:
            ref_ticket._request = lambda action, params, recs=recs: recs
            hp = type("HP", (), {"streams": streams, "feature_name": "global_pool"})()
            assert ref_ticket._get_candidate_features(splits, hp) == want
        # row order = insertion order of the reference's scores dict: first stream, split by split (ticket.py:146-160)
        ref_order = list(dict.fromkeys(c for p in splits for c in want[streams[0]][p]))
        ref_order += [c for c in dict.fromkeys(c for p in splits for c in want[streams[1]][p]) if c not in set(ref_order)]
        assert order == ref_order and got_splits == sorted({p for s in want for p in want[s] if want[s][p]})
        for si, s in enumerate(streams):
            for pi, p in enumerate(got_splits):
                for r, c in enumerate(order):
                    v = want[s][p].get(c)
                    assert present[r, si, pi] == (v is not None)
                    assert np.array_equal(X[r, si, pi], np.asarray(v, np.float32) if v is not None else np.zeros(16, np.float32))
        # append: clips a store already holds are dropped, the rest keep their first-appearance order
        held_ids = set(order[::3])
        o2, s2, X2, p2 = pack_feature_rows(recs, streams, "global_pool", held=lambda ids: np.array([c in held_ids for c in ids]))
        keep = [r for r, c in enumerate(order) if c not in held_ids]
        assert o2 == [order[r] for r in keep]
        at = [got_splits.index(p) for p in s2]
        assert np.array_equal(X2, X[keep][:, :, at]) and np.array_equal(p2, present[keep][:, :, at])
    empty = pack_feature_rows([{"dnn_stream_id": "audio", "dnn_stream_split": 1, "name": "global_pool", "feature_vector": [1.0],
                                "video_clip_id": 5}], streams, "global_pool")
    assert empty[0] == [] and empty[2].shape[0] == 0


class _RecordingLib:
    """Stands in for libvq_b200 in host-glue tests: every call succeeds and is recorded (arrays copied)."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            if name == "vq_device_count":
                args[0]._obj.value = 1
            self.calls.append((name,) + tuple(a.copy() if isinstance(a, np.ndarray) else a for a in args))
            return 0
        return fn

    def of(self, name):
        return [c for c in self.calls if c[0] == name]


def test_store_build_and_append_from_feature_records_issue_the_right_uploads(monkeypatch):
    """FeatureStore.from_feature_rows / append_feature_rows end to end on the host side (library calls recorded):
    what is uploaded, in which row order, with which split weights — the glue around pack_feature_rows."""
    from video_query_algorithms_b200 import store as ps
    fake = _RecordingLib()
    monkeypatch.setattr(ps, "lib", lambda: fake)
    monkeypatch.setattr(ps, "ptr", lambda a: a)
    monkeypatch.setattr(ps.FeatureStore, "_alloc_staging", lambda self, n: np.empty(int(n), np.float32))   # pinned in the product
    monkeypatch.setattr(ps.FeatureStore, "_free_staging", lambda self, a: None)
    monkeypatch.setattr(ps.FeatureStore, "CHUNK_BYTES", 7 * 2 * 3 * 8 * 4)                              # 7 rows per staging chunk
    streams, rng = ("rgb", "warped_optical_flow"), np.random.default_rng(4)
    recs = _records(rng, 30, [1, 2, 3], 8)
    st = ps.FeatureStore.from_feature_rows(recs, streams, "global_pool", devices=[0])
    order, splits, X, present = ps.pack_feature_rows(recs, streams, "global_pool")
    assert list(st.clip_ids) == order and st.splits == splits and st.n_rows == len(order) and st.dim == 8
    (create,) = fake.of("vq_store_create")
    assert create[2:] == (0, len(order), 2, 3, 8, 0)
    ups = fake.of("vq_store_upload_async")                      # the ingest pipeline: 7-row chunks through two staging buffers
    assert [u[2:4] for u in ups] == [(r, min(7, len(order) - r)) for r in range(0, len(order), 7)]
    assert np.array_equal(np.concatenate([u[4] for u in ups]), X.reshape(-1))
    names = [c[0] for c in fake.calls if c[0] in ("vq_store_upload_async", "vq_store_sync")]
    assert names[:2] == ["vq_store_upload_async"] * 2 and names[2] == "vq_store_sync" and names[-1] == "vq_store_sync"
    (sw,) = fake.of("vq_store_set_split_weights")
    assert np.allclose(sw[2], 1.0 / present.sum(axis=2)) and not present.all()
    # append: a response that repeats held clips and brings new ones, some of which lack splits
    more = _records(rng, 45, [1, 2], 8)                       # clips 100..144: 100..129 are held
    fake.calls.clear()
    added = st.append_feature_rows(more, "global_pool")
    o2, s2, X2, p2 = ps.pack_feature_rows(more, streams, "global_pool", held=lambda ids: np.array([c < 130 for c in ids]))
    assert added == len(o2) == 15 and list(st.clip_ids) == order + o2 and st.n_rows == len(order) + 15
    (ap,) = fake.of("vq_store_append")
    wide = np.zeros((15, 2, 3, 8), np.float32)
    wide[:, :, :2] = X2                                       # the store's third split slot stays zero / absent
    assert ap[2] == 15 and np.array_equal(ap[3], wide.reshape(15, -1))
    (sw2,) = fake.of("vq_store_set_split_weights")
    full = np.concatenate([present, np.concatenate([p2, np.zeros((15, 2, 1), bool)], axis=2)])
    assert np.allclose(sw2[2], 1.0 / full.sum(axis=2))
    assert st.append_feature_rows(more, "global_pool") == 0   # nothing new the second time
    with pytest.raises(ps.VQError):
        st.append_feature_rows(_records(rng, 50, [4], 8), "global_pool")          # a split the store does not have


def test_every_entry_point_refuses_null_arguments_without_crashing(built_lib):
    """The ABI contract: every function returns an int status and leaves a message — no crash, no exception across
    the boundary.  Each entry point is called with null pointers and zeros (in a child process, so that a regression
    shows as a failed test and not as a dead test run); the destroy calls accept null like free()."""
    code = (
        "import sys, ctypes as C\n"
        "sys.path.insert(0, %r)\n"
        "from video_query_algorithms_b200 import _ffi\n"
        "lib = _ffi.lib()\n"
        "for name, (res, args) in sorted(_ffi.PROTOTYPES.items()):\n"
        "    vals = [0 if a in (C.c_int, C.c_int32, C.c_int64, C.c_uint64) else 0.0 if a is C.c_double else None for a in args]\n"
        "    r = getattr(lib, name)(*vals)\n"
        "    if res is C.c_int and name != 'vq_abi_version':\n"
        "        print(name, r, len(lib.vq_last_error() or b''))\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-500:]
    seen = 0
    for line in p.stdout.splitlines():
        name, rc, msg_len = line.split()
        if name.endswith("_destroy") or name.endswith("_free"):
            assert int(rc) == 0, line
        else:
            assert int(rc) < 0 and int(msg_len) > 0, line
        seen += 1
    assert seen == len(built_lib.PROTOTYPES) - 2            # all but vq_last_error and vq_abi_version


def test_topk_merges_hold_for_arbitrary_lists(built_lib):
    """Property test (hypothesis) of vq_merge_topk / vq_merge_topk_batch: for any number of lists, any k, heavy ties,
    short and empty lists, ranked or not, the merge is the k best under (score descending, global row ascending) with
    -inf / -1 padding — the ranking rule of reference ticket.py:266."""
    from hypothesis import given, settings, strategies as hs
    from video_query_algorithms_b200.sharded import merge_payloads_host, pack_payload, unpack_payload
    from video_query_algorithms_b200.store import merge_topk_batch

    @settings(max_examples=150, deadline=None)
    @given(hs.integers(1, 6), hs.integers(1, 5), hs.integers(1, 24), hs.integers(0, 2 ** 31), hs.booleans())
    def prop(n_lists, n_q, k, seed, ranked):
        rng = np.random.default_rng(seed)
        levels = rng.random(max(2, k // 2)).astype(np.float32)                 # few distinct scores: many ties
        rows = np.full((n_lists, n_q, k), -1, np.int64)
        sc_ = np.full((n_lists, n_q, k), -np.inf, np.float32)
        for q in range(n_q):
            ids = rng.permutation(n_lists * k * 3)                             # distinct global rows per query
            for l in range(n_lists):
                n = int(rng.integers(0, k + 1))
                r = ids[l * k:l * k + n].astype(np.int64)
                s_ = rng.choice(levels, n)
                if ranked:
                    o = np.lexsort((r, -s_.astype(np.float64)))
                    r, s_ = r[o], s_[o]
                rows[l, q, :n], sc_[l, q, :n] = r, s_
        got_r, got_s = merge_topk_batch(rows, sc_)
        for q in range(n_q):
            m = rows[:, q].reshape(-1) >= 0
            ar, as_ = rows[:, q].reshape(-1)[m], sc_[:, q].reshape(-1)[m]
            o = np.lexsort((ar, -as_.astype(np.float64)))[:k]
            n = len(o)
            assert np.array_equal(got_r[q, :n], ar[o]) and np.array_equal(got_s[q, :n], as_[o])
            assert np.all(got_r[q, n:] == -1) and np.all(np.isneginf(got_s[q, n:]))
        # the single-query merge of per-rank payloads (counts summed) must agree with the batched one on query 0
        if ranked:
            pay = np.stack([pack_payload([3, 2, 1, int((rows[l, 0] >= 0).sum())], rows[l, 0][rows[l, 0] >= 0],
                                         sc_[l, 0][rows[l, 0] >= 0], k) for l in range(n_lists)])
            counts, r1, s1 = unpack_payload(merge_payloads_host(pay, n_lists, k), k)
            n = int((got_r[0] >= 0).sum())
            assert list(counts[:3]) == [3 * n_lists, 2 * n_lists, n_lists] and counts[3] == n
            assert np.array_equal(r1, got_r[0, :n]) and np.array_equal(s1, got_s[0, :n])

    prop()


def test_native_csv_reader_agrees_with_float_on_arbitrary_decimals(tmp_path, built_lib):
    """Property test (hypothesis): whatever finite double, in whatever of the spellings Python prints or accepts
    (repr, %.17g, %e, %f, fixed with many digits, with padding blanks), the library parses the cell to the double
    float() gives — bit for bit — because the store must hold exactly what the reference would have computed with."""
    from hypothesis import given, settings, strategies as hs
    from video_query_algorithms_b200 import ingest
    spell = [repr, lambda x: "%.17g" % x, lambda x: "%e" % x, lambda x: "%.30f" % x if abs(x) < 1e15 else repr(x),
             lambda x: " " + repr(x), lambda x: "%.3g" % x, lambda x: "%+.12E" % x]
    path = tmp_path / "flow_global_pool_features.csv"

    @settings(max_examples=60, deadline=None)
    @given(hs.lists(hs.floats(allow_nan=False, allow_infinity=False, width=64), min_size=8, max_size=8), hs.integers(0, 10 ** 6))
    def prop(values, salt):
        cells = [[spell[(i + j + salt) % len(spell)](v) for j, v in enumerate(values)] for i in range(5)]
        with open(path, "w") as f:
            f.write("video =v, video url =../x/, CNN stream =warped_optical_flow, feature blob =global_pool, caffe model =/m.caffemodel\n")
            for i, row in enumerate(cells):
                f.write(",".join([str(i)] + row) + "\n")
        _, clips, want = _python_csv(path)
        got = ingest.read_feature_csv(str(path), n_threads=1 + salt % 3)
        assert np.array_equal(got["clip_numbers"], clips)
        assert np.array_equal(got["features"].view(np.uint64), want.view(np.uint64))

    prop()


def test_vectorised_draws_match_python_for_arbitrary_parameters():
    """Property test (hypothesis) of _rng.choices_range / set_order: any seed (int or str), population size, draw
    count and number of consecutive calls give random.choices' indices, leave the generator in random.choices' state,
    and `set_order` gives list(set(...))'s order — the reference's resampling (target_clip.py:297-309) exactly."""
    from hypothesis import given, settings, strategies as hs
    from video_query_algorithms_b200._rng import choices_range, set_order

    @settings(max_examples=120, deadline=None)
    @given(hs.one_of(hs.integers(0, 2 ** 64), hs.text(min_size=1, max_size=12)), hs.integers(1, 3000), hs.integers(1, 400),
           hs.integers(1, 4))
    def prop(seed, n, k, repeats):
        a, b = random.Random(seed), random.Random(seed)
        want = [a.choices(range(n), k=k) for _ in range(repeats)]
        got = choices_range(b, n, k, repeats=repeats)
        assert np.array_equal(np.asarray(want).reshape(got.shape), got) and a.getstate() == b.getstate()
        assert a.random() == b.random()
        assert set_order(np.asarray(want[0]), n).tolist() == list(set(want[0]))

    prop()


def test_row_order_on_ragged_search_sets_is_the_reference_scores_order():
    """A search set whose clips lack arbitrary (stream, split) records: the store's row order must be the insertion order
    of the reference's `scores` dict (its seeded sampling walks that order, ticket.py:326-341), which the reference builds
    while walking the first stream split by split (ticket.py:146-160) — not simply the order of first appearance.  Spelled
    out on a small case and, where the reference tree is present, against `Ticket.compute_similarities` itself."""
    import types
    from video_query_algorithms_b200.store import pack_feature_rows
    streams = ("rgb", "warped_optical_flow")

    def response(rng, clips, splits, lacks):
        recs = []
        for p in splits:                                         # load_db.py order: split dirs, then one file per stream
            for s in streams:
                for c in clips:
                    if (c, s, p) not in lacks:
                        recs.append({"dnn_stream_id": s, "dnn_stream_split": p, "name": "global_pool",
                                     "feature_vector": rng.random(4).tolist(), "video_clip_id": c})
        return recs

    rng = np.random.default_rng(2)
    small = response(rng, (1, 2, 3, 4), (1, 2), {(3, "rgb", 1), (1, "rgb", 1), (1, "warped_optical_flow", 1)})
    assert pack_feature_rows(small, streams, "global_pool")[0] == [2, 4, 1, 3]       # by first appearance it would be 2, 4, 3, 1
    if not os.path.isdir("/root/reference/src"):
        return
    import test_flow_cpu
    test_flow_cpu.load_reference_driver()
    import models.ticket as rticket
    for trial in range(20):
        clips, splits = list(range(10, 40)), (1, 2, 3)
        lacks = set()
        for c in clips:
            for s in streams:
                drop = [p for p in splits if rng.random() < 0.3]
                for p in drop[:2]:                                # every clip keeps at least one split of every stream
                    lacks.add((c, s, p))
        recs = response(rng, clips, splits, lacks)
        t = object.__new__(rticket.Ticket)
        t.search_set, t._request = 1, (lambda action, params, recs=recs: recs)
        t.target = types.SimpleNamespace(splits=set(splits), target_features={s: {p: [1.0, 0.0, 0.0, 0.0] for p in splits} for s in streams})
        t.compute_similarities(types.SimpleNamespace(streams=streams, feature_name="global_pool"))
        order, got_splits, X, present = pack_feature_rows(recs, streams, "global_pool")
        assert order == list(t.similarities), trial
        # and the per-clip split counts the reference records next to each similarity (ticket.py:160)
        for r, c in enumerate(order):
            assert [int(present[r, si].sum()) for si in range(2)] == [t.similarities[c][s][1] for s in streams]
            assert [float(X[r, si, :, 0].sum() / present[r, si].sum()) for si in range(2)] == \
                pytest.approx([t.similarities[c][s][0] for s in streams], rel=1e-6)


# ---------------------------------------------------------------------------- round 2: native host pieces
def test_native_sample_equals_random_sample_and_leaves_the_same_state(built_lib):
    """vq_mt_sample_range (csrc/vq_rng.cu) against CPython's own random.sample(range(n), k) on the same generator state:
    same picks, same state afterwards (so every later draw of the tick is unchanged) — across the pool / taken-set
    switch of sample() (n <= 21 + 4**ceil(log4(3k))), bit lengths of n around powers of two (rejection loop of
    randbelow), wide populations (two 32-bit outputs per draw) and the finalize round's full permutation."""
    from video_query_algorithms_b200 import _rng
    cases = [(100, 100), (64, 64), (1000, 64), (50_000, 50_000), (90_000, 70), (5000, 1000), (70, 65), (2 ** 33, 100),
             (200_000, 199_999), (341, 64), (85, 64), (86, 64), (2 ** 20, 77), (2 ** 20 + 1, 77), (2 ** 32, 65), (2 ** 32 + 5, 65),
             (277, 64), (278, 64), (1045, 256), (1046, 256)]
    for n, k in cases:
        random.seed("73459912436")
        random.random()
        want = random.sample(range(n), k)
        state, nxt = random.getstate(), random.random()
        random.seed("73459912436")
        random.random()
        got = _rng.sample_range(random, n, k)
        assert got.tolist() == want, (n, k)
        assert random.getstate() == state and random.random() == nxt, (n, k)
    # small draws take Python's own path; a generator that is not CPython's MT19937 too
    random.seed(5)
    a = random.sample(range(1000), 10)
    random.seed(5)
    assert _rng.sample_range(random, 1000, 10).tolist() == a
    with pytest.raises(ValueError):
        _rng.sample_range(random, 10, 100)


def test_dict_order_follows_the_reference_for_any_target_split_order(built_lib):
    """FeatureStore.dict_order against a literal emulation of the reference's dict filling (ticket.py:142-160,374-381) on
    random ragged responses — shuffled records, duplicate slots, targets that walk the splits in any order and lack some:
    the order of the clips in the reference's `scores` dict, which its seeded sampling draws from."""
    from video_query_algorithms_b200 import store as ps
    rng = np.random.default_rng(1)
    S = ("a", "b")
    n_checked = n_moved = 0
    for trial in range(300):
        n, splits = int(rng.integers(3, 12)), [1, 2, 3][:int(rng.integers(1, 4))]
        recs = [{"dnn_stream_id": s, "dnn_stream_split": p, "name": "g", "video_clip_id": 100 + c, "feature_vector": [float(c), 1.0, 2.0, 3.0]}
                for c in range(n) for s in S for p in splits if rng.random() < 0.7]
        rng.shuffle(recs)
        if recs and rng.random() < 0.3:
            recs.append(dict(recs[0]))                          # a slot listed twice: the dict keeps its first place
        idx, _ = ps.index_feature_rows(recs, S, "g")
        if idx.n_rows == 0:
            continue
        st = ps.FeatureStore.__new__(ps.FeatureStore)
        st.streams, st.splits, st.n_rows, st.slot_pos = S, idx.splits, idx.n_rows, idx.slot_positions()
        t_splits = [int(p) for p in rng.permutation(idx.splits) if rng.random() < 0.8] or [int(idx.splits[0])]
        tf = {s: {p: [0.0] for p in t_splits} for s in S}
        cand = {s: {p: {} for p in t_splits} for s in S}
        for r in recs:
            if r["dnn_stream_split"] in cand[r["dnn_stream_id"]]:
                cand[r["dnn_stream_id"]][r["dnn_stream_split"]][r["video_clip_id"]] = 1
        want = {}
        for s in S:
            sims = {}
            for p in tf[s]:
                for c in cand[s][p]:
                    sims[c] = 1
            for c in sims:
                want.setdefault(c, 1)
        place = st.dict_order(tf)
        ids = np.asarray(idx.order)
        got = [int(c) for c in (ids if place is None else ids[np.argsort(place)]) if c in want]
        assert got == list(want), (trial, t_splits)
        n_checked += 1
        n_moved += place is not None
    assert n_checked > 250 and n_moved > 50
    # complete data listed consistently: the row order serves every target, nothing is kept
    recs = [{"dnn_stream_id": s, "dnn_stream_split": p, "name": "g", "video_clip_id": c, "feature_vector": [1.0, 2.0, 3.0, 4.0]}
            for p in (1, 2) for s in S for c in range(20)]
    assert ps.index_feature_rows(recs, S, "g")[0].slot_positions() is None

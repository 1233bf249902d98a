"""End-to-end replay of every golden scenario through the PRODUCT's host path on CPU: `compute_matches` -> Ticket /
TargetClip / Hyperparameter -> store, against the in-memory fake API, with the device behind the store played by
`OracleStore` — a test double that answers the FeatureStore calls with the float64 oracle (same interface, same result
conventions: fp32 scores, lists in database order, float64 labelled similarities).  It is test infrastructure: the
product has no such path and fails without its CUDA library.  What it checks is everything around the kernels —
store construction from API records, lazy score views, the weight update, target bootstrapping, selection, forced
clips, persistence, the final report — for all seven scenarios (the GPU replays are tests/test_gpu_parity.py; the same
double is the CPU side of tests/test_gpu_differential.py)."""
import os
import random
import types

import numpy as np
import pytest

from oracle import bootstrap as ob
from oracle import scoring as sc
from scenarios import ORACLE_ONLY_SCENARIOS, SCENARIOS, Scenario

os.environ.setdefault("COMPUTE_EPS", ".000003")
EPS = 3e-6


def make_oracle_store_class():
    from video_query_algorithms_b200 import store as ps

    class OracleStore(ps.FeatureStore):
        def __init__(self, n_rows, streams, splits, dim=1024, devices=None, clip_ids=None, first_global_row=0):
            self.streams, self.splits = tuple(streams), [int(p) for p in splits]
            self.dim, self.n_rows, self.first_global_row = int(dim), int(n_rows), int(first_global_row)
            self.row_shape = (len(self.streams), len(self.splits), self.dim)
            self.shards = [types.SimpleNamespace(device=0, first=self.first_global_row, n_rows=self.n_rows, handle=None,
                                                 close=lambda: None)]
            self.clip_ids, self._row_of, self._row_memo, self.present, self.last = None, None, {}, None, None
            if clip_ids is not None:
                self.set_clip_ids(clip_ids)
            self.X = np.zeros((self.n_rows,) + self.row_shape, np.float32)
            self.n_scans = 0

        # ---- what the library would hold
        def upload(self, first_row, rows):
            rows = np.asarray(rows, np.float32).reshape((-1,) + self.row_shape)
            self.X[first_row:first_row + len(rows)] = rows

        def append(self, rows, clip_ids=None, present=None):
            """The product's own bookkeeping (ids, row counts, split presence) with the one library call answered here."""
            rows32 = np.asarray(rows, np.float32).reshape((-1,) + self.row_shape)
            real = ps.lib
            ps.lib = lambda: types.SimpleNamespace(vq_store_append=lambda handle, n, p: 0)
            try:
                super().append(rows, clip_ids=clip_ids, present=present)
            finally:
                ps.lib = real
            self.X = np.concatenate([self.X, rows32])

        def set_present(self, present):
            present = np.asarray(present, bool)
            self.present = None if present.all() else present

        def _alloc_staging(self, n_floats):
            return np.empty(int(n_floats), np.float32)

        def _free_staging(self, a):
            pass

        def _upload_async(self, first_row, flat):
            self.upload(first_row, np.array(flat))

        def _sync_uploads(self):
            pass

        def _sync_split_weights_for_target(self, have):
            """no device table to load; the product's refusal of a job in which some clip shares no split with the target
            for some stream (the reference raises KeyError there, ticket.py:177) is kept"""
            present = np.ones((self.n_rows,) + self.row_shape[:2], bool) if self.present is None else self.present
            if self.n_rows and ((present & have[None]).sum(axis=2) == 0).any():
                raise ps.VQError("a clip has no feature row at all for one stream; the reference raises KeyError "
                                 "for such a search set (ticket.py:177)")

        def close(self):
            pass

        # ---- the scan and its results (conventions of vq_scan: fp32 scores, comparisons on (double)score)
        def scan(self, target_features, weights, threshold, lower_limit, eps, topk=0, want_sims=False, lists=True, packed=None):
            T, have = self.pack_target(target_features, np.float32)
            self._sync_split_weights_for_target(have)
            present = np.ones((self.n_rows,) + self.row_shape[:2], bool) if self.present is None else self.present
            sims, _ = sc.similarities(self.X, T.astype(np.float64), present & have[None])
            w = [weights[s] for s in self.streams] if isinstance(weights, dict) else list(weights)
            self._sims32 = sims.astype(np.float32)
            self._scores32 = sc.scores(sims, w).astype(np.float32)
            s = self._scores32.astype(np.float64)
            self._l = {"matches": np.flatnonzero(s >= threshold), "near_misses": np.flatnonzero((s >= lower_limit) & (s < threshold)),
                       "ties": np.flatnonzero((np.abs(s - threshold) < eps) | (np.abs(s - lower_limit) < eps))}
            self._k, self._lists, self.n_scans = int(topk), bool(lists), self.n_scans + 1
            self.last = ps.ScanResult(len(self._l["matches"]), len(self._l["near_misses"]), len(self._l["ties"]),
                                      min(self._k, self.n_rows), 0.0)
            return self.last

        def _list(self, which):
            rows = self._l[which]
            return (self.first_global_row + rows).astype(np.int64), self._scores32[rows]

        def matches(self, copy=True):
            assert self._lists, "whole lists read after a lists=False scan"
            return self._list("matches")

        def near_misses(self, copy=True):
            assert self._lists, "whole lists read after a lists=False scan"
            return self._list("near_misses")

        def ties(self, copy=True):
            return self._list("ties")

        def gather(self, which, positions):
            rows, scores = self._list(which)
            pos = np.asarray(positions, np.int64)
            return rows[pos], scores[pos]

        def gather_many(self, requests):
            return [self.gather(w, p) for w, p in requests]

        def near_best(self):
            rows, scores = self._list("near_misses")
            if not len(rows):
                return None
            j = int(np.argmax(scores))
            return j, int(rows[j]), float(scores[j])

        def topk(self):
            rows = sc.topk_stable(self._scores32, self._k)
            return (self.first_global_row + rows).astype(np.int64), self._scores32[rows]

        def ranked(self, which="matches"):
            rows, scores = self._list(which)
            o = np.argsort(-scores, kind="stable")
            return rows[o], scores[o]

        def scores(self):
            return self._scores32

        def scores_at(self, global_rows):
            return self._scores32[np.asarray(global_rows, np.int64) - self.first_global_row]

        def rank_list(self, which, place):
            _, scores = self._list(which)
            order = np.lexsort((np.asarray(place), -scores.astype(np.float64)))
            return np.asarray(place)[order], scores[order]

        def sims(self):
            return self._sims32

        # ---- labelled subset (float64, like K4 / K6)
        def labelled_sims(self, target_features, global_rows):
            T, have = self.pack_target(target_features, np.float64)
            self._sync_split_weights_for_target(have)
            rows = np.asarray(global_rows, np.int64) - self.first_global_row
            present = np.ones((len(rows),) + self.row_shape[:2], bool) if self.present is None else self.present[rows]
            return sc.similarities(self.X[rows], T, present & have[None])[0]

        def bootstrap_target(self, valid_rows, invalid_rows, mu, slots=None):
            X = self.X.astype(np.float64)
            v = np.asarray(valid_rows, np.int64) - self.first_global_row
            iv = np.asarray(invalid_rows if invalid_rows is not None else [], np.int64) - self.first_global_row
            out = np.empty(self.row_shape, np.float64)
            for s in range(self.row_shape[0]):
                for p in range(self.row_shape[1]):
                    try:
                        with np.errstate(all="ignore"):
                            out[s, p] = ob.solve_valid_invalid(X[v, s, p], X[iv, s, p], mu) if len(iv) else ob.solve_valid(X[v, s, p])
                    except np.linalg.LinAlgError:            # the kernel's LU has no singularity check: that slot is not finite
                        out[s, p] = np.nan
            return out

    return OracleStore


@pytest.fixture
def cpu_product(monkeypatch):
    """The product package with the device side of the store, the loss grid and the one-score fetch played by the oracle."""
    import video_query_algorithms_b200 as vq
    from video_query_algorithms_b200 import store as ps
    from video_query_algorithms_b200 import ticket as pt
    monkeypatch.setattr(ps, "FeatureStore", make_oracle_store_class())
    monkeypatch.setattr(ps, "loss_grid", lambda sims, labels, wg, tg, ballast, replicates=None, device=0:
                        sc.loss_grid(np.asarray(sims, np.float64), np.asarray(labels, bool), wg, tg, ballast)[None])
    ps.invalidate()
    yield vq
    ps._REGISTRY.clear()


def close(a, b, rel=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= rel * np.maximum(np.abs(b), 0.05)))


@pytest.mark.parametrize("name", SCENARIOS + ORACLE_ONLY_SCENARIOS)
def test_product_host_path_replays_reference_rounds_end_to_end(cpu_product, name, tmp_path, monkeypatch):
    vq = cpu_product
    from fake_api import FakeRepository
    scn = Scenario(name)
    api, qid = scn.build_api()
    (tmp_path / "work").mkdir()
    monkeypatch.chdir(tmp_path / "work")                     # the final report goes to ../final_reports/
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
        assert np.array_equal(t.feature_store().clip_ids, scn.arr(i, "clip_order"))
        assert close(t.scores.array(), scn.arr(i, "scores"))
        T = np.array([[t.target.target_features[s][p] for p in r["splits"]] for s in scn.streams])
        assert np.abs(T - scn.arr(i, "target")).max() <= 1e-6 * np.abs(scn.arr(i, "target")).max()
        if "r%d_losses" % i in scn.arrays.files:
            assert np.abs(hp.losses - scn.arr(i, "losses")).max() < 1e-6
        got_ids, want_ids = list(t.matches), [k for k, _ in r["selected"]]
        if not t.tie_band:
            assert got_ids == want_ids
            assert close([t.matches[k] for k in got_ids], [v for _, v in r["selected"]])
        assert list(t.ranked[0]) == [int(scn.clip_ids[j]) for j in sc.topk_stable(t.scores.array(), len(t.ranked[0]))]
        # persisted: one match entity per selected clip on the new query result, with the user's earlier label
        res_id = max(api.query_results)
        stored = [m for m in api.matches.values() if m["query_result"] == res_id]
        assert [m["video_clip"] for m in stored] == got_ids
    # the store is built once per search set and reused by later ticks (the reference re-downloads every job)
    assert api.calls.count(("search-sets", "features")) == 1
    if scn.rounds[-1]["kind"] == "finalize":
        assert len(api.uploaded_reports) == 1
        body, ref = api.uploaded_reports[0].splitlines(), scn.meta["final_report"].splitlines()
        assert len(body) == len(ref)
        n_sel = len(scn.rounds[-1]["selected"])
        # lines that carry float64-vs-fp32 digits (checked above) or names the two harnesses chose differently
        skip = ("number of reviews", "min score", "stream weights", "Search Set queried", "Query:")
        for a, b in zip(body[:-n_sel], ref[:-n_sel]):
            assert a == b or a.startswith(skip), (a, b)
        cols = lambda ln: ln.split(",")
        ref_score = {str(k): v for k, v in scn.rounds[-1]["selected"]}
        assert sorted(cols(ln)[4] for ln in body[-n_sel:]) == sorted(cols(ln)[4] for ln in ref[-n_sel:]) or tickets[-1].tie_band
        for a, b in zip(body[-n_sel:], ref[-n_sel:]):
            ca, cb = cols(a), cols(b)
            if ca[4] == cb[4]:
                assert ca[:5] == cb[:5] and ca[6:] == cb[6:] and float(ca[5]) == pytest.approx(float(cb[5]), rel=1e-5)
            else:                                            # equal-score clips may swap within COMPUTE_EPS
                assert abs(ref_score[ca[4]] - ref_score[cb[4]]) < EPS


def test_job_error_states_follow_the_reference_on_cpu(cpu_product, tmp_path, monkeypatch):
    """compute_matches.py:47-52,92-94 through the product's host path: a fatal query error -> state 5 + note; a
    recoverable one -> note and the round goes on; an empty selection -> state 5 'No matches were found'."""
    vq = cpu_product
    from fake_api import FakeRepository
    from video_query_algorithms_b200 import store as ps
    scn = Scenario("A_brooklyn_bagging")
    monkeypatch.chdir(tmp_path)
    factory = lambda api: (lambda job, url: vq.Ticket(job, url, client=api.client(), devices=[0]))
    api, qid = scn.build_api()                               # (1) reference time outside the video: no ref clip
    api.queries[qid]["ref_clip_id"] = None
    api.request(qid, "new")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=factory(api))
    assert api.queries[qid]["process_state"] == 5 and "Fatal Error" in api.queries[qid]["notes"]
    ps.invalidate()                                          # (2) revise, dynamic target adjustment, nothing confirmed
    api, qid = scn.build_api()
    api.request(qid, "new")
    random.seed(a=scn.seed)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=factory(api))
    api.label_latest_round(qid, lambda m: False)
    api.request(qid, "revise")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=factory(api))
    assert api.queries[qid]["process_state"] == 4
    assert "Changing dynamic target adjustment to False" in api.queries[qid]["notes"]
    ps.invalidate()                                          # (3) nothing selectable
    api, qid = scn.build_api()
    ss = api.queries[qid]["search_set_to_query"]
    api.search_sets[ss]["clip_ids"] = [c for c in api.search_sets[ss]["clip_ids"] if c != api.queries[qid]["ref_clip_id"]]
    api.request(qid, "new")
    hp = dict(scn.hp(), default_threshold=5.0, near_miss_default=0.0)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**hp), ticket_factory=factory(api))
    assert api.queries[qid]["process_state"] == 5 and "No matches were found" in api.queries[qid]["notes"]
    # (4) the search set grows between ticks (load_db.py adds a video): the resident store appends the new clips only
    ps.invalidate()
    both = Scenario("B_both_simple_mu")
    api, qid = both.build_api()
    ss = api.queries[qid]["search_set_to_query"]
    all_ids = list(api.search_sets[ss]["clip_ids"])
    api.search_sets[ss]["clip_ids"] = all_ids[:120]
    api.request(qid, "new")
    random.seed(a=both.seed)
    made = []
    fac = lambda job, url: made.append(vq.Ticket(job, url, client=api.client(), devices=[0])) or made[-1]
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)
    st = made[-1].feature_store()
    assert st.n_rows == 120
    # load_db.py adds clips: the next job sees them without anybody telling the store (the search-set record changed)
    api.search_sets[ss]["clip_ids"] = all_ids[:150]
    api.request(qid, "new")
    random.seed(a=both.seed)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)
    assert made[-1].feature_store() is st and st.n_rows == 150 and list(st.clip_ids) == all_ids[:150] and st.X.shape[0] == 150
    assert api.calls.count(("search-sets", "features")) == 2
    api.request(qid, "new")                                   # nothing changed: no second download
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)
    assert api.calls.count(("search-sets", "features")) == 2 and made[-1].feature_store() is st
    # an API whose search-set record does not change: VQ_STORE_FRESHNESS=always re-reads per job like the reference
    api.search_set_record = "bare"
    api.request(qid, "new")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)   # record changed shape: one re-read
    n_calls = api.calls.count(("search-sets", "features"))
    api.search_sets[ss]["clip_ids"] = all_ids[:170]
    api.request(qid, "new")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)
    assert st.n_rows == 150 and api.calls.count(("search-sets", "features")) == n_calls       # the probe cannot see it ...
    monkeypatch.setenv("VQ_STORE_FRESHNESS", "always")
    api.request(qid, "new")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)
    assert made[-1].feature_store() is st and st.n_rows == 170                                  # ... `always` does
    monkeypatch.delenv("VQ_STORE_FRESHNESS")
    # ... and so does a job that names a clip the store does not hold (a label on a new clip)
    api.search_sets[ss]["clip_ids"] = all_ids[:180]
    api.request(qid, "new")
    job_holder = {}
    real_status = FakeRepository.get_status
    def with_label(self):
        status = real_status(self)
        status["new"]["user_matches"] = {str(all_ids[175]): True}
        return status
    monkeypatch.setattr(FakeRepository, "get_status", with_label)
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)
    monkeypatch.setattr(FakeRepository, "get_status", real_status)
    assert st.n_rows == 180 and all_ids[175] in made[-1].matches
    # clips removed (or reordered): the store is rebuilt from the response
    api.search_set_record = "members"
    api.search_sets[ss]["clip_ids"] = all_ids[5:100]
    api.request(qid, "new")
    vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**both.hp()), ticket_factory=fac)
    st2 = made[-1].feature_store()
    assert st2 is not st and list(st2.clip_ids) == all_ids[5:100]
    st = st2
    api.search_sets[ss]["clip_ids"] = all_ids
    rows = api.client().action(None, ["search-sets", "features"], params={"id": ss})
    assert st.append_feature_rows(rows, both.hp()["feature_name"]) == len(all_ids) - 95
    assert sorted(st.clip_ids) == sorted(all_ids) and st.X.shape[0] == len(all_ids)


def test_a_full_device_evicts_the_least_recently_used_store_and_retries(cpu_product, tmp_path, monkeypatch):
    """A broker serving several search sets: building a store when the devices are full of other search sets' stores
    (the library's allocation error says "out of memory") closes the least recently used of those and tries again; the
    evicted search set is rebuilt from the API the next time a job names it.  An allocation failure with nothing left
    to evict, and any other error, reach the caller (the broker logs it, broker.py:88-89)."""
    vq = cpu_product
    from fake_api import FakeRepository
    from video_query_algorithms_b200 import store as ps
    monkeypatch.chdir(tmp_path)
    scn = Scenario("A_brooklyn_bagging")
    apis = []
    for i in range(3):                                        # three brokers' worth of search sets: distinct API urls -> distinct keys
        api, qid = scn.build_api()
        api.url = "http://api-%d/" % i
        apis.append((api, qid))

    def run(i):
        api, qid = apis[i]
        api.request(qid, "new")
        random.seed(a=scn.seed)
        made = []
        fac = lambda job, url: made.append(vq.Ticket(job, "http://api-%d/" % i, client=api.client(), devices=[0])) or made[-1]
        vq.compute_matches(FakeRepository(api), vq.Hyperparameter(**scn.hp()), ticket_factory=fac)
        assert api.queries[qid]["process_state"] == 4
        return made[-1].feature_store()

    st0, st1 = run(0), run(1)
    assert len(ps._REGISTRY) == 2
    run(0)                                                    # search set 0 was used last: search set 1 is the LRU one
    real = ps.FeatureStore.from_feature_rows.__func__
    budget = {"fail": 1}

    def flaky(cls, *a, **kw):
        if budget["fail"] > 0:
            budget["fail"] -= 1
            raise ps.VQError("vq_store_create failed (-3): vq_store: cudaMalloc(8192000000 bytes) for s->rows -> out of memory")
        return real(cls, *a, **kw)

    monkeypatch.setattr(ps.FeatureStore, "from_feature_rows", classmethod(flaky))
    closed = []
    monkeypatch.setattr(type(st1), "close", lambda self: closed.append(self))
    st2 = run(2)                                              # first attempt fails -> evict search set 1 -> retry succeeds
    assert closed == [st1] and st2.n_rows == st0.n_rows
    keys = {k[0] for k in ps._REGISTRY}
    assert keys == {"http://api-0/", "http://api-2/"}
    n_calls = apis[1][0].calls.count(("search-sets", "features"))
    assert run(1) is not st1 and apis[1][0].calls.count(("search-sets", "features")) == n_calls + 1   # rebuilt on demand
    # nothing left to evict: the error reaches the caller
    ps.invalidate()
    budget["fail"] = 1
    with pytest.raises(ps.VQError, match="out of memory"):
        run(0)
    # another error is not an eviction matter
    run(1)
    monkeypatch.setattr(ps.FeatureStore, "from_feature_rows",
                        classmethod(lambda cls, *a, **kw: (_ for _ in ()).throw(ps.VQError("vq_store_create: dim must be a multiple of 4"))))
    with pytest.raises(ps.VQError, match="multiple of 4"):
        run(2)
    assert len(ps._REGISTRY) == 1

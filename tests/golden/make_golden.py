"""Generate golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Outputs (committed): tests/golden/fixture_*.npz (inputs) and tests/golden/scn_*.npz/.json (outputs).

The reference's modules are imported in place from /root/reference/src with the three shims of
SURVEY.md §8(c): a stub `coreapi`, `authenticate` patched out, and `random.sample` accepting
Set populations by tuple()-ing them (what Python 3.7, the reference's interpreter, did; on
>= 3.11 `random.sample(dict.items(), k)` at ticket.py:333,341 raises TypeError).
Nothing of the reference is copied: we call its public `compute_matches` and record what its
objects hold afterwards.
"""
from __future__ import annotations

import collections.abc
import csv
import hashlib
import json
import os
import random
import time
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RANDOM_SEED = "73459912436"      # reference Dockerfile:15
COMPUTE_EPS = ".000003"


def install_shims(fake_api_holder):
    os.environ["COMPUTE_EPS"] = COMPUTE_EPS
    os.environ["RANDOM_SEED"] = RANDOM_SEED
    os.environ.setdefault("API_CLIENT_USERNAME", "x")
    os.environ.setdefault("API_CLIENT_PASSWORD", "x")
    coreapi = types.ModuleType("coreapi")

    class Client:
        def __init__(self, auth=None):
            self._c = fake_api_holder["api"].client()

        def get(self, url):
            return self._c.get(url)

        def action(self, schema, keys, params=None, encoding=None, **kw):
            return self._c.action(schema, keys, params=params, encoding=encoding)

    coreapi.Client = Client
    coreapi.auth = types.SimpleNamespace(TokenAuthentication=lambda **kw: None)
    sys.modules["coreapi"] = coreapi
    sys.path.insert(0, os.path.join(REF, "src"))
    import warnings
    warnings.simplefilter("ignore", SyntaxWarning)
    import models.ticket as rticket
    import api.api_repository as rrepo
    rticket.authenticate = lambda url=None: None
    rrepo.authenticate = lambda url=None: None
    _sample = random.sample

    def sample(population, k, **kw):
        if isinstance(population, collections.abc.Set):
            population = tuple(population)
        return _sample(population, k, **kw)

    random.sample = sample


def read_fixture(video_dir):
    """-> clip numbers, {stream: {split: ndarray[n, 1024]}} from the reference's CSV tree
    (layout: reference src/api/api_load_records.py:41-58)."""
    arrays = {}
    clip_numbers = None
    for split_dir in sorted(os.listdir(video_dir)):
        split = int(split_dir[-1])
        for fn in sorted(os.listdir(os.path.join(video_dir, split_dir))):
            if not fn.endswith(".csv"):
                continue
            with open(os.path.join(video_dir, split_dir, fn)) as f:
                rd = csv.reader(f)
                header = next(rd)
                stream = header[2].split("=")[-1]
                rows = [(int(r[0]), [float(x) for x in r[1:]]) for r in rd]
            nums = [r[0] for r in rows]
            if clip_numbers is None:
                clip_numbers = nums
            assert nums == clip_numbers
            arrays.setdefault(stream, {})[split] = np.array([r[1] for r in rows], np.float64)
    return clip_numbers, arrays


def rng_digest():
    return hashlib.sha256(repr(random.getstate()).encode()).hexdigest()[:16]


class Recorder:
    """Wraps reference methods to copy out their results; does not change behaviour."""

    def __init__(self):
        import models.ticket as rticket
        import models.hyperparameter as rhp
        import models.target_clip as rtc
        self.rounds = []
        self.cur = None
        rec = self
        T, H, C = rticket.Ticket, rhp.Hyperparameter, rtc.TargetClip
        o_sim, o_sel, o_opt = T.compute_similarities, T.select_clips_to_review, H.optimize_weights
        o_get, o_low = C.get_target_features, T.lowest_scoring_user_match

        def get_target_features(self_):
            rec.cur = {"rng_before": rng_digest()}
            rec.rounds.append(rec.cur)
            r = o_get(self_)
            rec.cur["target"] = {s: {int(p): list(map(float, v)) for p, v in by.items()}
                                 for s, by in self_.target_features.items()}
            rec.cur["rng_after_target"] = rng_digest()
            return r

        def compute_similarities(self_, hp):
            r = o_sim(self_, hp)
            rec.cur["clip_order"] = [int(c) for c in self_.similarities]
            rec.cur["sims"] = {s: [float(self_.similarities[c][s][0]) for c in self_.similarities]
                               for s in hp.streams}
            rec.cur["nsplits"] = {s: [int(self_.similarities[c][s][1]) for c in self_.similarities]
                                  for s in hp.streams}
            return r

        def optimize_weights(self_, ticket):
            captured = {}
            o_argmin = np.argmin

            def argmin(a, axis=None, **kw):
                captured["losses"] = np.array(a, copy=True)
                return o_argmin(a, axis=axis, **kw)

            np.argmin = argmin
            try:
                r = o_opt(self_, ticket)
            finally:
                np.argmin = o_argmin
            rec.cur["losses"] = captured["losses"]
            rec.cur["match_status_input"] = [
                {"video_clip": m["video_clip"], "user_match": m["user_match"],
                 "is_match": m["is_match"]} for m in ticket.matches]
            return r

        def lowest(self_):
            r = o_low(self_)
            rec.cur["lowest_user_match"] = [float(r[0]), r[1]]
            return r

        def select(self_, threshold=0.8, max_number_matches=20, near_miss=0.5):
            rec.cur["scores"] = [float(self_.scores[c]) for c in self_.scores]
            rec.cur["score_order"] = [int(c) for c in self_.scores]
            rec.cur["select_args"] = [float(threshold), float(max_number_matches), float(near_miss)]
            rec.cur["rng_before_select"] = rng_digest()
            r = o_sel(self_, threshold, max_number_matches, near_miss)
            rec.cur["selected"] = [[int(k), float(v)] for k, v in self_.matches.items()]
            rec.cur["rng_after_select"] = rng_digest()
            rec.cur["user_matches"] = [[k, v] for k, v in self_.user_matches.items()]  # ordered
            rec.cur["ref_clip_id"] = self_.ref_clip_id
            return r

        T.compute_similarities, T.select_clips_to_review = compute_similarities, select
        T.lowest_scoring_user_match = lowest
        H.optimize_weights, C.get_target_features = optimize_weights, get_target_features


ONLY = os.environ.get("VQ_GOLDEN_ONLY")     # e.g. F_shrp2_bagging_mu: write just that scenario, leave the other files alone


def save_scenario(name, rounds, meta):
    if ONLY and name != ONLY:
        return
    if os.environ.get("VQ_GOLDEN_TIMING_ONLY"):
        print("timings after", name, json.dumps(TIMINGS[-len(rounds):]))
        return
    arrays, js = {}, {"meta": meta, "rounds": []}
    for i, r in enumerate(rounds):
        streams = meta["streams"]
        splits = sorted({p for s in r["target"] for p in r["target"][s]})
        arrays["r%d_target" % i] = np.array([[r["target"][s][p] for p in splits] for s in streams])
        arrays["r%d_sims" % i] = np.array([r["sims"][s] for s in streams]).T
        arrays["r%d_nsplits" % i] = np.array([r["nsplits"][s] for s in streams]).T
        arrays["r%d_scores" % i] = np.array(r["scores"])
        arrays["r%d_clip_order" % i] = np.array(r["clip_order"], np.int64)
        if "losses" in r:
            arrays["r%d_losses" % i] = r["losses"]
        assert r["clip_order"] == r["score_order"]
        js["rounds"].append({k: r[k] for k in r if k not in
                             ("target", "sims", "nsplits", "scores", "clip_order", "score_order",
                              "losses")} | {"splits": splits})
    np.savez_compressed(os.path.join(HERE, "scn_%s.npz" % name), **arrays)
    with open(os.path.join(HERE, "scn_%s.json" % name), "w") as f:
        json.dump(js, f, indent=1, sort_keys=True)


TIMINGS = []           # wall time of every reference compute_matches call (VQ_GOLDEN_TIMING_ONLY=1: print, write nothing)


def run_rounds(api, repo_cls, hp_kwargs, plan, rec, qid):
    """plan: list of (kind, label_rule or None applied BEFORE that round)."""
    from models.compute_matches import compute_matches
    from models import Hyperparameter
    out_meta = []
    for kind, rule in plan:
        if rule is not None:
            api.label_latest_round(qid, rule)
        api.request(qid, kind)
        repo = repo_cls("http://fake/")
        hp = Hyperparameter(**hp_kwargs)
        random.seed(a=os.environ["RANDOM_SEED"])       # reference src/broker.py:83-84
        t0 = time.perf_counter()
        compute_matches(repo, hp)
        TIMINGS.append({"query": qid, "kind": kind, "reference_compute_matches_s": time.perf_counter() - t0,
                        "n_clips": len(getattr(api, "clip_ids", [])) or None})
        q = api.queries[qid]
        rec.cur.update({"kind": kind, "weights": [float(hp.weights[s]) for s in hp.streams],
                        "threshold": float(hp.threshold), "process_state": q["process_state"],
                        "notes": q["notes"]})
        out_meta.append(kind)
    return out_meta


def main():
    from fake_api import FakeAPI
    holder = {}
    install_shims(holder)
    # The reference's own APIRepository cannot serve revise jobs on Python >= 3.8: its
    # `_load_plus_convert_split_key` (api_repository.py:72-76) pops and re-adds keys while iterating
    # ("dictionary keys changed during iteration").  It is an HTTP client outside the scoring path,
    # so the harness uses the fake repository, which performs the same int-ification of split keys.
    from fake_api import FakeRepository
    APIRepository = lambda url: FakeRepository(holder["api"], url)
    from oracle import synth
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "work"))
    os.chdir(os.path.join(tmp, "work"))      # reference writes ../final_reports/ (ticket.py:203-205)
    rec = Recorder()

    fx = {}
    for key, rel in (("brooklyn", "stock-video-clips_features/DowntownBrooklynDrive_480p"),
                     ("shrp2", "SHRP2_Forward_clips_features/S06NDS_Sample_120406_1451_00186_Forward")):
        nums, arrays = read_fixture(os.path.join(REF, "data/features", rel))
        fx[key] = (os.path.basename(rel), nums, arrays)
        streams = ("rgb", "warped_optical_flow")
        splits = sorted(arrays["rgb"])
        if not ONLY:
            np.savez_compressed(os.path.join(HERE, "fixture_%s.npz" % key),
                                clip_numbers=np.array(nums, np.int64), splits=np.array(splits, np.int64),
                                X=np.array([[[arrays[s][p][i] for p in splits] for s in streams]
                                            for i in range(len(nums))]))

    streams = ("rgb", "warped_optical_flow")
    broker_defaults = dict(default_weights={"rgb": 1.0, "warped_optical_flow": 1.5},
                           default_threshold=0.8, ballast=0.0, near_miss_default=0.35, mu=0.0,
                           streams=streams, feature_name="global_pool", f_bootstrap=1,
                           f_memory=0.7, bootstrap_type="bagging", nbags=3)   # broker.py:36-59

    # ---- A: brooklyn, broker defaults, new -> revise -> finalize
    api = FakeAPI(); holder["api"] = api
    name, nums, arrays = fx["brooklyn"]
    vid, cids = api.load_feature_arrays(name, nums, arrays)
    ss = api.add_search_set("brooklyn", list(cids.values()))
    qid = api.add_query("qA", vid, cids[10], ss, max_matches=20, dynamic_target_adjustment=True)
    rec.rounds = []
    rule = lambda m: bool(m["score"] >= 0.83)
    run_rounds(api, APIRepository, broker_defaults,
               [("new", None), ("revise", rule), ("finalize", rule)], rec, qid)
    save_scenario("A_brooklyn_bagging", rec.rounds,
                  {"fixture": ["brooklyn"], "ref_clip_number": 10, "hp": _js(broker_defaults),
                   "streams": streams, "label_rule": "score>=0.83", "max_matches": 20,
                   "dynamic_target_adjustment": True, "seed": RANDOM_SEED, "eps": COMPUTE_EPS,
                   "final_report": api.uploaded_reports[-1] if api.uploaded_reports else None})

    # ---- B: both videos in one search set, 'simple' bootstrap on valid+invalid labels with
    #         mu/ballast non-zero and f_bootstrap 0.5 (sampling without replacement), 4 rounds
    api = FakeAPI(page_size=7); holder["api"] = api
    n1, nums1, arr1 = fx["brooklyn"]; n2, nums2, arr2 = fx["shrp2"]
    v1, c1 = api.load_feature_arrays(n1, nums1, arr1)
    v2, c2 = api.load_feature_arrays(n2, nums2, arr2)
    ss = api.add_search_set("both", list(c1.values()) + list(c2.values()))
    qid = api.add_query("qB", v2, c2[nums2[40]], ss, max_matches=30, dynamic_target_adjustment=True)
    hpB = dict(broker_defaults, ballast=0.3, mu=0.3, f_bootstrap=0.5, f_memory=0.7,
               bootstrap_type="simple", near_miss_default=0.5)
    rec.rounds = []
    ruleB = lambda m: bool(m["score"] >= 0.86)
    run_rounds(api, APIRepository, hpB,
               [("new", None), ("revise", ruleB), ("revise", ruleB), ("finalize", ruleB)], rec, qid)
    save_scenario("B_both_simple_mu", rec.rounds,
                  {"fixture": ["brooklyn", "shrp2"], "ref_clip_number": int(nums2[40]),
                   "ref_video": "shrp2", "hp": _js(hpB), "streams": streams, "page_size": 7,
                   "label_rule": "score>=0.86", "max_matches": 30,
                   "dynamic_target_adjustment": True, "seed": RANDOM_SEED, "eps": COMPUTE_EPS,
                   "final_report": api.uploaded_reports[-1] if api.uploaded_reports else None})

    # ---- C: shrp2, 'simple' bootstrap, only positive labels (valid-only branch, with a draw)
    api = FakeAPI(); holder["api"] = api
    v2, c2 = api.load_feature_arrays(n2, nums2, arr2)
    ss = api.add_search_set("shrp2", list(c2.values()))
    qid = api.add_query("qC", v2, c2[nums2[24]], ss, max_matches=16, dynamic_target_adjustment=True)
    hpC = dict(broker_defaults, bootstrap_type="simple", f_bootstrap=0.5, ballast=0.1)
    rec.rounds = []
    ruleC = lambda m: True if m["score"] >= 0.84 else None
    run_rounds(api, APIRepository, hpC,
               [("new", None), ("revise", ruleC), ("finalize", ruleC)], rec, qid)
    save_scenario("C_shrp2_simple", rec.rounds,
                  {"fixture": ["shrp2"], "ref_clip_number": int(nums2[24]), "hp": _js(hpC),
                   "streams": streams, "label_rule": "True if score>=0.84 else None",
                   "max_matches": 16, "dynamic_target_adjustment": True, "seed": RANDOM_SEED,
                   "eps": COMPUTE_EPS,
                   "final_report": api.uploaded_reports[-1] if api.uploaded_reports else None})

    # ---- E: 'partial_update'.  Through compute_matches the reference crashes on this bootstrap
    # type as soon as a previous target exists: avg_new_old_targets (target_clip.py:80-82) leaves
    # ndarrays in target_features and json.dumps (ticket.py:296) raises TypeError.  So the blend is
    # recorded by calling the reference's TargetClip directly on the round-2 job of scenario C's DB.
    import models.ticket as rticket
    import models.target_clip as rtc
    from models import Hyperparameter
    from fake_api import FakeRepository
    api.label_latest_round(qid, lambda m: bool(m["score"] >= 0.9))
    api.request(qid, "finalize")
    job = FakeRepository(api).get_status()["finalize"]
    hpE = dict(broker_defaults, bootstrap_type="partial_update", f_bootstrap=0.6, f_memory=0.7,
               mu=0.2)
    t = rticket.Ticket(job, "http://fake/")
    random.seed(a=os.environ["RANDOM_SEED"])
    tc = rtc.TargetClip(t, Hyperparameter(**hpE))
    rec.rounds = []
    tc.get_target_features()
    tgt = np.array([[np.asarray(tc.target_features[s][p], np.float64) for p in sorted(tc.target_features[s])]
                    for s in streams])
    if not ONLY:
      np.savez_compressed(os.path.join(HERE, "scn_E_partial_update.npz"), target=tgt,
                        previous=np.array([[job["latest_query_result"]["bootstrapped_target"][s][p]
                                            for p in sorted(job["latest_query_result"]["bootstrapped_target"][s])]
                                           for s in streams]))
      with open(os.path.join(HERE, "scn_E_partial_update.json"), "w") as f:
        json.dump({"hp": _js(hpE), "fixture": ["shrp2"], "seed": RANDOM_SEED,
                   "matches": [{"video_clip": m["video_clip"], "user_match": m["user_match"]}
                               for m in api.matches.values()
                               if m["query_result"] == job["latest_query_result"]["id"]],
                   "rng_after": rng_digest(), "first_clip_id": min(c2.values())}, f, indent=1)
    api.request(qid, None)

    # ---- D: BASELINE config 1 — VQSYN-1 10k clips, one split, new + revise (no target adjustment)
    N = 10000
    seed = synth.DEFAULT_SEED
    X = synth.database(seed, N)                       # [N, 2, 1024] float32
    api = FakeAPI(); holder["api"] = api
    vid, cids = api.load_feature_arrays("synthetic10k", list(range(N)),
                                        {"rgb": {1: X[:, 0]}, "warped_optical_flow": {1: X[:, 1]}})
    ss = api.add_search_set("syn", list(cids.values()))
    ref_row = synth.pick_reference_row(seed, N)
    qid = api.add_query("qD", vid, cids[ref_row], ss, max_matches=20,
                        dynamic_target_adjustment=False)
    rec.rounds = []
    ruleD = lambda m: bool(m["score"] >= 0.85)
    run_rounds(api, APIRepository, broker_defaults, [("new", None), ("revise", ruleD)], rec, qid)
    save_scenario("D_synth10k", rec.rounds,
                  {"fixture": [], "synthetic": {"seed": seed, "n_clips": N, "ref_row": ref_row,
                                                "first_clip_id": cids[0]},
                   "hp": _js(broker_defaults), "streams": streams, "label_rule": "score>=0.85",
                   "max_matches": 20, "dynamic_target_adjustment": False, "seed": RANDOM_SEED,
                   "eps": COMPUTE_EPS, "final_report": None})

    # ---- F: shrp2, bagging x4 on valid + invalid labels with mu > 0 (the `_bootstrap_valid_plus_invalid` solve inside
    #         `target_by_bagging`, target_clip.py:145-159,201-261), f_bootstrap 0.8, ballast 0.2, a narrow near-miss band,
    #         four rounds with two label rules
    api = FakeAPI(page_size=11); holder["api"] = api
    v2, c2 = api.load_feature_arrays(n2, nums2, arr2)
    ss = api.add_search_set("shrp2F", list(c2.values()))
    qid = api.add_query("qF", v2, c2[nums2[70]], ss, max_matches=24, dynamic_target_adjustment=True)
    hpF = dict(broker_defaults, bootstrap_type="bagging", nbags=4, mu=0.25, f_bootstrap=0.8, ballast=0.2,
               near_miss_default=0.2, default_threshold=0.75)
    rec.rounds = []
    ruleF = lambda m: bool(m["score"] >= 0.85)
    run_rounds(api, APIRepository, hpF,
               [("new", None), ("revise", ruleF), ("revise", ruleF), ("finalize", ruleF)], rec, qid)
    save_scenario("F_shrp2_bagging_mu", rec.rounds,
                  {"fixture": ["shrp2"], "ref_clip_number": int(nums2[70]), "hp": _js(hpF), "streams": streams,
                   "page_size": 11, "label_rule": "score>=0.85", "max_matches": 24,
                   "dynamic_target_adjustment": True, "seed": RANDOM_SEED, "eps": COMPUTE_EPS,
                   "final_report": api.uploaded_reports[-1] if api.uploaded_reports else None})

    # ---- G: brooklyn, the user rejects every clip shown (all labels False): no valid match -> the target stays the
    #         reference clip although dynamic adjustment is on (ticket.py:98-107), optimize_weights sees only negatives,
    #         finalize finds no user match at all (lowest_scoring_user_match keeps its initial 1, ticket.py:301-309)
    api = FakeAPI(); holder["api"] = api
    name, nums, arrays = fx["brooklyn"]
    vid, cids = api.load_feature_arrays(name, nums, arrays)
    ss = api.add_search_set("brooklynG", list(cids.values()))
    qid = api.add_query("qG", vid, cids[nums[33]], ss, max_matches=10, dynamic_target_adjustment=True)
    rec.rounds = []
    ruleG = lambda m: False
    run_rounds(api, APIRepository, broker_defaults,
               [("new", None), ("revise", ruleG), ("finalize", ruleG)], rec, qid)
    save_scenario("G_brooklyn_all_rejected", rec.rounds,
                  {"fixture": ["brooklyn"], "ref_clip_number": int(nums[33]), "hp": _js(broker_defaults), "streams": streams,
                   "label_rule": "False", "max_matches": 10, "dynamic_target_adjustment": True, "seed": RANDOM_SEED,
                   "eps": COMPUTE_EPS, "final_report": api.uploaded_reports[-1] if api.uploaded_reports else None})
    print("golden written to", HERE)


def _js(d):
    return {k: (list(v) if isinstance(v, tuple) else v) for k, v in d.items()}


if __name__ == "__main__":
    main()

"""Loaders for the golden scenarios recorded by tests/golden/make_golden.py (test helper)."""
from __future__ import annotations

import hashlib
import json
import os
import random

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = ["A_brooklyn_bagging", "B_both_simple_mu", "C_shrp2_simple", "D_synth10k", "F_shrp2_bagging_mu",
             "G_brooklyn_all_rejected"]
ORACLE_ONLY_SCENARIOS = []         # every recorded scenario is replayed on the GPU (tests/test_gpu_parity.py walks SCENARIOS)
FIXTURE_VIDEO = {"brooklyn": "DowntownBrooklynDrive_480p",
                 "shrp2": "S06NDS_Sample_120406_1451_00186_Forward"}


def rng_digest(rng=random):
    return hashlib.sha256(repr(rng.getstate()).encode()).hexdigest()[:16]


class Scenario:
    """Inputs (X [N,S,P,D] float64, clip_ids, clip numbers) + recorded reference outputs."""

    def __init__(self, name):
        self.name = name
        with open(os.path.join(GOLDEN, "scn_%s.json" % name)) as f:
            js = json.load(f)
        self.meta = js["meta"]
        self.rounds = js["rounds"]
        for r in self.rounds:                     # stored as ordered pairs (dict order matters)
            r["user_matches"] = {k: v for k, v in r["user_matches"]}
        self.arrays = np.load(os.path.join(GOLDEN, "scn_%s.npz" % name))
        self.streams = tuple(self.meta["streams"])
        self.eps = float(self.meta["eps"])
        self.seed = self.meta["seed"]
        self.videos = []          # (name, clip_numbers, X[n,S,P,D])
        if self.meta["fixture"]:
            for fx in self.meta["fixture"]:
                z = np.load(os.path.join(GOLDEN, "fixture_%s.npz" % fx))
                self.videos.append((FIXTURE_VIDEO[fx], z["clip_numbers"], z["X"]))
                self.splits = [int(p) for p in z["splits"]]
            first_id = 1
        else:
            from oracle import synth
            syn = self.meta["synthetic"]
            X = synth.database(syn["seed"], syn["n_clips"]).astype(np.float64)[:, :, None, :]
            self.videos.append(("synthetic10k", np.arange(syn["n_clips"]), X))
            self.splits = [1]
            first_id = syn["first_clip_id"]
        self.X = np.concatenate([v[2] for v in self.videos], axis=0)
        self.clip_ids = np.arange(first_id, first_id + self.X.shape[0], dtype=np.int64)
        self.row_of = {int(c): i for i, c in enumerate(self.clip_ids)}

    def ref_clip_id(self):
        return self.rounds[0]["ref_clip_id"]

    def arr(self, rnd, key):
        return self.arrays["r%d_%s" % (rnd, key)]

    def hp(self):
        d = dict(self.meta["hp"])
        d["streams"] = tuple(d["streams"])
        return d

    def build_api(self, page_size=50):
        """A FakeAPI holding this scenario's database and query, in the same insertion order as
        when the golden was recorded."""
        from fake_api import FakeAPI
        api = FakeAPI(page_size=self.meta.get("page_size", page_size))
        ids, vids = [], []
        for name, nums, X in self.videos:
            arrays = {s: {p: X[:, si, pi, :] for pi, p in enumerate(self.splits)}
                      for si, s in enumerate(self.streams)}
            vid, cids = api.load_feature_arrays(name, [int(n) for n in nums], arrays)
            vids.append(vid)
            ids.extend(cids.values())
        assert ids == [int(c) for c in self.clip_ids]
        ss = api.add_search_set(self.name, ids)
        ref_id = self.ref_clip_id()
        vid = api.clips[ref_id]["video"]
        qid = api.add_query("q" + self.name[0], vid, ref_id, ss,
                            max_matches=self.meta["max_matches"],
                            dynamic_target_adjustment=self.meta["dynamic_target_adjustment"])
        return api, qid

    def label_rule(self):
        return {"score>=0.83": lambda m: bool(m["score"] >= 0.83),
                "score>=0.86": lambda m: bool(m["score"] >= 0.86),
                "score>=0.85": lambda m: bool(m["score"] >= 0.85),
                "True if score>=0.84 else None": lambda m: True if m["score"] >= 0.84 else None,
                "False": lambda m: False,
                }[self.meta["label_rule"]]

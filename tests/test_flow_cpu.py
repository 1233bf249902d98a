"""Flow parity of the driver `compute_matches` on CPU: the product's driver and the
reference's own `src/models/compute_matches.py` are run against the same recording stand-ins for Ticket / TargetClip /
Hyperparameter / APIRepository over every branch (job kind x fatal / recoverable error x previous matches x empty
selection x finalize parameters), and must make the same calls with the same arguments in the same order.

The differential half needs the reference tree (present in the build container, absent on the GPU box: skipped there);
the expected call sequences of a few key cases are also written out below, so the product is checked without it too."""
import importlib
import itertools
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
os.environ.setdefault("COMPUTE_EPS", ".000003")


class Log(list):
    def call(self, who, name, *args):
        self.append((who, name) + tuple(repr(a) for a in args))


def make_stubs(log, case):
    class StubTicket:
        def __init__(self, job, url):
            log.call("ticket", "__init__", sorted(job), url)
            self.query_id = job["query_id"]
            self.number_of_matches_to_review = job["number_of_matches_to_review"]
            self.latest_query_result = job.get("latest_query_result")
            self.matches = job.get("matches")
            self.target = None

        def change_process_state(self, state, message=None):
            log.call("ticket", "change_process_state", state, message)

        def catch_errors(self, kind):
            log.call("ticket", "catch_errors", kind)
            return case["fatal"], case["recoverable"]

        def add_note(self, note):
            log.call("ticket", "add_note", note)

        def compute_similarities(self, hp):
            log.call("ticket", "compute_similarities", type(hp).__name__, type(self.target).__name__)

        def create_query_result(self, nround, hp):
            log.call("ticket", "create_query_result", nround, hp.weights, hp.threshold)
            return 4711

        def compute_scores(self, weights):
            log.call("ticket", "compute_scores", weights)

        def lowest_scoring_user_match(self):
            log.call("ticket", "lowest_scoring_user_match")
            return case["low_score"], 12

        def select_clips_to_review(self, threshold=0.8, max_number_matches=20, near_miss=0.5):
            log.call("ticket", "select_clips_to_review", threshold, max_number_matches, near_miss)
            self.matches = dict(case["selected"])

        def add_matches_to_database(self, rid):
            log.call("ticket", "add_matches_to_database", rid)

        def create_final_report(self, hp, rid):
            log.call("ticket", "create_final_report", type(hp).__name__, rid)

    class StubTarget:
        def __init__(self, ticket, hp):
            log.call("target", "__init__", type(ticket).__name__, type(hp).__name__)

        def get_target_features(self):
            log.call("target", "get_target_features")

    class StubHP:
        def __init__(self):
            self.default_weights = {"rgb": 1.0, "warped_optical_flow": 1.5}
            self.default_threshold = 0.8
            self.near_miss_default = 0.35
            self.weights, self.threshold = None, None

        def optimize_weights(self, ticket):
            log.call("hp", "optimize_weights", type(ticket).__name__)
            self.weights, self.threshold = {"rgb": 1.0, "warped_optical_flow": 0.9}, case["opt_threshold"]

    class StubRepo:
        url = "http://fake/"

        def get_status(self):
            log.call("repo", "get_status")
            return case["jobs"]

    return StubTicket, StubTarget, StubHP, StubRepo


def job(kind, with_matches=True, with_result=True):
    j = {"query_id": 7, "video_id": 1, "ref_clip": 3, "ref_clip_id": 30, "search_set": 2,
         "number_of_matches_to_review": 20, "dynamic_target_adjustment": True}
    if kind != "new":
        j["matches"] = [{"video_clip": 30, "user_match": True, "is_match": True}] if with_matches else []
        j["latest_query_result"] = {"id": 5, "round": 2, "bootstrapped_target": {}} if with_result else None
        j["user_matches"] = {"30": True}
    return j


def cases():
    out = []
    for kind, fatal, rec, with_m, sel, low, opt_th in itertools.product(
            ("new", "revise", "finalize"), ("", "*** Fatal"), ("", "*** Error"), (True, False),
            ((), ((30, 0.9), (31, 0.7))), (0.5, 0.95, 1), (0.8, 1.0 - 1e-7)):
        jobs = {"revise": None, "new": None, "finalize": None}
        jobs[kind] = job(kind, with_m)
        out.append({"kind": kind, "fatal": fatal, "recoverable": rec, "selected": sel, "low_score": low,
                    "opt_threshold": opt_th, "jobs": jobs})
    # several jobs in one tick, in the repository's order; and an empty selection on a job without a previous result
    out.append({"kind": "all", "fatal": "", "recoverable": "", "selected": ((1, 0.9),), "low_score": 0.7, "opt_threshold": 0.82,
                "jobs": {"revise": job("revise"), "new": job("new"), "finalize": job("finalize")}})
    out.append({"kind": "revise", "fatal": "", "recoverable": "", "selected": (), "low_score": 0.7, "opt_threshold": 0.82,
                "jobs": {"revise": job("revise", True, False), "new": None, "finalize": None}})
    return out


def run_product(case):
    cm = importlib.import_module("video_query_algorithms_b200.compute_matches")
    log = Log()
    T, C, H, R = make_stubs(log, case)
    old = cm.TargetClip
    cm.TargetClip = C
    try:
        err = None
        try:
            cm.compute_matches(R(), H(), ticket_factory=lambda j, url: T(j, url))
        except Exception as e:                    # the reference raises in some corner cases: the product must do the same
            err = type(e).__name__
        log.call("end", "exception", err)
    finally:
        cm.TargetClip = old
    return log


def load_reference_driver():
    if not os.path.isdir(REF_SRC):
        pytest.skip("reference tree not present (GPU box)")
    if "coreapi" not in sys.modules:
        stub = types.ModuleType("coreapi")
        stub.Client = object
        stub.auth = types.SimpleNamespace(TokenAuthentication=lambda **kw: None)
        sys.modules["coreapi"] = stub
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        return importlib.import_module("models.compute_matches")


def run_reference(case, ref):
    log = Log()
    T, C, H, R = make_stubs(log, case)
    old = ref.Ticket, ref.TargetClip
    ref.Ticket, ref.TargetClip = T, C
    try:
        err = None
        try:
            ref.compute_matches(R(), H())
        except Exception as e:
            err = type(e).__name__
        log.call("end", "exception", err)
    finally:
        ref.Ticket, ref.TargetClip = old
    return log


def test_driver_makes_the_same_calls_as_the_reference_on_every_branch():
    ref = load_reference_driver()
    n = 0
    for case in cases():
        want, got = run_reference(case, ref), run_product(case)
        assert got == want, (case["kind"], case["fatal"], case["recoverable"], case["selected"], case["low_score"])
        n += 1
    assert n >= 288


def test_driver_key_sequences_without_the_reference():
    """Spelled-out expectations for the three job kinds (they hold on the GPU box too, where the reference is absent)."""
    base = {"fatal": "", "recoverable": "", "selected": ((30, 0.9),), "low_score": 0.5, "opt_threshold": 0.9}
    names = lambda log: [c[1] for c in log if c[0] in ("ticket", "target", "hp")]
    new = run_product(dict(base, kind="new", jobs={"revise": None, "new": job("new"), "finalize": None}))
    assert names(new) == ["__init__", "change_process_state", "catch_errors", "__init__", "get_target_features",
                          "compute_similarities", "create_query_result", "compute_scores", "select_clips_to_review",
                          "add_matches_to_database", "change_process_state"]
    assert ("ticket", "create_query_result", "1", repr({"rgb": 1.0, "warped_optical_flow": 1.5}), "0.8") in new
    assert ("ticket", "select_clips_to_review", "0.8", "20", "0.35") in new and new[-2][2] == "4"
    fin = run_product(dict(base, kind="finalize", jobs={"revise": None, "new": None, "finalize": job("finalize")}))
    assert "optimize_weights" in names(fin) and names(fin)[-3:] == ["add_matches_to_database", "create_final_report",
                                                                    "change_process_state"]
    # finalize: every clip down to the lowest user match: near_miss = max(th - low, 0) / max(1 - th, eps) = 0.4 / 0.1
    sel = [c for c in fin if c[1] == "select_clips_to_review"][0]
    assert sel[2] == "0.9" and sel[3] == "inf" and float(sel[4]) == pytest.approx(4.0) and fin[-2][2] == "7"
    assert ("ticket", "create_query_result", "3", repr({"rgb": 1.0, "warped_optical_flow": 0.9}), "0.9") in fin
    fatal = run_product(dict(base, kind="revise", fatal="*** Fatal", jobs={"revise": job("revise"), "new": None, "finalize": None}))
    assert names(fatal) == ["__init__", "change_process_state", "catch_errors", "change_process_state"]
    assert fatal[-2][2:] == ("5", repr("*** Fatal"))
    empty = run_product(dict(base, kind="revise", selected=(), jobs={"revise": job("revise"), "new": None, "finalize": None}))
    assert empty[-2][1:] == ("change_process_state", "5",
                             repr("*** Error: No matches were found for round 2 of query 7! ***"))


def test_ticket_error_texts_and_hyperparameter_defaults_equal_the_reference():
    """`Ticket.catch_errors` (ticket.py:80-110) on every combination of its three conditions, and the constructor
    defaults / grids of `Hyperparameter` (hyperparameter.py:9-27), against the reference's own classes."""
    load_reference_driver()
    import models.hyperparameter as rhp
    import models.ticket as rticket
    from video_query_algorithms_b200 import Hyperparameter, Ticket
    import numpy as np
    for ref_clip_id, kind, matches, dyn in itertools.product(
            (None, 30), ("new", "revise", "finalize"),
            ([], [{"user_match": True}], [{"user_match": False}, {"user_match": None}]), (True, False)):
        out = []
        for cls in (rticket.Ticket, Ticket):
            t = object.__new__(cls)
            t.ref_clip_id, t.matches, t.dynamic_target_adjustment = ref_clip_id, list(matches), dyn
            out.append((t.catch_errors(kind), t.dynamic_target_adjustment))
        assert out[0] == out[1], (ref_clip_id, kind, matches, dyn)
    a, b = rhp.Hyperparameter({"rgb": 1.0, "warped_optical_flow": 1.5}), Hyperparameter({"rgb": 1.0, "warped_optical_flow": 1.5})
    for name, value in vars(a).items():
        got = getattr(b, name)
        assert np.array_equal(got, value) if isinstance(value, np.ndarray) else got == value, name
    assert len(b.weight_grid) == 40 and len(b.threshold_grid) == 31

"""Pin the CPU oracle (oracle/*.py) to outputs recorded from the reference itself.

CPU-only.  Every round of every golden scenario is replayed with the oracle's array functions,
consuming Python's `random` exactly as the reference does, and compared with what the reference's
objects held (tests/golden/scn_*.npz/json, produced by tests/golden/make_golden.py).
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import bootstrap as ob
from oracle import loop_port
from oracle import scoring as sc
from oracle import synth
from scenarios import GOLDEN, ORACLE_ONLY_SCENARIOS, SCENARIOS, Scenario, rng_digest


def replay_round(scn, i, check):
    r = scn.rounds[i]
    hp = scn.hp()
    kind = r["kind"]
    X, ids = scn.X, scn.clip_ids
    ref_row = scn.row_of[r["ref_clip_id"]]
    prev_matches = r.get("match_status_input")
    random.seed(a=scn.seed)                         # reference src/broker.py:83-84
    assert rng_digest() == r["rng_before"]

    # --- target (A1 / A10)
    dyn = scn.meta["dynamic_target_adjustment"]
    if kind != "new" and dyn and prev_matches and not any(m["user_match"] is True for m in prev_matches):
        dyn = False                                  # ticket.py:98-107
    valid = invalid = None
    if kind != "new" and prev_matches:
        valid = X[[scn.row_of[m["video_clip"]] for m in prev_matches if m["user_match"] is True]]
        invalid = X[[scn.row_of[m["video_clip"]] for m in prev_matches if m["user_match"] is False]]
    T = ob.get_target_features(X[ref_row], valid, invalid, None, dyn, kind != "new",
                               hp["bootstrap_type"], hp["f_bootstrap"], hp["f_memory"],
                               hp["nbags"], hp["mu"], random)
    check("target", T, scn.arr(i, "target"), 1e-9)
    assert rng_digest() == r["rng_after_target"]

    # --- similarities (A3)
    sims, cnt = sc.similarities(X, T)
    check("sims", sims, scn.arr(i, "sims"), 1e-12)
    assert np.array_equal(cnt, scn.arr(i, "nsplits"))
    assert np.array_equal(ids, scn.arr(i, "clip_order"))

    # --- weights (A8/A9)
    if kind == "new" or not prev_matches:
        weights = [hp["default_weights"][s] for s in scn.streams]
        th = hp["default_threshold"]
    else:
        w, th, losses = sc.optimize_weights(sims, ids, prev_matches, scn.eps, hp["ballast"])
        check("losses", losses, scn.arr(i, "losses"), 1e-12)
        weights = [1.0, w]
    assert weights == pytest.approx(r["weights"], rel=1e-10)
    assert th == pytest.approx(r["threshold"], rel=1e-10)

    # --- scores (A4) and selection (A5-A7)
    score = sc.scores(sims, weights)
    check("scores", score, scn.arr(i, "scores"), 1e-12)
    um = r["user_matches"]
    if kind == "finalize":
        low, low_clip = sc.lowest_scoring_user_match(score, ids, um)
        assert [low, low_clip] == pytest.approx(r["lowest_user_match"])
        near, mx = sc.finalize_near_miss(th, low, scn.eps), float("inf")
    else:
        near, mx = hp["near_miss_default"], scn.meta["max_matches"]
    assert [th, mx, near] == pytest.approx(r["select_args"], rel=1e-9)
    assert rng_digest() == r["rng_before_select"]
    sel = sc.select_clips_to_review(score, ids, random, th, mx, near, r["ref_clip_id"], um)
    assert [k for k in sel] == [k for k, _ in r["selected"]]
    assert [sel[k] for k in sel] == pytest.approx([v for _, v in r["selected"]], rel=1e-12)
    assert rng_digest() == r["rng_after_select"]


def _check(name, got, want, rtol):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, name
    err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
    assert err <= rtol, "%s: max rel err %.3e" % (name, err)


@pytest.mark.parametrize("name", SCENARIOS + ORACLE_ONLY_SCENARIOS)
def test_oracle_replays_reference_rounds(name):
    scn = Scenario(name)
    for i in range(len(scn.rounds)):
        replay_round(scn, i, _check)


def test_partial_update_blend_matches_reference():
    """Scenario E: `avg_new_old_targets` (target_clip.py:75-82) recorded by calling the reference's
    TargetClip directly (through compute_matches the reference crashes on this bootstrap type)."""
    with open(os.path.join(GOLDEN, "scn_E_partial_update.json")) as f:
        js = json.load(f)
    z = np.load(os.path.join(GOLDEN, "scn_E_partial_update.npz"))
    X = np.load(os.path.join(GOLDEN, "fixture_shrp2.npz"))["X"]
    row = lambda c: c - js["first_clip_id"]
    valid = X[[row(m["video_clip"]) for m in js["matches"] if m["user_match"] is True]]
    invalid = X[[row(m["video_clip"]) for m in js["matches"] if m["user_match"] is False]]
    hp = js["hp"]
    random.seed(a=js["seed"])
    T = ob.get_target_features(None, valid, invalid, z["previous"], True, True, hp["bootstrap_type"],
                               hp["f_bootstrap"], hp["f_memory"], hp["nbags"], hp["mu"], random)
    _check("target", T, z["target"], 1e-9)
    assert rng_digest() == js["rng_after"]


def test_loop_port_equals_reference_on_golden():
    """The loop-faithful port (what bench.py times as the CPU arm) reproduces the reference's
    similarities and scores bit-for-bit on scenario A round 0 and to 1e-15 on D."""
    for name, tol in (("A_brooklyn_bagging", 0.0), ("D_synth10k", 1e-15)):
        scn = Scenario(name)
        X = scn.X if name[0] == "A" else scn.X[:2000]
        ids = scn.clip_ids[:X.shape[0]]
        T = scn.arr(0, "target")
        cand = loop_port.make_candidates(X, ids, scn.streams, scn.splits)
        tf = loop_port.make_target(T, scn.streams, scn.splits)
        w = dict(zip(scn.streams, scn.rounds[0]["weights"]))
        sims, score, hits, near = loop_port.scoring_step(tf, cand, w, 0.8, 0.35)
        got = np.array([score[int(c)] for c in ids])
        want = scn.arr(0, "scores")[:len(ids)]
        assert np.max(np.abs(got - want)) <= tol * np.max(np.abs(want))
        m, nm = sc.classify(want, 0.8, 0.35)
        assert list(hits) == [int(ids[i]) for i in m]
        assert list(near) == [int(ids[i]) for i in nm]


def test_vectorised_equals_loop_port():
    X = synth.database(7, 500).astype(np.float64)[:, :, None, :]
    T = sc.scale_target(X[3])
    sims, _ = sc.similarities(X, T)
    score = sc.scores(sims, (1.0, 1.5))
    ids = np.arange(100, 600)
    lp = loop_port.scoring_step(loop_port.make_target(T, ("a", "b"), [1]),
                                loop_port.make_candidates(X, ids, ("a", "b"), [1]),
                                {"a": 1.0, "b": 1.5}, 0.8, 0.35)
    got = np.array([lp[1][int(c)] for c in ids])
    assert np.max(np.abs(got - score)) < 1e-14


def test_missing_splits_average_over_present_only():
    """ticket.py:155-157: a clip's per-stream similarity is the mean over the splits it HAS."""
    rng = np.random.default_rng(0)
    X = rng.random((6, 2, 3, 16))
    T = rng.random((2, 3, 16))
    present = np.ones((6, 2, 3), bool)
    present[2, 0, 1] = False
    present[4, 1, 0] = present[4, 1, 2] = False
    sims, cnt = sc.similarities(X, T, present)
    assert cnt[2, 0] == 2 and cnt[4, 1] == 1
    assert sims[2, 0] == pytest.approx((X[2, 0, 0] @ T[0, 0] + X[2, 0, 2] @ T[0, 2]) / 2)
    assert sims[4, 1] == pytest.approx(X[4, 1, 1] @ T[1, 1])


def test_grids_have_reference_sizes():
    assert len(sc.weight_grid()) == 40 and len(sc.threshold_grid()) == 31


def test_topk_stable_tie_rule():
    s = np.array([0.5, 0.9, 0.9, 0.1, 0.9])
    assert list(sc.topk_stable(s, 3)) == [1, 2, 4]


def test_synth_is_deterministic_and_shaped_like_real_features():
    a = synth.rows(synth.DEFAULT_SEED, [5, 999999999, 5])
    assert np.array_equal(a[0], a[2]) and a.dtype == np.float32 and a.min() >= 0
    db = synth.database(synth.DEFAULT_SEED, 64, first_row=999999990)
    assert np.array_equal(db[9], a[1])
    m = synth.database(3, 2000).mean(axis=(0, 2))
    assert 2.2 < m[0] < 2.8 and 0.75 < m[1] < 1.05

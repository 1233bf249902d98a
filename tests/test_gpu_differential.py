"""Differential test on the GPU box: the product's `compute_matches` with the REAL store (CUDA kernels behind the C ABI)
and the same product code with the device played by the float64 oracle (`OracleStore`, tests/test_rounds_cpu.py) run the
same randomly drawn jobs side by side — the jobs of tests/test_live_differential_cpu.py: small complete and ragged search
sets, 1-3 splits, random hyperparameters, bagging / simple bootstrap, mu, skipped labels, reference clips outside the
search set, targets that walk the splits in any order — through new / revise / finalize rounds.  On CPU that generator
pins the product's host path to the unmodified reference; here it pins the kernels (K1 generic shapes, K2 lists and
gathers, K4-K6 in fp64, K7 report order) to the oracle under the very same host logic, where the reference cannot
travel.  Rounds whose scores touch a selection boundary within COMPUTE_EPS (reported as the tie band) or whose loss grid
has a plateau of minima are counted, not compared (see the CPU twin for why)."""
import os
import random

import numpy as np
import pytest

from oracle import scoring as sc
from test_live_differential_cpu import build_api, draw_job, snapshot
from test_rounds_cpu import close, make_oracle_store_class

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("devices,big", [([0], False), ([0, 0, 0], False), ([0], True), ([0, 0], True)],
                         ids=["one-shard", "three-shards", "1024d-one-shard", "1024d-two-shards"])
def test_gpu_store_and_oracle_store_agree_on_random_jobs(tmp_path, monkeypatch, devices, big):
    """big: search sets in the production layout — 2 streams x 1 split x 1024-d, 300-2500 clips — which take the register
    variant of the scan kernel (K1 `scan_rows_reg<2,8>`), selection chunks of 4096 rows and review rounds that sample from
    lists of hundreds of clips; fewer jobs (the fake API holds every feature as a Python list).
    devices = [0, 0, 0]: the search set sharded three ways (all shards on device 0, so that a single-GPU box runs it): the
    one-call multi-shard scan, the multi-shard gathers, per-shard device sorts merged for the report, labelled rows
    gathered across shards for the solve — under the same jobs.  VQ_DIFF_DEVICES=0,1 spreads the shards over real devices."""
    import video_query_algorithms_b200 as vq
    if os.environ.get("VQ_DIFF_DEVICES") and len(devices) > 1:
        devices = [int(d_) for d_ in os.environ["VQ_DIFF_DEVICES"].split(",")]
    from fake_api import FakeRepository
    from video_query_algorithms_b200 import store as ps
    real_store, real_loss_grid = ps.FeatureStore, ps.loss_grid
    oracle_store = make_oracle_store_class()
    oracle_loss_grid = (lambda sims, labels, wg, tg, ballast, replicates=None, device=0:
                        sc.loss_grid(np.asarray(sims, np.float64), np.asarray(labels, bool), wg, tg, ballast)[None])
    for d in ("gpu/work", "cpu/work"):
        (tmp_path / d).mkdir(parents=True)
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(int(os.environ.get("VQ_DIFF_SEED", "4711")) + len(devices))
    n_trials = int(os.environ.get("VQ_DIFF_TRIALS", "40"))
    if big:
        n_trials = max(3, n_trials // 8)
    compared = ties = plateaus = reports = errors = 0
    ps.invalidate()
    registries = {"gpu": {}, "cpu": {}}

    def side(name):
        """the store class, the loss grid and the store registry of one side"""
        ps.FeatureStore, ps.loss_grid = (real_store, real_loss_grid) if name == "gpu" else (oracle_store, oracle_loss_grid)
        ps._REGISTRY = registries[name]
        os.chdir(tmp_path / name / "work")

    try:
        for trial in range(n_trials):
            job = draw_job(rng, (int(rng.integers(300, 2500)), 1024, 1)) if big else draw_job(rng)
            if trial % 3 == 2:                                 # every third job: hyperparameters at the edges of their ranges
                job["hp"].update(default_threshold=float(rng.choice([0.3, 0.5, 0.999, 1.0])), near_miss_default=float(rng.choice([0.0, 0.05, 1.0, 3.0])),
                                 mu=float(rng.choice([0.0, 0.001, 5.0])), f_bootstrap=float(rng.choice([0.01, 0.3, 1])),
                                 nbags=int(rng.integers(1, 7)), ballast=float(rng.choice([0.0, 2.0])))
                job["max_matches"] = int(rng.choice([0, 1, 2, 3, 5, 20]))
            kinds = ["new", "revise", "finalize"][:int(rng.integers(2, 4))]
            api_g, q_g = build_api(job, str(trial))
            api_c, q_c = build_api(job, str(trial))
            for reg in registries.values():
                for st in reg.values():
                    st.close()
                reg.clear()
            tickets = {"gpu": [], "cpu": []}
            for i, kind in enumerate(kinds):
                if i > 0:                                      # the user labels what the GPU side showed; same labels for both
                    shown = {m["video_clip"]: m["score"] for m in api_g.matches.values()
                             if m["query_result"] == api_g._latest_result(q_g)["id"]}
                    if not shown:
                        break
                    cut = float(np.quantile(list(shown.values()), job["label_quantile"])) + 1e-4
                    labels = {c: bool(v >= cut) for c, v in shown.items()}
                    for c in sorted(labels):
                        if rng.random() < job["unlabelled"]:
                            labels[c] = None
                    api_g.label_latest_round(q_g, lambda m: labels[m["video_clip"]])
                    api_c.label_latest_round(q_c, lambda m: labels[m["video_clip"]])
                out = {}
                for name, api, q in (("gpu", api_g, q_g), ("cpu", api_c, q_c)):
                    api.request(q, kind)
                    hp = vq.Hyperparameter(**job["hp"])
                    side(name)
                    random.seed(a=job["seed"])
                    err = None
                    factory = (lambda j, url, name=name, api=api:
                               tickets[name].append(vq.Ticket(j, url, client=api.client(), devices=devices)) or tickets[name][-1])
                    try:
                        vq.compute_matches(FakeRepository(api), hp, ticket_factory=factory)
                    except Exception as e:                     # e.g. a singular labelled set: both sides must refuse it
                        err = "%s: %s" % (type(e).__name__, str(e)[:300])
                    out[name] = (err, snapshot(api, q, hp), random.getstate(), hp)
                (err_g, a, state_g, hp_g), (err_c, b, state_c, hp_c) = out["gpu"], out["cpu"]
                where = (trial, kind, job["hp"]["bootstrap_type"], job["ragged"], job["X"].shape)
                if err_g or err_c:
                    # the library reports a singular solve as an error; the oracle double returns a non-finite slot, which
                    # the product's host code turns into the same error — either way both sides must fail
                    assert err_g is not None and err_c is not None, (where, err_g, err_c)
                    errors += 1
                    break
                assert a["state"] == b["state"] and a["round"] == b["round"], where
                assert a["notes"] == b["notes"], where
                if tickets["gpu"] and (tickets["gpu"][-1].tie_band or tickets["cpu"][-1].tie_band):
                    ties += 1
                    break
                L = hp_c.losses
                if L is not None and kind != "new" and int(np.sum(L - L.min() < 1e-9)) > 1:
                    plateaus += 1
                    break
                assert a["clips"] == b["clips"], (where, a["weights"], b["weights"], a["threshold"], b["threshold"])
                assert close(a["scores"], b["scores"]), where
                if a["weights"] is not None:
                    assert a["weights"] == pytest.approx(b["weights"], rel=1e-5) and a["threshold"] == pytest.approx(b["threshold"], rel=1e-5), where
                assert state_g == state_c, where               # both sides drew the same numbers
                if kind == "finalize" and a["state"] == 7:
                    rep_g, rep_c = api_g.uploaded_reports[-1].splitlines(), api_c.uploaded_reports[-1].splitlines()
                    assert len(rep_g) == len(rep_c), where
                    n_rows = len(a["clips"])
                    sc_c = dict(zip(b["clips"], b["scores"]))
                    for x, y in zip(rep_g[-n_rows:], rep_c[-n_rows:]):
                        cx, cy = x.split(","), y.split(",")
                        if cx[4] == cy[4]:
                            assert cx[:5] == cy[:5] and cx[6:] == cy[6:] and float(cx[5]) == pytest.approx(float(cy[5]), rel=1e-5), where
                        else:                                  # equal-score clips may swap within rounding
                            assert abs(sc_c[int(cx[4])] - sc_c[int(cy[4])]) < 3e-6, (where, x, y)
                    reports += 1
                compared += 1
    finally:
        ps.FeatureStore, ps.loss_grid = real_store, real_loss_grid
        for reg in registries.values():
            for st in reg.values():
                st.close()
        ps._REGISTRY = {}
    print("random jobs: %d rounds compared, %d reports, %d tie-band, %d plateau, %d refused by both"
          % (compared, reports, ties, plateaus, errors))
    assert compared >= n_trials and (reports >= 1 or big), (compared, ties, plateaus, errors, reports)

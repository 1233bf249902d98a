"""BASELINE config 1 end to end: wall time of every compute_matches round of the golden scenarios through this
repo's Ticket / TargetClip / Hyperparameter (fake API in memory, so HTTP is excluded on both sides).  The reference's
own times for the same rounds come from `VQ_GOLDEN_TIMING_ONLY=1 python tests/golden/make_golden.py` (CPU container)."""
import json
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq  # noqa: E402
from fake_api import FakeRepository  # noqa: E402
from scenarios import SCENARIOS, Scenario  # noqa: E402

out = []
for name in SCENARIOS:
    scn = Scenario(name)
    if not scn.rounds or "kind" not in scn.rounds[0]:
        continue
    for rep in range(2):                                  # second repetition: warm library, same work
        os.chdir(tempfile.mkdtemp())                      # the final report goes to ../final_reports/, one file per query name and second
        os.makedirs("work", exist_ok=True)
        os.chdir("work")
        api, qid = scn.build_api()
        vq.invalidate()
        rule = scn.label_rule()
        factory = lambda job, url: vq.Ticket(job, url, client=api.client(), devices=[0])
        times = []
        for i, r in enumerate(scn.rounds):
            if i > 0:
                api.label_latest_round(qid, rule)
            api.request(qid, r["kind"])
            hp = vq.Hyperparameter(**scn.hp())
            random.seed(a=scn.seed)
            t0 = time.perf_counter()
            vq.compute_matches(FakeRepository(api), hp, ticket_factory=factory)
            times.append({"kind": r["kind"], "compute_matches_s": time.perf_counter() - t0})
    out.append({"scenario": name, "rounds": times, "note": "round 1 builds the resident store from the search-sets/features response "
                "(the reference re-fetches and re-parses it every round); later rounds reuse it"})
print(json.dumps(out))

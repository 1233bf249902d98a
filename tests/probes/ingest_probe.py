"""Store build rates on the GPU box: (a) from `search-sets/features`-shaped records — a list of dicts with
Python-float lists, what the API client hands over — through index_feature_rows + the pinned two-buffer ingest pipeline;
(b) from a CSV tree as the TSN extractor writes it (vq_csv_read on all cores + pipelined upload).  Both compared bit for bit with
the plain path (pack_feature_rows + upload).  Prints one JSON object.

  python tests/probes/ingest_probe.py [n_clips] [csv_clips]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("COMPUTE_EPS", ".000003")
import video_query_algorithms_b200 as vq  # noqa: E402
from video_query_algorithms_b200 import ingest, store as ps  # noqa: E402

S = ("rgb", "warped_optical_flow")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    n_csv = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
    rng = np.random.default_rng(0)
    out = {"cores": len(os.sched_getaffinity(0))}
    # ---- (a) API-shaped records
    X = rng.random((n, 2, 1024))
    t0 = time.perf_counter()
    recs = [{"dnn_stream_id": s, "dnn_stream_split": 1, "name": "global_pool", "video_clip_id": c + 1,
             "feature_vector": X[c, si].tolist()} for c in range(n) for si, s in enumerate(S)]
    out["make_records_s"] = time.perf_counter() - t0
    vq.FeatureStore.from_feature_rows(recs[:2000], S, "global_pool", devices=[0]).close()        # warm: context, staging
    t0 = time.perf_counter()
    idx, _ = ps.index_feature_rows(recs, S, "global_pool")
    t_index = time.perf_counter() - t0
    t0 = time.perf_counter()
    st = vq.FeatureStore.from_feature_rows(recs, S, "global_pool", devices=[0])
    t_build = time.perf_counter() - t0
    got = st.download(0, n)
    same = bool(np.array_equal(got[:, :, 0, :], X.astype(np.float32)))
    st.close()
    out["api_records"] = {"clips": n, "records": len(recs), "index_pass_s": t_index, "build_total_s": t_build,
                          "clips_per_s": n / t_build, "seconds_per_1M_clips": t_build * 1e6 / n, "bits_equal_to_source": same,
                          "what": "FeatureStore.from_feature_rows: one native pass over the record dicts, then vectors unboxed chunk by "
                                  "chunk into pinned staging by %d threads while the previous chunk is copied to HBM" % min(out["cores"], 16)}
    del recs
    # ---- (b) CSV tree
    with tempfile.TemporaryDirectory() as tmp:
        Xc = rng.random((n_csv, 2, 1024)) * 3
        for split in (1,):
            d = os.path.join(tmp, "video_a", "split%d" % split)
            os.makedirs(d)
            for si, s in enumerate(S):
                with open(os.path.join(d, "%s_global_pool_features.csv" % s), "w") as f:
                    f.write("video =video_a, video url =x, CNN stream =%s, feature blob =global_pool, caffe model =m\n" % s)
                    for c in range(n_csv):
                        f.write("%d," % c + ",".join(repr(float(v)) for v in Xc[c, si]) + "\n")
        size = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(tmp) for f in fs)
        path = os.path.join(tmp, "video_a", "split1", "rgb_global_pool_features.csv")
        rates = {}
        for th in (1, 8, 0):
            t0 = time.perf_counter()
            rec = ingest.read_feature_csv(path, n_threads=th)
            rates["threads_%s" % (th or "all")] = n_csv / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        st, ids = ingest.store_from_feature_tree(tmp, S, "global_pool", devices=[0])
        t_tree = time.perf_counter() - t0
        same = bool(np.array_equal(st.download(0, n_csv)[:, :, 0, :], Xc.astype(np.float32)))
        st.close()
        out["csv_tree"] = {"clips": n_csv, "bytes": size, "parse_rows_per_s": rates, "build_total_s": t_tree,
                           "clips_per_s": n_csv / t_tree, "mb_per_s": size / t_tree / 1e6, "bits_equal_to_source": same}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

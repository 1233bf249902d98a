"""Ingest of the real feature-database formats into the HBM store.

Two sources carry the same records:
  * the CSV tree written by the TSN extractor and read by `load_db.py`
    (`<src>/<video>/<split dir>/<stream>_<blob>_features.csv`; header = 5 `key =value` fields, rows =
    `clip_no, f0 .. f1023`; the split number is the last character of the directory name —
    reference src/load_db.py:10-28, src/api/api_load_records.py:41-58);
  * the `search-sets/features` API response (list of feature dicts) — handled by
    `FeatureStore.from_feature_rows` (reference src/models/ticket.py:363-382).
Parsing is host work; the arrays are then uploaded once and stay resident.
"""
from __future__ import annotations

import csv
import os

import numpy as np

from .store import FeatureStore


def read_feature_csv(path, n_threads=0):
    """-> dict(video, stream, feature_name, weights_uri, clip_numbers [n], features float64 [n, dim]).
    Parsed by the library (vq_csv_read: all cores, same doubles as float()); a malformed file raises VQError."""
    import ctypes as C
    from ._ffi import check, lib, ptr
    bpath = os.fsencode(path)
    n, dim, hdr = C.c_int64(), C.c_int32(), C.create_string_buffer(1 << 16)
    check(lib().vq_csv_shape(bpath, C.byref(n), C.byref(dim), hdr, len(hdr)), "vq_csv_shape")
    clips = np.empty(n.value, np.int64)
    feats = np.empty((n.value, dim.value), np.float64)
    got = C.c_int64()
    check(lib().vq_csv_read(bpath, int(n_threads), n.value, dim.value, ptr(clips), ptr(feats), C.byref(got)), "vq_csv_read")
    header = next(csv.reader([hdr.value.decode()]))
    field = lambda i: header[i].split("=")[-1]
    return {"video": field(0), "stream": field(2), "feature_name": field(3), "weights_uri": field(4),
            "clip_numbers": clips, "features": feats}


def read_feature_tree(src_dir):
    """Walk `<src>/<video>/<split dir>/*.csv` like load_db.py.  Returns
    {video: {"clip_numbers": [n], "features": {stream: {split: float64 [n, dim]}}, "feature_name": str}}."""
    out = {}
    for video_dir in sorted(os.scandir(src_dir), key=lambda e: e.name):
        if not video_dir.is_dir():
            continue
        for split_dir in sorted(os.scandir(video_dir.path), key=lambda e: e.name):
            if not split_dir.is_dir():
                continue
            split = int(split_dir.name[-1])
            for entry in sorted(os.scandir(split_dir.path), key=lambda e: e.name):
                if not (entry.is_file() and entry.name.endswith(".csv") and not entry.name.startswith(".")):
                    continue
                rec = read_feature_csv(entry.path)
                v = out.setdefault(rec["video"], {"clip_numbers": rec["clip_numbers"], "features": {},
                                                  "feature_name": rec["feature_name"]})
                if not np.array_equal(v["clip_numbers"], rec["clip_numbers"]):
                    # files of one video may list different clips: align on the union, NaN marks "absent"
                    union = np.union1d(v["clip_numbers"], rec["clip_numbers"])
                    for s_ in v["features"].values():
                        for p_ in list(s_):
                            full = np.full((len(union), s_[p_].shape[1]), np.nan)
                            full[np.searchsorted(union, v["clip_numbers"])] = s_[p_]
                            s_[p_] = full
                    feat = np.full((len(union), rec["features"].shape[1]), np.nan)
                    feat[np.searchsorted(union, rec["clip_numbers"])] = rec["features"]
                    v["clip_numbers"], rec["features"] = union, feat
                v["features"].setdefault(rec["stream"], {})[split] = rec["features"]
    return out


def store_from_feature_tree(src_dir, streams=("rgb", "warped_optical_flow"), feature_name="global_pool",
                            devices=None, first_clip_id=1):
    """Build a FeatureStore from a CSV tree.  Clip ids are assigned first_clip_id.. in (video, clip) order —
    the order `load_db.py` creates them on an empty database.  Returns (store, {(video, clip_no): clip_id})."""
    tree = read_feature_tree(src_dir)
    blocks, present, ids, next_id, splits = [], [], {}, first_clip_id, set()
    for v in tree.values():
        if v["feature_name"] != feature_name:
            continue
        for s in streams:
            splits.update(v["features"].get(s, {}))
    splits = sorted(splits)
    dim = None
    for name, v in tree.items():
        if v["feature_name"] != feature_name:
            continue
        n = len(v["clip_numbers"])
        some = next(iter(next(iter(v["features"].values())).values()))
        dim = some.shape[1]
        X = np.zeros((n, len(streams), len(splits), dim), np.float32)
        P = np.zeros((n, len(streams), len(splits)), bool)
        for si, s in enumerate(streams):
            for pi, p in enumerate(splits):
                arr = v["features"].get(s, {}).get(p)
                if arr is None:
                    continue
                ok = ~np.isnan(arr[:, 0])
                X[ok, si, pi] = arr[ok]
                P[ok, si, pi] = True
        for c in v["clip_numbers"]:
            ids[(name, int(c))] = next_id
            next_id += 1
        blocks.append(X)
        present.append(P)
    X = np.concatenate(blocks)
    st = FeatureStore(X.shape[0], streams, splits, dim, devices=devices,
                      clip_ids=np.arange(first_clip_id, first_clip_id + X.shape[0]))
    st.upload_pipelined(0, X)
    st.set_present(np.concatenate(present))
    return st, ids

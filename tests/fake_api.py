"""In-memory stand-in for the external video-query-api (Django REST) server.

Test infrastructure only.  It speaks the coreapi-shaped surface the scoring path uses
(`client.get(url)`, `client.action(schema, [..], params=..., encoding=...)`) and keeps the
record layout that `load_db.py` writes (reference `src/api/api_load_records.py:104-113`) and
that the path reads (`src/models/ticket.py:374-381`, `src/models/target_clip.py:279-285`).

The same fake serves two users:
  * `tests/golden/make_golden.py`, which drives the UNMODIFIED reference against it to record
    golden outputs (only in the build container, where /root/reference exists);
  * the parity tests, which drive this repo's drop-in path against the same data.

The real API server is not part of the reference repo, so its job-assembly rules (which
matches / user_matches a "revise" job carries) are restated here from the field list the
reference consumes (`src/api/api_repository.py:25-43`, `src/models/ticket.py:38-54`).
"""
from __future__ import annotations

import copy
import json


class FakeSchema:
    """Opaque token handed back by client.get(<url>/docs)."""


class FakeAPI:
    def __init__(self, page_size=50):
        self.search_set_record = "members"                    # or "bare": search-sets/read returns id and name only
        self.videos = {}          # id -> dict
        self.clips = {}           # id -> dict(id, clip, video, duration, notes)
        self.features = []        # list of feature rows, in insertion (= response) order
        self.features_by_clip = {}
        self.search_sets = {}     # id -> dict(id, name, clip_ids)
        self.queries = {}         # id -> dict
        self.query_results = {}   # id -> dict
        self.matches = {}         # id -> dict
        self.page_size = page_size
        self.calls = []           # (action tuple) log, for asserting the API contract
        self.uploaded_reports = []
        self._next = {"video": 1, "clip": 1, "ss": 1, "query": 1, "qr": 1, "match": 1}

    # ------------------------------------------------------------------ loading
    def _id(self, kind):
        v = self._next[kind]
        self._next[kind] = v + 1
        return v

    def add_video(self, name, path="", first_clip_id=None):
        vid = self._id("video")
        self.videos[vid] = {"id": vid, "name": name, "path": path}
        if first_clip_id is not None:
            self._next["clip"] = max(self._next["clip"], first_clip_id)
        return vid

    def add_clip(self, video_id, clip_no, duration=10):
        cid = self._id("clip")
        self.clips[cid] = {"id": cid, "clip": clip_no, "video": video_id, "duration": duration,
                           "notes": ""}
        self.features_by_clip[cid] = []
        return cid

    def add_feature(self, clip_id, stream, split, vector, name="global_pool"):
        row = {
            "dnn_stream_id": stream,          # the stream NAME, as the API returns it
            "dnn_stream_split": int(split),
            "feature_vector": vector,          # Python list of float, like decoded JSON
            "name": name,
            "video_clip_id": clip_id,
        }
        self.features.append(row)
        self.features_by_clip[clip_id].append(row)

    def add_search_set(self, name, clip_ids):
        sid = self._id("ss")
        self.search_sets[sid] = {"id": sid, "name": name, "clip_ids": list(clip_ids)}
        return sid

    def add_query(self, name, video_id, ref_clip_id, search_set, max_matches=20,
                  dynamic_target_adjustment=True, reference_time="00:01:40"):
        qid = self._id("query")
        self.queries[qid] = {
            "id": qid, "name": name, "video": video_id, "ref_clip_id": ref_clip_id,
            "search_set_to_query": search_set, "max_matches_for_review": max_matches,
            "use_dynamic_target_adjustment": dynamic_target_adjustment,
            "reference_time": reference_time, "notes": "", "process_state": 1,
            "final_report_file": None, "pending": "new",   # which compute-* list it shows up in
        }
        return qid

    def load_feature_arrays(self, video_name, clip_numbers, arrays, duration=10):
        """arrays: {stream: {split: ndarray[n_clips, D]}} in CSV row order.

        Insertion order mimics load_db.py walking split dirs then CSV files: all rows of one
        (split, stream) file, then the next file (reference `src/load_db.py:10-28`).
        """
        vid = self.add_video(video_name)
        clip_ids = {}
        for stream, by_split in arrays.items():
            for split, arr in by_split.items():
                for i, cno in enumerate(clip_numbers):
                    if cno not in clip_ids:
                        clip_ids[cno] = self.add_clip(vid, int(cno), duration)
        for split in sorted({sp for by in arrays.values() for sp in by}):
            for stream, by_split in arrays.items():
                if split not in by_split:
                    continue
                arr = by_split[split]
                for i, cno in enumerate(clip_numbers):
                    self.add_feature(clip_ids[cno], stream, split, [float(x) for x in arr[i]])
        return vid, clip_ids

    # ------------------------------------------------------------------ user simulation
    def label_latest_round(self, query_id, rule):
        """Set user_match on the matches of the latest round: rule(match_dict) -> True/False/None."""
        qr = self._latest_result(query_id)
        for m in self.matches.values():
            if m["query_result"] == qr["id"]:
                m["user_match"] = rule(m)

    def request(self, query_id, kind):
        """Mark the query as waiting for 'new' | 'revise' | 'finalize' work."""
        self.queries[query_id]["pending"] = kind

    # ------------------------------------------------------------------ coreapi surface
    def client(self):
        return FakeClient(self)

    def _latest_result(self, query_id):
        rs = [r for r in self.query_results.values() if r["query"] == query_id]
        return max(rs, key=lambda r: r["round"]) if rs else None

    def _job(self, kind):
        for q in self.queries.values():
            if q.get("pending") != kind:
                continue
            clip = self.clips.get(q["ref_clip_id"])
            job = {
                "query_id": q["id"],
                "video_id": q["video"],
                "ref_clip": clip["clip"] if clip else None,
                "ref_clip_id": q["ref_clip_id"],
                "search_set": q["search_set_to_query"],
                "number_of_matches_to_review": q["max_matches_for_review"],
                "dynamic_target_adjustment": q["use_dynamic_target_adjustment"],
            }
            if kind != "new":
                qr = self._latest_result(q["id"])
                job["latest_query_result"] = {
                    "id": qr["id"], "round": qr["round"],
                    "match_criterion": qr["match_criterion"], "weights": qr["weights"],
                    "bootstrapped_target": qr["bootstrapped_target"],   # JSON string
                }
                job["matches"] = [
                    {"video_clip": m["video_clip"], "user_match": m["user_match"],
                     "is_match": m["is_match"], "score": m["score"]}
                    for m in self.matches.values() if m["query_result"] == qr["id"]
                ]
                um = {}
                rounds = sorted((r for r in self.query_results.values() if r["query"] == q["id"]),
                                key=lambda r: r["round"])
                for r in rounds:
                    for m in self.matches.values():
                        if m["query_result"] == r["id"] and m["user_match"] is not None:
                            um[str(m["video_clip"])] = m["user_match"]
                job["user_matches"] = um
            return copy.deepcopy(job)
        return None

    def action(self, keys, params=None, encoding=None):
        keys = tuple(keys)
        params = params or {}
        self.calls.append(keys)
        if keys[0] == "query-state":
            kind = {"compute-new": "new", "compute-revised": "revise",
                    "compute-finalize": "finalize"}[keys[1]]
            return self._job(kind)
        if keys == ("search-sets", "features"):
            ids = set(self.search_sets[params["id"]]["clip_ids"])
            return [r for r in self.features if r["video_clip_id"] in ids]
        if keys == ("search-sets", "read"):
            s = self.search_sets[params["id"]]
            if self.search_set_record == "bare":              # an API whose record says nothing about the members
                return {"id": s["id"], "name": s["name"]}
            videos = sorted({self.clips[c]["video"] for c in s["clip_ids"]})
            return {"id": s["id"], "name": s["name"], "videos": videos, "number_of_clips": len(s["clip_ids"])}
        if keys == ("video-clips", "features"):
            return list(self.features_by_clip[params["id"]])
        if keys == ("video-clips", "read"):
            return dict(self.clips[params["id"]])
        if keys == ("videos", "read"):
            return dict(self.videos[params["id"]])
        if keys == ("queries", "read"):
            return dict(self.queries[params["id"]])
        if keys == ("queries", "partial_update"):
            q = self.queries[params["id"]]
            for k, v in params.items():
                if k == "id":
                    continue
                if k == "final_report_file":
                    self.uploaded_reports.append(v.read())
                    q[k] = "uploaded"
                    continue
                q[k] = v
                if k == "process_state" and v in (4, 5, 7):
                    q["pending"] = None
            return dict(q)
        if keys == ("query-results", "create"):
            rid = self._id("qr")
            self.query_results[rid] = {
                "id": rid, "round": params["round"],
                "match_criterion": float(params["match_criterion"]),
                "weights": [float(w) for w in params["weights"]], "query": params["query"],
                "bootstrapped_target": params["bootstrapped_target"],
            }
            assert isinstance(params["bootstrapped_target"], str)
            json.loads(params["bootstrapped_target"])
            return dict(self.query_results[rid])
        if keys == ("query-results", "read"):
            return dict(self.query_results[params["id"]])
        if keys == ("matches", "create"):
            mid = self._id("match")
            qr = self.query_results[params["query_result"]]
            clip = self.clips[params["video_clip"]]
            start = clip["clip"] * clip["duration"]
            self.matches[mid] = {
                "id": mid, "query_result": params["query_result"], "score": float(params["score"]),
                "user_match": params["user_match"], "video_clip": params["video_clip"],
                "is_match": bool(float(params["score"]) >= qr["match_criterion"]),
                "match_video_time_span": "{},{}".format(start, start + clip["duration"]),
            }
            return dict(self.matches[mid])
        if keys == ("matches", "list"):
            rows = [m for m in self.matches.values() if m["query_result"] == params["query_result"]]
            if "video_clip" in params:
                rows = [m for m in rows if m["video_clip"] == params["video_clip"]]
                return {"results": [dict(m) for m in rows], "pagination": {"nextPage": None}}
            page = int(params.get("page", 1))
            lo, hi = (page - 1) * self.page_size, page * self.page_size
            nxt = page + 1 if hi < len(rows) else None
            return {"results": [dict(m) for m in rows[lo:hi]], "pagination": {"nextPage": nxt}}
        raise KeyError("fake API: unknown action {}".format(keys))


class FakeClient:
    def __init__(self, api):
        self.api = api

    def get(self, url):
        return FakeSchema()

    def action(self, schema, keys, params=None, encoding=None, **kw):
        return self.api.action(keys, params=params, encoding=encoding)


class FakeRepository:
    """Stand-in for APIRepository (`src/api/api_repository.py:12-78`): .url, .get_status()."""

    def __init__(self, api, url="http://fake/"):
        self.api = api
        self.url = url
        self.client = api.client()
        self.schema = self.client.get(url + "docs")

    def get_status(self):
        out = {}
        for kind, key in (("revise", "compute-revised"), ("new", "compute-new"),
                          ("finalize", "compute-finalize")):
            job = self.client.action(self.schema, ["query-state", key, "list"])
            if job and kind != "new":
                bt = job["latest_query_result"]["bootstrapped_target"]
                if bt:
                    d = json.loads(bt)
                    job["latest_query_result"]["bootstrapped_target"] = {
                        s: {int(p): v for p, v in by.items()} for s, by in d.items()}
            out[kind] = job
        return out

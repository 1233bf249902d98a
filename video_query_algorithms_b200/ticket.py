"""Job ticket: scoring and selection for one query update, backed by the HBM store.

Mirror of reference `src/models/ticket.py`: same constructor, attributes and method names, same
API actions.  What changes is where the arithmetic runs:
  * `_get_candidate_features` returns the resident FeatureStore instead of a nested dict built
    from a full HTTP download (ticket.py:358-382);
  * `compute_similarities` / `compute_scores` only record the target and the weights; the work is
    one fused GPU scan issued by `select_clips_to_review` (or by the first read of `.scores` /
    `.similarities`, which are lazy read-only mappings over device results);
  * the two candidate scans (ticket.py:326-327) come back as row lists in database order, and the
    seeded sampling (:333, :341) runs on the host with Python's `random`, consuming it exactly as
    the reference does.
"""
from __future__ import annotations

import csv
import itertools
import json
import logging
import os
import random
import time
from collections.abc import Mapping
from datetime import datetime, timedelta
from time import sleep

import numpy as np

from . import store as _store
from ._ffi import VQError
from ._rng import sample_range

try:
    from requests import ConnectionError
except Exception:                           # pragma: no cover
    ConnectionError = OSError


def _default_client(api_url):
    """coreapi client with token auth, as the reference builds it (ticket.py:36-37,
    api/authenticate.py:6-24).  Only used when no client is injected."""
    import coreapi
    import requests
    response = requests.post(os.path.join(api_url, 'api-token-auth/'),
                             data={'username': os.environ['API_CLIENT_USERNAME'],
                                   'password': os.environ['API_CLIENT_PASSWORD']})
    auth = None
    if response.status_code == requests.codes.ok:
        auth = coreapi.auth.TokenAuthentication(scheme='Token', token=response.json()['token'])
    else:
        print("Authentication failed: ")
    return coreapi.Client(auth=auth)


class _ScoreView(Mapping):
    """Read-only {video_clip_id: score} over the device score array, in database order."""

    def __init__(self, ticket):
        self._t = ticket
        self._arr = None

    def _scores(self):
        if self._arr is None:
            self._t._ensure_scan()
            self._arr = self._t.feature_store().scores()
        return self._arr

    def __getitem__(self, clip):
        st = self._t.feature_store()
        if not st.has_clip(clip):
            raise KeyError(clip)
        return float(self._scores()[st.row_of(clip)])

    def __contains__(self, clip):
        return self._t.feature_store().has_clip(clip)

    def __iter__(self):
        return (int(c) for c in self._t._ids_in_dict_order())

    def __len__(self):
        return self._t.feature_store().n_rows

    def array(self):
        return self._scores()


class _SimilarityView(Mapping):
    """Read-only {video_clip_id: {stream: [avg similarity, n_splits]}} (ticket.py:124)."""

    def __init__(self, ticket):
        self._t = ticket
        self._arr = None

    def _sims(self):
        if self._arr is None:
            self._t._ensure_scan(want_sims=True)
            self._arr = self._t.feature_store().sims()
        return self._arr

    def __getitem__(self, clip):
        st = self._t.feature_store()
        if not st.has_clip(clip):
            raise KeyError(clip)
        r = st.row_of(clip)
        have = self._t._target_have
        out = {}
        for si, s in enumerate(st.streams):
            pres = have[si] if st.present is None else (st.present[r, si] & have[si])
            out[s] = [float(self._sims()[r, si]), int(np.sum(pres))]
        return out

    def __iter__(self):
        return (int(c) for c in self._t._ids_in_dict_order())

    def __len__(self):
        return self._t.feature_store().n_rows

    def array(self):
        return self._sims()


class Ticket:
    def __init__(self, update_object, api_url, client=None, schema=None, store=None, devices=None):
        """update_object: the job dict of `APIRepository.get_status()` (ticket.py:18-33).
        client/schema: an injected coreapi-style client (tests, embedding); store: a pre-built
        FeatureStore for the job's search set (otherwise built once from `search-sets/features`
        and cached for later ticks)."""
        self.api_url = api_url
        self.client = client if client is not None else _default_client(api_url)
        self.schema = schema if schema is not None else self.client.get(os.path.join(api_url, "docs"))
        self.query_id = update_object["query_id"]
        self.video_id = update_object["video_id"]
        self.ref_clip = update_object["ref_clip"]
        self.ref_clip_id = update_object["ref_clip_id"]
        self.search_set = update_object["search_set"]
        self.number_of_matches_to_review = update_object["number_of_matches_to_review"]
        self.dynamic_target_adjustment = update_object["dynamic_target_adjustment"]
        self.latest_query_result = update_object.get("latest_query_result")
        if "matches" in update_object:
            self.matches = update_object["matches"]
        self.user_matches = update_object.get("user_matches", {})
        self.target = None
        self.similarities = {}
        self.scores = {}
        # --- B200 path state
        self.devices = devices
        self.features_from_store = False   # True: read clip features back from HBM, not the API
        self._store = store
        self._hp = None
        self._weights = None
        self._target_have = None
        self._scanned = False
        self._have_sims = False
        self.topk = 0                    # optional ranked list size (BASELINE "top-100")
        self.ranked = None               # (clip ids, scores) of the last scan's top-k
        self.tie_band = []               # clip ids within COMPUTE_EPS of a set boundary
        self.last_scan = None

    # ------------------------------------------------------------------ API plumbing
    def add_matches_to_database(self, new_result_id):
        for video_clip, score in self.matches.items():
            self.create_match(new_result_id, score, self.user_matches.get(str(video_clip)), video_clip)

    def add_note(self, note):
        result = self._request(["queries", "read"], {"id": self.query_id})
        new_notes = result["notes"] + '\n\n' + note if result["notes"] else note
        self._request(["queries", "partial_update"], {"id": self.query_id, "notes": new_notes})

    def catch_errors(self, job_type):
        """Fatal / recoverable query errors (ticket.py:80-110)."""
        fatal, recoverable = [], []
        if self.ref_clip_id is None:
            fatal.append("*** Fatal Error: A video clip corresponding to the reference time does "
                         "not exist in the database. ***")
        if job_type != "new" and not self.matches:
            # byte for byte the note the reference writes: its string literal continues across lines inside the quotes
            # (ticket.py:92-94), so the text carries two stray quote pairs with 34 and 35 blanks between them
            fatal.append("*** Fatal Error: This is not a new query but there are 0 matches computed '" + " " * 34 +
                         "'for the previous round. Cannot update without matches. Check database consistency '" + " " * 35 +
                         "'for this query")
        if job_type != "new" and self.dynamic_target_adjustment is True:
            if not any(match["user_match"] is True for match in self.matches):
                recoverable.append('*** Error: Dynamic target adjustment is True but there are no user matches '
                                   'provided for the previous round. Changing dynamic target adjustment to False')
                self.dynamic_target_adjustment = False
        return "\n".join(fatal), "\n".join(recoverable)

    def change_process_state(self, process_state, message=None):
        result = self._request(["queries", "partial_update"], {"id": self.query_id, "process_state": process_state})
        if message:
            self.add_note(message)
        return result["process_state"]

    def create_match(self, qresult, score, user_match, video_clip):
        self._request(["matches", "create"], {"query_result": qresult, "score": score,
                                              "user_match": user_match, "video_clip": video_clip})

    def create_query_result(self, nround, hyperparameters):
        weights_values = [hyperparameters.weights[stream] for stream in hyperparameters.streams]
        params = {"round": nround, "match_criterion": hyperparameters.threshold, "weights": weights_values,
                  "query": self.query_id, "bootstrapped_target": json.dumps(self.target.target_features)}
        return self._request(["query-results", "create"], params)["id"]

    def _request(self, action, params):
        while True:
            try:
                return self.client.action(self.schema, action, params=params)
            except ConnectionError:
                sleep(0.05)
                logging.warning('Try API request again: action = {}, params = {}'.format(action, params))

    def _post_file(self, action, params):
        while True:
            try:
                return self.client.action(self.schema, action, params=params, encoding="multipart/form-data")
            except ConnectionError:
                sleep(0.05)
                logging.warning('Try API file post by Ticket again: action = {}, params = {}'.format(action, params))

    # ------------------------------------------------------------------ store
    def feature_store(self, optional=False):
        if self._store is None and not optional:
            if self._hp is None:
                raise VQError("ticket has no feature store yet: call attach_store / compute_similarities first")
            self._get_candidate_features(None, self._hp)
        return self._store

    def _get_candidate_features(self, splits, hyperparameters):
        """The resident store of this job's search set: built on first use from the same API action the reference calls
        every job (ticket.py:363-365), then kept across ticks — and kept CURRENT: the reference re-reads the search set
        per job, so clips that load_db.py added since the last tick (reference load_db.py:10-28) must be scored too."""
        if self._store is not None:
            return self._store
        key = (self.api_url, self.search_set, tuple(hyperparameters.streams), hyperparameters.feature_name)
        st = _store._REGISTRY.get(key)
        if st is None:
            st = self._build_store(key, hyperparameters, self._request(["search-sets", "features"], {"id": self.search_set}))
        elif self._store_is_stale(st):
            rows = self._request(["search-sets", "features"], {"id": self.search_set})
            added = st.sync_feature_rows(rows, hyperparameters.feature_name)
            if added is None:                                 # not "what we hold plus new clips at the end": rebuild from the response
                logging.info("search set %s changed in place: rebuilding its feature store", self.search_set)
                _store.invalidate(key)
                st = self._build_store(key, hyperparameters, rows)
            else:
                if added:
                    logging.info("search set %s grew by %d clip(s): appended to the resident store", self.search_set, added)
                st.freshness = self._freshness_probe()
                st.built_at = time.monotonic()
        st.last_used = time.monotonic()
        self._store = st
        return st

    def _build_store(self, key, hyperparameters, rows):
        """Build the search set's store; when the devices are full of OTHER search sets' stores, the least recently used
        of those make room (they are rebuilt from the API the next time a job names them)."""
        if not isinstance(rows, list):
            rows = list(rows)
        while True:
            try:
                st = _store.FeatureStore.from_feature_rows(rows, hyperparameters.streams, hyperparameters.feature_name,
                                                           devices=self.devices)
                break
            except VQError as e:
                if "out of memory" not in str(e) or not _store.evict_least_recently_used(keep=key):
                    raise
                logging.warning("device memory is full: evicted the least recently used feature store to build search set %s",
                                self.search_set)
        st.freshness = self._freshness_probe()
        st.built_at = time.monotonic()
        _store.register_store(key, st)
        return st

    def _freshness_probe(self):
        """A cheap API fact that changes when the search set does: the `search-sets/read` record (ticket.py:197-199
        reads it for the report), canonicalised.  Whatever the API puts there — member videos, counts, a modification
        stamp — takes part; fields that never change cost nothing."""
        if os.environ.get("VQ_STORE_FRESHNESS", "probe") != "probe":
            return None
        try:
            return json.dumps(self._request(["search-sets", "read"], {"id": self.search_set}), sort_keys=True, default=str)
        except ConnectionError:
            raise
        except Exception:                                     # an API without that action: the other signals still apply
            return None

    def _store_is_stale(self, st):
        """VQ_STORE_FRESHNESS = probe (default): stale when the search-set record changed, when this job names a clip
        the store does not hold (the previous round's matches and the user's labels come from this search set), or when
        the store is older than VQ_STORE_MAX_AGE_S (unset: no age limit).  `always` re-reads the search set every job
        like the reference (only new clips are uploaded); `never` trusts the resident store."""
        mode = os.environ.get("VQ_STORE_FRESHNESS", "probe")
        if mode == "never":
            return False
        if mode == "always":
            return True
        prev = getattr(self, "matches", None)
        named = [m["video_clip"] for m in prev] if isinstance(prev, list) else []
        named += [int(c) for c in self.user_matches]
        if named and (st._lookup(named) < 0).any():
            return True
        max_age = os.environ.get("VQ_STORE_MAX_AGE_S")
        if max_age and time.monotonic() - getattr(st, "built_at", 0.0) > float(max_age):
            return True
        return self._freshness_probe() != getattr(st, "freshness", None)

    def attach_store(self, hyperparameters):
        """Resolve the store before the target is built (TargetClip reads labelled rows from it)."""
        self._hp = hyperparameters
        return self._get_candidate_features(None, hyperparameters)

    # ------------------------------------------------------------------ scoring (A3, A4)
    def compute_similarities(self, hyperparameters):
        """Bind target and store; similarities materialise on first read (ticket.py:120-163)."""
        self._hp = hyperparameters
        st = self._get_candidate_features(self.target.splits, hyperparameters)
        T, self._target_have = st.pack_target(self.target.target_features, np.float32)
        self._packed_target = (np.ascontiguousarray(T), self._target_have)     # the job's target does not change from here on
        # place of every row in the reference's `scores` dict when that is not the store's row order (a ragged search set
        # and a target whose splits do not come in ascending order, FeatureStore.dict_order): the seeded sampling, the
        # "first maximum" and the "last confirmed clip" of this job then walk that order
        self._place = st.dict_order(self.target.target_features)
        self._scanned = self._have_sims = False
        self.similarities = _SimilarityView(self)

    def compute_scores(self, weights):
        """Record the stream weights; scores materialise with the next scan (ticket.py:165-180)."""
        self._weights = dict(weights)
        self._scanned = self._have_sims = False
        self.scores = _ScoreView(self)

    def _ids_in_dict_order(self):
        ids, place = self.feature_store().clip_ids, getattr(self, "_place", None)
        return ids if place is None else ids[np.argsort(place)]

    def _eps(self):
        return float(os.environ["COMPUTE_EPS"])

    def _scan(self, threshold, lower_limit, want_sims=False, lists=True):
        st = self.feature_store()
        weights = self._weights if self._weights is not None else (
            self._hp.weights or self._hp.default_weights)
        want_sims = want_sims or self._have_sims
        self.last_scan = st.scan(self.target.target_features, [weights[s] for s in st.streams],
                                 threshold, lower_limit, self._eps(), topk=self.topk, want_sims=want_sims, lists=lists,
                                 packed=getattr(self, "_packed_target", None))
        self._scanned, self._have_sims = True, want_sims
        if isinstance(self.scores, _ScoreView):
            self.scores._arr = None
        if isinstance(self.similarities, _SimilarityView):
            self.similarities._arr = None
        if self.topk:
            rows, sc = st.topk()
            self.ranked = (st.clip_ids[rows - st.first_global_row], sc)
        return self.last_scan

    def _ensure_scan(self, want_sims=False):
        """Run the scan if `.scores` / `.similarities` are read before (or without) a selection."""
        if not self._scanned or (want_sims and not self._have_sims):
            th = self._hp.threshold if self._hp is not None else float("nan")
            self._scan(th, float("nan"), want_sims=want_sims)

    def lowest_scoring_user_match(self):
        """(min(1, scores of user-confirmed clips), last such clip in database order) — ticket.py:301-309.
        Uses the float64 labelled-subset path: no full scan is needed for a handful of clips."""
        st = self.feature_store()
        clips = [int(c) for c, v in self.user_matches.items() if v is True and st.has_clip(int(c))]
        if not clips:
            return 1, None
        place = getattr(self, "_place", None)
        clips.sort(key=st.row_of if place is None else (lambda c: int(place[st.row_of(c)])))
        sims = st.labelled_sims(self.target.target_features, st.rows_of(clips))
        weights = self._weights if self._weights is not None else self._hp.weights
        ssum, denom = np.zeros(len(clips)), 0.0
        for si, s in enumerate(st.streams):
            ssum = ssum + (weights[s] * (1 - sims[:, si])) ** 2
            denom += weights[s] ** 2
        sc = 1 - np.sqrt(ssum / denom)
        return min(1, float(sc.min())), clips[-1]

    # ------------------------------------------------------------------ selection (A5)
    def select_clips_to_review(self, threshold=0.8, max_number_matches=20, near_miss=0.5):
        """Matches and near misses for review (ticket.py:311-356): half from {score >= threshold},
        the rest from {lower <= score < threshold} with the best near miss always kept; the reference
        clip and previously confirmed clips are forced in.  Result: self.matches = {clip: score}."""
        lower_limit = threshold - near_miss * (1 - threshold)
        st = self.feature_store()
        # A review round samples a few dozen clips: the lists stay on the device and only the sampled entries are
        # fetched.  The finalize round (max = inf) returns every clip above the criterion: whole lists come back, and
        # everything below is array work — no per-clip Python objects until the final dict.
        place = getattr(self, "_place", None)                  # not None: the reference's dict order differs from the row order
        sampled = max_number_matches != float("inf") and place is None
        res = self._scan(threshold, lower_limit, lists=not sampled)
        ids, first = st.clip_ids, st.first_global_row
        t_rows, _ = st.ties(copy=False)
        self.tie_band = ids[t_rows - first].tolist()
        if self.tie_band:
            logging.info("query %s: %d clip(s) within COMPUTE_EPS of a selection boundary: %s",
                         self.query_id, len(self.tie_band), self.tie_band[:20])
        mscores = int(min(max_number_matches / 2, res.n_match))
        m_near_scores = int(min(max_number_matches - mscores, res.n_near))
        # random.sample(population, k) draws depend only on (len(population), k): sampling the index
        # range consumes the generator exactly like sampling the reference's dict items (:333).
        picked = sample_range(random, res.n_match, mscores)
        order_m = order_n = None
        if not sampled:
            all_rows, all_sc = st.matches(copy=False)     # views of the scan's host mirror: consumed before the next scan
            nm_rows, nm_sc = st.near_misses(copy=False)
            if place is not None:                          # the lists in the reference's dict order
                order_m = np.argsort(place[all_rows - first], kind="stable")
                order_n = np.argsort(place[nm_rows - first], kind="stable")
                all_rows, all_sc, nm_rows, nm_sc = all_rows[order_m], all_sc[order_m], nm_rows[order_n], nm_sc[order_n]
            m_rows, m_sc = all_rows[picked], all_sc[picked]
        jbest, best_row, best_sc = None, None, None
        n_left = res.n_near
        if m_near_scores > 0:
            m_near_scores -= 1
            n_left -= 1
            if sampled:
                jbest, best_row, best_sc = st.near_best()  # first maximum in database order (:338), found on the device
            else:
                jbest = int(np.argmax(nm_sc))
                best_row, best_sc = int(nm_rows[jbest]), float(nm_sc[jbest])
        picked_n = sample_range(random, n_left, m_near_scores)
        # positions in the list with the best near miss deleted (:340) -> positions in the full list
        pos = picked_n if jbest is None else picked_n + (picked_n >= jbest)
        if sampled:                                        # one round trip for both lists' sampled entries
            (m_rows, m_sc), (n_rows_p, n_sc_p) = st.gather_many([("matches", picked), ("near_misses", pos)])
        else:
            n_rows_p, n_sc_p = nm_rows[pos], nm_sc[pos]
        sel_rows = np.concatenate([m_rows, n_rows_p] + ([[best_row]] if best_row is not None else [])).astype(np.int64)
        sel_sc = np.concatenate([m_sc, n_sc_p] + ([[best_sc]] if best_row is not None else []))
        sel_ids = ids[sel_rows - first]
        self.matches = dict(zip(sel_ids.tolist(), sel_sc.tolist()))
        forced = []
        if st.has_clip(self.ref_clip_id):
            forced.append(self.ref_clip_id)
        if self.user_matches:
            forced += [int(clip) for clip, value in self.user_matches.items() if value is True]
        if forced:
            f_sc = st.scores_at(st.rows_of(forced))        # one round trip for all forced clips (:346-356)
            self.matches.update(zip(forced, f_sc.tolist()))
        # what the final report needs to rank on the device (ranked_selection): the lists' entries in selection order
        # (as indices into the device's row-ordered lists; a review round on the whole-list path keeps nothing)
        self._selection = None
        if max_number_matches == float("inf"):
            row_idx = lambda order, j: j if order is None else order[j]
            self._selection = {"n_match": res.n_match, "n_near": res.n_near, "picked": row_idx(order_m, picked),
                               "pos": row_idx(order_n, pos), "jbest": None if jbest is None else int(row_idx(order_n, jbest)),
                               "n_listed": len(sel_ids)}

    def ranked_selection(self):
        """[(clip id, score)] of the last finalize selection in REPORT order (see ranked_selection_arrays); the list of
        pairs costs one Python tuple per clip, so consumers that only walk the order take the arrays."""
        r_ids, r_sc = self.ranked_selection_arrays()
        return list(zip(r_ids.tolist(), r_sc.tolist()))

    def ranked_selection_arrays(self):
        """(clip ids, scores) of the last finalize selection in REPORT order: score descending, ties in the order the
        selection inserted the clips (the reference's stable sort of its dict's items, ticket.py:266).  The match and
        near-miss lists are ranked on the device (vq_rank_list, K7) with each entry's place in the selection order as the
        tie-break; forced clips that sit in neither list (a reference clip or a confirmed clip below the band) are few
        and are merged in here.  Every match outranks every near miss, so the two ranked lists are simply concatenated."""
        sel = getattr(self, "_selection", None)
        if sel is None:
            raise VQError("ranked_selection: the last selection was not a finalize round (max_number_matches = inf)")
        st = self.feature_store()
        ids, first = st.clip_ids, st.first_global_row
        n_m, n_n = sel["n_match"], sel["n_near"]
        place_m = np.empty(n_m, np.uint32)
        place_m[sel["picked"]] = np.arange(n_m, dtype=np.uint32)              # list entry -> place in the selection order
        place_n = np.empty(n_n, np.uint32)
        if n_n:
            place_n[sel["pos"]] = np.arange(n_n - 1, dtype=np.uint32)
            place_n[sel["jbest"]] = n_n - 1                                    # the best near miss is inserted last (:343)
        m_rows, m_sc = st.matches(copy=False)
        nm_rows, nm_sc = st.near_misses(copy=False)
        om, sm = st.rank_list("matches", place_m)
        on, sn = st.rank_list("near_misses", place_n)
        inv_m = np.empty(n_m, np.int64)
        inv_m[place_m] = np.arange(n_m)
        inv_n = np.empty(n_n, np.int64)
        inv_n[place_n] = np.arange(n_n)
        r_ids = np.concatenate([ids[m_rows[inv_m[om]] - first], ids[nm_rows[inv_n[on]] - first]])
        r_sc = np.concatenate([sm, sn])
        listed = n_m + n_n
        if len(self.matches) > listed:                                         # forced clips outside both lists, in dict order
            extra = list(itertools.islice(self.matches.items(), listed, None))
            x_ids = np.array([c for c, _ in extra], dtype=r_ids.dtype)
            x_sc = np.array([v for _, v in extra], dtype=r_sc.dtype)
            # an extra clip lands after every listed clip of equal score (it was inserted later) and extras of equal score
            # keep their own order: a stable sort of the few extras, then one merge into the already ordered listed part
            xo = np.argsort(-x_sc.astype(np.float64), kind="stable")
            x_ids, x_sc = x_ids[xo], x_sc[xo]
            at = np.searchsorted(-r_sc.astype(np.float64), -x_sc.astype(np.float64), side="right")
            r_ids, r_sc = np.insert(r_ids, at, x_ids), np.insert(r_sc, at, x_sc)
        return r_ids, r_sc

    def _score_of(self, clip):
        st = self.feature_store()
        if not st.has_clip(clip):
            raise KeyError(clip)
        return float(st.scores_at(st.rows_of([clip]))[0])

    # ------------------------------------------------------------------ final report (A7)
    def create_final_report(self, hyperparameters, query_result_id):
        """CSV of every selected clip ranked by score, stable and descending (ticket.py:182-274)."""
        query = self._request(["queries", "read"], {"id": self.query_id})
        video = self._request(["videos", "read"], {"id": self.video_id})
        query_result = self._request(["query-results", "read"], {"id": query_result_id})
        search_set = self._request(["search-sets", "read"], {"id": query["search_set_to_query"]})
        out_dir = '../final_reports/'
        os.makedirs(out_dir, exist_ok=True)
        path = os.path.join(out_dir, 'final_report_query_{}_{}.csv'.format(
            query["name"], datetime.now().strftime('%m-%d-%Y_%Hh%Mm%Ss')))
        hp = hyperparameters
        header = [
            ['Query:', query["name"], 'Query pk:', self.query_id],
            ['Search Set queried:', search_set["name"], 'Search set pk:', search_set["id"]],
            ['Reference Video:', video["name"], 'Video pk:', self.video_id],
            ['Reference time:', query["reference_time"]],
            ['number of reviews:', query_result["round"] - 1],
            ['min score for a match:', query_result["match_criterion"]],
            ["max matches to review:", query["max_matches_for_review"]],
            ['streams:', str(hp.streams)],
            ['stream weights:', str(query_result["weights"])],
            ['Target bootstrapping:', query["use_dynamic_target_adjustment"]],
            ['query notes:', query["notes"]],
            ['Hyperparameters:'],
            ['', 'default weights:', str(hp.default_weights)],
            ['', 'default threshold:', str(hp.default_threshold)],
            ['', 'near miss default:', str(hp.near_miss_default)],
            ['', 'feature name:', str(hp.feature_name)],
            ['', 'ballast:', str(hp.ballast)],
            ['', 'mu:', str(hp.mu)],
            ['', 'f_bootstrap:', str(hp.f_bootstrap)],
            ['', 'f_memory:', str(hp.f_memory)],
            ['', 'bootstrap type:', str(hp.bootstrap_type)],
        ]
        if hp.bootstrap_type == "bagging":
            header.append(['', 'number of bags:', str(hp.nbags)])
        header += [[''],
                   ['List of all clips with scores greater than min(threshold, score of lowest scoring'
                    ' user validated match)'],
                   ['clip #', 'start time', 'match type', 'video pk', 'video clip id', 'score', 'duration', 'notes']]
        # report order = stable descending sort of the selection (ticket.py:266), ranked on the device for a finalize
        # selection; any other producer of self.matches gets the same order from the host sort below
        if getattr(self, "_selection", None) is not None and self._selection["n_listed"] <= len(self.matches):
            r_ids, r_sc = self.ranked_selection_arrays()
            ordered, presorted = zip(r_ids.tolist(), r_sc.tolist()), True
        else:
            ordered, presorted = self.matches.items(), False
        rows = []
        for video_clip_id, score in ordered:
            label = self.user_matches.get(str(video_clip_id))
            if str(video_clip_id) in self.user_matches:
                match_type = "user-identified match" if label is True else "user-identified non-match"
            elif score >= query_result["match_criterion"]:
                match_type = "inferred match"
            else:
                match_type = "inferred non-match"
            clip = self._request(["video-clips", "read"], {"id": video_clip_id})
            match = self._request(["matches", "list"], {"query_result": query_result_id, "video_clip": video_clip_id})
            start = int(match["results"][0]["match_video_time_span"].split(",")[0])
            rows.append([clip['clip'], str(timedelta(seconds=start)), match_type, clip['video'], video_clip_id,
                         score, clip['duration'], clip['notes']])
        if not presorted:
            rows.sort(key=lambda r: r[5], reverse=True)
        with open(path, 'x', newline='') as f:
            w = csv.writer(f)
            w.writerows(header)
            w.writerows(rows)
        with open(path, 'r') as f:
            self._post_file(["queries", "partial_update"], {"id": self.query_id, "final_report_file": f})

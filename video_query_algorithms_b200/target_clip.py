"""Target construction and bootstrapping from user-labelled clips.

Mirror of reference `src/models/target_clip.py` (same class, attributes and method names).  The
reference pulls every labelled clip's features over HTTP (one call per clip, :131) and solves on
the CPU; here the labelled rows are already in the HBM store, so only their row numbers travel and
`vq_bootstrap_target` does the Gram matrices, the solve and the recombination in float64 on the
GPU.  The seeded resampling stays on the host with Python's `random`, in the reference's order.
"""
from __future__ import annotations

import logging
import random
from time import sleep

import numpy as np

from ._rng import choices_range

try:                                        # requests is only needed to recognise its ConnectionError
    from requests import ConnectionError
except Exception:                           # pragma: no cover
    ConnectionError = OSError


class TargetClip:
    def __init__(self, ticket, hyperparameters):
        self.client = ticket.client
        self.schema = ticket.schema
        self.ticket = ticket
        self.bootstrap_target = ticket.dynamic_target_adjustment
        self.latest_query_result = ticket.latest_query_result
        self.hyperparameters = hyperparameters
        ticket.attach_store(hyperparameters)        # labelled rows are read from the resident store
        self.ref_clip_features, self.splits = self._get_clip_features(ticket.ref_clip_id)
        self.previous_target_features = None
        self.target_features = {}
        if ticket.latest_query_result:
            if ticket.latest_query_result["bootstrapped_target"]:
                self.previous_target_features = ticket.latest_query_result["bootstrapped_target"]

    # ------------------------------------------------------------------ public contract
    def get_target_features(self):
        """Fill self.target_features = {stream: {split: [floats]}} — case analysis of
        target_clip.py:26-73."""
        if not self.bootstrap_target or self.latest_query_result is None:
            self.target_features = self.scaled_ref_clip_features()
            return
        valid_rows, splits = self.rows_for_matches(user_match_value=True)
        invalid_rows, __ = self.rows_for_matches(user_match_value=False)
        if len(valid_rows) == 0:
            self.target_features = self.scaled_ref_clip_features()
            return
        kind = self.hyperparameters.bootstrap_type
        if kind == 'simple':
            self.target_features = self.dynamic_target_adjustment(
                valid_rows, invalid_rows, splits, self.hyperparameters.f_bootstrap, replacement=False)
        elif kind == 'partial_update':
            self.target_features = self.dynamic_target_adjustment(
                valid_rows, invalid_rows, splits, self.hyperparameters.f_bootstrap, replacement=False)
            self.avg_new_old_targets(splits)
        elif kind == 'bagging':
            self.target_by_bagging(valid_rows, invalid_rows, splits)
        else:
            raise Exception("Error: bootstrap_type should be one of 'simple', 'partial_update', or 'bagging'")

    def scaled_ref_clip_features(self):
        """f / (f . f) per (stream, split) — target_clip.py:137-143, :311-313."""
        return {stream: {split: self._scale_feature(np.asarray(f, dtype=np.float64)).tolist()
                         for split, f in by_split.items()}
                for stream, by_split in self.ref_clip_features.items()}

    def avg_new_old_targets(self, splits):
        """f_memory * new + (1 - f_memory) * old (target_clip.py:75-82).  Stored as lists: the
        reference leaves ndarrays here, which makes its own json.dumps (ticket.py:296) raise."""
        if not self.previous_target_features:
            return
        f = self.hyperparameters.f_memory
        for stream in self.hyperparameters.streams:
            for split in splits:
                new = np.multiply(f, self.target_features[stream][split])
                old = np.multiply((1 - f), self.previous_target_features[stream][split])
                self.target_features[stream][split] = (new + old).tolist()

    def dynamic_target_adjustment(self, valid_rows, invalid_rows, splits, b_fraction, replacement=False):
        """New target from labelled rows (target_clip.py:84-103).  RNG consumption matches the
        reference: with invalid labels both lists are resampled, valid first (:227-230); without,
        a draw happens only if b_fraction != 1 or replacement (:181-182)."""
        store = self.ticket.feature_store()
        mu = self.hyperparameters.mu
        if len(invalid_rows):
            valid_rows = self._random_fraction(valid_rows, b_fraction, replacement)
            invalid_rows = self._random_fraction(invalid_rows, b_fraction, replacement)
        elif b_fraction != 1 or replacement is True:
            valid_rows = self._random_fraction(valid_rows, b_fraction, replacement)
        w = self._solve_slots(store, valid_rows, invalid_rows, mu)        # [S, P, dim] float64
        return self._as_feature_dict(w, store, splits)

    @staticmethod
    def _solve_slots(store, valid_rows, invalid_rows, mu):
        """One `vq_bootstrap_target` call solves every (stream, split) slot from the same labelled rows.  On a ragged
        search set a labelled clip may lack a slot (a zero row in the store); the reference then solves that slot with
        the clips that do have it — its per-slot feature lists only receive the rows a clip has
        (target_clip.py:184-187, 232-240) — so the slot's rows are filtered the same way, one call per slot with only
        that slot solved (the other slots of such a call would see zero rows: singular systems)."""
        valid_rows, invalid_rows = np.asarray(valid_rows, np.int64), np.asarray(invalid_rows, np.int64)
        if store.present is None:
            return store.bootstrap_target(valid_rows, invalid_rows, mu)
        has_v = store.present[valid_rows - store.first_global_row]
        has_i = store.present[invalid_rows - store.first_global_row]
        if has_v.all() and has_i.all():
            return store.bootstrap_target(valid_rows, invalid_rows, mu)
        w = np.zeros(has_v.shape[1:] + (store.dim,), np.float64)
        for si in range(has_v.shape[1]):
            for pi in range(has_v.shape[2]):
                v, iv = valid_rows[has_v[:, si, pi]], invalid_rows[has_i[:, si, pi]]
                if len(v):                                                 # no labelled clip has the slot: it stays zero
                    only = np.zeros(has_v.shape[1:], bool)
                    only[si, pi] = True
                    w[si, pi] = store.bootstrap_target(v, iv, mu, slots=only)[si, pi]
        return w

    def target_by_bagging(self, valid_rows, invalid_rows, splits):
        """Mean of nbags targets, each from a with-replacement resample (target_clip.py:145-159)."""
        store = self.ticket.feature_store()
        bags = []
        for _ in range(self.hyperparameters.nbags):
            d = self.dynamic_target_adjustment(valid_rows, invalid_rows, splits, b_fraction=1, replacement=True)
            bags.append(d)
        self.target_features = {}
        for stream in self.hyperparameters.streams:
            self.target_features[stream] = {}
            for split in splits:
                stack = [bags[b][stream][split] for b in range(self.hyperparameters.nbags)]
                self.target_features[stream][split] = np.average(stack, axis=0).tolist()

    def rows_for_matches(self, user_match_value=True):
        """Store rows of the latest round's matches labelled `user_match_value`, in the order the API
        lists them (the reference's features_for_matches, target_clip.py:105-135, minus the
        per-clip feature download), plus the splits those clips have."""
        page, matches = 1, []
        while page is not None:
            results = self._request(["matches", "list"],
                                    {"query_result": self.latest_query_result["id"], "page": page})
            matches.extend(results["results"])
            page = results["pagination"]["nextPage"]
        store = self.ticket.feature_store()
        rows, splits = [], set()
        for match in matches:
            if match["user_match"] is user_match_value:
                r = store.first_global_row + store.row_of(match["video_clip"])
                rows.append(r)
                if store.present is None:
                    splits.update(store.splits)
                else:
                    loc = r - store.first_global_row
                    splits.update(p for pi, p in enumerate(store.splits) if store.present[loc, :, pi].any())
        return np.array(rows, dtype=np.int64), splits

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _as_feature_dict(w, store, splits):
        return {stream: {split: w[si, store.splits.index(split)].tolist() for split in splits}
                for si, stream in enumerate(store.streams)}

    def _get_clip_features(self, clip_id):
        """{stream: {split: [floats]}} and the set of splits for one clip (target_clip.py:263-286).
        One API call, as in the reference; with ticket.features_from_store the row is read back from
        HBM instead (fp32-rounded), which is what the synthetic benchmarks use."""
        streams = self.hyperparameters.streams
        store = self.ticket.feature_store(optional=True)
        if self.ticket.features_from_store and store is not None and store.has_clip(clip_id):
            loc = store.row_of(clip_id)
            row = store.download(loc, 1)[0].astype(np.float64)
            results, splits = {s: {} for s in streams}, set()
            for si, s in enumerate(store.streams):
                for pi, p in enumerate(store.splits):
                    if store.present is None or store.present[loc, si, pi]:
                        results[s][p] = row[si, pi].tolist()
                        splits.add(p)
            return results, splits
        results, splits = {s: {} for s in streams}, set()
        for fo in self._request(["video-clips", "features"], {"id": clip_id}):
            s = fo["dnn_stream_id"]
            if s in streams and fo["name"] == self.hyperparameters.feature_name:
                splits.add(fo["dnn_stream_split"])
                results[s][fo["dnn_stream_split"]] = fo["feature_vector"]
        return results, splits

    def _request(self, action, params):
        while True:
            try:
                return self.client.action(self.schema, action, params=params)
            except ConnectionError:
                sleep(0.05)
                logging.warning('Try API request by Target again: action = {}, params = {}'.format(action, params))

    @staticmethod
    def _random_fraction(flist, fraction, replacement):
        """Seeded subsample, duplicates dropped (target_clip.py:297-309)."""
        n = len(flist)
        t = max(round(n * fraction), 1)
        if replacement is False:
            picks = random.sample(range(n), t)
        else:
            picks = choices_range(random, n, t).tolist()          # same stream as random.choices, vectorised
        picks = list(set(picks))
        return np.asarray([flist[m] for m in picks], dtype=np.int64)

    @staticmethod
    def _scale_feature(f):
        return f / np.dot(f, f)

"""Weights / threshold holder and their update from user-labelled matches.

Mirror of reference `src/models/hyperparameter.py`: same constructor, same attributes, same
`optimize_weights(ticket)` contract.  The 40 x 31 loss grid over the labelled clips
(hyperparameter.py:56-65) runs on the GPU in float64 (`vq_labelled_sims` + `vq_loss_grid`); it needs
only the labelled rows, not the 40 full-database rescoring passes the reference makes (:58).
`argmin`, the border rule and the parabola fit (:66-114) stay on the host in float64.
"""
from __future__ import annotations

import logging
import os

import numpy as np

from ._rng import choices_range, set_order

from . import store as _store


def eps_threshold():
    """COMPUTE_EPS, read from the environment like the reference (hyperparameter.py:5)."""
    return float(os.environ["COMPUTE_EPS"])


class Hyperparameter:
    def __init__(self, default_weights, default_threshold=0.8, ballast=0.3, near_miss_default=0.5, mu=.3,
                 streams=('rgb', 'warped_optical_flow'), feature_name='global_pool', f_bootstrap=0.5,
                 f_memory=0.5, bootstrap_type='simple', nbags=3):
        self.default_weights = default_weights
        self.weights = {}
        self.default_threshold = default_threshold
        self.threshold = self.default_threshold
        self.near_miss_default = near_miss_default
        self.streams = streams
        self.feature_name = feature_name
        self.ballast = ballast
        self.weight_grid = np.arange(0.5, 2.5, 0.05)        # 40 points, hyperparameter.py:20
        self.threshold_grid = np.arange(0.5, 1.1, 0.02)     # 31 points, hyperparameter.py:21
        self.mu = mu
        self.f_bootstrap = f_bootstrap
        self.f_memory = f_memory
        self.bootstrap_type = bootstrap_type
        self.nbags = nbags
        self.losses = None            # last loss grid(s), kept for inspection

    # ------------------------------------------------------------------ A8
    @staticmethod
    def match_status(matches):
        """Ordered {clip: label}: the user's label where given, else is_match (hyperparameter.py:45-50)."""
        status = {}
        for match in matches:
            if match["user_match"] is not None:
                status[match["video_clip"]] = match["user_match"]
            else:
                status[match["video_clip"]] = match["is_match"]
        return status

    def labelled_similarities(self, ticket, clips):
        fs = ticket.feature_store()
        return fs.labelled_sims(ticket.target.target_features, fs.rows_of(clips))

    def optimize_weights(self, ticket):
        """Grid-search (flow weight, threshold) minimising the labelled-match loss, then fine tune.
        Sets self.weights and self.threshold exactly as hyperparameter.py:66-76 does."""
        status = self.match_status(ticket.matches)
        clips = list(status)
        labels = np.array([bool(status[c]) for c in clips])
        sims = self.labelled_similarities(ticket, clips)
        device = ticket.feature_store().shards[0].device
        losses = _store.loss_grid(sims, labels, self.weight_grid, self.threshold_grid, self.ballast,
                                  device=device)[0]
        self.losses = losses
        w, th = self.optimum(losses)
        self.threshold = th - eps_threshold()
        self.weights = {self.streams[0]: 1.0, self.streams[1]: w}

    def optimize_weights_replicates(self, ticket, replicates):
        """Bootstrap search (BASELINE config 5): the same update for R resampled index sets at once.
        replicates: list of index arrays into the ordered labelled-clip list (drawn by the caller
        from Python's `random`, see resample_labelled).  Returns (weights [R], thresholds [R])."""
        status = self.match_status(ticket.matches)
        clips = list(status)
        labels = np.array([bool(status[c]) for c in clips])
        sims = self.labelled_similarities(ticket, clips)
        device = ticket.feature_store().shards[0].device
        losses = _store.loss_grid(sims, labels, self.weight_grid, self.threshold_grid, self.ballast,
                                  replicates=replicates, device=device)
        self.losses = losses
        out_w, out_th = np.empty(len(replicates)), np.empty(len(replicates))
        for r in range(len(replicates)):
            w, th = self.optimum(losses[r])
            out_w[r], out_th[r] = w, th - eps_threshold()
        return out_w, out_th

    # ------------------------------------------------------------------ A9
    def optimum(self, losses):
        iw0, ith0 = np.unravel_index(np.argmin(losses, axis=None), losses.shape)
        nw, nth = len(self.weight_grid), len(self.threshold_grid)
        if iw0 == 0 or ith0 == 0 or iw0 == nw - 1 or ith0 == nth - 1:
            return self.weight_grid[iw0], self.threshold_grid[ith0]
        return self.fine_tune(iw0, ith0, losses)

    def fine_tune(self, iw0, ith0, losses):
        wg, tg = self.weight_grid, self.threshold_grid
        x = [(wg[iw0 - 1], wg[iw0], wg[iw0 + 1]), (tg[ith0 - 1], tg[ith0], tg[ith0 + 1])]
        y = [losses[iw0 - 1, ith0], losses[iw0, ith0 - 1], losses[iw0, ith0], losses[iw0, ith0 + 1],
             losses[iw0 + 1, ith0]]
        return self._quad_fit(x, y)

    @staticmethod
    def _quad_fit(x, y):
        """Vertex of a0 (w - w0)^2 + b0 (th - th0)^2 + c0 through the five losses around the grid
        minimum, clamped to the neighbouring grid points; falls back to the grid point when the fit
        misses the data by more than 1e-6 (hyperparameter.py:85-114)."""
        def vertex(xs, y_lo, y_mid, y_hi):
            xl, xm, xh = xs
            num = (y_hi - y_lo) * xm ** 2 + (y_mid - y_hi) * xl ** 2 - (y_mid - y_lo) * xh ** 2
            den = (y_hi - y_lo) * xm + (y_mid - y_hi) * xl - (y_mid - y_lo) * xh
            v = 0.5 * num / den
            curv = (y_mid - y_lo) / ((xm - v) ** 2 - (xl - v) ** 2)
            return v, curv

        w0, a0 = vertex(x[0], y[0], y[2], y[4])
        th0, b0 = vertex(x[1], y[1], y[2], y[3])
        c0 = y[2] - a0 * (x[0][1] - w0) ** 2 - b0 * (x[1][1] - th0) ** 2
        w0 = max(min(w0, x[0][2]), x[0][0])
        th0 = max(min(th0, x[1][2]), x[1][0])
        model = lambda w, th: a0 * (w - w0) ** 2 + b0 * (th - th0) ** 2 + c0
        fit = [model(x[0][0], x[1][1]), model(x[0][1], x[1][0]), model(x[0][1], x[1][1]),
               model(x[0][1], x[1][2]), model(x[0][2], x[1][1])]
        if sum(abs(y[i] - fit[i]) for i in range(5)) > 10 ** -6:
            logging.warning("hyperparameter quadratic fine tuning failed - resort to selecting optimum on grid "
                            "without further interpolation")
            w0, th0 = x[0][1], x[1][1]
        return w0, th0


def resample_labelled(n_labelled, n_replicates, rng):
    """Replicate index sets exactly as the reference's bagging draws them: `random.choices(range(n), k=n)`
    then `list(set(...))` (target_clip.py:297-309 with fraction 1, replacement True)."""
    draws = choices_range(rng, n_labelled, max(round(n_labelled * 1), 1), repeats=n_replicates)   # same stream as the
    draws = draws.reshape(n_replicates, -1)                                                        # reference's choices calls
    return [set_order(d, n_labelled).astype(np.int32) for d in draws]                            # the reference's set order

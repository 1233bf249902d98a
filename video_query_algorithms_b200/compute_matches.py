"""Public entry point of the scoring path: `compute_matches(query_updates, hyperparameters)`, the call the broker
makes once per tick (reference src/broker.py:87) with the signature and the observable behaviour of the reference's
driver (src/models/compute_matches.py:8-107) — the same Ticket / TargetClip / Hyperparameter calls with the same
arguments in the same order on every branch, which `tests/test_flow_cpu.py` checks against the reference's own driver.

A tick is a sequence of independent jobs (revise, new, finalize, in the repository's order, sharing one RNG stream);
each job is a short pipeline whose stages are the functions below.  Everything heavy happens behind the ticket:
`compute_similarities` / `compute_scores` bind the target and the weights, `select_clips_to_review` issues the fused
GPU scan (ticket.py), `optimize_weights` the labelled-subset kernels (hyperparameter.py).
"""
from __future__ import annotations

import os

from .target_clip import TargetClip
from .ticket import Ticket

# process states of a query in the API database (compute_matches.py:43,49,101,104; ticket.py:113-118)
IN_PROGRESS, PROCESSED, FAILED, FINALIZED = 3, 4, 5, 7


def compute_matches(query_updates, hyperparameters, ticket_factory=None):
    """Run every pending job of this tick.  `query_updates` needs `.get_status()` -> {kind: job dict or None} and
    `.url` (APIRepository); `ticket_factory(job, url)` lets embedders inject an API client or a pre-built store."""
    make_ticket = ticket_factory or (lambda job, url: Ticket(job, url))
    for kind, job in query_updates.get_status().items():
        if job is not None:
            _run_job(kind, job, make_ticket(job, query_updates.url), hyperparameters)


def _run_job(kind, job, ticket, hp):
    ticket.change_process_state(IN_PROGRESS)
    fatal, recoverable = ticket.catch_errors(kind)                   # compute_matches.py:47-53
    if fatal:
        ticket.change_process_state(FAILED, message=fatal)
        return
    if recoverable:
        ticket.add_note(recoverable)

    ticket.target = TargetClip(ticket, hp)                           # the reference clip, or bootstrapped from labels
    ticket.target.get_target_features()
    ticket.compute_similarities(hp)
    _set_criteria(kind, job, ticket, hp)
    result_id = ticket.create_query_result(_round_number(kind, ticket), hp)

    ticket.compute_scores(hp.weights)
    max_matches, near_miss = _review_window(kind, ticket, hp)
    ticket.select_clips_to_review(hp.threshold, max_matches, near_miss)
    if not ticket.matches:
        catch_no_matches_error(ticket)
        return

    ticket.add_matches_to_database(result_id)
    if kind == "finalize":
        ticket.create_final_report(hp, result_id)
    ticket.change_process_state(FINALIZED if kind == "finalize" else PROCESSED)


def _set_criteria(kind, job, ticket, hp):
    """Weights and threshold for this round (compute_matches.py:61-67): the defaults for a new query or when the
    previous round left no matches, otherwise the optimum over the user-labelled matches."""
    if kind == "new" or not job["matches"]:
        hp.weights, hp.threshold = hp.default_weights, hp.default_threshold
    elif kind in ("revise", "finalize"):
        hp.optimize_weights(ticket)
    else:
        raise Exception('update type is invalid')


def _round_number(kind, ticket):
    return 1 if kind == "new" else ticket.latest_query_result["round"] + 1      # compute_matches.py:70-73


def _review_window(kind, ticket, hp):
    """(max clips to select, near-miss band) — compute_matches.py:78-88.  A review round shows the configured number
    of clips; the final report takes every clip down to min(threshold, lowest user-confirmed score): no size limit
    and a band that reaches that score (COMPUTE_EPS guards the division when the threshold sits at 1)."""
    if kind != "finalize":
        return ticket.number_of_matches_to_review, hp.near_miss_default
    low_score, _ = ticket.lowest_scoring_user_match()
    band = max(hp.threshold - low_score, 0) / max(1 - hp.threshold, float(os.environ["COMPUTE_EPS"]))
    return float("inf"), band


def catch_no_matches_error(ticket):
    """An empty selection is a failed round (compute_matches.py:91-93,110-113)."""
    mround = ticket.latest_query_result["round"] if ticket.latest_query_result else 1
    error_message = "*** Error: No matches were found for round {} of query {}! ***".format(mround, ticket.query_id)
    ticket.change_process_state(FAILED, message=error_message)

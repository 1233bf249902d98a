"""Public entry point of the scoring path — same signature and flow as the reference's
`compute_matches(query_updates, hyperparameters)` (src/models/compute_matches.py:8), called by the
broker once per tick (src/broker.py:87)."""
from __future__ import annotations

import os

from .target_clip import TargetClip
from .ticket import Ticket


def compute_matches(query_updates, hyperparameters, ticket_factory=None):
    """For every pending job (revise, new, finalize — one shared RNG stream, in that order): build
    the ticket and target, score the search set, update weights/threshold from labelled matches,
    store the new query result, select clips for review and persist them; finalize jobs also get
    their report.  `query_updates` needs `.get_status()` and `.url` (APIRepository).
    ticket_factory(update_object, url) lets embedders inject an API client or a pre-built store."""
    make_ticket = ticket_factory or (lambda job, url: Ticket(job, url))
    updates_needed = query_updates.get_status()
    for update_type, update_object in updates_needed.items():
        if update_object is None:
            continue
        ticket = make_ticket(update_object, query_updates.url)
        ticket.change_process_state(3)                               # in progress

        fatal_error_message, error_message = ticket.catch_errors(update_type)
        if fatal_error_message:
            ticket.change_process_state(5, message=fatal_error_message)
            continue
        if error_message:
            ticket.add_note(error_message)

        ticket.target = TargetClip(ticket, hyperparameters)
        ticket.target.get_target_features()
        ticket.compute_similarities(hyperparameters)

        if (update_type == "new") or not update_object["matches"]:
            hyperparameters.weights = hyperparameters.default_weights
            hyperparameters.threshold = hyperparameters.default_threshold
        elif update_type == "revise" or update_type == "finalize":
            hyperparameters.optimize_weights(ticket)
        else:
            raise Exception('update type is invalid')

        new_round = 1 if update_type == 'new' else ticket.latest_query_result["round"] + 1
        new_result_id = ticket.create_query_result(new_round, hyperparameters)

        ticket.compute_scores(hyperparameters.weights)
        if update_type == "finalize":
            max_number_matches = float("inf")
            low_score, __ = ticket.lowest_scoring_user_match()
            near_miss = max(hyperparameters.threshold - low_score, 0) / \
                max(1 - hyperparameters.threshold, float(os.environ["COMPUTE_EPS"]))
        else:
            max_number_matches = ticket.number_of_matches_to_review
            near_miss = hyperparameters.near_miss_default
        ticket.select_clips_to_review(hyperparameters.threshold, max_number_matches, near_miss)

        if not ticket.matches:
            catch_no_matches_error(ticket)
            continue

        ticket.add_matches_to_database(new_result_id)
        if update_type == "finalize":
            ticket.create_final_report(hyperparameters, new_result_id)
            ticket.change_process_state(7)
        else:
            ticket.change_process_state(4)


def catch_no_matches_error(ticket):
    mround = ticket.latest_query_result["round"] if ticket.latest_query_result else 1
    error_message = "*** Error: No matches were found for round {} of query {}! ***".format(mround, ticket.query_id)
    ticket.change_process_state(5, message=error_message)

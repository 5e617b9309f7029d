"""Lifecycle state of one shard and how a cluster of shards reports it.

Behavioural mirror of the reference's `IndexState` (distributed_faiss/index_state.py:11-36):
same member names and values, same aggregation answers
(tests/test_index_state.py:14-22); search is only legal in TRAINED (index.py:247).
"""
from enum import Enum
from typing import Iterable


class IndexState(Enum):
    NOT_TRAINED = 1
    TRAINING = 2
    ADD = 3
    TRAINED = 4

    @staticmethod
    def get_aggregated_states(states: Iterable["IndexState"]) -> "IndexState":
        """One state for the whole cluster: unanimous -> that state; otherwise the
        'least finished' one wins in the order TRAINING, NOT_TRAINED, ADD, TRAINED."""
        seen = set(states)
        if not seen:
            raise AssertionError("no shard states to aggregate")
        for candidate in (IndexState.TRAINING, IndexState.NOT_TRAINED, IndexState.ADD):
            if candidate in seen and len(seen) > 1:
                return candidate
        return next(iter(seen)) if len(seen) == 1 else IndexState.TRAINED

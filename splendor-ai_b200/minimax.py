"""The reference's search agent on the CUDA engine: ``splendor.agents.our_agents.minmax.MiniMaxAgent`` (minmax.py:22-139).

Same class name, constructor and ``SelectAction(actions, game_state, game_rule)`` contract, and the module exports
``myAgent`` as the reference's agent loader expects (general_game_runner.py:174-175).  The depth-2 walk the reference does
one ``generateSuccessor`` / ``generatePredecessor`` pair at a time runs as ONE kernel call (``spl_minimax2``): a thread
per root action plays it, enumerates the opponent's replies in registers and keeps the minimum of the evaluation.

Differences worth knowing: the value of every action is bit-identical to the reference's evaluation of the unpermuted
state; among exactly equal best actions the reference returns the first in (type name, ``random.shuffle``) order, this
class the first in (type name, ALL_ACTIONS index) order - deterministic, and one of the reference's possible answers.
For whole batches use ``BatchedSplendor.minimax2()``."""
from __future__ import annotations

from . import template
from .game_rule import action_index

EXPECTED_AMOUNT_OF_PLAYERS = 2  # minmax.py:16
DEPTH = 2                       # minmax.py:17


class MiniMaxAgent(template.Agent):
    def SelectAction(self, actions, game_state, game_rule):
        from . import _single

        assert len(game_state.agents) == EXPECTED_AMOUNT_OF_PLAYERS
        best, self.last_value, self.last_actions, self.last_values = _single.minimax2(
            game_state._words, game_state._deck_lists, game_rule.num_of_agent, self.id)
        for a in actions:
            if action_index(a, game_state, self.id) == best:
                return a
        raise ValueError(f"the search chose action {best}, which is not among the actions offered")

    def evaluate(self, game_state) -> float:
        """_evaluation_function(state) (minmax.py:90-136) from this agent's point of view."""
        from . import _single

        return _single.minimax_eval(game_state._words, len(game_state.agents), self.id)


myAgent = MiniMaxAgent  # pylint: disable=invalid-name

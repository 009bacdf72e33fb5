"""The agent / rule contract that callers of the reference program against (reference: ``splendor/template.py:17-89``).

Only the surface is shared with the reference - class names, attribute names and method names - so that agent modules
written for it (``SelectAction(actions, game_state, game_rule)``, a module-level ``myAgent`` factory) and drivers such as
``Game.Run`` work unchanged on the B200 rule classes:

  GameRule   num_of_agent, current_agent_index, current_game_state, action_counter;
             update(action) = generateSuccessor on the current state for the seat to move, then advance the seat
  Agent      ``Agent(_id)``, ``.id``, ``SelectAction`` (uniform random by default)
  GameState  marker base class of rule-specific states

An abstract hook that a subclass forgot to provide reports itself and terminates the process with status 1, which is what
the reference's ``utils.raiseNotDefined`` does (``splendor/utils.py:7-13``).
"""
from __future__ import annotations

import random
import sys
import traceback
from typing import Any, Sequence


def raiseNotDefined() -> None:
    caller = traceback.extract_stack(limit=2)[0]
    print(f"*** Method not implemented: {caller.name} at line {caller.lineno} of {caller.filename}")
    sys.exit(1)


class GameState:
    """Base class of game states; takes the reference's constructor arguments and ignores them."""

    def __init__(self, num_of_agent: int, agent_id: int):
        del num_of_agent, agent_id


class Action:
    """Placeholder kept for import compatibility; Splendor actions are plain dicts."""


class GameRule:
    """Turn bookkeeping shared by all rule objects.  Subclasses supply the five game-specific hooks."""

    def __init__(self, num_of_agent: int = 2):
        self.num_of_agent = num_of_agent
        self.current_agent_index = 0
        self.action_counter = 0
        self.current_game_state = self.initialGameState()

    # ---- hooks a concrete rule must implement -------------------------------------------------------------------
    def initialGameState(self) -> Any:
        raiseNotDefined()

    def getLegalActions(self, game_state: Any, agent_id: int) -> Sequence[Any]:
        raiseNotDefined()
        return ()

    def generateSuccessor(self, game_state: Any, action: Any, agent_id: int) -> Any:
        raiseNotDefined()

    def gameEnds(self) -> bool:
        raiseNotDefined()
        return False

    def calScore(self, game_state: Any, agent_id: int) -> float:
        raiseNotDefined()
        return 0.0

    # ---- turn order -----------------------------------------------------------------------------------------------
    def getCurrentAgentIndex(self) -> int:
        return self.current_agent_index

    def getNextAgentIndex(self) -> int:
        nxt = self.current_agent_index + 1
        return 0 if nxt >= self.num_of_agent else nxt

    def update(self, action: Any) -> None:
        """Apply `action` for the seat to move, hand the turn to the next seat and count the move."""
        mover = self.current_agent_index
        self.current_game_state = self.generateSuccessor(self.current_game_state, action, mover)
        self.current_agent_index = self.getNextAgentIndex()
        self.action_counter += 1


class Agent:
    """An agent occupies seat `_id`; the default policy is a uniform choice among the offered actions."""

    def __init__(self, _id: int):
        super().__init__()
        self.id = _id

    def SelectAction(self, actions: Sequence[Any], game_state: Any, game_rule: GameRule) -> Any:
        del game_state, game_rule
        return random.choice(actions)


class RandomAgent(Agent):
    """Uniform random play (reference: ``agents/generic/random.py:21-27``); identical to the default policy."""


class Displayer:
    """No-op display hooks for drivers that accept a displayer object."""

    def InitDisplayer(self, runner: Any) -> None:
        del runner

    def ExcuteAction(self, agent_id: int, move: Any, game_state: Any) -> None:
        del agent_id, move, game_state

    def TimeOutWarning(self, runner: Any, agent_id: int) -> None:
        del runner, agent_id

    def EndGame(self, game_state: Any, scores: Any) -> None:
        del game_state, scores


myAgent = RandomAgent

"""The framework contract of the reference (``splendor/template.py:17-89``): GameState, GameRule, Agent.

Same class names, attributes and method names, so agents written against the reference (``SelectAction(actions,
game_state, game_rule)``, modules exporting ``myAgent``) run unchanged on the B200 engine.  Abstract-method misuse
prints and exits with status 1 like the reference's ``utils.raiseNotDefined`` (``splendor/utils.py:7-13``).
"""
from __future__ import annotations

import inspect
import random
import sys


def raiseNotDefined():
    frame = inspect.stack()[1]
    print("*** Method not implemented: %s at line %s of %s" % (frame[3], frame[2], frame[1]))
    sys.exit(1)


class GameState:
    def __init__(self, num_of_agent, agent_id):
        pass


class GameRule:
    def __init__(self, num_of_agent=2):
        self.current_agent_index = 0
        self.num_of_agent = num_of_agent
        self.current_game_state = self.initialGameState()
        self.action_counter = 0

    def initialGameState(self):
        raiseNotDefined()
        return 0

    def generateSuccessor(self, game_state, action, agent_id):
        raiseNotDefined()
        return 0

    def getNextAgentIndex(self):
        return (self.current_agent_index + 1) % self.num_of_agent

    def getLegalActions(self, game_state, agent_id):
        raiseNotDefined()
        return []

    def calScore(self, game_state, agent_id):
        raiseNotDefined()
        return 0

    def gameEnds(self):
        raiseNotDefined()
        return False

    def update(self, action):
        temp_state = self.current_game_state
        self.current_game_state = self.generateSuccessor(temp_state, action, self.current_agent_index)
        self.current_agent_index = self.getNextAgentIndex()
        self.action_counter += 1

    def getCurrentAgentIndex(self):
        return self.current_agent_index


class Agent:
    def __init__(self, _id):
        self.id = _id
        super().__init__()

    def SelectAction(self, actions, game_state, game_rule):
        return random.choice(actions)


class RandomAgent(Agent):
    """agents/generic/random.py:21-27"""

    def SelectAction(self, actions, game_state, game_rule):
        return random.choice(actions)


myAgent = RandomAgent

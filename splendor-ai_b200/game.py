"""Tournament-side compatibility: the reference's ``Game`` / ``GameReplayer`` turn loop and
history format (``splendor/game.py:29-258``) on the B200 rule classes, plus the batched equivalent of
``splendor -m N`` for random agents (``general_game_runner.py:352-490``): N whole games in one fused kernel, results in
the runner's ``matches`` dictionary (total_scores / wins / ties / loses / win_rates).

``Game`` keeps the reference's behaviour knobs: ``random.seed(seed)`` then 1000 per-step seeds, agents receive deep copies
of the action list and the state, the global RNG is reseeded before and after every ``update`` (``game.py:183,196``), and
scores come from ``calScore`` (``_EndGame``, ``game.py:81-102``).  ``FREEDOM`` is True as in the reference (no time-outs).
"""
from __future__ import annotations

import copy
import random

import numpy as np

from .template import Agent as DummyAgent

FREEDOM = True


class Game:
    def __init__(self, GameRule, agent_list, num_of_agent, seed=1, time_limit=1, warning_limit=3, displayer=None,
                 agents_namelist=("Alice", "Bob"), interactive=False):
        self.seed = seed
        random.seed(self.seed)
        self.seed_list = [random.randint(0, int(1e10)) for _ in range(1000)]
        self.seed_idx = 0
        for i, plyr in enumerate(agent_list):
            assert plyr.id == i
        self.game_rule = GameRule(num_of_agent)
        self.gamemaster = DummyAgent(num_of_agent)
        self.agents = agent_list
        self.agents_namelist = list(agents_namelist)
        self.time_limit, self.warning_limit = time_limit, warning_limit
        self.warnings = [0] * len(agent_list)
        self.warning_positions = []
        self.displayer = displayer
        if self.displayer is not None:
            self.displayer.InitDisplayer(self)
        self.interactive = interactive

    def _EndGame(self, num_of_agent, history, isTimeOut=True, id=None):
        history.update({"seed": self.seed, "num_of_agent": num_of_agent, "agents_namelist": self.agents_namelist,
                        "warning_positions": self.warning_positions, "warning_limit": self.warning_limit})
        history["scores"] = {i: 0 for i in range(num_of_agent)}
        if isTimeOut:
            history["scores"][id] = -1
        else:
            for i in range(num_of_agent):
                history["scores"][i] = self.game_rule.calScore(self.game_rule.current_game_state, i)
        if self.displayer is not None:
            self.displayer.EndGame(self.game_rule.current_game_state, history["scores"])
        return history

    def Run(self):
        history = {"actions": []}
        action_counter = 0
        while not self.game_rule.gameEnds():
            agent_index = self.game_rule.getCurrentAgentIndex()
            agent = self.agents[agent_index] if agent_index < len(self.agents) else self.gamemaster
            game_state = self.game_rule.current_game_state
            game_state.agent_to_move = agent_index
            actions = self.game_rule.getLegalActions(game_state, agent_index)
            actions_copy, gs_copy = copy.deepcopy(actions), copy.deepcopy(game_state)
            if action_counter == 0 and self.displayer is not None:
                self.displayer._DisplayState(self.game_rule.current_game_state)
            selected = agent.SelectAction(actions_copy, gs_copy, self.game_rule)
            random.seed(self.seed_list[self.seed_idx % len(self.seed_list)])
            self.seed_idx += 1
            history["actions"].append({action_counter: {"agent_id": self.game_rule.current_agent_index, "action": selected}})
            action_counter += 1
            self.game_rule.update(selected)
            random.seed(self.seed_list[self.seed_idx % len(self.seed_list)])
            self.seed_idx += 1
            if self.displayer is not None:
                self.displayer.ExcuteAction(agent_index, selected, self.game_rule.current_game_state)
        return self._EndGame(self.game_rule.num_of_agent, history, isTimeOut=False)


class GameReplayer:
    """Re-applies a recorded history (``game.py:216-258``).  Because the first thing a game does with ``random`` after
    ``random.seed(seed)`` is draw its nobles and decks, replaying under the recorded seed reproduces the same board."""

    def __init__(self, GameRule, replay, displayer=None):
        self.replay = replay
        self.seed = replay["seed"]
        random.seed(self.seed)
        self.seed_list = [random.randint(0, int(1e10)) for _ in range(1000)]
        self.seed_idx = 0
        self.num_of_agent = replay["num_of_agent"]
        self.agents_namelist = replay["agents_namelist"]
        self.warning_limit = replay["warning_limit"]
        self.warnings = [0] * self.num_of_agent
        self.warning_positions = replay["warning_positions"]
        self.game_rule = GameRule(self.num_of_agent)
        self.scores = replay["scores"]
        self.displayer = displayer
        if self.displayer is not None:
            self.displayer.InitDisplayer(self)

    def Run(self):
        for item in self.replay["actions"]:
            ((index, info),) = item.items()
            self.game_rule.current_agent_index = info["agent_id"]
            random.seed(self.seed_list[self.seed_idx % len(self.seed_list)])
            self.seed_idx += 1
            self.game_rule.update(info["action"])
            random.seed(self.seed_list[self.seed_idx % len(self.seed_list)])
            self.seed_idx += 1
            if self.displayer is not None:
                self.displayer.ExcuteAction(info["agent_id"], info["action"], self.game_rule.current_game_state)
        if self.displayer is not None:
            self.displayer.EndGame(self.game_rule.current_game_state, self.scores)
        return self.game_rule.current_game_state


def tally(scores_per_game) -> dict:
    """The runner's cumulative statistics (``general_game_runner.py:381-431,465-490``) from per-game score lists."""
    scores = np.asarray(scores_per_game, dtype=np.float64)
    n_games, P = scores.shape
    best = scores.max(axis=1, keepdims=True)
    is_best = scores == best
    shared = is_best.sum(axis=1, keepdims=True) > 1
    wins = (is_best & ~shared).sum(axis=0)
    ties = (is_best & shared).sum(axis=0)
    loses = (~is_best).sum(axis=0)
    return {"total_scores": scores.sum(axis=0).tolist(), "wins": wins.tolist(), "ties": ties.tolist(), "loses": loses.tolist(),
            "win_rates": (wins / max(n_games, 1) * 100).tolist(), "succ": True}


def run_random_tournament(num_games: int, num_of_agents: int = 2, seed: int = 1234, first_game: int = 0,
                          limit_rounds: bool = False, agent_names=None) -> dict:
    """`splendor -m num_games` for RandomAgents on every seat, all games at once on the GPU (spl_playout).

    Returns the reference runner's ``matches`` dictionary: per-game scores under "games", and the cumulative
    total_scores / wins / ties / loses / win_rates (percent) per agent."""
    from .engine import counters_dict, playout_games

    out = playout_games(num_games, num_of_agents, seed=seed, first_game=first_game, limit_rounds=limit_rounds)
    score = out["score2"].cpu().numpy().astype(np.float64) / 2.0
    names = list(agent_names or [f"random{i}" for i in range(num_of_agents)])
    matches = {"games": [{"scores": {i: float(s) for i, s in enumerate(row)}} for row in score], "teams": {},
               "num_of_games": num_games, "agent_names": names}
    matches.update(tally(score))
    c = counters_dict(out["counters"])  # the kernel's own win/tie counters must agree with the host-side tally
    assert [c[f"wins{i}"] for i in range(num_of_agents)] == matches["wins"]
    assert [c[f"ties{i}"] for i in range(num_of_agents)] == matches["ties"]
    matches["steps"] = out["steps"].cpu().numpy()
    matches["end_kind"] = out["end_kind"].cpu().numpy()
    return matches

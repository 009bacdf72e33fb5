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


class _SeedSchedule:
    """`random.seed(seed)` followed by 1000 per-move seeds; the global RNG is re-seeded from the list around every update
    (twice per move), exactly the schedule of ``game.py:43-45,183,196``.  Wraps around after 500 moves."""

    def __init__(self, seed):
        random.seed(seed)
        self.values = [random.randint(0, int(1e10)) for _ in range(1000)]
        self.cursor = 0

    def reseed(self):
        random.seed(self.values[self.cursor % len(self.values)])
        self.cursor += 1


class Game:
    """One game between `agent_list` under `GameRule`; `Run()` returns the reference's history dictionary."""

    def __init__(self, GameRule, agent_list, num_of_agent, seed=1, time_limit=1, warning_limit=3, displayer=None,
                 agents_namelist=("Alice", "Bob"), interactive=False):
        self.seed = seed
        self._schedule = _SeedSchedule(seed)
        self.seed_list, self.seed_idx = self._schedule.values, 0
        if any(agent.id != seat for seat, agent in enumerate(agent_list)):
            raise AssertionError("agent ids must be 0..N-1 in list order")
        self.game_rule = GameRule(num_of_agent)  # draws nobles and decks from the RNG state left by the schedule
        self.gamemaster = DummyAgent(num_of_agent)
        self.agents = agent_list
        self.agents_namelist = list(agents_namelist)
        self.time_limit, self.warning_limit, self.interactive = time_limit, warning_limit, interactive
        self.warnings, self.warning_positions = [0] * len(agent_list), []
        self.displayer = displayer
        if displayer is not None:
            displayer.InitDisplayer(self)

    def _EndGame(self, num_of_agent, history, isTimeOut=True, id=None):
        rule = self.game_rule
        scores = {seat: 0 for seat in range(num_of_agent)}
        if isTimeOut:
            scores[id] = -1
        else:
            scores = {seat: rule.calScore(rule.current_game_state, seat) for seat in range(num_of_agent)}
        history.update(seed=self.seed, num_of_agent=num_of_agent, agents_namelist=self.agents_namelist,
                       warning_positions=self.warning_positions, warning_limit=self.warning_limit, scores=scores)
        if self.displayer is not None:
            self.displayer.EndGame(rule.current_game_state, scores)
        return history

    def _play_one_move(self, move_number):
        rule = self.game_rule
        seat = rule.getCurrentAgentIndex()
        state = rule.current_game_state
        state.agent_to_move = seat
        offered = rule.getLegalActions(state, seat)
        mover = self.agents[seat] if seat < len(self.agents) else self.gamemaster
        if move_number == 0 and self.displayer is not None:
            self.displayer._DisplayState(state)
        # agents get private copies of the offer and of the state, but the live rule object (game.py:117-118,142)
        chosen = mover.SelectAction(copy.deepcopy(offered), copy.deepcopy(state), rule)
        self._schedule.reseed()
        record = {move_number: {"agent_id": seat, "action": chosen}}
        rule.update(chosen)
        self._schedule.reseed()
        self.seed_idx = self._schedule.cursor
        if self.displayer is not None:
            self.displayer.ExcuteAction(seat, chosen, rule.current_game_state)
        return record

    def Run(self):
        history, move_number = {"actions": []}, 0
        while not self.game_rule.gameEnds():
            history["actions"].append(self._play_one_move(move_number))
            move_number += 1
        return self._EndGame(self.game_rule.num_of_agent, history, isTimeOut=False)


class GameReplayer:
    """Re-applies a recorded history (``game.py:216-258``).  The first thing a game does with ``random`` after seeding is
    draw its nobles and decks, so replaying under the recorded seed reproduces the same board."""

    def __init__(self, GameRule, replay, displayer=None):
        self.replay = replay
        self.seed = replay["seed"]
        self._schedule = _SeedSchedule(self.seed)
        self.seed_list, self.seed_idx = self._schedule.values, 0
        self.num_of_agent = replay["num_of_agent"]
        self.agents_namelist = replay["agents_namelist"]
        self.warning_limit, self.warning_positions = replay["warning_limit"], replay["warning_positions"]
        self.warnings = [0] * self.num_of_agent
        self.scores = replay["scores"]
        self.game_rule = GameRule(self.num_of_agent)
        self.displayer = displayer
        if displayer is not None:
            displayer.InitDisplayer(self)

    def Run(self):
        rule = self.game_rule
        for entry in self.replay["actions"]:
            ((_, move),) = entry.items()
            rule.current_agent_index = move["agent_id"]
            self._schedule.reseed()
            rule.update(move["action"])
            self._schedule.reseed()
            if self.displayer is not None:
                self.displayer.ExcuteAction(move["agent_id"], move["action"], rule.current_game_state)
        if self.displayer is not None:
            self.displayer.EndGame(rule.current_game_state, self.scores)
        return rule.current_game_state


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

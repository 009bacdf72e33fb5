"""`-m gpu`: the tournament-side formats against histories produced by the reference's own
`Game.Run()` / `GameReplayer` (tests/golden/game_histories.json, oracle/gen_golden_game.py; splendor/game.py:29-258).

  * `splendor_ai_b200.game.Game` with index-scripted agents and the same seed plays the reference's game move for move:
    same deck orders out of `random.seed(seed)`, same (agent id, ALL_ACTIONS index) sequence, the same action dicts
    (keys and values, cards / nobles by code) and the same `scores` / bookkeeping entries in the history dictionary;
  * `GameReplayer` replays every reference history (RandomAgent games included) to the reference's final state."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def games():
    with open(os.path.join(GOLDEN, "game_histories.json")) as fh:
        return json.load(fh)["games"]


def render(action):
    out = {}
    for k, v in action.items():
        if k == "card":
            out[k] = v.code
        elif k == "noble":
            out[k] = None if v is None else v[0]
        elif k == "card_position":
            out[k] = list(v)
        else:
            out[k] = dict(v) if isinstance(v, dict) else v
    return out


def scripted_choice(idxs, t):  # oracle/gen_golden_game.py, same definition
    buys = [i for i in idxs if i >= 3420]
    pool = buys if (buys and t % 4 != 3) else list(idxs)
    return pool[(5 * t + 1) % len(pool)]


def snapshot_of(state, P):
    from splendor_ai_b200.state import snapshots

    return snapshots(state._words.reshape(-1, 1, 4), P)[0]


def test_game_plays_the_reference_scripted_games(games):
    from splendor_ai_b200 import game_rule as G
    from splendor_ai_b200 import template
    from splendor_ai_b200.game import Game

    class Scripted(template.Agent):
        def SelectAction(self, actions, game_state, game_rule):
            by_index = {G.action_index(a, game_state, self.id): a for a in actions}
            t = sum(len(ag.agent_trace.action_reward) for ag in game_state.agents)
            return by_index[scripted_choice(sorted(by_index), t)]

    n = 0
    for g in (x for x in games if x["agents"] == "scripted"):
        P = g["P"]
        game = Game(G.SplendorGameRule, [Scripted(i) for i in range(P)], P, seed=g["seed"], agents_namelist=g["agents_namelist"])
        assert game.game_rule.current_game_state._deck_lists.tolist() == g["decks"], "random.seed(seed) must give the reference's decks"
        history = game.Run()
        assert sorted(history) == g["history_keys"]
        for key in ("num_of_agent", "agents_namelist", "warning_positions", "warning_limit"):
            assert history[key] == g[key], key
        assert history["seed"] == g["seed"]
        assert {str(k): v for k, v in history["scores"].items()} == g["scores"]
        assert len(history["actions"]) == len(g["actions"])
        for i, (item, want) in enumerate(zip(history["actions"], g["actions"])):
            ((idx, info),) = item.items()
            assert idx == i and info["agent_id"] == want["agent_id"], i
            assert render(info["action"]) == want["action"], (i, render(info["action"]), want["action"])
        assert snapshot_of(game.game_rule.current_game_state, P).tolist() == g["final_snapshot"]
        n += 1
    assert n >= 4


def test_replayer_replays_reference_histories(games):
    from splendor_ai_b200 import game_rule as G
    from splendor_ai_b200.game import GameReplayer

    def to_action(r):
        a = dict(r)
        if "card" in a:
            a["card"] = G.card_of(G.CARD_ID[a["card"]])
        if a.get("noble") is not None:
            a["noble"] = G.NOBLES[G.NOBLE_ID[a["noble"]]]
        if "card_position" in a:
            a["card_position"] = tuple(a["card_position"])
        return a

    for g in games:
        P = g["P"]
        replay = {"seed": g["seed"], "num_of_agent": g["num_of_agent"], "agents_namelist": g["agents_namelist"],
                  "warning_limit": g["warning_limit"], "warning_positions": g["warning_positions"],
                  "scores": {int(k): v for k, v in g["scores"].items()},
                  "actions": [{i: {"agent_id": e["agent_id"], "action": to_action(e["action"])}} for i, e in enumerate(g["actions"])]}
        r = GameReplayer(G.SplendorGameRule, replay)
        assert r.game_rule.current_game_state._deck_lists.tolist() == g["decks"]
        final = r.Run()
        assert snapshot_of(final, P).tolist() == g["replayed_snapshot"] == g["final_snapshot"], (g["agents"], g["seed"])
        assert {str(i): r.game_rule.calScore(final, i) for i in range(P)} == g["scores"]
        # the applied indices are the ones the reference's create_action_mapping gave to its own action dicts
        assert [idx for _, idx in final._applied] == [e["index"] for e in g["actions"]]

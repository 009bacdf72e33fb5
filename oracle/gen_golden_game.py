"""Generate tests/golden/game_histories.json by running the UNMODIFIED reference `Game` / `GameReplayer`
(splendor/game.py:29-258) on `SplendorGameRule` - the tournament-side formats of SURVEY.md 8(f) rank 2.

TEST INFRASTRUCTURE.  Run:  PYTHONHASHSEED=0 python oracle/gen_golden_game.py              (about two minutes)

Per recorded game: the history dictionary `Game.Run()` returned (its keys, seed, names, warnings, `scores` from
`_EndGame`, and every `actions` entry rendered to JSON: agent id, the action dict with cards / nobles by code, and the
action's ALL_ACTIONS index taken with the reference's own create_action_mapping), the initial deck orders and nobles,
the final 108-byte snapshot, and the snapshot the reference's `GameReplayer` reaches from that history (equal).

Agents: the reference's `RandomAgent` (its trajectory depends on PYTHONHASHSEED through the order of getLegalActions, so
it can only be REPLAYED elsewhere) and an index-scripted agent whose choice is a function of the sorted legal index set,
which any implementation that offers the same legal sets reproduces move for move.
"""
import json
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

ref_harness.install()
if "func_timeout" not in sys.modules:  # absent here; game.py only uses it when FREEDOM is False
    ft = types.ModuleType("func_timeout")
    ft.FunctionTimedOut = type("FunctionTimedOut", (Exception,), {})
    ft.func_timeout = lambda timeout, func, args=(), kwargs=None: func(*args, **(kwargs or {}))
    sys.modules["func_timeout"] = ft

from gen_golden import NOBLE_ID, initial_deck_lists, snapshot  # noqa: E402
from splendor.agents.generic.random import myAgent as RefRandomAgent  # noqa: E402
from splendor.game import Game, GameReplayer  # noqa: E402
from splendor.splendor.gym.envs.utils import create_action_mapping  # noqa: E402
from splendor.splendor.splendor_model import SplendorGameRule  # noqa: E402
from splendor.template import Agent  # noqa: E402


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


def index_of(chosen, actions, state, agent_id):
    mapping = create_action_mapping(actions, state, agent_id)
    hits = [i for i, a in mapping.items() if a == chosen]
    assert len(hits) == 1
    return hits[0]


def scripted_choice(idxs, t):
    """Prefer a purchase three turns out of four, else any legal action: ends by score in ~60-90 moves."""
    buys = [i for i in idxs if i >= 3420]
    pool = buys if (buys and t % 4 != 3) else list(idxs)
    return pool[(5 * t + 1) % len(pool)]


class LoggingRandom(RefRandomAgent):
    def __init__(self, _id, log):
        super().__init__(_id)
        self.log = log

    def SelectAction(self, actions, game_state, game_rule):
        chosen = super().SelectAction(actions, game_state, game_rule)
        self.log.append(index_of(chosen, actions, game_state, self.id))
        return chosen


class Scripted(Agent):
    def __init__(self, _id, log):
        super().__init__(_id)
        self.log = log

    def SelectAction(self, actions, game_state, game_rule):
        mapping = create_action_mapping(actions, game_state, self.id)
        t = sum(len(a.agent_trace.action_reward) for a in game_state.agents)
        i = scripted_choice(sorted(mapping), t)
        self.log.append(i)
        return mapping[i]


def record(kind, P, seed):
    log = []
    cls = LoggingRandom if kind == "random" else Scripted
    names = ["Alice", "Bob", "Carol", "Dave"][:P]
    game = Game(SplendorGameRule, [cls(i, log) for i in range(P)], P, seed=seed, agents_namelist=names)
    decks = initial_deck_lists(game.game_rule)
    nobles = [NOBLE_ID[n[0]] for n in game.game_rule.current_game_state.board.nobles]
    history = game.Run()
    final = snapshot(game.game_rule)
    assert len(log) == len(history["actions"])
    entries = []
    for n, item in enumerate(history["actions"]):
        ((idx, info),) = item.items()
        assert idx == n
        entries.append({"agent_id": info["agent_id"], "index": int(log[n]), "action": render(info["action"])})
    # game.py:222 calls random.randint(0, 1e10): a float bound, accepted by the Python the reference pins (3.10, with a
    # deprecation warning) and a TypeError on this container's 3.12.  The shim restores the old coercion for that call only.
    real_randint = random.randint
    random.randint = lambda lo, hi: real_randint(int(lo), int(hi))
    try:
        replayer = GameReplayer(SplendorGameRule, history)
    finally:
        random.randint = real_randint
    replayer.Run()
    replayed = snapshot(replayer.game_rule)
    # GameReplayer never advances to a terminal check; its seat bookkeeping equals the game's after the last update
    assert np.array_equal(replayed, final)
    return {
        "agents": kind, "P": P, "seed": seed, "history_keys": sorted(history), "num_of_agent": history["num_of_agent"],
        "agents_namelist": history["agents_namelist"], "warning_positions": history["warning_positions"],
        "warning_limit": history["warning_limit"], "scores": {str(k): v for k, v in history["scores"].items()},
        "actions": entries, "decks": decks.tolist(), "nobles": nobles, "final_snapshot": final.tolist(),
        "replayed_snapshot": replayed.tolist(),
    }


if __name__ == "__main__":
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0"
    games = []
    for kind, P, seed in [("random", 2, 1), ("random", 2, 7), ("random", 3, 11), ("scripted", 2, 3), ("scripted", 2, 42),
                          ("scripted", 3, 5), ("scripted", 4, 9)]:
        games.append(record(kind, P, seed))
        g = games[-1]
        print(kind, P, seed, "moves", len(g["actions"]), "scores", g["scores"])
    path = os.path.join(ROOT, "tests", "golden", "game_histories.json")
    with open(path, "w") as fh:
        json.dump({"games": games}, fh, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path) // 1024, "KB")

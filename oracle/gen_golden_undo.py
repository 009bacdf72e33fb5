"""tests/golden/undo.npz: the reference's generatePredecessor (splendor_model.py:156-220) recorded on real games.
Run:  PYTHONHASHSEED=0 python oracle/gen_golden_undo.py      TEST INFRASTRUCTURE (needs /root/reference).

Every step of every game: snapshot, successor, snapshot, PREDECESSOR, snapshot (kept), successor again to move on.
The undo inputs the engine needs beyond the action index are stored too: the card bought/reserved, the noble claimed,
and the action's returned_gems."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_golden as G  # noqa: E402  (installs the reference on sys.path)
from splendor.splendor.gym.envs.utils import create_action_mapping  # noqa: E402
from splendor.splendor.splendor_model import SplendorGameRule  # noqa: E402

COLS = G.COLS


def record(P, policy, seed):
    random.seed(seed)
    rule = SplendorGameRule(P)
    rng = random.Random(seed + 5)
    out = dict(P=P, decks=G.initial_deck_lists(rule),
               nobles=np.array([G.NOBLE_ID[n[0]] for n in rule.current_game_state.board.nobles] + [G.NONE] * (4 - P), np.uint8),
               action=[], card=[], noble=[], returned=[], after_undo=[], after_step=[])
    while not rule.gameEnds() and len(out["action"]) < 400:
        st, seat = rule.current_game_state, rule.current_agent_index
        mapping = create_action_mapping(rule.getLegalActions(st, seat), st, seat)
        a = G.choose(policy, rng, mapping)
        act = mapping[a]
        rule.generateSuccessor(st, act, seat)
        rule.generatePredecessor(st, act, seat)
        out["after_undo"].append(G.snapshot(rule))
        rule.update(act)
        out["after_step"].append(G.snapshot(rule))
        out["action"].append(a)
        out["card"].append(G.CARD_ID[act["card"].code] if act.get("card") is not None else G.NONE)
        out["noble"].append(G.NOBLE_ID[act["noble"][0]] if act.get("noble") is not None else G.NONE)
        out["returned"].append([act.get("returned_gems", {}).get(c, 0) for c in COLS])
    return out


if __name__ == "__main__":
    games = [record(P, pol, 7000 + i) for i, (P, pol) in enumerate(
        [(2, "random"), (2, "greedy"), (2, "greedy"), (2, "hoard"), (2, "reserve"), (3, "greedy"), (4, "random"), (4, "greedy")])]
    n = np.array([len(g["action"]) for g in games], np.int32)
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "undo.npz"),
        P=np.array([g["P"] for g in games], np.uint8), decks=np.stack([g["decks"] for g in games]),
        nobles=np.stack([np.concatenate([g["nobles"], [G.NONE]])[:5] for g in games]).astype(np.uint8), n_steps=n,
        action=np.concatenate([np.array(g["action"], np.int16) for g in games]),
        card=np.concatenate([np.array(g["card"], np.uint8) for g in games]),
        noble=np.concatenate([np.array(g["noble"], np.uint8) for g in games]),
        returned=np.concatenate([np.array(g["returned"], np.uint8).reshape(-1, 6) for g in games]),
        after_undo=np.concatenate([np.stack(g["after_undo"]) for g in games]),
        after_step=np.concatenate([np.stack(g["after_step"]) for g in games]))
    print("undo golden:", len(games), "games", int(n.sum()), "steps; nobles undone:", int(sum((np.array(g["noble"]) != G.NONE).sum() for g in games)))

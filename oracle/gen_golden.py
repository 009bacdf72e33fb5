"""Generate tests/golden/replays_*.npz by running the UNMODIFIED reference engine in the build container.

TEST INFRASTRUCTURE.  Run:  PYTHONHASHSEED=0 python oracle/gen_golden.py          (about 5-8 minutes on one core)

For every recorded game the reference's own code produces every number stored:
  * SplendorGameRule / LimitRoundsGameRule build the state, `random` shuffles the decks and samples the nobles
    (splendor_model.py:69-93) - the resulting deck orders are stored so the game can be re-played elsewhere;
  * getLegalActions (splendor_model.py:404-576) + create_action_mapping (gym/envs/utils.py:167-181) give the legal
    ALL_ACTIONS index set - the reference's own dict->index conversion, not a restatement;
  * GameRule.update (template.py:47-53) applies the chosen action;
  * gameEnds of both rule classes, calScore, and features.extract_metrics_with_cards(...).astype(float32)
    (splendor_env.py:206) are recorded at every step.
Nothing here depends on set/dict iteration order: legal actions are stored as sorted index sets.

Action-choice policies (to reach the rule's corners, not to play well): uniform random; buy-greedy (nobles, 7-card
cap, score endings); hoard (10-gem returns, empty bank, reserve-without-yellow, pass, deadlock); reserve-first.
"""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

ref_harness.install()
from splendor.splendor import features  # noqa: E402
from splendor.splendor.gym.envs.utils import create_action_mapping, create_legal_actions_mask  # noqa: E402
from splendor.splendor.splendor_model import SplendorGameRule  # noqa: E402
from splendor.splendor.splendor_utils import NOBLES  # noqa: E402
from splendor.splendor.utils import LimitRoundsGameRule  # noqa: E402

TABLES = json.load(open(os.path.join(ROOT, "tests", "golden", "tables.json")))
CARD_ID = {c["code"]: i for i, c in enumerate(TABLES["cards"])}
NOBLE_ID = {n["code"]: i for i, n in enumerate(TABLES["nobles"])}
COLS = ["black", "red", "green", "blue", "white", "yellow"]
NONE = 0xFF


def snapshot(rule) -> np.ndarray:
    """Reference object graph -> the canonical 108-byte field vector (same layout as oracle.State.snapshot)."""
    st = rule.current_game_state
    b = st.board
    out = np.zeros(108, np.uint8)
    out[0:6] = [b.gems[c] for c in COLS]
    out[6:18] = [CARD_ID[c.code] if c else NONE for row in b.dealt for c in row]
    out[18:23] = [NOBLE_ID[n[0]] for n in b.nobles] + [NONE] * (5 - len(b.nobles))
    out[23] = len(b.nobles)
    out[24:27] = [len(d) for d in b.decks]
    out[27] = rule.current_agent_index
    for i, a in enumerate(st.agents):
        o = 28 + 20 * i
        out[o:o + 6] = [a.gems[c] for c in COLS]
        out[o + 6:o + 11] = [len(a.cards[c]) for c in COLS[:5]]
        res = [CARD_ID[c.code] for c in a.cards["yellow"]]
        out[o + 11:o + 14] = res + [NONE] * (3 - len(res))
        out[o + 14] = len(res)
        out[o + 15] = a.score
        out[o + 16] = 1 if a.passed else 0
        out[o + 17] = len(a.nobles)
        turns = len(a.agent_trace.action_reward)
        out[o + 18], out[o + 19] = turns & 0xFF, turns >> 8
    return out


def initial_deck_lists(rule) -> np.ndarray:
    """The three shuffled lists as they were BEFORE the 12 initial deals (deal() pops from the end)."""
    b = rule.current_game_state.board
    out = []
    for t in range(3):
        lst = [CARD_ID[c.code] for c in b.decks[t]] + [CARD_ID[b.dealt[t][j].code] for j in (3, 2, 1, 0)]
        assert len(lst) == (40, 30, 20)[t]
        out += lst
    return np.array(out, np.uint8)


def choose(policy, rng, mapping):
    idxs = sorted(mapping)
    kind = lambda i: mapping[i]["type"]
    if policy == "greedy" and rng.random() < 0.85:
        buys = [i for i in idxs if kind(i).startswith("buy")]
        if buys:
            nob = [i for i in buys if mapping[i]["noble"] is not None]
            return rng.choice(nob or buys)
    if policy == "hoard" and rng.random() < 0.9:
        pool = [i for i in idxs if kind(i).startswith("collect")] or [i for i in idxs if kind(i) == "reserve"]
        if pool:
            return rng.choice(pool)
    if policy == "reserve" and rng.random() < 0.7:
        pool = [i for i in idxs if kind(i) == "reserve"] or [i for i in idxs if kind(i).startswith("buy")]
        if pool:
            return rng.choice(pool)
    return rng.choice(idxs)


def record_game(P, policy, seed, limit_rounds, max_steps=600, check_mask_every=7):
    random.seed(seed)
    rule = LimitRoundsGameRule(P)  # a SplendorGameRule subclass: both gameEnds variants can be asked of one object
    ended = (lambda: rule.gameEnds()) if limit_rounds else (lambda: SplendorGameRule.gameEnds(rule))
    rng = random.Random(seed * 7919 + 13)
    g = dict(P=P, policy=policy, seed=seed, limit=int(limit_rounds), decks=initial_deck_lists(rule),
             nobles=np.array([NOBLE_ID[n[0]] for n in rule.current_game_state.board.nobles] + [NONE] * (4 - P), np.uint8),
             seat=[], action=[], reward=[], ends_plain=[], ends_limit=[], legal=[], snap=[snapshot(rule)], obs=[])
    steps = 0
    while not ended() and steps < max_steps:
        st, seat = rule.current_game_state, rule.current_agent_index
        legal = rule.getLegalActions(st, seat)
        mapping = create_action_mapping(legal, st, seat)
        assert len(mapping) == len(legal), "list->index must be a bijection"
        if steps % check_mask_every == 0:
            mask = create_legal_actions_mask(legal, st, seat)
            assert sorted(np.nonzero(mask)[0].tolist()) == sorted(mapping)
        g["obs"].append(features.extract_metrics_with_cards(st, seat).astype(np.float32))
        a = choose(policy, rng, mapping)
        before = st.agents[seat].score
        rule.update(mapping[a])
        g["seat"].append(seat)
        g["action"].append(a)
        g["reward"].append(st.agents[seat].score - before)
        g["legal"].append(np.array(sorted(mapping), np.uint16))
        g["snap"].append(snapshot(rule))
        g["ends_plain"].append(int(SplendorGameRule.gameEnds(rule)))
        g["ends_limit"].append(int(rule.gameEnds()))
        steps += 1
    g["score2"] = np.array([int(round(2 * rule.calScore(rule.current_game_state, i))) for i in range(P)] + [0] * (4 - P), np.int16)
    g["final_obs"] = np.stack([features.extract_metrics_with_cards(rule.current_game_state, i).astype(np.float32) for i in range(P)]
                              + [np.zeros(265, np.float32)] * (4 - P))
    return g


def pack(games):
    n_steps = np.array([len(g["action"]) for g in games], np.int32)
    legal_len = np.array([len(l) for g in games for l in g["legal"]], np.int32)
    return dict(
        P=np.array([g["P"] for g in games], np.uint8),
        limit=np.array([g["limit"] for g in games], np.uint8),
        seed=np.array([g["seed"] for g in games], np.int64),
        policy=np.array([g["policy"] for g in games]),
        decks=np.stack([g["decks"] for g in games]),
        nobles=np.stack([np.concatenate([g["nobles"], [NONE]])[:5] for g in games]).astype(np.uint8),
        n_steps=n_steps,
        seat=np.concatenate([np.array(g["seat"], np.uint8) for g in games]),
        action=np.concatenate([np.array(g["action"], np.int16) for g in games]),
        reward=np.concatenate([np.array(g["reward"], np.int8) for g in games]),
        ends_plain=np.concatenate([np.array(g["ends_plain"], np.uint8) for g in games]),
        ends_limit=np.concatenate([np.array(g["ends_limit"], np.uint8) for g in games]),
        legal_len=legal_len,
        legal_flat=np.concatenate([l for g in games for l in g["legal"]]),
        snap0=np.stack([g["snap"][0] for g in games]),
        snap=np.concatenate([np.stack(g["snap"][1:]) for g in games]),
        obs=np.concatenate([np.stack(g["obs"]) for g in games]),
        score2=np.stack([g["score2"] for g in games]),
        final_obs=np.stack([g["final_obs"] for g in games]),
    )


PLAN = [
    # name, list of (P, policy, limit_rounds, n_games)
    ("replays_2p", [(2, "random", False, 14), (2, "greedy", False, 8), (2, "hoard", False, 5), (2, "reserve", False, 3),
                    (2, "random", True, 4), (2, "hoard", True, 2)]),
    ("replays_3p", [(3, "random", False, 4), (3, "greedy", False, 3), (3, "hoard", False, 2)]),
    ("replays_4p", [(4, "random", False, 6), (4, "greedy", False, 4), (4, "hoard", False, 2), (4, "random", True, 2),
                    (4, "reserve", False, 1)]),
]

# a second batch with other seeds (added later in round 1; the first three files are left untouched)
PLAN_B = [
    ("replays_2p_b", [(2, "random", False, 20), (2, "greedy", False, 10), (2, "hoard", False, 6), (2, "reserve", False, 4),
                      (2, "random", True, 6), (2, "greedy", True, 4)]),
    ("replays_3p_b", [(3, "random", False, 6), (3, "greedy", False, 4), (3, "hoard", True, 2)]),
    ("replays_4p_b", [(4, "random", False, 8), (4, "greedy", False, 6), (4, "hoard", False, 3), (4, "reserve", True, 2)]),
]

if __name__ == "__main__":
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0 for reproducible action choices"
    seed = 1000
    plan = PLAN
    if len(sys.argv) > 1 and sys.argv[1] == "b":
        plan, seed = PLAN_B, 5000
    for name, groups in plan:
        t0, games = time.time(), []
        for P, policy, limit, n in groups:
            for _ in range(n):
                games.append(record_game(P, policy, seed, limit))
                seed += 1
        arrs = pack(games)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **arrs)
        kinds = {}
        print("%s: %d games, %d steps, %.1f s, %.0f KB; max legal %d; nobles won %d; passes %d" % (
            name, len(games), int(arrs["n_steps"].sum()), time.time() - t0, os.path.getsize(path) / 1024,
            int(arrs["legal_len"].max()), int(sum(g["snap"][-1][28 + 17::20][:g["P"]].sum() for g in games)),
            int((arrs["action"] < 6).sum())))

"""Generate tests/golden/env_*.npz by driving the UNMODIFIED reference `SplendorEnv` (gym/envs/splendor_env.py:27-245).

TEST INFRASTRUCTURE.  Run:  PYTHONHASHSEED=0 python oracle/gen_golden_env.py            (a few minutes on one core)

What is the reference's and what is the harness':
  * the env object, its reset / step / _simulate_opponents / get_legal_actions_mask, LimitRoundsGameRule, the feature
    vector and the dict -> ALL_ACTIONS index conversion are the reference's own code, imported from /root/reference;
  * the harness decides the deck orders and nobles (a stand-in for the `random` module seen by splendor_model.py, so
    that the game the reference builds is this repo's synthetic game (seed, game id), oracle.synth_game) and supplies the
    opponents: `Agent` objects whose SelectAction reproduces the in-kernel choice of spl_env_step - the Philox word of
    (seed, game, total turns) scaled to the number of legal actions, taken over the index-sorted legal list - or, in the
    "scripted" episodes, a deterministic hoarding policy that drives the game into the 100-round cap.

Recorded per env call (reset counts as call 0): observation bits, reward, terminated, the learner's legal index set
(from get_legal_actions_mask), the 108-byte state snapshot; per episode: players, learner seat (`my_id`), seed, game
id, deck lists, nobles, and for scripted episodes every action of every seat in turn order.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

ref_harness.install()
import oracle as O  # noqa: E402  (Philox and the synthetic deck orders only)
from gen_golden import CARD_ID, NOBLE_ID, snapshot  # noqa: E402
from splendor.splendor import splendor_model  # noqa: E402
from splendor.splendor.gym.envs.splendor_env import SplendorEnv  # noqa: E402
from splendor.splendor.gym.envs.utils import create_action_mapping  # noqa: E402
from splendor.splendor.splendor_utils import NOBLES  # noqa: E402
from splendor.template import Agent  # noqa: E402

NONE = 0xFF
ID_CARD = {v: k for k, v in CARD_ID.items()}
NOBLE_BY_ID = {NOBLE_ID[n[0]]: n for n in NOBLES}


class InjectedRandom:
    """Stand-in for the `random` module inside splendor_model.py while ONE SplendorState is built
    (BoardState.__init__, splendor_model.py:69-93: one sample() for the nobles, then one shuffle() per tier)."""

    def __init__(self, deck_lists, nobles):
        self.lists, self.nobles, self.tier = [int(c) for c in deck_lists], [int(n) for n in nobles], 0

    def sample(self, population, k):
        assert population is NOBLES and k == len(self.nobles)
        return [NOBLE_BY_ID[i] for i in self.nobles]

    def shuffle(self, deck):
        base, size = (0, 40, 70)[self.tier], (40, 30, 20)[self.tier]
        want = [ID_CARD[c] for c in self.lists[base:base + size]]
        by_code = {card.code: card for card in deck}
        assert sorted(by_code) == sorted(want)
        deck[:] = [by_code[code] for code in want]
        self.tier += 1


class injected_decks:
    def __init__(self, deck_lists, nobles):
        self.fake = InjectedRandom(deck_lists, nobles)

    def __enter__(self):
        self.real = splendor_model.random
        splendor_model.random = self.fake

    def __exit__(self, *exc):
        splendor_model.random = self.real


def total_turns(state):
    return sum(len(a.agent_trace.action_reward) for a in state.agents)


class PhiloxAgent(Agent):
    """The opponent spl_env_step plays in-kernel: k = (philox(seed, game, total turns).x * n_legal) >> 32 over the legal
    actions in ascending ALL_ACTIONS index."""

    def __init__(self, _id, seed, game, log):
        super().__init__(_id)
        self.seed, self.game, self.log = seed, game, log

    def SelectAction(self, actions, game_state, game_rule):
        mapping = create_action_mapping(actions, game_state, self.id)
        idxs = sorted(mapping)
        r = int(O.philox(self.seed, self.game, total_turns(game_state), 0)[0])
        a = idxs[(r * len(idxs)) >> 32]
        self.log.append(a)
        return mapping[a]


def scripted_choice(idxs, t):
    """Deterministic hoarding policy on a sorted legal index set: collect if possible, else reserve, else anything."""
    pool = [i for i in idxs if 510 <= i < 3420] or [i for i in idxs if 6 <= i < 510] or list(idxs)
    return pool[(7 * t + 3) % len(pool)]


class ScriptedAgent(Agent):
    def __init__(self, _id, log):
        super().__init__(_id)
        self.log = log

    def SelectAction(self, actions, game_state, game_rule):
        mapping = create_action_mapping(actions, game_state, self.id)
        a = scripted_choice(sorted(mapping), total_turns(game_state))
        self.log.append(a)
        return mapping[a]


def learner_choice(policy, rng, idxs, t):
    if policy == "scripted":
        return scripted_choice(idxs, t)
    if policy == "greedy" and rng.random() < 0.85:
        buys = [i for i in idxs if i >= 3420]
        if buys:
            return int(rng.choice(buys))
    return int(rng.choice(idxs))


def record_episode(P, seat, seed, game, policy, opponents, shuffle_turns, max_calls=400, pad=0, np_seed=None):
    """seat=None: the env draws the learner's seat itself (fixed_turn=None -> np.random.randint, splendor_env.py:107-111)."""
    lists, nobles = O.synth_game(seed, game, P)
    log = []
    if opponents == "philox":
        agents = [PhiloxAgent(0, seed, game, log) for _ in range(P - 1)]
    else:
        agents = [ScriptedAgent(0, log) for _ in range(P - 1)]
    env = SplendorEnv(agents, shuffle_turns=shuffle_turns, fixed_turn=seat)
    rng = np.random.default_rng(seed * 1_000_003 + game)
    ep = dict(P=P, my_id=seat, seed=seed, game=game, policy=policy, opponents=opponents, pad=pad, decks=np.asarray(lists, np.uint8),
              nobles=np.concatenate([nobles, [NONE] * (5 - len(nobles))]).astype(np.uint8),
              action=[], reward=[], terminated=[], obs=[], legal=[], snap=[], all_actions=log)

    def observe_call(obs, reward, terminated, action):
        mask = env.get_legal_actions_mask()
        ep["action"].append(action)
        ep["reward"].append(reward)
        ep["terminated"].append(int(terminated))
        ep["obs"].append(np.asarray(obs, np.float32))
        ep["legal"].append(np.nonzero(mask)[0].astype(np.uint16))
        ep["snap"].append(snapshot(env.game_rule))

    if np_seed is not None:
        np.random.seed(np_seed)
    with injected_decks(lists, nobles):
        obs, info = env.reset()
    if seat is None:
        seat = ep["my_id"] = int(info["my_id"])
    assert info["my_id"] == seat == env.my_turn
    if pad:
        # Harness edit of the reference's STATE (not its code): every agent's action trace gets `pad` placeholder entries
        # right after reset, as if `pad` full rounds had been played without changing anything else.  Only the trace
        # LENGTH is read by the rules (LimitRoundsGameRule.gameEnds, splendor/utils.py:19-23) and by observation feature 1
        # (features.py:100-104).  With more than two players hoarding games end by score long before turn 100, so this is
        # how the 100-round cap is reached for 3 and 4 players.
        for ag in env.state.agents:
            ag.agent_trace.action_reward.extend([({"type": "pass", "noble": None}, 0)] * pad)
        obs = env._vectorize(env.state, env.my_turn)
    observe_call(obs, 0, env.game_rule.gameEnds(), -1)
    calls = 0
    while not ep["terminated"][-1] and calls < max_calls:
        a = learner_choice(policy, rng, [int(i) for i in ep["legal"][-1]], total_turns(env.state))
        log.append(a)
        obs, reward, terminated, truncated, _ = env.step(a)
        assert truncated is False
        observe_call(obs, reward, terminated, a)
        calls += 1
    assert ep["terminated"][-1], "episode did not end"
    return ep


def pack(eps):
    cat = lambda key, dt: np.concatenate([np.asarray(e[key], dt) for e in eps])
    return dict(
        P=np.array([e["P"] for e in eps], np.uint8), my_id=np.array([e["my_id"] for e in eps], np.uint8),
        seed=np.array([e["seed"] for e in eps], np.int64), game=np.array([e["game"] for e in eps], np.int64),
        policy=np.array([e["policy"] for e in eps]), opponents=np.array([e["opponents"] for e in eps]),
        decks=np.stack([e["decks"] for e in eps]), nobles=np.stack([e["nobles"] for e in eps]),
        n_calls=np.array([len(e["action"]) for e in eps], np.int32),
        action=cat("action", np.int16), reward=cat("reward", np.int8), terminated=cat("terminated", np.uint8),
        obs=np.concatenate([np.stack(e["obs"]) for e in eps]), snap=np.concatenate([np.stack(e["snap"]) for e in eps]),
        legal_len=np.array([len(l) for e in eps for l in e["legal"]], np.int32),
        legal_flat=np.concatenate([l for e in eps for l in e["legal"]]),
        n_all=np.array([len(e["all_actions"]) for e in eps], np.int32), all_actions=cat("all_actions", np.int16),
        pad=np.array([e["pad"] for e in eps], np.int32),
    )


# per file: (P, seed, first game id, episodes per learner seat, learner policies in rotation, opponents)
PLAN = [
    ("env_2p", 2, 31, 100, 6, ("random", "greedy", "random"), "philox"),
    ("env_3p", 3, 32, 200, 3, ("random", "greedy", "random"), "philox"),
    ("env_4p", 4, 33, 300, 2, ("random", "greedy"), "philox"),
    ("env_cap", None, 34, 400, 1, ("scripted",), "scripted"),  # 100-round cap: 2, 3 and 4 players, learner seat 0 / 1 / 3
    ("env_rand", -3, 35, 500, 6, ("random", "greedy"), "philox"),  # 3 players, the env draws the learner's seat and shuffles the opponents
]

if __name__ == "__main__":
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0"
    only = sys.argv[1:]
    for name, P, seed, first, per_seat, policies, opponents in PLAN:
        if only and name not in only:
            continue
        t0, eps, game = time.time(), [], first
        if P is not None and P < 0:
            for k in range(per_seat):
                eps.append(record_episode(-P, None, seed, game, policies[k % len(policies)], opponents, shuffle_turns=True, np_seed=7000 + k))
                game += 1
        elif P is None:
            # scripted hoarders on every seat.  2 players: a game that ends by score (id 400) and one that runs into the
            # 100-round cap on its own (id 401; found with the oracle).  3 and 4 players: turn counters preset to 95 (pad).
            for PP, seat, g, pad in ((2, 0, 400, 0), (2, 1, 401, 0), (2, 0, 402, 0), (3, 1, 403, 95), (4, 3, 404, 95), (4, 0, 405, 95)):
                eps.append(record_episode(PP, seat, seed, g, "scripted", "scripted", False, pad=pad))
        else:
            for seat in range(P):  # game ids of one seat are contiguous: SplendorVecEnv(fixed_turn=seat, first_game=..) replays them
                for k in range(per_seat):
                    eps.append(record_episode(P, seat, seed, game, policies[k % len(policies)], opponents, shuffle_turns=bool(k & 1)))
                    game += 1
        arrs = pack(eps)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **arrs)
        ended_on_opponent = sum(1 for e in eps if e["snap"][-1][27] == 0 and e["my_id"] != e["P"] - 1)
        capped = sum(1 for e in eps if all(e["snap"][-1][28 + 20 * i + 18] == 100 for i in range(e["P"])))
        print("%s: %d episodes, %d env calls, %d actions in all, %.1f s, %.0f KB; ended after an opponent's move %d; at the round cap %d" % (
            name, len(eps), int(arrs["n_calls"].sum()), int(arrs["n_all"].sum()), time.time() - t0, os.path.getsize(path) / 1024,
            ended_on_opponent, capped))

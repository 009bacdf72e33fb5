"""tests/golden/minimax.npz: the reference's MiniMaxAgent (agents/our_agents/minmax.py) recorded on real 2-player games.
Run:  PYTHONHASHSEED=0 python oracle/gen_golden_minimax.py      TEST INFRASTRUCTURE (needs /root/reference).

Games are driven by a mix of policies; at sampled steps the unmodified agent code gives
  * _evaluation_function(state) for both agents (float64 bit patterns),
  * the depth-2 value of EVERY root action of the seat to move, computed with the reference's own generateSuccessor /
    generatePredecessor / getLegalActions / _evaluation_function (the recursion of :43-88 without its alpha-beta cut),
  * the action MiniMaxAgent.SelectAction actually returns (with its shuffle, type sort and alpha-beta)."""
import copy
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_golden as G  # noqa: E402  (installs the reference on sys.path)
from splendor.agents.our_agents.minmax import MiniMaxAgent  # noqa: E402
from splendor.splendor.gym.envs.utils import create_action_mapping  # noqa: E402
from splendor.splendor.splendor_model import SplendorGameRule  # noqa: E402


def root_values(rule, state, me):
    agent = MiniMaxAgent(me)
    mapping = create_action_mapping(rule.getLegalActions(state, me), state, me)
    out = {}
    for idx, act in mapping.items():
        # a fresh copy per root action: generatePredecessor re-appends an un-bought reserved card at the END of the
        # agent's list, so inside the reference's own walk later siblings add their float terms in a different order
        # (last-bit noise that depends on its random.shuffle); the recorded values are those of the unpermuted state
        st = copy.deepcopy(state)
        act = create_action_mapping(rule.getLegalActions(st, me), st, me)[idx]
        rule.generateSuccessor(st, act, me)
        best = np.inf
        for reply in rule.getLegalActions(st, 1 - me):
            rule.generateSuccessor(st, reply, 1 - me)
            best = min(best, agent._evaluation_function(st))
            rule.generatePredecessor(st, reply, 1 - me)
        out[idx] = float(best)
    return out, mapping


def record(policy, seed, every):
    random.seed(seed)
    rule = SplendorGameRule(2)
    rng = random.Random(seed + 5)
    g = dict(decks=G.initial_deck_lists(rule),
             nobles=np.array([G.NOBLE_ID[n[0]] for n in rule.current_game_state.board.nobles] + [G.NONE] * 2, np.uint8),
             action=[], samples=[])
    t = 0
    while not rule.gameEnds() and t < 400:
        st, seat = rule.current_game_state, rule.current_agent_index
        mapping = create_action_mapping(rule.getLegalActions(st, seat), st, seat)
        if t % every == every - 1 or len(mapping) == 1:
            evals = [float(MiniMaxAgent(i)._evaluation_function(st)) for i in range(2)]
            before = G.snapshot(rule)
            vals, mp = root_values(rule, st, seat)
            random.seed(seed * 1000 + t)
            chosen = MiniMaxAgent(seat).SelectAction(list(mp.values()), copy.deepcopy(st), rule)  # Game hands agents a copy
            assert np.array_equal(before, G.snapshot(rule))
            chosen_idx = list(create_action_mapping([chosen], st, seat))  # the agent enumerates its own action dicts
            assert len(chosen_idx) == 1 and chosen_idx[0] in mp
            g["samples"].append(dict(step=t, evals=evals, actions=sorted(vals), values=[vals[i] for i in sorted(vals)],
                                     chosen=chosen_idx[0]))
        a = G.choose(policy, rng, mapping)
        rule.update(mapping[a])
        g["action"].append(a)
        t += 1
    # the final position too: its evaluation takes the max_score >= 15 branch
    g["final_evals"] = [float(MiniMaxAgent(i)._evaluation_function(rule.current_game_state)) for i in range(2)]
    return g


if __name__ == "__main__":
    plan = [("random", 9100, 9), ("greedy", 9101, 6), ("greedy", 9102, 7), ("hoard", 9103, 11), ("reserve", 9104, 8), ("greedy", 9105, 5)]
    games = [record(pol, seed, every) for pol, seed, every in plan]
    S = [(gi, s) for gi, g in enumerate(games) for s in g["samples"]]
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "minimax.npz"),
        decks=np.stack([g["decks"] for g in games]), nobles=np.stack([g["nobles"][:5] for g in games]).astype(np.uint8),
        n_steps=np.array([len(g["action"]) for g in games], np.int32),
        action=np.concatenate([np.array(g["action"], np.int16) for g in games]),
        final_evals=np.array([g["final_evals"] for g in games], np.float64),
        s_game=np.array([gi for gi, _ in S], np.int32), s_step=np.array([s["step"] for _, s in S], np.int32),
        s_evals=np.array([s["evals"] for _, s in S], np.float64), s_chosen=np.array([s["chosen"] for _, s in S], np.int16),
        s_n=np.array([len(s["actions"]) for _, s in S], np.int32),
        s_actions=np.concatenate([np.array(s["actions"], np.int16) for _, s in S]),
        s_values=np.concatenate([np.array(s["values"], np.float64) for _, s in S]))
    print("minimax golden:", len(games), "games,", len(S), "sampled roots,", int(sum(len(s["actions"]) for _, s in S)), "root actions")

import time, random, numpy as np, sys
sys.path.insert(0, "/root/repo")
from splendor_ai_b200 import template, game_rule as G
from splendor_ai_b200.env import SplendorEnv
random.seed(1); np.random.seed(1)
env = SplendorEnv(agents=[template.RandomAgent(0)], fixed_turn=1)
obs, info = env.reset(seed=1)
n = 0; t0 = time.perf_counter()
for ep in range(3):
    done = False
    while not done:
        mask = env.get_legal_actions_mask()
        a = int(np.random.choice(np.nonzero(mask)[0]))
        obs, r, done, _, _ = env.step(a); n += 1
    env.reset()
dt = time.perf_counter() - t0
print("SplendorEnv (host opponents): %.0f env.step/s (%d steps)" % (n / dt, n))
rule = G.SplendorGameRule(2); n = 0; t0 = time.perf_counter()
while not rule.gameEnds():
    acts = rule.getLegalActions(rule.current_game_state, rule.current_agent_index)
    rule.update(random.choice(acts)); n += 1
dt = time.perf_counter() - t0
print("SplendorGameRule loop: %.0f engine steps/s (%d steps)" % (n / dt, n))

"""Random ARBITRARY (not necessarily reachable) Splendor states for fuzzing the rule code: every corner the reference's
rules distinguish - agents holding 8/9/10 gems, drained bank colours, empty board slots and decks, 7 cards of a colour,
0-3 reserved cards, several nobles earned at once, passed flags, the 100-turn cap.  Returns both the oracle's struct and
the field arrays that `state.pack` turns into the packed device layout."""
import numpy as np

NONE = 0xFF
TIER_BASE, TIER_SIZE = (0, 40, 70), (40, 30, 20)


def random_state(rng, oracle, P=None):
    P = int(P or rng.integers(2, 5))
    n = {2: 4, 3: 5, 4: 7}[P]
    s = oracle.State()
    s.P, s.cur = P, int(rng.integers(0, P))
    f = dict(bank=np.zeros((1, 6), np.uint8), dealt=np.full((1, 12), NONE, np.uint8), nobles=np.full((1, 5), NONE, np.uint8),
             deck_n=np.zeros((1, 3), np.uint8), cur=np.array([s.cur]), gems=np.zeros((1, P, 6), np.uint8),
             cards=np.zeros((1, P, 5), np.uint8), resv=np.full((1, P, 3), NONE, np.uint8), score=np.zeros((1, P), np.uint8),
             passed=np.zeros((1, P), np.uint8), turns=np.zeros((1, P), np.uint16), nobles_mask=np.zeros((1, P), np.uint16))
    # gems: distribute each colour's supply between the bank and the agents, biased towards rich agents
    target = rng.integers(0, 11, P) if rng.random() < 0.7 else np.full(P, 10)
    held = np.zeros(P, int)
    for c in rng.permutation(6):
        supply = 5 if c == 5 else n
        for _ in range(supply):
            cand = [p for p in range(P) if held[p] < target[p]]
            if cand and rng.random() < 0.75:
                p = int(rng.choice(cand))
                f["gems"][0, p, c] += 1
                held[p] += 1
            else:
                f["bank"][0, c] += 1
    # cards on the board, in reserve and in the decks
    lists = np.zeros(90, np.uint8)
    for t in range(3):
        ids = rng.permutation(np.arange(TIER_BASE[t], TIER_BASE[t] + TIER_SIZE[t]))
        k = 0
        for j in range(4):
            if rng.random() < 0.9:
                f["dealt"][0, 4 * t + j] = ids[k]; k += 1
        for p in range(P):
            for r in range(3):
                if f["resv"][0, p, r] == NONE and (r == 0 or f["resv"][0, p, r - 1] != NONE) and rng.random() < 0.25:
                    f["resv"][0, p, r] = ids[k]; k += 1
        left = 0 if rng.random() < 0.15 else int(rng.integers(0, TIER_SIZE[t] - k + 1))
        lists[TIER_BASE[t]:TIER_BASE[t] + left] = ids[k:k + left]
        f["deck_n"][0, t] = left
    for p in range(P):  # compact each reserved list (ids must be a prefix)
        r = [x for x in f["resv"][0, p] if x != NONE]
        f["resv"][0, p] = r + [NONE] * (3 - len(r))
    if rng.random() < 0.06:  # deadlock corner: nothing to collect, nothing to reserve, (almost) nothing affordable
        f["bank"][0, :5] = 0
        f["bank"][0, 5] = rng.integers(0, 2)
        p = s.cur
        f["gems"][0, p] = 0
        f["gems"][0, p, rng.integers(0, 6)] = rng.integers(0, 2)
        pool = [int(x) for t in range(3) for x in range(TIER_BASE[t], TIER_BASE[t] + TIER_SIZE[t])
                if x not in set(f["dealt"][0].tolist()) and all(x not in set(f["resv"][0, q].tolist()) for q in range(P)) and x not in set(lists.tolist())]
        have = [x for x in f["resv"][0, p] if x != NONE]
        while len(have) < 3:
            have.append(pool.pop())
        f["resv"][0, p] = have
    nn = int(rng.integers(0, P + 2))
    f["nobles"][0, :nn] = rng.permutation(10)[:nn]
    for p in range(P):
        rich = rng.random() < 0.4
        f["cards"][0, p] = rng.integers(0, 8 if rich else 4, 5) if rng.random() < 0.85 else 0
        f["score"][0, p] = rng.integers(0, 22)
        f["passed"][0, p] = rng.random() < 0.15
        f["turns"][0, p] = 100 if rng.random() < 0.05 else rng.integers(0, 100)
        f["nobles_mask"][0, p] = 0
    # mirror into the oracle struct
    for c in range(6):
        s.bank[c] = int(f["bank"][0, c])
    for i in range(12):
        s.dealt[i] = int(f["dealt"][0, i])
    for k in range(5):
        s.nobles[k] = int(f["nobles"][0, k])
    s.n_nobles = nn
    for t in range(3):
        s.deck_n[t] = int(f["deck_n"][0, t])
        for i in range(TIER_SIZE[t]):
            s.deck[t][i] = int(lists[TIER_BASE[t] + i])
    for p in range(P):
        pl = s.pl[p]
        for c in range(6):
            pl.gems[c] = int(f["gems"][0, p, c])
        for c in range(5):
            pl.cards[c] = int(f["cards"][0, p, c])
        for r in range(3):
            pl.resv[r] = int(f["resv"][0, p, r])
        pl.n_resv = int((f["resv"][0, p] != NONE).sum())
        pl.score, pl.passed, pl.turns = int(f["score"][0, p]), int(f["passed"][0, p]), int(f["turns"][0, p])
        pl.n_nobles, pl.nobles_mask = 0, 0
    return s, f, lists

"""The fixed discrete action space of the reference's gym adapter, in closed form.

Mirrors ``splendor/splendor/gym/envs/actions.py`` (ActionEnum :29-39, CardPosition :42-50, Action :56-106,
ALL_ACTIONS :222-237) and ``gym/envs/utils.py`` (build_action :43-145): same names, same field meanings, same 3510
entries in the same order - but every entry is DECODED from its index arithmetically instead of being generated with
itertools, and ``action_index`` is the O(1) inverse of the reference's ``ALL_ACTIONS.index(to_action_element(...))``
linear search.  The kernels use the same index arithmetic (csrc/engine.cuh).

Layout (n = noble slot 0..5: 0 none, k+1 = board noble at list position k; p = tier*4 + column):
  PASS           0    + n
  RESERVE        6    + p*42 + n*7 + v      v: 0 no yellow | 1 yellow | 2+c yellow and return NORMAL_COLORS[c]
  COLLECT_SAME   510  + c*126 + n*21 + v    v: 0 | 1..10 pairs of the other colours | 11..15 one of | 16..20 two of
  COLLECT_DIFF   1140 + ...                 per length L=1,2,3: combos of NORMAL_COLORS x n x (none + multisets of 1..L)
  BUY_RESERVE    3420 + r*6 + n
  BUY_AVAILABLE  3438 + p*6 + n
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum, auto
from functools import lru_cache

NORMAL_COLORS = ["black", "red", "green", "blue", "white"]          # constants.py:12
ALL_GEMS_COLORS = ["black", "red", "yellow", "green", "blue", "white"]  # splendor_utils.py:124-131
NUM_ACTIONS = 3510
MAX_NOBLES, MAX_RESERVED, MAX_TIER_CARDS, NUMBER_OF_TIERS = 5, 3, 4, 3

BASE_PASS, BASE_RESERVE, BASE_SAME, BASE_DIFF, BASE_BUY_RESERVE, BASE_BUY_AVAILABLE = 0, 6, 510, 1140, 3420, 3438
_DIFF_BASE = {1: 0, 2: 180, 3: 1080}
_DIFF_V = {1: 6, 2: 15, 3: 20}


class ActionEnum(Enum):
    PASS = auto()
    COLLECT_SAME = auto()
    COLLECT_DIFF = auto()
    RESERVE = auto()
    BUY_AVAILABLE = auto()
    BUY_RESERVE = auto()


@dataclass
class CardPosition:
    tier: int
    card_index: int
    reserved_index: int


@dataclass
class Action:
    type_enum: ActionEnum
    collected_gems: dict | None = None
    returned_gems: dict | None = None
    position: CardPosition | None = None
    noble_index: int | None = None


def _combos(L):
    """itertools.combinations(range(5), L) in lexicographic order, written out."""
    out = []
    for a in range(5):
        if L == 1:
            out.append((a,))
            continue
        for b in range(a + 1, 5):
            if L == 2:
                out.append((a, b))
                continue
            for c in range(b + 1, 5):
                out.append((a, b, c))
    return out


def _multisets(k, r):
    """combinations_with_replacement(range(k), r) in lexicographic order."""
    if r == 0:
        return [()]
    out = []

    def rec(start, prefix):
        if len(prefix) == r:
            out.append(tuple(prefix))
            return
        for x in range(start, k):
            rec(x, prefix + [x])

    rec(0, [])
    return out


def _count(items):
    d = {}
    for x in items:
        d[x] = d.get(x, 0) + 1
    return d


def decode(index: int) -> Action:
    """ALL_ACTIONS[index] without the list (a fresh object per call; the field values are cached)."""
    t, collected, returned, pos, noble = _decode_fields(int(index))
    return Action(t, None if collected is None else dict(collected), None if returned is None else dict(returned),
                  None if pos is None else CardPosition(*pos), noble)


@lru_cache(maxsize=None)
def _decode_fields(index: int):
    a = _decode_uncached(index)
    pos = None if a.position is None else (a.position.tier, a.position.card_index, a.position.reserved_index)
    return a.type_enum, a.collected_gems, a.returned_gems, pos, a.noble_index


def _decode_uncached(index: int) -> Action:
    if not 0 <= index < NUM_ACTIONS:
        raise ValueError(f"The action {index} isn't a valid action")
    noble = lambda n: None if n == 0 else n - 1
    if index < BASE_RESERVE:
        return Action(ActionEnum.PASS, noble_index=noble(index))
    if index < BASE_SAME:
        r = index - BASE_RESERVE
        p, n, v = r // 42, (r % 42) // 7, r % 7
        return Action(ActionEnum.RESERVE, {} if v == 0 else {"yellow": 1}, {} if v < 2 else {NORMAL_COLORS[v - 2]: 1},
                      CardPosition(p // 4, p % 4, -1), noble(n))
    if index < BASE_DIFF:
        r = index - BASE_SAME
        c, n, v = r // 126, (r % 126) // 21, r % 21
        others = [x for x in ALL_GEMS_COLORS if x != NORMAL_COLORS[c]]
        if v == 0:
            ret = {}
        elif v <= 10:
            i, j = _multiset_pairs()[v - 1]
            ret = {others[i]: 1, others[j]: 1}
        elif v <= 15:
            ret = {others[v - 11]: 1}
        else:
            ret = {others[v - 16]: 2}
        return Action(ActionEnum.COLLECT_SAME, {NORMAL_COLORS[c]: 2}, ret, None, noble(n))
    if index < BASE_BUY_RESERVE:
        r = index - BASE_DIFF
        L = 1 if r < 180 else (2 if r < 1080 else 3)
        r -= _DIFF_BASE[L]
        V = _DIFF_V[L]
        combo = _combos(L)[r // (6 * V)]
        n, v = (r % (6 * V)) // V, r % V
        others = [x for x in ALL_GEMS_COLORS if x not in [NORMAL_COLORS[i] for i in combo]]
        ret, k = {}, len(others)
        if v > 0:
            v -= 1
            for size in range(1, L + 1):
                ms = _multisets(k, size)
                if v < len(ms):
                    ret = _count(others[i] for i in ms[v])
                    break
                v -= len(ms)
        return Action(ActionEnum.COLLECT_DIFF, {NORMAL_COLORS[i]: 1 for i in combo}, ret, None, noble(n))
    if index < BASE_BUY_AVAILABLE:
        r = index - BASE_BUY_RESERVE
        return Action(ActionEnum.BUY_RESERVE, None, None, CardPosition(-1, -1, r // 6), noble(r % 6))
    r = index - BASE_BUY_AVAILABLE
    p = r // 6
    return Action(ActionEnum.BUY_AVAILABLE, None, None, CardPosition(p // 4, p % 4, -1), noble(r % 6))


@lru_cache(maxsize=None)
def _multiset_pairs():
    return [(i, j) for i in range(5) for j in range(i + 1, 5)]


def encode(type_name: str, noble_index: int | None = None, position: int | None = None,
           collected: dict | None = None, returned: dict | None = None) -> int:
    """Index of an action given by its fields (type as the reference's string, e.g. "collect_diff").

    position: board slot tier*4+col (reserve / buy_available) or reserved index (buy_reserve)."""
    n = 0 if noble_index is None else noble_index + 1
    collected, returned = collected or {}, returned or {}
    t = type_name.lower()
    if t == "pass":
        return n
    if t == "reserve":
        v = 0 if not collected else 1
        for c, cnt in returned.items():
            if cnt:
                v = 2 + NORMAL_COLORS.index(c)
        return BASE_RESERVE + position * 42 + n * 7 + v
    if t == "collect_same":
        (colour, _), = collected.items()
        c = NORMAL_COLORS.index(colour)
        others = [x for x in ALL_GEMS_COLORS if x != colour]
        pos = sorted(others.index(k) for k, cnt in returned.items() for _ in range(cnt))
        if not pos:
            v = 0
        elif len(pos) == 1:
            v = 11 + pos[0]
        elif pos[0] == pos[1]:
            v = 16 + pos[0]
        else:
            v = 1 + _multiset_pairs().index((pos[0], pos[1]))
        return BASE_SAME + c * 126 + n * 21 + v
    if t == "collect_diff":
        combo = tuple(sorted(NORMAL_COLORS.index(c) for c in collected))
        L, V = len(combo), _DIFF_V[len(combo)]
        others = [x for x in ALL_GEMS_COLORS if x not in collected]
        pos = tuple(sorted(others.index(k) for k, cnt in returned.items() for _ in range(cnt)))
        v = 0
        if pos:
            v = 1 + sum(len(_multisets(len(others), s)) for s in range(1, len(pos))) + _multisets(len(others), len(pos)).index(pos)
        return BASE_DIFF + _DIFF_BASE[L] + _combos(L).index(combo) * 6 * V + n * V + v
    if t == "buy_reserve":
        return BASE_BUY_RESERVE + position * 6 + n
    if t == "buy_available":
        return BASE_BUY_AVAILABLE + position * 6 + n
    raise ValueError(f"Unknown action type: {type_name}")


class _AllActions:
    """Sequence view with the reference's ALL_ACTIONS interface (len, indexing, iteration, .index)."""

    def __len__(self):
        return NUM_ACTIONS

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [decode(j) for j in range(*i.indices(NUM_ACTIONS))]
        return decode(i if i >= 0 else NUM_ACTIONS + i)

    def __iter__(self):
        return (decode(i) for i in range(NUM_ACTIONS))

    def index(self, action: Action) -> int:
        pos = None
        if action.position is not None:
            pos = action.position.reserved_index if action.type_enum == ActionEnum.BUY_RESERVE else \
                action.position.tier * 4 + action.position.card_index
        idx = encode(action.type_enum.name, action.noble_index, pos, action.collected_gems, action.returned_gems)
        if decode(idx) != action:
            raise ValueError(f"{action} is not in list")
        return idx


ALL_ACTIONS = _AllActions()

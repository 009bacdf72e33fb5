"""Single-game drop-in for the reference's rule object, backed by the CUDA kernels.

``SplendorGameRule`` / ``SplendorState`` / ``Card`` keep the names, attributes and method signatures of
``splendor/splendor/splendor_model.py`` (Card :24-48, SplendorState :52-138, SplendorGameRule :142-576) and
``LimitRoundsGameRule`` those of ``splendor/splendor/utils.py:9-25``, so callers such as ``Game.Run`` (game.py:104-213),
the gym env and search agents work unchanged:

  * ``getLegalActions`` runs the legal-mask kernel and returns the reference's action dicts, in ascending ALL_ACTIONS
    order (the reference's own list order depends on PYTHONHASHSEED; the index order is its only canonical one);
  * ``generateSuccessor`` converts the dict to its ALL_ACTIONS index and runs the step kernel; an action that is not
    legal raises ``ValueError`` (the reference would silently corrupt the state);
  * ``initialGameState`` draws the nobles and shuffles the decks with Python's ``random`` in the reference's call order,
    so ``random.seed(s)`` gives the same game as the reference.

Every rule decision is computed on the GPU.  The Python objects only carry what the packed state does not - WHICH
cards an agent bought, the noble tuples, the action/reward trace - as presentation for callers that read them.
"""
from __future__ import annotations

import random
from functools import lru_cache

import numpy as np

from . import actions as A
from . import template
from .state import NONE, TIER_BASE, TIER_SIZE, state_words, unpack_one
from .tables import CARD_TABLE, NOBLE_TABLE

COLOURS = {"B": "black", "r": "red", "y": "yellow", "g": "green", "b": "blue", "w": "white"}
_COLS = ["black", "red", "green", "blue", "white", "yellow"]  # colour ids
ROUNDS_LIMIT = 100


class Card:
    def __init__(self, colour, code, cost, deck_id, points):
        self.colour, self.code, self.cost, self.deck_id, self.points = colour, code, cost, deck_id, points

    def __str__(self):
        gems = ", ".join(f"{n} {c}" for c, n in self.cost.items())
        return f"Tier {self.deck_id + 1} {self.colour} card worth {self.points} points and costing {gems}"

    def __repr__(self):
        return self.code

    def __eq__(self, other):
        return hasattr(other, "code") and other.code == self.code and self.points == other.points

    def __hash__(self):
        return hash(self.code)


@lru_cache(maxsize=None)
def card_of(card_id: int) -> Card:
    colour, cost, tier, points, code = CARD_TABLE[card_id]
    return Card(_COLS[colour], code, {_COLS[c]: n for c, n in enumerate(cost) if n}, tier, points)


CARD_ID = {row[4]: i for i, row in enumerate(CARD_TABLE)}
NOBLES = [(code, {_COLS[c]: n for c, n in enumerate(cost) if n}) for code, cost in NOBLE_TABLE]
NOBLE_ID = {code: i for i, (code, _) in enumerate(NOBLES)}


class AgentTrace:
    def __init__(self, pid):
        self.id = pid
        self.action_reward = []


def draw_initial_decks(num_agents):
    """BoardState.__init__ (splendor_model.py:83-90) restated on ids: sample the nobles first, then shuffle tier 1, 2, 3 -
    the same `random` calls on lists of the same lengths, hence the same game as the reference for the same seed."""
    nobles = random.sample(range(len(NOBLES)), k=num_agents + 1)
    lists = []
    for t in range(3):
        deck = list(range(TIER_BASE[t], TIER_BASE[t] + TIER_SIZE[t]))
        random.shuffle(deck)
        lists += deck
    return lists, nobles


class SplendorState(template.GameState):
    """The reference's object graph, decoded from the packed state words."""

    class BoardState:
        pass

    class AgentState:
        pass

    def __init__(self, num_agents, _deck_lists=None, _nobles=None):
        from . import _single

        self.num_agents = num_agents
        if _deck_lists is None:
            _deck_lists, _nobles = draw_initial_decks(num_agents)
        self._deck_lists = np.asarray(_deck_lists, np.uint8)
        self._words = _single.inject(num_agents, self._deck_lists, _nobles)
        self._bought = [{c: [] for c in COLOURS.values()} for _ in range(num_agents)]
        self._noble_tuples = [[] for _ in range(num_agents)]
        self._traces = [AgentTrace(i) for i in range(num_agents)]
        self._last_action = [None] * num_agents
        self._applied = []  # (agent_id, ALL_ACTIONS index) of every successor, so a predecessor can be taken in LIFO order
        self.agent_to_move = 0
        self._refresh()

    def _refresh(self):
        u = unpack_one(self._words, self.num_agents)
        b = SplendorState.BoardState()
        bank = dict(zip(_COLS, u["bank"]))
        b.gems = {c: bank[c] for c in COLOURS.values()}
        b.dealt = [[None if u["dealt"][t * 4 + j] == NONE else card_of(u["dealt"][t * 4 + j]) for j in range(4)] for t in range(3)]
        b.nobles = [NOBLES[i] for i in u["nobles"] if i != NONE]
        lists = self._deck_lists
        b.decks = [[card_of(int(c)) for c in lists[TIER_BASE[t]:TIER_BASE[t] + u["deck_n"][t]]] for t in range(3)]
        b.dealt_list = lambda: [card for deck in b.dealt for card in deck if card]
        self.board = b
        self.agents = []
        for i in range(self.num_agents):
            a = SplendorState.AgentState()
            a.id = i
            a.score = u["score"][i]
            held = dict(zip(_COLS, u["gems"][i]))
            a.gems = {c: held[c] for c in COLOURS.values()}
            a.cards = {c: list(v) for c, v in self._bought[i].items()}
            a.cards["yellow"] = [card_of(c) for c in u["resv"][i] if c != NONE]
            a.nobles = list(self._noble_tuples[i])
            a.passed = bool(u["passed"][i])
            a.agent_trace = self._traces[i]
            a.last_action = self._last_action[i]
            self.agents.append(a)

    def __str__(self):
        out = "\nAvailable Gems:\n" + str(self.board.gems) + "\nDealt Card List: \n"
        out += "".join("\t" + str(c) + "\n" for c in self.board.dealt_list())
        out += "\nNoble List \n" + str(self.board.nobles) + "\n"
        for a in self.agents:
            out += "Agent (%d): \n\tscore: %d,\n\tgems: %s\n\tcards: %s\n\tnobles: %s.\n" % (a.id, a.score, a.gems, a.cards, a.nobles)
        return out


def action_index(action: dict, state: SplendorState, agent_id: int) -> int:
    """Action.to_action_element + ALL_ACTIONS.index (gym/envs/actions.py:68-106, utils.py:160-162) in O(1)."""
    noble_index = None
    if action.get("noble") is not None:
        noble_index = [n[0] for n in state.board.nobles].index(action["noble"][0])
    position = None
    if action.get("card") is not None:
        if action["type"] == "buy_reserve":
            position = state.agents[agent_id].cards["yellow"].index(action["card"])
        else:
            tier = action["card"].deck_id
            position = tier * 4 + state.board.dealt[tier].index(action["card"])
    buy = action["type"].startswith("buy")
    return A.encode(action["type"], noble_index, position, None if buy else action.get("collected_gems"),
                    None if buy else action.get("returned_gems"))


def _payment(agent, card) -> dict:
    """The returned_gems entry of a buy action dict (resources_sufficient's return combo, :371-394) - presentation only;
    whether the purchase is legal and what the step does with the gems is decided by the kernels."""
    out, wild = {}, 0
    for colour, cost in card.cost.items():
        gem_cost = max(cost - len(agent.cards[colour]), 0)
        short = max(gem_cost - agent.gems[colour], 0)
        if gem_cost - short:
            out[colour] = gem_cost - short
        wild += short
    if wild:
        out["yellow"] = wild
    return out


def build_action(index: int, state: SplendorState, agent_id: int) -> dict:
    """ALL_ACTIONS index -> the reference's action dict for this state (gym/envs/utils.py:43-145)."""
    a = A.decode(index)
    noble = state.board.nobles[a.noble_index] if a.noble_index is not None and a.noble_index < len(state.board.nobles) else None
    t = a.type_enum.name.lower()
    if t == "pass":
        return {"type": "pass", "noble": noble}
    if t in ("collect_same", "collect_diff"):
        return {"type": t, "collected_gems": a.collected_gems, "returned_gems": a.returned_gems, "noble": noble}
    agent = state.agents[agent_id]
    if t == "buy_reserve":
        if a.position.reserved_index >= len(agent.cards["yellow"]):
            raise ValueError(f"Can't build action {a} since there is not card to buy!")
        card = agent.cards["yellow"][a.position.reserved_index]
        pos = (3, a.position.reserved_index)
    else:
        card = state.board.dealt[a.position.tier][a.position.card_index]
        pos = (a.position.tier, a.position.card_index)
        if card is None and t == "buy_available":
            raise ValueError(f"Can't build action {a} since there is not card to buy!")
    if t == "reserve":
        return {"type": "reserve", "card": card, "card_position": pos, "collected_gems": a.collected_gems,
                "returned_gems": a.returned_gems, "noble": noble}
    return {"type": t, "card": card, "card_position": pos, "returned_gems": _payment(agent, card), "noble": noble}


class SplendorGameRule(template.GameRule):
    LIMIT_ROUNDS = False

    def __init__(self, num_of_agent):
        super().__init__(num_of_agent)
        self.private_information = None

    def initialGameState(self):
        return SplendorState(self.num_of_agent)

    def legal_action_indices(self, game_state, agent_id) -> np.ndarray:
        from . import _single

        return _single.legal_indices(game_state._words, self.num_of_agent, agent_id)

    def getLegalActions(self, game_state, agent_id):
        return [build_action(int(i), game_state, agent_id) for i in self.legal_action_indices(game_state, agent_id)]

    def generateSuccessor(self, state, action, agent_id):
        from . import _single

        idx = action if isinstance(action, (int, np.integer)) else action_index(action, state, agent_id)
        act = build_action(int(idx), state, agent_id) if isinstance(action, (int, np.integer)) else action
        reward, illegal = _single.step(state._words, state._deck_lists, self.num_of_agent, agent_id, int(idx))
        if illegal:
            raise ValueError(f"the action {idx} ({act}) is illegal!")
        if "card" in act and act["type"].startswith("buy"):
            state._bought[agent_id][act["card"].colour].append(act["card"])
        if act.get("noble") is not None:
            state._noble_tuples[agent_id].append(act["noble"])
        state._traces[agent_id].action_reward.append((act, reward))
        state._last_action[agent_id] = act
        state._applied.append((agent_id, int(idx)))
        state._refresh()
        return state

    def generatePredecessor(self, state, action, agent_id):
        """Undo `action` of `agent_id` (splendor_model.py:156-220), most recent first as search agents do
        (agents/our_agents/minmax.py:43-88).  Runs the undo kernel; the reference's list-order quirks are kept."""
        from . import _single

        if not state._applied or state._applied[-1][0] != agent_id:
            raise ValueError("generatePredecessor must undo the most recent action of that agent")
        _, idx = state._applied.pop()
        card = CARD_ID[action["card"].code] if action.get("card") is not None else NONE
        noble = NOBLE_ID[action["noble"][0]] if action.get("noble") is not None else NONE
        returned = [int((action.get("returned_gems") or {}).get(c, 0)) for c in _COLS]
        _single.undo(state._words, self.num_of_agent, agent_id, idx, card, noble, returned)
        if action["type"].startswith("buy"):
            state._bought[agent_id][action["card"].colour].remove(action["card"])
        if action.get("noble") is not None:
            state._noble_tuples[agent_id].remove(action["noble"])
        state._traces[agent_id].action_reward.pop()
        state._last_action[agent_id] = action
        state._refresh()
        return state

    def gameEnds(self):
        from . import _single

        st = self.current_game_state
        return _single.game_ends(st._words, self.num_of_agent, self.current_agent_index, self.LIMIT_ROUNDS) != 0

    def calScore(self, game_state, agent_id):
        from . import _single

        return _single.cal_score2(game_state._words, self.num_of_agent)[agent_id] / 2


class LimitRoundsGameRule(SplendorGameRule):
    """splendor/utils.py:9-25: also ends when every agent has made ROUNDS_LIMIT turns."""

    LIMIT_ROUNDS = True

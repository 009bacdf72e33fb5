"""Gym-style environments on the B200 engine.

``SplendorEnv`` keeps the constructor, spaces, ``reset`` / ``step`` / ``get_legal_actions_mask`` signatures, properties
and ValueErrors of the reference's single-learner env (``splendor/splendor/gym/envs/splendor_env.py:20-245``): the
learner picks an ``ALL_ACTIONS`` index, the opponents (any ``Agent`` objects) are simulated until it is the learner's
turn again, the observation is the 265-float feature vector, the reward is the learner's score delta, and the game is
capped at 100 rounds (``LimitRoundsGameRule``).  Every rule decision goes through the CUDA kernels.

``SplendorVecEnv`` is the batched form the PPO rollout needs (BASELINE.json config 4): N games resident in HBM, one fused
kernel per call = learner moves + in-kernel ``RandomAgent`` opponents + next legal-action masks + observations, with
gymnasium's next-step auto-reset.  Nothing crosses the host boundary unless the caller asks for it.
(gymnasium itself is not a dependency: the spaces below are two tiny stand-ins with the attributes callers read.)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from . import _single
from .engine import MASK_WORDS, NUM_ACTIONS, OBS_SIZE, BatchedSplendor
from .game_rule import LimitRoundsGameRule, SplendorState


class Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = n, (), np.int64

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class SplendorEnv:
    """Single-learner env with host-side opponents (drop-in for gym.make("splendor-v1", agents=...).unwrapped)."""

    metadata = {"render_modes": []}

    def __init__(self, agents, shuffle_turns: bool = True, fixed_turn: int | None = None, render_mode: str | None = None):
        _ = render_mode
        self.fixed_turn = fixed_turn
        self.shuffle_turns = shuffle_turns
        self.agents = np.array(agents)
        self.number_of_players = len(self.agents) + 1
        self.game_rule = LimitRoundsGameRule(self.number_of_players)
        self.action_space = Discrete(NUM_ACTIONS)
        self.observation_space = Box(-np.inf, np.inf, (OBS_SIZE,))
        self.my_turn = -1

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        if seed is not None:
            np.random.seed(seed)
        self.game_rule = LimitRoundsGameRule(self.number_of_players)
        if self.shuffle_turns:
            np.random.shuffle(self.agents)
        if self.fixed_turn is None or self.fixed_turn not in range(self.number_of_players):
            self.my_turn = int(np.random.randint(0, self.number_of_players))
        else:
            self.my_turn = self.fixed_turn
        self._set_opponents_ids()
        self._simulate_opponents()
        return self._vectorize(self.state, self.my_turn), {"my_id": self.my_turn}

    def step(self, action: int):
        if action not in range(NUM_ACTIONS):
            raise ValueError(f"The action {action} isn't a valid action")
        previous_score = self.state.agents[self.my_turn].score
        legal = self.game_rule.legal_action_indices(self.state, self.my_turn)
        if action not in legal:
            raise ValueError(f"the action {action} is illegal!")
        self.game_rule.update(int(action))
        current_score = self.state.agents[self.my_turn].score
        terminated, next_state = self._simulate_opponents()
        return self._vectorize(next_state, self.my_turn), current_score - previous_score, terminated, False, {}

    def render(self) -> None:
        pass

    def get_legal_actions_mask(self) -> np.ndarray:
        mask = np.zeros(NUM_ACTIONS)
        mask[self.game_rule.legal_action_indices(self.state, self.my_turn)] = 1
        return mask

    def _get_opponent_by_turn(self, turn: int):
        for agent in self.agents:
            if agent.id == turn:
                return agent
        raise ValueError(f"Invalid turn ({turn})")

    @staticmethod
    def _vectorize(state: SplendorState, turn: int) -> np.ndarray:
        return _single.observe(state._words, state.num_agents, turn)

    def _set_opponents_ids(self) -> None:
        ids = list(range(self.number_of_players))
        ids.remove(self.my_turn)
        for agent, agent_id in zip(self.agents, ids, strict=True):
            agent.id = agent_id

    def _simulate_opponents(self):
        while self.turn != self.my_turn and not self.game_rule.gameEnds():
            available_actions = self.game_rule.getLegalActions(self.state, self.turn)
            agent = self._get_opponent_by_turn(self.turn)
            action = agent.SelectAction(available_actions, self.state, self.game_rule)
            self.game_rule.update(action)
        return self.game_rule.gameEnds(), self.state

    @property
    def turn(self) -> int:
        return self.game_rule.current_agent_index

    @property
    def state(self) -> SplendorState:
        return self.game_rule.current_game_state


def unpack_mask(mask_words: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """[N,110] int32 bit-packed legal mask -> [N,3510] dense 0/1 tensor (what the reference's policy multiplies by)."""
    shifts = torch.arange(32, device=mask_words.device, dtype=torch.int32)
    bits = (mask_words.unsqueeze(-1) >> shifts) & 1
    return bits.reshape(mask_words.shape[0], -1)[:, :NUM_ACTIONS].to(dtype)


class SplendorVecEnv:
    """N single-learner envs stepped by one fused kernel per call, RandomAgent opponents in-kernel, auto-reset.

    self_play=True: no in-kernel opponents - every call applies the action of the seat to move (info["acted"]) and returns
    the next mover's observation and mask (info["my_id"]), so one policy can play all seats.
    opponent="minimax" (2 players): the opponent is the reference's MiniMaxAgent (agents/our_agents/minmax.py), the default
    PPO test opponent (arguments_parsing.py:107), searched on the GPU for all envs at once (spl_minimax2).
    opponent=<callable>: every seat that is not the learner's is played by a (frozen) POLICY, as the reference does for
    `--opponent itself` and for saved PPO opponents (ppo/ppo.py:176-196: PPOAgent objects that share or load a network;
    ppo_agent.py:32-64: observation of the agent's own seat -> masked network -> arg-max).  The callable receives
    (obs float32 [N,265] of the seats to move, packed legal masks int32 [N,110], seats int8 [N]) for ALL envs and returns
    either float32 logits [N,>=3510] - the opponent then takes the arg-max over its legal actions (opponent_greedy=True, the
    reference's behaviour) or a masked categorical sample (opponent_greedy=False; spl_masked_sample keyed by the env's game
    id and turn) - or an int16 [N] tensor of action indices.  Rows of envs that are at their learner or already over are
    ignored.  Everything stays on the device; a call makes at most P-1 policy evaluations.

    reset()  -> (obs [N,265] f32, {"my_id": int8 [N]})
    step(a)  -> (obs, reward f32 [N], terminated bool [N], truncated bool [N] (always False), info)
                info: "legal_mask" [N,110] int32 bit-packed (see unpack_mask), "illegal" bool [N], "my_id"
    An env whose game ended is reset by the NEXT step call, which ignores that env's action and returns the first
    observation of the new game (gymnasium's next-step auto-reset).  An illegal action leaves the env untouched and is
    reported in info["illegal"] (raise_on_illegal=True turns it into the reference's ValueError).
    """

    def __init__(self, num_envs: int, num_agents: int = 2, seed: int = 1234, first_game: int = 0, shuffle_turns: bool = True,
                 fixed_turn: int | None = None, device="cuda", raise_on_illegal: bool = False, self_play: bool = False,
                 opponent="random", opponent_greedy: bool = True):
        if not callable(opponent) and opponent not in ("random", "minimax"):
            raise ValueError(f"unknown opponent {opponent!r}")
        if opponent == "minimax" and (num_agents != 2 or self_play):
            raise ValueError("the minimax opponent supports only a game of 2 players (and no self-play)")
        if callable(opponent) and self_play:
            raise ValueError("a policy opponent and self_play exclude each other (self_play already hands every seat to the caller)")
        self.opponent = "policy" if callable(opponent) else opponent
        self.opponent_policy = opponent if callable(opponent) else None
        self.opponent_greedy = opponent_greedy
        self.opponent_seed = seed
        self.num_envs, self.number_of_players = int(num_envs), int(num_agents)
        self.engine = BatchedSplendor(num_envs, num_agents, seed=seed, first_game=first_game, limit_rounds=True, device=device)
        self.device = self.engine.device
        self.action_space = Discrete(NUM_ACTIONS)
        self.single_action_space = self.action_space
        self.observation_space = Box(-np.inf, np.inf, (num_envs, OBS_SIZE))
        self.single_observation_space = Box(-np.inf, np.inf, (OBS_SIZE,))
        self.shuffle_turns, self.fixed_turn, self.raise_on_illegal = shuffle_turns, fixed_turn, raise_on_illegal
        self.self_play = self_play  # the caller's policy plays every seat (ppo.py --opponent itself); info["my_id"] = seat to move
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self._base = first_game
        self._episode = torch.zeros(num_envs, dtype=torch.int64, device=self.device)
        self._ids = None
        self._done = torch.zeros(num_envs, dtype=torch.bool, device=self.device)
        self._policy_calls = 0  # the `step` key of the opponents' sampled draws (with the env's game id)
        self.my_turn = torch.zeros(num_envs, dtype=torch.int8, device=self.device)
        self.legal_mask = None

    def _draw_seats(self, n):
        if self.fixed_turn is not None and self.fixed_turn in range(self.number_of_players):
            return torch.full((n,), self.fixed_turn, dtype=torch.int8)
        return torch.randint(0, self.number_of_players, (n,), generator=self._gen, dtype=torch.int8)

    def _game_ids(self):
        return self._base + torch.arange(self.num_envs, device=self.device, dtype=torch.int64) + self._episode * self.num_envs

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        if seed is not None:
            self._gen.manual_seed(seed)
        self._episode.zero_()
        self._ids = self._game_ids()
        self.engine.reset(game_ids=self._ids)
        self.my_turn = self._draw_seats(self.num_envs).to(self.device)
        self._done.zero_()
        skip = torch.full((self.num_envs,), -1, dtype=torch.int16, device=self.device)
        if self.opponent == "minimax":
            self._minimax_opponents_move()
        elif self.opponent == "policy":
            self._policy_opponents_move()
        obs, _, term, _, mask = self.engine.env_step(self.my_turn, skip, self_play=self.self_play)
        if self.self_play:
            self.my_turn = self._seat_to_move()
        self._done = term != 0
        self.legal_mask = mask
        return obs, {"my_id": self.my_turn, "legal_mask": mask}

    def step(self, actions: torch.Tensor):
        a = actions.to(device=self.device, dtype=torch.int16)
        if self.raise_on_illegal and self.legal_mask is not None:
            # the reference raises before it mutates anything (splendor_env.py:139-153): test the action bits of the current
            # masks first, so that a caught ValueError leaves every env, _done and legal_mask as they were
            idx = a.to(torch.int64)
            in_range = (idx >= 0) & (idx < NUM_ACTIONS)
            word = torch.gather(self.legal_mask, 1, (idx.clamp(0, NUM_ACTIONS - 1) >> 5).unsqueeze(1)).squeeze(1)
            legal = in_range & (((word >> (idx.clamp(0, NUM_ACTIONS - 1) & 31).to(torch.int32)) & 1) != 0)
            bad_rows = torch.nonzero(~legal & ~self._done & (idx != -1))
            if bad_rows.numel():
                bad = int(bad_rows[0])
                raise ValueError(f"the action {int(a[bad])} is illegal! (env {bad})")
        if bool(self._done.any()):  # next-step auto-reset: a fresh game id for every finished env
            self._episode += self._done.to(torch.int64)
            self._ids = self._game_ids()
            self.engine.reset(where=self._done, game_ids=self._ids)
            n = int(self._done.sum())
            self.my_turn = self.my_turn.clone()
            self.my_turn[self._done] = self._draw_seats(n).to(self.device)
            a = torch.where(self._done, torch.full_like(a, -1), a)
        if self.opponent in ("minimax", "policy"):
            # the learner's move alone (it is the seat to move wherever a >= 0), the opponents' replies, then the learner's rows
            _, reward, _, illegal, _ = self.engine.env_step(None, a, want_mask=False, want_obs=False, self_play=True)
            if self.opponent == "minimax":
                self._minimax_opponents_move()
            else:
                self._policy_opponents_move()
            skip = torch.full_like(a, -1)
            obs, _, term, _, mask = self.engine.env_step(self.my_turn, skip)
        else:
            obs, reward, term, illegal, mask = self.engine.env_step(self.my_turn, a, self_play=self.self_play)
        acted = self.my_turn
        if self.self_play:
            self.my_turn = self._seat_to_move()
        if self.raise_on_illegal and bool(illegal.any()):
            bad = int(torch.nonzero(illegal)[0])
            raise ValueError(f"the action {int(a[bad])} is illegal! (env {bad})")
        self._done = term != 0
        self.legal_mask = mask
        info = {"legal_mask": mask, "illegal": illegal != 0, "my_id": self.my_turn, "end_kind": term, "acted": acted}
        return obs, reward.to(torch.float32), self._done.clone(), torch.zeros_like(self._done), info

    def _minimax_opponents_move(self):
        """Where the game goes on and it is not the learner's turn: one depth-2 minimax move of the opponent."""
        skip = torch.full((self.num_envs,), -1, dtype=torch.int16, device=self.device)
        _, _, term, _, _ = self.engine.env_step(None, skip, want_mask=False, want_obs=False, self_play=True)
        need = (term == 0) & (self._seat_to_move() != self.my_turn)
        if bool(need.any()):
            best = self.engine.minimax2()[0]
            self.engine.env_step(None, torch.where(need, best, skip), want_mask=False, want_obs=False, self_play=True)

    def _policy_opponents_move(self):
        """_simulate_opponents (splendor_env.py:219-231) with policy-driven agents: until every env is back at its learner
        or over, the seats to move get their own observation and mask, the policy answers, the moves are applied."""
        from .sampling import masked_sample

        skip = torch.full((self.num_envs,), -1, dtype=torch.int16, device=self.device)
        for _ in range(self.number_of_players - 1):
            obs, _, term, _, mask = self.engine.env_step(None, skip, self_play=True)  # rows of the seats to move
            seats = self._seat_to_move()
            need = (term == 0) & (seats != self.my_turn)
            if not bool(need.any()):
                break
            out = self.opponent_policy(obs, mask, seats)
            if out.dtype in (torch.int16, torch.int32, torch.int64):
                actions = out.to(torch.int16)
            else:
                step = int(self._policy_calls)
                actions, _ = masked_sample(out.to(torch.float32).contiguous(), mask, self.opponent_seed, step, ids=self._ids,
                                           greedy=self.opponent_greedy, want_log_prob=False)
            self._policy_calls += 1
            self.engine.env_step(None, torch.where(need, actions, skip), want_mask=False, want_obs=False, self_play=True)

    def _seat_to_move(self) -> torch.Tensor:
        return ((self.engine.state[1, :, 3] >> 20) & 3).to(torch.int8)

    def get_legal_actions_mask(self, packed: bool = False, dtype=torch.float64) -> torch.Tensor:
        """The learners' current masks: bit-packed [N,110] int32, or dense [N,3510] like the reference's float64 mask."""
        return self.legal_mask if packed else unpack_mask(self.legal_mask, dtype)

    @property
    def state(self):
        return self.engine.state

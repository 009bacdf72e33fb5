"""BatchedSplendor: N Splendor games resident in HBM, driven through the C ABI (include/splendor_b200.h).

torch is used for device memory and streams only; every rule computation happens in the CUDA kernels of
csrc/kernels.cu.  Method names follow the reference's rule object (SplendorGameRule, splendor_model.py:142-576) where a
batched equivalent exists: legal_mask ~ getLegalActions, step ~ update/generateSuccessor, game_ends, cal_score.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .state import state_shape

NUM_ACTIONS = 3510
MASK_WORDS = 110
OBS_SIZE = 265
DEFAULT_SEED = 1234  # the reference's SEED (agents/our_agents/ppo/constants.py:12)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class BatchedSplendor:
    """N independent games of `num_agents` players.

    Decks: by default counter-based (Philox keyed by (seed, first_game + env index)); `inject` switches the batch to
    caller-supplied deck orders (the parity contract: same deck orders + same actions => same trajectory as the
    reference engine).
    """

    def __init__(self, num_envs: int, num_agents: int = 2, seed: int = DEFAULT_SEED, first_game: int = 0,
                 limit_rounds: bool = False, device: str | torch.device = "cuda"):
        if not torch.cuda.is_available():
            raise nat.SplendorNativeError("no CUDA device: the B200 engine has no CPU fallback")
        self.lib = nat.lib()
        self.N, self.P = int(num_envs), int(num_agents)
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seed, self.first_game = int(seed), int(first_game)
        self.flags = nat.F_LIMIT_ROUNDS if limit_rounds else 0
        self.state = torch.zeros(state_shape(self.N, self.P), dtype=torch.int32, device=self.device)
        self.decks = None  # explicit deck lists [N, 90] uint8 after inject()
        self.game_ids = None  # per-env game ids (int64 [N]) once envs are reset individually; None = first_game + index
        self.reset()

    # -- construction -------------------------------------------------------------------------------------------
    def reset(self, first_game: int | None = None, want_lists: bool = False, where: torch.Tensor | None = None,
              game_ids: torch.Tensor | None = None):
        """SplendorGameRule(P).initialGameState() for every env (counter-based decks), or only for envs with where != 0.

        game_ids (int64 [N]): per-env game ids; once given they are used by every later call on this batch."""
        if first_game is not None:
            self.first_game = int(first_game)
        if game_ids is not None:
            self.game_ids = game_ids.to(device=self.device, dtype=torch.int64).contiguous()
        elif where is None:
            self.game_ids = None
        w = None if where is None else where.to(device=self.device, dtype=torch.uint8).contiguous()
        if where is None:
            self.decks = None
        # a partial reset of an injected batch keeps the deck lists: the kernels pick the deck form per env (an env that
        # was reset draws from its counter-based deck, the others keep popping their explicit lists)
        lists = torch.empty((self.N, 90), dtype=torch.uint8, device=self.device) if want_lists else None
        nobles = torch.empty((self.N, 5), dtype=torch.uint8, device=self.device) if want_lists else None
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_reset(_ptr(self.state), self.seed, self.first_game, _ptr(self.game_ids), _ptr(w), self.N, self.P,
                                         _ptr(lists), _ptr(nobles), _stream(self.device)), "spl_reset")
        return (lists, nobles) if want_lists else None

    def inject(self, deck_lists, nobles, game_ids: torch.Tensor | None = None):
        """Initial states from explicit deck orders ([N,90] card ids, reference list order) and noble ids ([N,>=P+1]).
        game_ids (int64 [N]): the ids that key the in-kernel random choices (opponents, playouts) of these games."""
        if game_ids is not None:
            self.game_ids = game_ids.to(device=self.device, dtype=torch.int64).contiguous()
        d = torch.as_tensor(np.ascontiguousarray(deck_lists, dtype=np.uint8)).to(self.device)
        n = np.full((self.N, 5), 0xFF, np.uint8)
        nb = np.asarray(nobles, dtype=np.uint8)
        n[:, :nb.shape[1]] = nb
        n = torch.as_tensor(n).to(self.device)
        assert d.shape == (self.N, 90)
        self.decks = d
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_inject(_ptr(self.state), _ptr(d), _ptr(n), self.N, self.P, _stream(self.device)), "spl_inject")

    # -- the rule path ------------------------------------------------------------------------------------------
    def legal_mask(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """Bit-packed legal-action mask [N, 110] int32 (bit i%32 of word i/32 = ALL_ACTIONS[i] legal for the seat to move)."""
        mask = out if out is not None else torch.empty((self.N, MASK_WORDS), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_legal_mask(_ptr(self.state), _ptr(mask), self.N, self.P, _stream(self.device)), "spl_legal_mask")
        return mask

    def step(self, actions: torch.Tensor):
        """GameRule.update for one ALL_ACTIONS index per env (int16; <0 = skip).  Returns (reward i8, done u8, illegal u8)."""
        a = actions.to(device=self.device, dtype=torch.int16).contiguous()
        reward = torch.empty(self.N, dtype=torch.int8, device=self.device)
        done = torch.empty(self.N, dtype=torch.uint8, device=self.device)
        illegal = torch.empty(self.N, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_step(_ptr(self.state), _ptr(self.decks), self.seed, self.first_game, _ptr(self.game_ids), _ptr(a), _ptr(reward),
                                        _ptr(done), _ptr(illegal), self.N, self.P, self.flags, _stream(self.device)), "spl_step")
        return reward, done, illegal

    def count_legal(self) -> torch.Tensor:
        """len(getLegalActions(state, seat to move)) per env (int32 [N])."""
        counts = torch.empty(self.N, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_count_legal(_ptr(self.state), _ptr(counts), self.N, self.P, _stream(self.device)), "spl_count_legal")
        return counts

    def expand(self):
        """Every successor of every env (one ply): returns (children: BatchedSplendor of M envs, action int16 [M], reward int8 [M],
        parent int32 [M], offsets int64 [N+1]); children of env e are envs offsets[e]..offsets[e+1]-1 in ascending action order."""
        counts = self.count_legal()
        offsets = torch.zeros(self.N + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=offsets[1:])
        M = int(offsets[-1].item())
        child = BatchedSplendor.__new__(BatchedSplendor)
        child.lib, child.N, child.P, child.device = self.lib, M, self.P, self.device
        child.seed, child.first_game, child.flags = self.seed, self.first_game, self.flags
        child.state = torch.empty(state_shape(M, self.P), dtype=torch.int32, device=self.device)
        action = torch.empty(M, dtype=torch.int16, device=self.device)
        reward = torch.empty(M, dtype=torch.int8, device=self.device)
        parent = torch.empty(M, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_expand(_ptr(self.state), _ptr(self.decks), self.seed, self.first_game, _ptr(self.game_ids), _ptr(offsets),
                                          _ptr(child.state), _ptr(action), _ptr(reward), _ptr(parent), self.N, M, self.P,
                                          _stream(self.device)), "spl_expand")
        idx = parent.to(torch.int64)
        child.decks = None if self.decks is None else self.decks[idx].contiguous()
        base_ids = self.game_ids if self.game_ids is not None else \
            self.first_game + torch.arange(self.N, device=self.device, dtype=torch.int64)
        child.game_ids = base_ids[idx].contiguous()
        return child, action, reward, parent, offsets

    def minimax_eval(self, seat: torch.Tensor | None = None) -> torch.Tensor:
        """MiniMaxAgent._evaluation_function (minmax.py:90-136) per env as float64 [N]; seat int8 per env (None = seat to move)."""
        out = torch.empty(self.N, dtype=torch.float64, device=self.device)
        s = None if seat is None else seat.to(device=self.device, dtype=torch.int8).contiguous()
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_minimax_eval(_ptr(self.state), _ptr(s), _ptr(out), self.N, self.P, _stream(self.device)), "spl_minimax_eval")
        return out

    def minimax2(self):
        """MiniMaxAgent.SelectAction (minmax.py:32-88, depth 2) for the seat to move of every 2-player env.  Returns
        (best_action int16 [N], best_value float64 [N], child_action int16 [M], child_value float64 [M], offsets int64 [N+1]):
        the children of env e are entries offsets[e] .. offsets[e+1]-1, one per legal action in ascending index order."""
        counts = self.count_legal()
        offsets = torch.zeros(self.N + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=offsets[1:])
        M = int(offsets[-1].item())
        child_value = torch.empty(M, dtype=torch.float64, device=self.device)
        child_action = torch.empty(M, dtype=torch.int16, device=self.device)
        best_action = torch.empty(self.N, dtype=torch.int16, device=self.device)
        best_value = torch.empty(self.N, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_minimax2(_ptr(self.state), _ptr(self.decks), self.seed, self.first_game, _ptr(self.game_ids), _ptr(offsets),
                                            _ptr(child_value), _ptr(child_action), _ptr(best_action), _ptr(best_value), self.N, M, self.P,
                                            _stream(self.device)), "spl_minimax2")
        return best_action, best_value, child_action, child_value, offsets

    def undo(self, seat: torch.Tensor, actions: torch.Tensor, card: torch.Tensor, noble: torch.Tensor, returned: torch.Tensor):
        """generatePredecessor for one action per env (see spl_undo): seat int8, action int16 (<0 = skip), card / noble
        uint8 ids (255 = none), returned uint8 [N,6] by colour id."""
        t = lambda x, dt: x.to(device=self.device, dtype=dt).contiguous()
        s, a, c, n, r = t(seat, torch.int8), t(actions, torch.int16), t(card, torch.uint8), t(noble, torch.uint8), t(returned, torch.uint8)
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_undo(_ptr(self.state), _ptr(s), _ptr(a), _ptr(c), _ptr(n), _ptr(r), self.N, self.P,
                                        _stream(self.device)), "spl_undo")

    def observe(self, seat: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """extract_metrics_with_cards(...).astype(float32) -> [N, 265]; seat: int8 per env (None = seat to move)."""
        obs = out if out is not None else torch.empty((self.N, OBS_SIZE), dtype=torch.float32, device=self.device)
        s = None if seat is None else seat.to(device=self.device, dtype=torch.int8).contiguous()
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_observe(_ptr(self.state), _ptr(s), _ptr(obs), self.N, self.P, _stream(self.device)), "spl_observe")
        return obs

    def final_scores(self):
        """(2*calScore [N,P] int8, win2/tie1/loss0 [N,P] uint8, gameEnds kind [N] uint8)."""
        score2 = torch.empty((self.N, self.P), dtype=torch.int8, device=self.device)
        outcome = torch.empty((self.N, self.P), dtype=torch.uint8, device=self.device)
        kind = torch.empty(self.N, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_final_scores(_ptr(self.state), _ptr(score2), _ptr(outcome), _ptr(kind), self.N, self.P,
                                                self.flags, _stream(self.device)), "spl_final_scores")
        return score2, outcome, kind

    def playout(self, max_steps: int = 1000, trace_len: int = 0, counters: torch.Tensor | None = None):
        """Random playouts to the end from the CURRENT states (RandomAgent on every seat).  Returns a dict of tensors."""
        out = dict(
            score2=torch.empty((self.N, self.P), dtype=torch.int8, device=self.device),
            steps=torch.empty(self.N, dtype=torch.int16, device=self.device),
            end_kind=torch.empty(self.N, dtype=torch.uint8, device=self.device),
            counters=counters if counters is not None else torch.zeros(16, dtype=torch.int64, device=self.device),
            trace=torch.empty((trace_len, self.N), dtype=torch.int16, device=self.device) if trace_len else None,
        )
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_playout_from(_ptr(self.state), _ptr(self.decks), self.seed, self.first_game, _ptr(self.game_ids), self.N, self.P,
                                                self.flags, max_steps, _ptr(out["score2"]), _ptr(out["steps"]),
                                                _ptr(out["end_kind"]), _ptr(out["counters"]), _ptr(out["trace"]), trace_len,
                                                _stream(self.device)), "spl_playout_from")
        return out

    def env_step(self, learner: torch.Tensor | None, actions: torch.Tensor, want_mask: bool = True, want_obs: bool = True,
                 self_play: bool = False):
        """SplendorEnv.step for every env: learner move, random opponents, next mask + observation (see spl_env_step).
        self_play=True: the action is for the seat to move, no opponents are simulated, and the mask / observation are the
        next mover's (learner is ignored and may be None)."""
        l = None if learner is None else learner.to(device=self.device, dtype=torch.int8).contiguous()
        a = actions.to(device=self.device, dtype=torch.int16).contiguous()
        reward = torch.empty(self.N, dtype=torch.int8, device=self.device)
        term = torch.empty(self.N, dtype=torch.uint8, device=self.device)
        illegal = torch.empty(self.N, dtype=torch.uint8, device=self.device)
        mask = torch.empty((self.N, MASK_WORDS), dtype=torch.int32, device=self.device) if want_mask else None
        obs = torch.empty((self.N, OBS_SIZE), dtype=torch.float32, device=self.device) if want_obs else None
        with torch.cuda.device(self.device):
            nat.check(self.lib.spl_env_step(_ptr(self.state), _ptr(self.decks), self.seed, self.first_game, _ptr(self.game_ids), _ptr(l), _ptr(a),
                                            _ptr(reward), _ptr(term), _ptr(illegal), _ptr(mask), _ptr(obs), self.N, self.P,
                                            self.flags | (nat.F_SELF_PLAY if self_play else 0), _stream(self.device)), "spl_env_step")
        return obs, reward, term, illegal, mask

    def state_numpy(self) -> np.ndarray:
        return self.state.cpu().numpy().view(np.uint32)


def playout_games(num_games: int, num_agents: int = 2, seed: int = DEFAULT_SEED, first_game: int = 0,
                  limit_rounds: bool = False, max_steps: int = 1000, trace_len: int = 0, want_final_state: bool = False,
                  counters: torch.Tensor | None = None, device: str | torch.device = "cuda", outputs: bool = True):
    """Fused reset + random playout + scoring for games first_game .. first_game+num_games-1 (spl_playout): nothing but the
    results ever touches HBM."""
    if not torch.cuda.is_available():
        raise nat.SplendorNativeError("no CUDA device: the B200 engine has no CPU fallback")
    lib = nat.lib()
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    G, P = int(num_games), int(num_agents)
    out = dict(
        score2=torch.empty((G, P), dtype=torch.int8, device=dev) if outputs else None,
        steps=torch.empty(G, dtype=torch.int16, device=dev) if outputs else None,
        end_kind=torch.empty(G, dtype=torch.uint8, device=dev) if outputs else None,
        counters=counters if counters is not None else torch.zeros(16, dtype=torch.int64, device=dev),
        trace=torch.empty((trace_len, G), dtype=torch.int16, device=dev) if trace_len else None,
        final_state=torch.empty(state_shape(G, P), dtype=torch.int32, device=dev) if want_final_state else None,
    )
    with torch.cuda.device(dev):
        nat.check(lib.spl_playout(int(seed), int(first_game), G, P, nat.F_LIMIT_ROUNDS if limit_rounds else 0, max_steps,
                                  _ptr(out["score2"]), _ptr(out["steps"]), _ptr(out["end_kind"]), _ptr(out["counters"]),
                                  _ptr(out["trace"]), trace_len, _ptr(out["final_state"]), _stream(dev)), "spl_playout")
    return out


def counters_dict(counters) -> dict:
    c = counters.cpu().numpy() if hasattr(counters, "cpu") else np.asarray(counters)
    return {k: int(v) for k, v in zip(nat.CTR_NAMES, c)}

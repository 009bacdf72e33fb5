/* splendor_b200.h -- C ABI of the B200-native batched Splendor engine (libsplendor_b200.so).
 *
 * This is the drop-in boundary for the reference's game-rule path.  The reference has no FFI:
 * the "interface" is a set of Python classes (all paths relative to /root/reference/src/splendor/).  Each entry
 * point below names the reference call it replaces; INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer into caller-owned memory unless the function
 *     name ends in _host; `stream` is a cudaStream_t passed as void* (0 = default stream); no hidden host sync
 *     except in the *_host functions;
 *   - return value: SPL_OK (0) or a negative SPL_E_* code; nothing throws across the ABI;
 *   - devices and threads: every call works on the CURRENT CUDA device (cudaSetDevice before the call); the library's own
 *     state - the staging memory and events of the *_host calls, the cached SM count, the opt-in to large shared memory -
 *     is kept PER DEVICE, so one process may drive several GPUs.  Calls may come from several host threads (the *_host
 *     calls serialise on an internal mutex while they enqueue); spl_last_cuda_error is per thread;
 *   - N = number of environments (games) in the batch, P = players per game (2, 3 or 4);
 *   - action indices are positions in the reference's ALL_ACTIONS list (gym/envs/actions.py:222-237), 0..3509;
 *   - card ids are tier-major (include/spl_tables.h), noble ids index NOBLES, colour ids black0 red1 green2 blue3
 *     white4 yellow5.
 *
 * State layout in HBM ("packed SoA of 128-bit chunks")
 *   One env = W = 8 + 4*P little-endian u32 words = C = W/4 chunks of 16 bytes.  `state` is uint4[C][N]: chunk c of
 *   env e lives at ((c * N) + e) * 16 bytes, so a warp of consecutive envs reads/writes 512 contiguous bytes per chunk.
 *     word 0   bank            six 5-bit fields, field f at bits 5f..5f+4, in the reference's COLOURS order black, red,
 *                              yellow, green, blue, white (splendor_utils.py:124-131); values 0..7
 *     word 1-3 dealt[tier]     four bytes, byte col = card id of face-up slot (tier, col), 0xFF = empty slot
 *     word 4-6 deck            counter-based decks: bitmask of cards still in each deck
 *                                word4 = tier-1 bits 0..31, word5 = tier-2 (30 bits),
 *                                word6 = tier-3 (bits 0..19) | tier-1 bits 32..39 (at bits 20..27)
 *                              explicit decks (after spl_inject): word4 bytes 0..2 = cards left per tier, 5-6 zero
 *     word 7   meta            bits 0..19 board nobles, LIST ORDER, one nibble each (0xF = none) | bits 20..21 seat to
 *                              move | bits 22..23 P-2 | bit 24 explicit-deck flag | bit 26 sticky illegal-action flag
 *     player p at words 8+4p .. 11+4p
 *       +0 gems    six 5-bit fields (same order)  +1 bought cards per colour, same fields (the yellow field stays 0)
 *       +2 reserved card ids, bytes 0..2 in list order (0xFF = none), byte 3 = score
 *       +3 bits 0..15 turns taken | bits 16..25 nobles owned (bit = noble id) | bit 26 passed
 */
#ifndef SPLENDOR_B200_H
#define SPLENDOR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPL_NUM_ACTIONS_ 3510
#define SPL_MASK_WORDS_ 110
#define SPL_OBS_SIZE_ 265
#define SPL_DECK_BYTES 90 /* explicit deck lists per env: tier-1 (40) | tier-2 (30) | tier-3 (20) */

/* error codes */
#define SPL_OK 0
#define SPL_E_BADARG (-1)   /* null pointer / P not in 2..4 / negative size */
#define SPL_E_CUDA (-2)     /* a CUDA runtime call failed; spl_last_cuda_error() has the cudaError_t */
#define SPL_E_NODEVICE (-3) /* no CUDA device: there is deliberately no CPU fallback */
#define SPL_E_RANGE (-4)    /* max_steps / trace_len out of range */

/* rule flags */
#define SPL_F_LIMIT_ROUNDS 1u /* LimitRoundsGameRule: also end when every agent has made 100 turns (splendor/utils.py:14-25) */
#define SPL_F_SELF_PLAY 2u    /* spl_env_step: no in-kernel opponents; the caller's policy plays every seat (ppo.py --opponent itself) */

/* end kinds written by playouts / spl_game_ends */
#define SPL_END_NONE 0
#define SPL_END_SCORE 1     /* some agent has >=15 and the round is complete (splendor_model.py:303) */
#define SPL_END_DEADLOCK 2  /* every agent passed (splendor_model.py:305) */
#define SPL_END_ROUNDCAP 3  /* 100-turn cap */
#define SPL_END_TRUNCATED 4 /* max_steps reached first */

/* counters[16] (int64, ACCUMULATED into, so one buffer can span launches and be all-reduced with NCCL) */
#define SPL_CTR_GAMES 0
#define SPL_CTR_STEPS 1
#define SPL_CTR_WINS 2       /* [2..5]  unique best calScore, per seat (general_game_runner.py:416-430) */
#define SPL_CTR_TIES 6       /* [6..9]  shared best calScore, per seat */
#define SPL_CTR_END_SCORE 10
#define SPL_CTR_END_DEADLOCK 11
#define SPL_CTR_END_ROUNDCAP 12
#define SPL_CTR_END_TRUNCATED 13
#define SPL_CTR_SCORE2_SUM 14 /* sum over seats of 2*calScore */
#define SPL_CTR_NOBLES 15     /* nobles claimed */

int spl_abi_version(void);
int spl_state_words(int P);             /* 8 + 4*P, or SPL_E_BADARG */
int spl_last_cuda_error(void);          /* cudaError_t of the last failing runtime call on this thread */
const char *spl_error_string(int code);

/* SplendorGameRule(P).initialGameState() for N synthetic games (splendor_model.py:69-93), decks and nobles drawn from
 * Philox4x32-10 keyed (seed, game id) instead of Python's `random` (DESIGN.md "Synthetic games").  The game id of env e
 * is game_ids[e] if game_ids is non-null, else first_game + e (the same convention holds for every call below that
 * takes both).  where (nullable, N bytes): only envs with where[e] != 0 are reset - the vector env's auto-reset.
 * If deck_lists_out (N*90 bytes) / nobles_out (N*5 bytes) are non-null they receive the full deck orders in the
 * REFERENCE's list order (deal() pops from the end) and the noble ids, so the same games can be replayed through the
 * reference engine. */
int spl_reset(void *state, uint64_t seed, uint64_t first_game, const uint64_t *game_ids, const uint8_t *where,
              int64_t N, int P, uint8_t *deck_lists_out, uint8_t *nobles_out, void *stream);

/* Initial states from caller-supplied deck orders (reference list order, N*90 bytes) and noble ids (N*5 bytes,
 * first P+1 used): what BoardState.__init__ leaves after its 12 deals.  `deck_lists` must stay alive and be passed as
 * `decks` to every later call on these states. */
int spl_inject(void *state, const uint8_t *deck_lists, const uint8_t *nobles, int64_t N, int P, void *stream);

/* SplendorGameRule.getLegalActions(state, seat to move) as a 3510-bit mask: mask[e*110 + i/32] bit i%32
 * (splendor_model.py:404-576 + gym/envs/utils.py:148-164 create_legal_actions_mask).  One thread builds a row in registers,
 * the block stores the rows coalesced. */
int spl_legal_mask(const void *state, uint32_t *mask, int64_t N, int P, void *stream);

/* GameRule.update(action) = generateSuccessor + seat advance (splendor_model.py:222-295, template.py:47-53) for one
 * action index per env.  reward[e] = acting agent's score delta (splendor_env.py:142-161); done[e] = end kind after the
 * move (gameEnds, SPL_END_*); illegal[e] = 1 and the env is left untouched if the index is out of range or not legal
 * (the ValueError of splendor_env.py:139-153).  action[e] < 0 skips the env.  Output pointers may be null.
 * decks: null for counter-based decks, else the array given to spl_inject.  seed/first_game: as given to spl_reset. */
int spl_step(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
             const int16_t *action, int8_t *reward, uint8_t *done, uint8_t *illegal, int64_t N, int P, uint32_t flags,
             void *stream);

/* One-ply expansion for search-style agents (what minimax / the GA agent do one action at a time with
 * generateSuccessor + generatePredecessor, agents/our_agents/minmax.py:43-88, genetic_algorithm_agent.py:93-96):
 * spl_count_legal writes len(getLegalActions(state, seat to move)) per env; the caller turns the counts into exclusive
 * prefix sums offsets[0..N] (offsets[N] = M); spl_expand then writes every successor: child j of env e (its j-th legal
 * action in ascending ALL_ACTIONS order) becomes env offsets[e]+j of an M-env batch in the packed layout, with the action
 * index, the acting agent's reward and the parent env in the optional side arrays.  Children of explicit-deck envs read
 * the parent's deck list: pass decks gathered by out_parent to later calls on the children. */
int spl_count_legal(const void *state, int32_t *counts, int64_t N, int P, void *stream);
int spl_expand(const void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
               const int64_t *offsets, void *out_state, int16_t *out_action, int8_t *out_reward, int32_t *out_parent, int64_t N,
               int64_t M, int P, void *stream);

/* MiniMaxAgent (agents/our_agents/minmax.py), the reference's depth-2 search agent and the default PPO test opponent
 * (arguments_parsing.py:107), for 2-player envs.
 * spl_minimax_eval: _evaluation_function(state) (:90-136) from seat[e]'s point of view (null = the seat to move), float64,
 *   bit-exact with the reference (same operation order).  Any P.
 * spl_minimax2: SelectAction (:32-88) for the seat to move of every env.  offsets[0..N] are the exclusive prefix sums of
 *   spl_count_legal (offsets[N] = M).  child_value[offsets[e]+j] = the depth-2 value of env e's j-th legal action (ascending
 *   ALL_ACTIONS order) = min over the opponent's replies of the evaluation; child_action = its index.  best_action /
 *   best_value (optional): the maximum; exact ties resolve to the lowest type rank in the reference's sort (buy_available,
 *   buy_reserve, collect_diff, collect_same, pass, reserve), then the lowest index - the reference's own order inside a
 *   type is a random.shuffle.  Like the reference, no terminal test between the plies; replacement cards come from the
 *   env's deck (the search peeks at it, as generateSuccessor does).  P must be 2. */
int spl_minimax_eval(const void *state, const int8_t *seat, double *value, int64_t N, int P, void *stream);
int spl_minimax2(const void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
                 const int64_t *offsets, double *child_value /*[M]*/, int16_t *child_action /*[M]*/, int16_t *best_action /*[N]*/,
                 double *best_value /*[N]*/, int64_t N, int64_t M, int P, void *stream);

/* SplendorGameRule.generatePredecessor(state, action, agent_id) (splendor_model.py:156-220): undo one action per env for
 * seat[e].  Besides the ALL_ACTIONS index the undo needs what the reference reads from the action dict: card[e] = id of
 * the card that was bought / reserved (0xFF otherwise), noble[e] = id of the noble that was claimed (0xFF if none),
 * returned[e*6 .. +6] = returned_gems by colour id (for buys: the payment).  action[e] < 0 skips the env.  The seat to
 * move becomes seat[e] again (the inverse of spl_step); the reference's list-order quirks are kept. */
int spl_undo(void *state, const int8_t *seat, const int16_t *action, const uint8_t *card, const uint8_t *noble,
             const uint8_t *returned, int64_t N, int P, void *stream);

/* gameEnds() per env (SPL_END_*), and calScore per seat as 2*score (+1 = the half-point tie-break), plus the
 * win(2)/tie(1)/loss(0) class of general_game_runner.py:416-430.  Output pointers may be null. */
int spl_final_scores(const void *state, int8_t *score2 /*[N][P]*/, uint8_t *outcome /*[N][P]*/, uint8_t *end_kind /*[N]*/,
                     int64_t N, int P, uint32_t flags, void *stream);

/* features.extract_metrics_with_cards(state, seat).astype(float32) (features.py:470-480, splendor_env.py:206):
 * obs[e*265 .. +265].  seat: per-env seat (int8, null = the seat to move). */
/* (obs: a 16-byte aligned buffer is written with 128-bit stores; a buffer that is only float-aligned works, with scalar stores.) */
int spl_observe(const void *state, const int8_t *seat, float *obs, int64_t N, int P, void *stream);

/* Whole random playouts that never leave the GPU: Game.Run with RandomAgents (game.py:104-213,
 * agents/generic/random.py:21-27) for games first_game .. first_game+num_games-1, fused reset + legal-count + uniform
 * choice (word step & 3 of the Philox block (seed, game, step >> 2, stream 0) scaled to the count - one block serves four
 * consecutive moves -, k-th legal action in ascending ALL_ACTIONS order)
 * + step + terminal + calScore.  Outputs (all optional): score2 [G][P], steps [G], end_kind [G], counters[16]
 * (accumulated), trace [trace_len][G] = action index taken at step t (-1 after the end), final_state (packed layout,
 * N = num_games).  max_steps in 1..4096.
 * Kernel: playout_pool_kernel - every block owns a pool of games in shared memory, warps play 8 moves of 32 live games
 * out of registers, then the block compacts the survivors and tops the pool up with new games. */
int spl_playout(uint64_t seed, uint64_t first_game, int64_t num_games, int P, uint32_t flags, int max_steps,
                int8_t *score2, int16_t *steps, uint8_t *end_kind, int64_t *counters,
                int16_t *trace, int trace_len, void *final_state, void *stream);

/* Same, continuing from existing states (e.g. after spl_inject, with `decks`); states are updated in place.
 * `steps` and the SPL_CTR_STEPS counter count the moves made by THIS call (a state handed in mid-game does not bring its
 * earlier turns along); a state that is already terminal is scored with 0 steps. */
int spl_playout_from(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
                     int64_t N, int P, uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind,
                     int64_t *counters, int16_t *trace, int trace_len, void *stream);

/* The PPO rollout call: SplendorEnv.step(action) (splendor_env.py:127-172) for N envs at once -- the learner's move,
 * then in-kernel RandomAgent opponents until it is the learner's seat again or the game ends
 * (_simulate_opponents :219-231), then the learner's next legal mask and observation.  learner[e] = the learner's
 * seat.  action[e] < 0 = "reset step": only simulate opponents up to the learner's first turn (SplendorEnv.reset).
 * reward[e] int8, terminated[e] u8 (end kind), illegal[e] u8, mask [N][110] u32, obs [N][265] f32 (mask/obs optional).
 * With SPL_F_SELF_PLAY the action is applied for the seat to move (learner may be null), no opponent is simulated, and the
 * mask / observation returned are those of the NEXT seat to move: the caller's policy plays all seats.
 * Three kernels in stream order on `stream` (move, observation, mask); an in-kernel opponent's choice is word 0 of the Philox
 * block (seed, game, total turns so far, stream 0) scaled to its legal count. */
int spl_env_step(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
                 const int8_t *learner, const int16_t *action, int8_t *reward, uint8_t *terminated, uint8_t *illegal,
                 uint32_t *mask, float *obs, int64_t N, int P, uint32_t flags, void *stream);

/* The rollout's action choice under the legal mask (agents/our_agents/ppo/network.py:113-115, training.py:110-116):
 *   probs = softmax(where(mask == 0, HUGE_NEG, logits)); action ~ Categorical(probs); log_prob(action)
 * for N rows of 3510 float32 logits (row stride in floats >= 3510) and the BIT-PACKED masks of spl_legal_mask /
 * spl_env_step.  Only the legal logits are read.  The sample is an inverse CDF over weights floor(exp(l - max) * 2^24) with one Philox word per row keyed
 * (seed, row id = ids[e] or first_id + e, step, stream 6); greedy != 0 returns the arg-max (lowest index on ties) instead.
 * log_prob (optional) = log softmax of the chosen action over the legal ones, float32 (|error| < 1e-5 vs float64). */
int spl_masked_sample(const float *logits, int64_t row_stride, const uint32_t *mask, uint64_t seed, uint64_t first_id,
                      const uint64_t *ids, uint32_t step, int greedy, int16_t *action, float *log_prob, int64_t N, void *stream);

/* Host-buffer forms (the end-to-end path a reference-side caller uses): pageable or pinned HOST pointers in and out,
 * device staging buffers owned by the library (per device and slot, grow-only), all work on `stream`: two H2D copies,
 * ONE kernel that builds every game from its deck order and plays it to the end, four D2H copies - then a stream
 * synchronise.  deck_lists/nobles as in spl_inject; outputs as in spl_playout. */
int spl_playout_host(const uint8_t *deck_lists, const uint8_t *nobles, uint64_t seed, uint64_t first_game, int64_t N,
                     int P, uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind,
                     int64_t *counters, void *stream);

/* The same without the final synchronise, for callers that pipeline batches (a rollout loop that prepares batch k+1
 * while batch k plays): returns as soon as the copies and kernels are queued.  The output arrays and `counters`
 * (16 x int64, OVERWRITTEN, not accumulated) are valid once `stream` has reached the end of the queued work
 * (cudaStreamSynchronize / an event recorded after the call); all host buffers must stay alive until then and should be
 * pinned, or the copies degrade to synchronous ones.  `slot` (0 or 1) picks one of two independent staging areas: calls
 * in flight at the same time must use different slots and different streams; a slot is reused only after its previous
 * call has completed (the library orders that on the device if the caller does not). */
int spl_playout_host_async(const uint8_t *deck_lists, const uint8_t *nobles, uint64_t seed, uint64_t first_game, int64_t N,
                           int P, uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind,
                           int64_t *counters, int slot, void *stream);

/* Creates the CUDA context of the current device and caches its properties (optional; every call does it on demand). */
int spl_warmup(void);

#ifdef __cplusplus
}
#endif
#endif

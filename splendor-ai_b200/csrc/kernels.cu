// kernels.cu -- CUDA kernels (sm_100a) and the C ABI of include/splendor_b200.h.
//
//   playout_pool_kernel a block owns a POOL of up to 640 games: between rounds the live games sit in shared memory as a
//                       dense list, in a round each warp plays 8 moves of 32 of them out of registers (terminal test, legal
//                       count, choice by the round's Philox words, k-th legal action, apply, lazy deal), survivors are compacted into the next
//                       list, finished games are scored, a keeper warp tops the list up with new games.  No HBM traffic
//                       inside the loop except the optional trace.
//   legal_mask_kernel   persistent, software-pipelined state loads; one thread builds one env's 3510-bit row in registers
//                       (32/64-bit group patterns) and flushes finished words into its warp's dense shared-memory tile;
//                       lane 0 sends the warp's 32 rows (14,080 contiguous bytes) to HBM with one bulk copy (cp.async.bulk)
//                       while the warp moves on.
//   step_kernel         thread-per-env, persistent: the ALL_ACTIONS index decoded by table, membership test, apply,
//                       reward / done / illegal.
//   observe_kernel      same scheme for the 265-float observation: rows staged as DENSE bytes, the warp's 32 rows widened
//                       to float32 as one flat span with 128-bit stores (no block barrier).
//   env_move_kernel     the PPO rollout call's first part: learner move + in-kernel random opponents, thread-per-env,
//                       persistent; spl_env_step then runs observe_kernel and legal_mask_kernel for the learner's seat in
//                       stream order.
//   reset / inject / final_scores: thread-per-env.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "engine.cuh"
#include "minimax.cuh"

namespace spl {

__device__ const Tables g_tables = make_tables();
__device__ const uint32_t g_log1p_bits[32] = SPL_LOG1P_BITS;
__device__ const DecodeTable g_decode = make_decode_table(make_tables());  // 56 KB, read through L2 by the step kernels
__device__ __forceinline__ uint4 decode_row(int a) { return g_decode.row[min(max(a, 0), SPL_NUM_ACTIONS - 1)]; }

constexpr int BLOCK = 128;

// A kernel stages only the part of Tables it reads (engine.cuh: the struct is ordered for this).  Every boundary is a multiple
// of 16 bytes: the copies are 16-byte loads from L2 and 16-byte shared-memory stores.
constexpr int TABLES_HEAD_BYTES = (int)offsetof(Tables, combo_c);    // card + noble: the observation kernel (with cardvec)
constexpr int TABLES_STEP_BYTES = (int)offsetof(Tables, pext_diff);  // + action decode / apply: step
constexpr int TABLES_MASK_BYTES = (int)offsetof(Tables, poly_pow);   // + legality patterns: legal mask
constexpr int TABLES_RULE_BYTES = (int)offsetof(Tables, cardvec);    // + counting / selecting: playouts, env move, expansion
static_assert(TABLES_HEAD_BYTES % 16 == 0 && TABLES_STEP_BYTES % 16 == 0 && TABLES_MASK_BYTES % 16 == 0 && TABLES_RULE_BYTES % 16 == 0 &&
                  sizeof(Tables) % 16 == 0 && alignof(Tables) >= 16,
              "the parts of Tables must be copyable in 16-byte pieces");
// bytes [begin, end) of the tables -> the block's copy, WITHOUT a barrier: a caller issues its first env's loads, calls
// this (every thread has a few independent 16-byte loads in flight), then synchronises - the table copy and the state
// loads overlap instead of forming a serial prologue.
__device__ __forceinline__ void stage_tables_nosync(Tables &sT, int begin = 0, int end = TABLES_RULE_BYTES)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(&g_tables);
    uint4 *dst = reinterpret_cast<uint4 *>(&sT);
#pragma unroll 4
    for (int i = begin / 16 + threadIdx.x; i < end / 16; i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void stage_tables(Tables &sT, int bytes = TABLES_RULE_BYTES)
{
    stage_tables_nosync(sT, 0, bytes);
    __syncthreads();
}
constexpr int ENV_TB = 256;        // threads per block of the persistent thread-per-env kernels (step, env move)
// Resident blocks per SM of those kernels, enforced through __launch_bounds__ so that the persistent grids are what actually
// resides: three blocks = 80 registers per thread.  (Four step blocks = 64 registers spill 40-60 bytes with the decoded row
// in flight: 0.0143 ms against 0.0132 per 262,144 envs; unconstrained, ptxas took 126 registers and two blocks: 0.0153.)
constexpr int STEP_BLOCKS_PER_SM = 3, ENV_MOVE_BLOCKS_PER_SM = 3;

// ---------------------------------------------------------------------------------------------- reset / inject
template <int P>
__global__ void __launch_bounds__(BLOCK) reset_kernel(uint4 *state, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
                                                      const uint8_t *where, int64_t N, uint8_t *deck_lists_out, uint8_t *nobles_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || (where && !where[i])) return;
    Env<P> e;
    const uint64_t game = game_ids ? game_ids[i] : first_game + (uint64_t)i;
    reset_env(e, seed, game);
    store_env(e, state, N, i);
    if (nobles_out)
        for (int k = 0; k < 5; k++) {
            const uint32_t id = (e.meta >> (4 * k)) & 15u;
            nobles_out[i * 5 + k] = id == 15u ? 0xFF : (uint8_t)id;
        }
    if (deck_lists_out) {  // finish drawing every deck so the whole order can be handed to the reference engine
        uint8_t *out = deck_lists_out + i * SPL_DECK_BYTES;
        const DeckSrc src{nullptr, seed, game};
        for (int t = 0; t < 3; t++) {
            const int size = t == 0 ? 40 : (t == 1 ? 30 : 20), base = t == 0 ? 0 : (t == 1 ? 40 : 70);
            for (int c = 0; c < 4; c++) out[base + size - 1 - c] = (uint8_t)byte_of(e.d[t], c);
            for (int d = 4; d < size; d++) out[base + size - 1 - d] = (uint8_t)deal(e, t, src);
        }
    }
}

// BoardState.__init__ from caller-supplied deck orders and nobles: what the 12 initial deals leave (explicit-list deck form)
template <int P>
__device__ __forceinline__ void inject_env(Env<P> &e, const uint8_t *lists, const uint8_t *nobles)
{
    const uint32_t n = P == 2 ? 4 : (P == 3 ? 5 : 7);
    e.bank = n * (ONES6 & ~(1u << YSH)) | (5u << YSH);
    e.k[0] = 40u | 30u << 8 | 20u << 16; e.k[1] = 0; e.k[2] = 0;
    uint32_t list = 0xFFFFFu;
    for (int k = 0; k < P + 1; k++) list = (list & ~(15u << (4 * k))) | ((uint32_t)(nobles[k] & 15u) << (4 * k));
    e.meta = list | ((uint32_t)(P - 2) << 22) | META_EXPLICIT;
    const DeckSrc src{lists, 0, 0};
#pragma unroll
    for (int t = 0; t < 3; t++) {
        uint32_t w = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) w |= deal(e, t, src) << (8 * c);
        e.d[t] = w;
    }
#pragma unroll
    for (int p = 0; p < P; p++) { e.pg[p] = 0; e.pc[p] = 0; e.pr[p] = 0x00FFFFFFu; e.pi[p] = 0; }
}

template <int P>
__global__ void __launch_bounds__(BLOCK) inject_kernel(uint4 *state, const uint8_t *deck_lists, const uint8_t *nobles, int64_t N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Env<P> e;
    inject_env(e, deck_lists + i * SPL_DECK_BYTES, nobles + i * 5);
    store_env(e, state, N, i);
}

// ---------------------------------------------------------------------------------------------- playout
struct PlayoutArgs {
    uint4 *state;           // FRESH: optional final-state output; else in/out states
    const uint8_t *decks;   // explicit deck lists (only with states / POOL_LISTS) or nullptr
    const uint8_t *nobles;  // POOL_LISTS: noble ids, 5 bytes per game
    uint64_t seed, first_game;
    const uint64_t *game_ids;  // per-game ids (only with states) or nullptr = first_game + index
    int64_t n_games;
    uint32_t flags;
    int max_steps;
    int8_t *score2;
    int16_t *steps;
    uint8_t *end_kind;
    unsigned long long *counters;
    int16_t *trace;
    int trace_len;
    uint32_t rk[20];        // Philox key schedule of `seed` (filled by launch_playout)
};

template <int P>
__device__ __forceinline__ uint32_t turns_total(const Env<P> &e)
{
    uint32_t t = 0;
#pragma unroll
    for (int p = 0; p < P; p++) t += e.pi[p] & 0xFFFFu;
    return t;
}

// ---------------------------------------------------------------------------------------------- playout, block pools
// A block owns a POOL of up to CAP = 32 * NW games.  Between rounds the live games sit in shared memory as a dense list
// (packed state words, 64 B per 2-player game, chunk-major like the HBM layout); in a ROUND warp w loads the w-th group of
// 32 live games into registers and plays SPL_POOL_STEPS moves of each out of registers, exactly as a one-game-per-thread
// kernel would; the survivors are appended to the next round's list (one warp-aggregated shared-memory atomic), the games
// that ended are scored, and whenever 32 slots are free a spare warp starts 32 new games with all lanes active.
// So the lanes of a stepping warp are (nearly) always all alive: the one-game-per-thread kernel of round 1 kept 24 of 32
// (a warp waited for its longest game: profiles/r1q_playout.md), which was its largest single loss.  The price is one
// block barrier and ~60 instructions of list traffic per round, i.e. per SPL_POOL_STEPS moves.
// A second form that also SORTED the games by action family between counting and selecting (so that a warp selects and
// applies one family without divergence) was built first and measured slower: handing the counts from one phase to the
// other plus two more barriers per move cost more than the ~100 divergent instructions it saved (profiles/r2_playout.md).
// Decks are drawn lazily from the counter-based form (a deal = one Philox call + a bit select), so a game needs nothing on
// chip but its state words.  Results do not depend on where a game sits: every draw is keyed by (seed, game id, turn).
#ifndef SPL_POOL_WARPS
#define SPL_POOL_WARPS 20
#endif
#ifndef SPL_POOL_BLOCKS_PER_SM
#define SPL_POOL_BLOCKS_PER_SM 1
#endif
#ifndef SPL_POOL_STEPS
#define SPL_POOL_STEPS 8
#endif
#ifndef SPL_POOL_SPREAD
#define SPL_POOL_SPREAD 8  // at most this many live games in a block: one per warp (<= SPL_POOL_WARPS - 2)
#endif

template <int P, int NW>
struct PoolSmem {
    static constexpr int CAP = 32 * NW, C = 2 + P;
    uint4 st[2][C][CAP];   // live list of this round / of the next round
    uint32_t gid[2][CAP];  // game index inside the launch
    uint32_t t0[2][CAP];   // total turns when the game entered the pool (0 for fresh games)
    Tables T;
    uint4 rnd[CAP][3];     // per thread: the Philox blocks its game's action draws of this round come from (four moves per block)
    unsigned int ctr[16];  // block-level counters (32 bits are enough for one block of one launch)
    int n_alive[3];        // list lengths, rotating: read in round r, appended to for round r+1, zeroed for round r+2
};

enum PoolMode { POOL_FRESH = 0, POOL_STATES = 1, POOL_LISTS = 2 };  // new games: Philox reset | states from HBM | deck orders + nobles

template <int P, int NW>
__device__ __forceinline__ void pool_load(Env<P> &e, const uint4 (*st)[PoolSmem<P, NW>::CAP], int p)
{
    const uint4 a = st[0][p], b = st[1][p];
    e.bank = a.x; e.d[0] = a.y; e.d[1] = a.z; e.d[2] = a.w;
    e.k[0] = b.x; e.k[1] = b.y; e.k[2] = b.z; e.meta = b.w;
#pragma unroll
    for (int q = 0; q < P; q++) {
        const uint4 v = st[2 + q][p];
        e.pg[q] = v.x; e.pc[q] = v.y; e.pr[q] = v.z; e.pi[q] = v.w;
    }
}
template <int P, int NW>
__device__ __forceinline__ void pool_store(const Env<P> &e, uint4 (*st)[PoolSmem<P, NW>::CAP], int p)
{
    st[0][p] = make_uint4(e.bank, e.d[0], e.d[1], e.d[2]);
    st[1][p] = make_uint4(e.k[0], e.k[1], e.k[2], e.meta);
#pragma unroll
    for (int q = 0; q < P; q++) st[2 + q][p] = make_uint4(e.pg[q], e.pc[q], e.pr[q], e.pi[q]);
}

// a finished game: calScore, win / tie / loss, counters (block-level, shared memory), per-game outputs
template <int P>
__device__ __forceinline__ void pool_finalize(const Env<P> &e, const PlayoutArgs &a, unsigned int *sctr, int64_t g, int steps, int kind)
{
    int sc[P], oc[P];
    final_scores(e, sc, oc);
    unsigned int nob = 0, sum = 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        if (a.score2) a.score2[g * P + p] = (int8_t)sc[p];
        if (oc[p]) atomicAdd(&sctr[(oc[p] == 2 ? SPL_CTR_WINS : SPL_CTR_TIES) + p], 1u);
        sum += (unsigned)sc[p];
        nob += __popc(e.pi[p] >> 16 & 0x3FFu);
    }
    atomicAdd(&sctr[SPL_CTR_GAMES], 1u);
    atomicAdd(&sctr[SPL_CTR_STEPS], (unsigned)steps);
    atomicAdd(&sctr[SPL_CTR_END_SCORE - 1 + kind], 1u);
    atomicAdd(&sctr[SPL_CTR_SCORE2_SUM], sum);
    if (nob) atomicAdd(&sctr[SPL_CTR_NOBLES], nob);
    if (a.steps) a.steps[g] = (int16_t)steps;
    if (a.end_kind) a.end_kind[g] = (uint8_t)kind;
    if (a.state) store_env(e, a.state, a.n_games, g);
}

template <int P, int MODE, int NW>
__global__ void __launch_bounds__(32 * NW, SPL_POOL_BLOCKS_PER_SM) playout_pool_kernel(const __grid_constant__ PlayoutArgs a)
{
    using Sm = PoolSmem<P, NW>;
    constexpr int CAP = Sm::CAP;
    extern __shared__ __align__(16) unsigned char pool_raw[];
    Sm &S = *reinterpret_cast<Sm *>(pool_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 16) S.ctr[threadIdx.x] = 0;
    if (threadIdx.x < 3) S.n_alive[threadIdx.x] = 0;
    stage_tables(S.T);  // ends with a block barrier
    const Tables &T = S.T;

    // the launch's games are split evenly over the blocks; a block starts them as slots of its pool come free
    const int64_t g_lo = a.n_games * (int64_t)blockIdx.x / gridDim.x;
    const int n_mine = (int)(a.n_games * (int64_t)(blockIdx.x + 1) / gridDim.x - g_lo);
    int started = 0;

    // Warps 0 .. NW-2 step; the last warp is the KEEPER: it never steps, it carries the (at most 32) live games beyond the
    // steppers' 32 * (NW - 1) over to the next round untouched and starts new games in its remaining lanes - so the list is
    // topped up every round and the stepping warps stay full.  Stepping warps that have no group this round (start of the
    // launch, small launches, the drain at the end) start new games as well, 32 each.
    constexpr int SCAP = 32 * (NW - 1);
    for (int round = 0, r3 = 0;; round++, r3 = r3 == 2 ? 0 : r3 + 1) {
        const int b = round & 1, r3n = r3 == 2 ? 0 : r3 + 1, r3z = r3n == 2 ? 0 : r3n + 1;
        const int nA = S.n_alive[r3];
        // The last few games of a block are SPREAD: one game per warp (lane 0) instead of all in one warp - a warp runs the
        // union of its lanes' paths one after the other, and at the end of a launch the step latency of the longest games
        // is all that is left.
        const bool spread = nA <= SPL_POOL_SPREAD;
        const int n_step = min(nA, SCAP), chunks = spread ? nA : (n_step + 31) >> 5, carried = nA - n_step;
        // new games this round: the keeper's free lanes first, then 32 per idle stepping warp; never more than fit the list
        const int room = min(min(CAP - nA, n_mine - started), (32 - carried) + 32 * (NW - 1 - chunks));
        if (nA == 0 && room == 0) break;
        if (threadIdx.x == 0) S.n_alive[r3z] = 0;  // last read in the previous round, next appended to in the next one
        const bool keeper = warp == NW - 1, stepper = warp < chunks;
        const int p = keeper ? SCAP + lane : (spread ? warp : warp * 32 + lane);
        const bool valid = keeper ? lane < carried : (stepper && (spread ? lane == 0 : p < n_step));
        // fresh-game slot of this lane: keeper lanes carried .. 31 -> 0 .., idle stepping warp j -> (32 - carried) + 32 j + lane
        const int j_new = keeper ? lane - carried : (32 - carried) + (warp - chunks) * 32 + lane;
        const bool loading = !valid && !stepper && j_new >= 0 && j_new < room;
        const int i_new = started + j_new;
        started += room;
        // a scheduler (warp id mod 4) that holds one stepping warp fewer than the fullest one would idle at the barrier:
        // its warps play one move more per round (games are independent, so rounds need not be the same length for all)
        const int heavy = (chunks + 3) >> 2, mine = (chunks - (warp & 3) + 3) >> 2;
        const int n_steps = SPL_POOL_STEPS + ((chunks & 3) && mine < heavy ? 1 : 0);
        static_assert(SPL_POOL_STEPS + 1 + 3 <= 12, "three Philox blocks cover a round's action draws from any phase");
        if (stepper || keeper || room > 0) {  // warp-uniform: a warp with nothing to do goes straight to the barrier
            Env<P> e;
            uint32_t gid = 0, t0 = 0;
            int kind = SPL_END_NONE, t = 0;
            bool survive = false;
            if (valid) {
                pool_load<P, NW>(e, S.st[b], p);
                gid = S.gid[b][p]; t0 = S.t0[b][p];
                survive = true;
                if (!keeper) {
                    const uint64_t game = (MODE == POOL_STATES && a.game_ids) ? a.game_ids[gid] : a.first_game + gid;
                    const DeckSrc src{(MODE != POOL_FRESH && a.decks) ? a.decks + (int64_t)gid * SPL_DECK_BYTES : nullptr, a.seed, game, a.rk, &T};
                    t = (int)turns_total(e);
                    const int t_end = (int)t0 + a.max_steps;
                    if (MODE == POOL_STATES) kind = game_ends(e, a.flags);  // a game that was over when it was handed in
                    // The round's action draws up front, all lanes at once: move t takes word t & 3 of Philox block t >> 2, so
                    // two blocks (three when the game is mid-block or the round has nine moves) serve the whole round - the
                    // generator is out of the divergent step loop and runs 2-3 times per round instead of 8-9.
                    const uint32_t blk = (uint32_t)t >> 2;
                    S.rnd[threadIdx.x][0] = philox_rk(a.rk, game, blk, STREAM_ACTION);
                    S.rnd[threadIdx.x][1] = philox_rk(a.rk, game, blk + 1, STREAM_ACTION);
                    if ((t & 3) + n_steps > 8) S.rnd[threadIdx.x][2] = philox_rk(a.rk, game, blk + 2, STREAM_ACTION);
                    const uint32_t *draws = reinterpret_cast<const uint32_t *>(&S.rnd[threadIdx.x][0]);  // move t: draws[t - t_blk]
                    const int t_blk = (int)(4u * blk);
#pragma unroll 1
                    for (int s = 0; s < n_steps && kind == SPL_END_NONE; s++) {
                        Ctx c;
                        FamilyCounts f;
                        Choice ch;
                        make_ctx<P, true>(e, T, c);
                        const int total = count_legal(e, T, c, f);
                        const uint32_t r = draws[t - t_blk];
                        select_legal(e, T, c, f, total, (int)__umulhi(r, (uint32_t)(total ? total : c.nS)), ch);
                        if (__builtin_expect(t < a.trace_len, 0)) a.trace[(int64_t)t * a.n_games + gid] = (int16_t)ch.idx;
                        apply_action<P, false, true>(e, ch, src);
                        rotate_players_left(e);  // the next mover's words to the front (the pool keeps states mover-first)
                        t++;
                        kind = game_ends(e, a.flags);
                        if (kind == SPL_END_NONE && t >= t_end) kind = SPL_END_TRUNCATED;
                    }
                    survive = kind == SPL_END_NONE;
                }
            } else if (loading) {
                gid = (uint32_t)(g_lo + i_new);
                if (MODE == POOL_FRESH) reset_env(e, a.seed, a.first_game + gid, a.rk);
                else if (MODE == POOL_LISTS) inject_env(e, a.decks + (int64_t)gid * SPL_DECK_BYTES, a.nobles + (int64_t)gid * 5);
                else { load_env(e, a.state, a.n_games, (int64_t)gid); t0 = turns_total(e); to_mover_first(e); }
                survive = true;  // (fresh games start with seat 0 to move: seat order IS mover-first)
            }
            // survivors, carried and new games -> the next round's list (lanes with `survive` take consecutive slots)
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, survive);
            if (m) {
                const int leader = __ffs(m) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&S.n_alive[r3n], __popc(m));
                const int q = __shfl_sync(0xFFFFFFFFu, base, leader) + __popc(m & ((1u << lane) - 1u));
                if (survive) {
                    pool_store<P, NW>(e, S.st[b ^ 1], q);
                    S.gid[b ^ 1][q] = gid; S.t0[b ^ 1][q] = t0;
                }
            }
            if (valid && !survive) {
                to_seat_order(e);
                pool_finalize(e, a, S.ctr, (int64_t)gid, t - (int)t0, kind);
            }
        }
        __syncthreads();  // the next round's list is complete; this round's list may be overwritten
    }
    if (a.counters && threadIdx.x < 16 && S.ctr[threadIdx.x]) atomicAdd(&a.counters[threadIdx.x], (unsigned long long)S.ctr[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------- step
// Persistent: the grid is (SMs x resident blocks), every block stages the tables ONCE and strides over the envs; the first
// env's action and state loads are issued before the table copy and its barrier.
template <int P>
__global__ void __launch_bounds__(ENV_TB, STEP_BLOCKS_PER_SM) step_kernel(uint4 *state, const uint8_t *decks, uint64_t seed, uint64_t first_game,
                                                      const uint64_t *game_ids, const int16_t *action, int8_t *reward, uint8_t *done, uint8_t *illegal,
                                                      int64_t N, uint32_t flags)
{
    __shared__ __align__(16) unsigned char sT_bytes[TABLES_STEP_BYTES];  // the part of Tables that decode + apply read
    Tables &sT = *reinterpret_cast<Tables *>(sT_bytes);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    Env<P> e;
    int a = -1;
    uint4 row = make_uint4(0, 0, 0, 0);  // the action's decoded fields (engine.cuh DecodeTable), loaded with the state
    if (i < N) {
        a = action[i];
        load_env(e, state, N, i);
        row = decode_row(a);
    }
    stage_tables_nosync(sT, 0, TABLES_STEP_BYTES);  // decode + apply read a quarter of the tables
    __syncthreads();
    // (prefetching the next env's state across the loop was measured slower: 103 registers, 0.0167 -> 0.0211 ms per 262,144 envs)
    while (i < N) {
        int rew = 0, bad = 0;
        if (a >= 0) {
            Ctx c;
            Choice ch;
            make_ctx(e, sT, c);
            if (decode_and_check_row(e, sT, c, a, row, ch, g_tables)) {  // (a pass is checked against the complete tables in global memory)
                const DeckSrc src{decks ? decks + i * SPL_DECK_BYTES : nullptr, seed, game_ids ? game_ids[i] : first_game + (uint64_t)i};
                rew = apply_action(e, ch, src);
            } else {
                bad = 1;
                e.meta |= META_ILLEGAL;
            }
            store_env(e, state, N, i);
        }
        if (reward) reward[i] = (int8_t)rew;
        if (illegal) illegal[i] = (uint8_t)bad;
        if (done) done[i] = (uint8_t)game_ends(e, flags);
        i += stride;
        if (i < N) {
            a = action[i];
            load_env(e, state, N, i);
            row = decode_row(a);
        }
    }
}

template <int P>
__global__ void __launch_bounds__(BLOCK) undo_kernel(uint4 *state, const int8_t *seat, const int16_t *action, const uint8_t *card,
                                                     const uint8_t *noble, const uint8_t *returned, int64_t N)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || action[i] < 0) return;
    Env<P> e;
    load_env(e, state, N, i);
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < 6; c++) r |= (uint32_t)returned[i * 6 + c] << (5 * field_of_colour_id(c));
    undo_action(e, sT, (int)seat[i], (int)action[i], (uint32_t)card[i], (uint32_t)noble[i], r);
    store_env(e, state, N, i);
}

// ---------------------------------------------------------------------------------------------- one-ply expansion
template <int P>
__global__ void __launch_bounds__(BLOCK) count_legal_kernel(const uint4 *state, int32_t *counts, int64_t N)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Env<P> e;
    load_env(e, state, N, i);
    Ctx c;
    FamilyCounts f;
    make_ctx(e, sT, c);
    const int total = count_legal(e, sT, c, f);
    counts[i] = total ? total : c.nS;
}

// children of env i go to columns offsets[i] .. offsets[i+1]-1 of the M-env output batch, in ascending ALL_ACTIONS order
template <int P>
__global__ void __launch_bounds__(BLOCK) expand_kernel(const uint4 *state, const uint8_t *decks, uint64_t seed, uint64_t first_game,
                                                       const uint64_t *game_ids, const int64_t *offsets, uint4 *out_state,
                                                       int16_t *out_action, int8_t *out_reward, int32_t *out_parent, int64_t N, int64_t M)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Env<P> e;
    load_env(e, state, N, i);
    Ctx c;
    FamilyCounts f;
    make_ctx(e, sT, c);
    const int total = count_legal(e, sT, c, f), n = total ? total : c.nS;
    const DeckSrc src{decks ? decks + i * SPL_DECK_BYTES : nullptr, seed, game_ids ? game_ids[i] : first_game + (uint64_t)i};
    const int64_t base = offsets[i];
    for (int k = 0; k < n && base + k < M; k++) {
        Choice ch;
        select_legal(e, sT, c, f, total, k, ch);
        Env<P> child = e;
        const int rew = apply_action(child, ch, src);
        store_env(child, out_state, M, base + k);
        if (out_action) out_action[base + k] = (int16_t)ch.idx;
        if (out_reward) out_reward[base + k] = (int8_t)rew;
        if (out_parent) out_parent[base + k] = (int32_t)i;
    }
}

// ---------------------------------------------------------------------------------------------- depth-2 minimax
// MiniMaxAgent (agents/our_agents/minmax.py): one thread per ROOT ACTION.  Thread c finds its root env in the prefix sums,
// re-derives the root's c-th legal action, plays it, then walks every reply of the opponent (k-th legal action, apply on a
// register copy, evaluate) keeping the minimum - the grandchildren never exist in memory.  No terminal test between the
// plies and no alpha-beta cut: the cut changes neither the root value nor the action chosen (a cut child returns a bound
// <= the best so far and the root only replaces on a strict improvement).
__global__ void __launch_bounds__(BLOCK) minimax_children_kernel(const uint4 *state, const uint8_t *decks, uint64_t seed, uint64_t first_game,
                                                                 const uint64_t *game_ids, const int64_t *offsets, double *child_value,
                                                                 int16_t *child_action, int64_t N, int64_t M)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int64_t cidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= M) return;
    int64_t lo = 0, hi = N;  // the root env: largest i with offsets[i] <= cidx
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= cidx) lo = mid; else hi = mid;
    }
    const int64_t i = lo;
    Env<2> e;
    load_env(e, state, N, i);
    const int me = seat_of(e.meta);
    const DeckSrc src{decks ? decks + i * SPL_DECK_BYTES : nullptr, seed, game_ids ? game_ids[i] : first_game + (uint64_t)i};
    Choice ch;
    {
        Ctx c;
        FamilyCounts f;
        make_ctx(e, sT, c);
        const int total = count_legal(e, sT, c, f);
        select_legal(e, sT, c, f, total, (int)(cidx - offsets[i]), ch);
        apply_action(e, ch, src);
    }
    Ctx c;
    FamilyCounts f;
    make_ctx(e, sT, c);
    const int total = count_legal(e, sT, c, f), n = total ? total : c.nS;
    double best = __longlong_as_double(0x7FF0000000000000LL);
    for (int k = 0; k < n; k++) {
        Choice reply;
        select_legal(e, sT, c, f, total, k, reply);
        Env<2> g = e;
        apply_action(g, reply, src);
        best = fmin(best, minimax_eval(g, sT, me));
    }
    child_value[cidx] = best;
    if (child_action) child_action[cidx] = (int16_t)ch.idx;
}

// per root: the maximum child value; among exactly equal maxima the reference takes the first in its (type name, shuffle)
// order - here the lowest type rank, then the lowest ALL_ACTIONS index
__global__ void __launch_bounds__(BLOCK) minimax_root_kernel(const int64_t *offsets, const double *child_value, const int16_t *child_action,
                                                             int16_t *best_action, double *best_value, int64_t N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double bv = -__longlong_as_double(0x7FF0000000000000LL);
    int ba = -1, br = 99;
    for (int64_t k = offsets[i]; k < offsets[i + 1]; k++) {
        const double v = child_value[k];
        const int a = child_action[k], r = minimax_type_rank(a);
        if (v > bv || (v == bv && r < br)) { bv = v; ba = a; br = r; }
    }
    if (best_action) best_action[i] = (int16_t)ba;
    if (best_value) best_value[i] = bv;
}

template <int P>
__global__ void __launch_bounds__(BLOCK) minimax_eval_kernel(const uint4 *state, const int8_t *seat, double *value, int64_t N)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Env<P> e;
    load_env(e, state, N, i);
    value[i] = minimax_eval(e, sT, seat ? (int)seat[i] : seat_of(e.meta));
}

template <int P>
__global__ void __launch_bounds__(BLOCK) final_scores_kernel(const uint4 *state, int8_t *score2, uint8_t *outcome,
                                                             uint8_t *end_kind, int64_t N, uint32_t flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Env<P> e;
    load_env(e, state, N, i);
    int sc[P], oc[P];
    final_scores(e, sc, oc);
#pragma unroll
    for (int p = 0; p < P; p++) {
        if (score2) score2[i * P + p] = (int8_t)sc[p];
        if (outcome) outcome[i * P + p] = (uint8_t)oc[p];
    }
    if (end_kind) end_kind[i] = (uint8_t)game_ends(e, flags);
}

// ---------------------------------------------------------------------------------------------- tiles
// mask / observation kernels: one thread computes one env's row into a shared-memory tile (engine.cuh thread_fill_*).
// Mask rows are DENSE in the tile (440 B apart), so the 32 rows of a warp are one contiguous, 16-byte aligned 14,080-byte
// span of the output: lane 0 hands it to the bulk-copy engine (cp.async.bulk shared -> global, SASS UBLKCP) and the warp
// moves on - no store loop, no block barrier.  Observation rows are staged as bytes and widened to float32 on the way out
// by the block (coalesced 32-bit stores).
extern __shared__ __align__(16) uint32_t dyn_tile[];

__device__ __forceinline__ void bulk_store_async(void *gdst, const void *ssrc, uint32_t bytes)
{
    SPL_CHECK((bytes & 15u) == 0 && ((uintptr_t)gdst & 15u) == 0);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}
// the issuing thread: all its bulk stores have finished READING shared memory (the tile may be overwritten)
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make this thread's st.shared visible to the bulk-copy engine (async proxy); follow with the warp / block sync
__device__ __forceinline__ void fence_smem_to_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int MASK_ROW_BYTES = SPL_MASK_WORDS * 4;
static_assert((32 * MASK_ROW_BYTES) % 16 == 0 && MASK_ROW_BYTES % 8 == 0, "a warp's mask rows must be a 16-byte multiple");

// one thread's finished mask words [w0, w1) -> its dense tile row (64-bit stores: rows are 8-byte aligned; a half-warp's
// 16 rows start 14 banks apart, so the 32 banks are hit once each)
struct MaskRowSink {
    uint2 *row;
    const uint32_t *m;
    __device__ __forceinline__ void operator()(int w0, int w1) const
    {
#pragma unroll
        for (int w = w0; w < w1; w += 2) row[w >> 1] = make_uint2(m[w], m[w + 1]);
    }
};
// rows [0, n) of a warp's tile -> out; all 32 lanes call it after their rows are written
__device__ __forceinline__ void mask_tile_store(const uint32_t *wtile, uint32_t *out, int n, int lane)
{
    fence_smem_to_async_proxy();
    __syncwarp();
    if ((n & 1) == 0) {  // n * 440 is a multiple of 16
        if (lane == 0 && n > 0) bulk_store_async(out, wtile, (uint32_t)n * MASK_ROW_BYTES);
    } else {             // ragged last tile with an odd row count: plain coalesced stores
        for (int w = lane; w < n * SPL_MASK_WORDS; w += 32) out[w] = wtile[w];
    }
}

// Observation rows are staged as BYTES, DENSE: row j of a warp's tile starts at byte 265 j, so the 32 rows are one flat
// 8,480-byte array whose float32 image is one flat, 16-byte aligned 33,920-byte span of the output.  The warp widens it four
// entries per lane and step: one 32-bit shared-memory load, four byte permutes that drop each byte into 0x4B0000bb (the
// 2^23 trick: 8388608 + b as a float, exact), four FADD, one 128-bit store - 2.5 instructions per entry instead of 4 with
// scalar stores (the row-by-row loop was 37 % of the kernel's instructions and throttled the load/store pipe).
constexpr int OBS_TILE_BYTES = 32 * SPL_OBS_SIZE;
static_assert(OBS_TILE_BYTES % 16 == 0, "a warp's observation rows must be a 16-byte multiple");
__device__ __forceinline__ float byte_to_float(uint32_t b) { return __uint_as_float(0x4B000000u | b) - 8388608.0f; }
__device__ __forceinline__ float byte_of_word_to_float(uint32_t w, int j) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u + j)) - 8388608.0f; }
__device__ __forceinline__ void obs_tile_store(const uint32_t *wtile, float *out, int n, int lane)
{
    const int total = n * SPL_OBS_SIZE;
    if (((uintptr_t)out & 15u) == 0) {
        const int nq = total >> 2;
        float4 *out4 = reinterpret_cast<float4 *>(out);
        // eight quads per lane in flight: the loads of a batch are issued together, and a batch's eight 128-bit stores use
        // distinct registers (a store holds its source registers until the load/store unit has read them)
        constexpr int BATCH = 8;
        int q = lane;
        for (; q + 32 * (BATCH - 1) < nq; q += 32 * BATCH) {
            uint32_t w[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; u++) w[u] = wtile[q + 32 * u];
#pragma unroll
            for (int u = 0; u < BATCH; u++)
                out4[q + 32 * u] = make_float4(byte_of_word_to_float(w[u], 0), byte_of_word_to_float(w[u], 1), byte_of_word_to_float(w[u], 2),
                                               byte_of_word_to_float(w[u], 3));
        }
        for (; q < nq; q += 32) {
            const uint32_t w = wtile[q];
            out4[q] = make_float4(byte_of_word_to_float(w, 0), byte_of_word_to_float(w, 1), byte_of_word_to_float(w, 2), byte_of_word_to_float(w, 3));
        }
        const int f = (nq << 2) + lane;  // a ragged last tile: up to three entries left
        if (f < total) out[f] = byte_to_float(reinterpret_cast<const uint8_t *>(wtile)[f]);
    } else {  // a caller's buffer that is only float-aligned
        for (int f = lane; f < total; f += 32) out[f] = byte_to_float(reinterpret_cast<const uint8_t *>(wtile)[f]);
    }
}
// A lane's row starts (lane & 3) bytes into a word: its finished staged words go out shifted by that much.  Word k of the
// row's span holds the top of o[k - 1] and the bottom of o[k]; k = 0 (shared with the previous lane's last bytes) and
// k = 66 are written by the kernel after the row is complete.  thread_fill_obs finishes the card words [18, 66) in
// ascending ranges, then [0, 18), then the spill words [66, 69) - so a range's first word can use the word before it,
// except at 18 (word 17 is final only with [0, 18)).
struct ObsDenseSink {
    uint32_t *dst;  // the tile word that holds this row's byte 0
    const uint32_t *o;
    uint32_t sh;    // 8 * (lane & 3)
    __device__ __forceinline__ void operator()(int w0, int w1) const
    {
        const int k0 = (w0 == 0 || w0 == 18) ? w0 + 1 : w0, k1 = w1 == 18 ? 19 : (w1 < 66 ? w1 : 66);
#pragma unroll
        for (int k = k0; k < k1; k++) dst[k] = __funnelshift_l(o[k - 1], o[k], sh);
    }
};
// what thread_fill_obs reads of the tables: 3 KB instead of 13 (card and noble are the head of Tables)
struct alignas(16) ObsTables {
    uint2 card[128];
    uint32_t noble[16];
    uint4 cardvec[128];
};
static_assert(offsetof(ObsTables, cardvec) == TABLES_HEAD_BYTES && offsetof(Tables, noble) == offsetof(ObsTables, noble), "ObsTables mirrors the head of Tables");

constexpr int TILE_TB_MAX = 512;  // the row kernels' largest block (row_shape picks warps per block and blocks per SM)

// Software-pipelined over the warp's tiles: the NEXT tile's states (and seats) are loaded into registers before the current
// one is built - the tile, not the register file, caps residency here (70 of 146 registers used at 14 warps per SM), and
// waiting for the state loads was 17 % of all stall samples (profiles/r2_kernels.md, r3g).  The first loads are issued
// before the table copy; the wait for the previous bulk store comes after the context is built, just before the first
// write to the tile.
template <int P>
__global__ void __launch_bounds__(TILE_TB_MAX) legal_mask_kernel(const uint4 *state, const int8_t *seat, uint32_t *mask, int64_t N)
{
    __shared__ __align__(16) unsigned char sT_bytes[TABLES_MASK_BYTES];  // the part of Tables the row builder reads
    Tables &sT = *reinterpret_cast<Tables *>(sT_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
    Env<P> nxt;
    int nseat = 0;
    if (base + lane < N) {
        load_env(nxt, state, N, base + lane);
        if (seat) nseat = seat[base + lane];
    }
    stage_tables(sT, TABLES_MASK_BYTES);
    uint32_t *wtile = dyn_tile + warp * (32 * SPL_MASK_WORDS);  // this warp's 32 dense rows
    for (; base < N; base += stride) {
        const int64_t i = base + lane;
        Env<P> e = nxt;
        const int eseat = nseat;
        if (i + stride < N) {
            load_env(nxt, state, N, i + stride);
            if (seat) nseat = seat[i + stride];
        }
        Ctx c;
        if (i < N) {
            if (seat) e.meta = (e.meta & ~(3u << 20)) | ((uint32_t)eseat << 20);  // the mask of this seat, whoever is to move
            make_ctx(e, sT, c);
        }
        if (lane == 0) bulk_store_wait_read();  // the previous tile has left shared memory
        __syncwarp();
        if (i < N) {
            uint32_t m[MASK_REGS];
#pragma unroll
            for (int w = 0; w < MASK_REGS; w++) m[w] = 0;
            const MaskRowSink sink{reinterpret_cast<uint2 *>(wtile + lane * SPL_MASK_WORDS), m};
            // warp-uniform choice: the general row builder only if some env of this group has a noble waiting (rare)
            if (__all_sync(__activemask(), c.S == 0)) thread_fill_mask<P, MaskRowSink, true>(e, sT, c, m, sink);
            else thread_fill_mask<P, MaskRowSink, false>(e, sT, c, m, sink);
        }
        mask_tile_store(wtile, mask + base * SPL_MASK_WORDS, (int)min((int64_t)32, N - base), lane);
    }
    if (lane == 0) bulk_store_wait_read();
}

// The seven entries of a row that are not plain bytes - turns [1], the variance [6] (one IEEE division), log(1 + buying
// power) [19..23] - are written by the lane that OWNS the row from its registers, all 32 lanes at once, after the warp's
// plain stores.  Round 1 decoded them inside the row loop: the division's slow path ran once per ROW at one active lane
// (12 % of the kernel's instructions, 24 % with the other two cases; profiles/r2_kernels.md).
template <int P>
__global__ void __launch_bounds__(TILE_TB_MAX) observe_kernel(const uint4 *state, const int8_t *seat, float *obs, int64_t N)
{
    __shared__ ObsTables sT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
    Env<P> nxt;  // software-pipelined like the mask kernel: the next tile's states are in flight while this one is built
    int nseat = -1;
    if (base + lane < N) {
        load_env(nxt, state, N, base + lane);
        if (seat) nseat = seat[base + lane];
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&g_tables), *vec = g_tables.cardvec;
        uint4 *dst = reinterpret_cast<uint4 *>(&sT);
        for (int i = threadIdx.x; i < TABLES_HEAD_BYTES / 16; i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < 128; i += blockDim.x) sT.cardvec[i] = vec[i];
        __syncthreads();
    }
    uint32_t *wtile = dyn_tile + warp * (OBS_TILE_BYTES / 4);           // this warp's 32 dense byte rows
    uint32_t *wrow = wtile + ((lane * SPL_OBS_SIZE) >> 2);              // the word that holds this lane's byte 0
    const uint32_t sh = 8u * (lane & 3);
    for (; base < N; base += stride) {
        const int64_t i = base + lane;
        Env<P> e = nxt;
        const int eseat = nseat;
        if (i + stride < N) {
            load_env(nxt, state, N, i + stride);
            if (seat) nseat = seat[i + stride];
        }
        uint32_t first = 0, last = 0, spill0 = 0, spill1 = 0, head4 = 0, head5 = 0;
        if (i < N) {
            uint32_t o[OBS_WORDS + 1];
#pragma unroll
            for (int w = 0; w <= OBS_WORDS; w++) o[w] = 0;
            thread_fill_obs(e, sT, seat ? eseat : seat_of(e.meta), o, ObsDenseSink{wrow, o, sh});
            first = o[0] << sh;
            last = __funnelshift_l(o[65], o[66] & 0xFFu, sh);
            spill0 = o[66]; spill1 = o[67]; head4 = o[4]; head5 = o[5];  // the special entries' sources (bytes 265..268, 19..23)
        }
        // the two boundary words of each row: a row's last bytes share a word with the next row's first ones
        const uint32_t prev_last = __shfl_up_sync(0xFFFFFFFFu, last, 1);
        wrow[0] = (lane & 3) ? (first | prev_last) : first;
        if ((lane & 3) == 3) wrow[66] = last;
        __syncwarp();
        obs_tile_store(wtile, obs + base * SPL_OBS_SIZE, (int)min((int64_t)32, N - base), lane);
        __syncwarp();  // orders the plain stores of every lane before the owners' patches of the same addresses
        if (i < N) {
            float *dst = obs + i * SPL_OBS_SIZE;
            dst[1] = (float)((spill0 >> 24) | (spill1 & 0xFFu) << 8);                 // turns (bytes 267, 268)
            dst[6] = __fdiv_rn((float)((spill0 >> 8) & 0xFFFFu), 25.0f);              // np.var numerator (bytes 265, 266) / 25
            dst[19] = __uint_as_float(g_log1p_bits[(head4 >> 24) & 31u]);
#pragma unroll
            for (int n = 0; n < 4; n++) dst[20 + n] = __uint_as_float(g_log1p_bits[(head5 >> (8 * n)) & 31u]);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- PPO env step
// The learner's action (membership-checked), then RandomAgent opponents until the learner's seat or the end
// (_simulate_opponents, splendor_env.py:219-231): one thread per env out of registers, at full occupancy.  The learner's
// next observation and legal mask come from observe_kernel / legal_mask_kernel on the same stream: a fused version of the
// three was slower (0.214 ms against 0.16 ms per 262,144 envs) because the shared-memory tile of the row kernels caps
// residency at 14 warps per SM, and this latency-bound part then runs at that occupancy too.
template <int P>
__global__ void __launch_bounds__(ENV_TB, ENV_MOVE_BLOCKS_PER_SM) env_move_kernel(uint4 *state, const uint8_t *decks, uint64_t seed, uint64_t first_game,
                                                          const uint64_t *game_ids, const int8_t *learner, const int16_t *action,
                                                          int8_t *reward, uint8_t *terminated, uint8_t *illegal, int64_t N, uint32_t flags)
{
    __shared__ __align__(16) Tables sT;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool self_play = flags & SPL_F_SELF_PLAY;
    Env<P> e;
    int a = -1, me = 0;
    uint4 row = make_uint4(0, 0, 0, 0);
    if (i < N) {  // the first env's loads go out before the table copy and its barrier
        a = action[i];
        me = self_play ? 0 : (int)learner[i];
        load_env(e, state, N, i);
        row = decode_row(a);
    }
    stage_tables_nosync(sT);
    __syncthreads();
    while (i < N) {
        if (self_play) me = seat_of(e.meta);
        const DeckSrc src{decks ? decks + i * SPL_DECK_BYTES : nullptr, seed, game_ids ? game_ids[i] : first_game + (uint64_t)i, nullptr, &sT};  // (the table form of the k-th-bit select for lazy deals)
        int rew = 0, bad = 0;
        bool go = true;
        if (a >= 0) {
            Ctx c;
            Choice ch;
            make_ctx(e, sT, c);
            if (seat_of(e.meta) == me && game_ends(e, flags) == SPL_END_NONE && decode_and_check_row(e, sT, c, a, row, ch, sT))
                rew = apply_action(e, ch, src);
            else { bad = 1; go = false; e.meta |= META_ILLEGAL; }
        }
        for (int guard = 0; go && !self_play && guard < 4; guard++) {
            if (seat_of(e.meta) == me || game_ends(e, flags) != SPL_END_NONE) break;
            Ctx c;
            FamilyCounts f;
            Choice ch;
            const uint32_t r = philox(seed, src.game, turns_total(e), STREAM_ACTION).x;
            make_ctx(e, sT, c);
            const int total = count_legal(e, sT, c, f);
            select_legal(e, sT, c, f, total, (int)__umulhi(r, (uint32_t)(total ? total : c.nS)), ch);
            apply_action(e, ch, src);
        }
        store_env(e, state, N, i);
        if (reward) reward[i] = (int8_t)rew;
        if (illegal) illegal[i] = (uint8_t)bad;
        if (terminated) terminated[i] = (uint8_t)game_ends(e, flags);
        i += stride;
        if (i < N) {
            a = action[i];
            me = self_play ? 0 : (int)learner[i];
            load_env(e, state, N, i);
            row = decode_row(a);
        }
    }
}

// ---------------------------------------------------------------------------------------------- masked action sampling
// The PPO rollout's action choice (agents/our_agents/ppo/network.py:113-115 + training.py:110-116):
//   probs = softmax(where(mask == 0, HUGE_NEG, logits)); a ~ Categorical(probs); log_prob(a)
// consuming the BIT-PACKED legal mask the env kernels emit, and touching ONLY the legal logits (about 22 of 3510 per row):
// one warp per row, lane l owns mask words l, l+32, l+64, l+96 and walks their set bits.  A first version read the whole
// 14 KB row and reached 16 % of HBM bandwidth (3.7 ms per 262,144 rows); the gather reads ~0.7 KB per row.
// Sampling is an exact inverse CDF over fixed-point weights floor(p * 2^24), p = exp(l - max) in (0, 1]: per-word totals,
// a parallel prefix over the 110 totals to find the word, then the owner lane walks that word's bits.  One Philox word per
// row, keyed (seed, row id, step, stream 6).
constexpr int SAMPLE_WARPS = 4;
constexpr int SAMPLE_CACHE = 6;  // gathered logits parked per lane between the passes (a lane rarely owns more legal bits)

__device__ __forceinline__ uint32_t fixed24(float p) { return (uint32_t)(p * 16777216.f); }

__global__ void __launch_bounds__(32 * SAMPLE_WARPS) masked_sample_kernel(const float *logits, int64_t ld, const uint32_t *mask,
                                                                          uint64_t seed, uint64_t first_id, const uint64_t *ids,
                                                                          uint32_t step, int greedy, int16_t *action, float *log_prob,
                                                                          int64_t N)
{
    __shared__ uint32_t word_total[SAMPLE_WARPS][128];
    __shared__ float value_cache[SAMPLE_WARPS][SAMPLE_CACHE * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * SAMPLE_WARPS;
    uint32_t *T = word_total[warp];
    float *cache = value_cache[warp];
    for (int64_t r = (int64_t)blockIdx.x * SAMPLE_WARPS + warp; r < N; r += nwarps) {
        const float *L = logits + r * ld;
        const uint32_t *M = mask + r * SPL_MASK_WORDS;
        uint32_t mw[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int w = q * 32 + lane;
            mw[q] = w < SPL_MASK_WORDS ? M[w] : 0u;
            if (w == SPL_MASK_WORDS - 1) mw[q] &= (1u << (SPL_NUM_ACTIONS - 32 * (SPL_MASK_WORDS - 1))) - 1u;  // bits >= 3510
        }
        // pass 1: max over the legal logits (and its index, lowest on ties); the gathered values are parked in shared
        // memory (SAMPLE_CACHE per lane, conflict-free layout) so that the later passes do not gather again
        float mx = -INFINITY;
        int arg = 0x7FFFFFFF, n_seen = 0;
#pragma unroll
        for (int q = 0; q < 4; q++)
            for (uint32_t bits = mw[q]; bits; bits &= bits - 1) {
                const int i = (q * 32 + lane) * 32 + __ffs(bits) - 1;
                const float x = L[i];
                if (n_seen < SAMPLE_CACHE) cache[n_seen * 32 + lane] = x;
                n_seen++;
                if (x > mx) { mx = x; arg = i; }  // indices ascend within a lane
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float om = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
            const int oa = __shfl_xor_sync(0xFFFFFFFFu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        // pass 2: partition function and per-word fixed-point totals
        float z = 0.f;
        int n_done = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t tot = 0;
            for (uint32_t bits = mw[q]; bits; bits &= bits - 1) {
                const float x = n_done < SAMPLE_CACHE ? cache[n_done * 32 + lane] : L[(q * 32 + lane) * 32 + __ffs(bits) - 1];
                n_done++;
                const float p = expf(x - mx);
                z += p;
                tot += fixed24(p);
            }
            T[q * 32 + lane] = tot;  // words 110..127 get 0
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xFFFFFFFFu, z, o);
        int chosen = arg;
        float p_chosen = 1.f;
        if (!greedy && arg != 0x7FFFFFFF) {
            __syncwarp();
            // lane l sums words 4l .. 4l+3; prefix over lanes; the lane whose range holds the target resolves the word
            const uint32_t t0 = T[4 * lane], t1 = T[4 * lane + 1], t2 = T[4 * lane + 2], t3 = T[4 * lane + 3];
            unsigned long long inc = (unsigned long long)t0 + t1 + t2 + t3;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += up;
            }
            const unsigned long long total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            const uint64_t id = ids ? ids[r] : first_id + (uint64_t)r;
            const uint32_t u = philox(seed, id, step, 6).x;
            const unsigned long long target = __umul64hi(total, (unsigned long long)u << 32);  // floor(total*u/2^32) < total
            const int finder = __ffs(__ballot_sync(0xFFFFFFFFu, inc > target)) - 1;
            unsigned long long before = inc - ((unsigned long long)t0 + t1 + t2 + t3);
            int word = 4 * lane;
            if (before + t0 <= target) { before += t0; word++; if (before + t1 <= target) { before += t1; word++; if (before + t2 <= target) { before += t2; word++; } } }
            word = __shfl_sync(0xFFFFFFFFu, word, finder);
            before = __shfl_sync(0xFFFFFFFFu, before, finder);
            // the lane that owns `word` walks its bits in ascending index order
            const int owner = word & 31;
            if (lane == owner) {
                const uint32_t bits0 = (word >> 5) == 0 ? mw[0] : ((word >> 5) == 1 ? mw[1] : ((word >> 5) == 2 ? mw[2] : mw[3]));
                for (uint32_t bits = bits0; bits; bits &= bits - 1) {
                    const int i = word * 32 + __ffs(bits) - 1;
                    const float p = expf(L[i] - mx);
                    before += fixed24(p);
                    chosen = i; p_chosen = p;
                    if (before > target) break;
                }
            }
            chosen = __shfl_sync(0xFFFFFFFFu, chosen, owner);
            p_chosen = __shfl_sync(0xFFFFFFFFu, p_chosen, owner);
            __syncwarp();
        }
        if (lane == 0) {
            const bool none_legal = arg == 0x7FFFFFFF;  // cannot happen for Splendor masks (a pass is always legal)
            action[r] = none_legal ? (int16_t)-1 : (int16_t)chosen;
            if (log_prob) log_prob[r] = none_legal ? -INFINITY : logf(p_chosen) - logf(z);
        }
    }
}

}  // namespace spl

// ================================================================================================ C ABI
using namespace spl;

static thread_local int g_last_cuda_error = 0;
#define CK(call)                                   \
    do {                                           \
        cudaError_t err__ = (call);                \
        if (err__ != cudaSuccess) {                \
            g_last_cuda_error = (int)err__;        \
            return SPL_E_CUDA;                     \
        }                                          \
    } while (0)

static int check_device()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return SPL_E_NODEVICE;
    return SPL_OK;
}
static inline unsigned grid_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }
constexpr int SPL_MAX_DEVICES = 64;
static int sm_count()  // of the CURRENT device (cached per device: a process may drive several GPUs)
{
    static int sms[SPL_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= SPL_MAX_DEVICES) return 148;
    if (!sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}

// persistent thread-per-env kernels: one block per ENV_TB envs, capped at what is resident at once (the blocks stride)
static unsigned persistent_grid(int64_t n, int blocks_per_sm)
{
    const int64_t want = (n + ENV_TB - 1) / ENV_TB, cap = (int64_t)sm_count() * blocks_per_sm;
    return (unsigned)(want < cap ? want : cap);
}

#define DISPATCH_P(P, ...)             \
    switch (P) {                       \
    case 2: { constexpr int PP = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int PP = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int PP = 4; __VA_ARGS__; } break; \
    default: return SPL_E_BADARG;      \
    }

template <int P, int MODE, int NW>
static int launch_pool_nw(const PlayoutArgs &a, cudaStream_t stream)
{
    using Sm = PoolSmem<P, NW>;
    auto kernel = playout_pool_kernel<P, MODE, NW>;
    static bool configured[SPL_MAX_DEVICES] = {};  // per device: opt in to > 48 KB of dynamic shared memory once
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= SPL_MAX_DEVICES) return SPL_E_BADARG;
    if (!configured[dev]) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm)));
        configured[dev] = true;
    }
    const int64_t cap = (int64_t)sm_count() * SPL_POOL_BLOCKS_PER_SM, want = (a.n_games + 31) / 32;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    kernel<<<grid, 32 * NW, sizeof(Sm), stream>>>(a);
    return SPL_OK;
}
// Two pool sizes: 20 warps per block (19 stepping) for launches that keep the pools refilled, and 19 warps (a budget of
// 104 instead of 96 registers per thread) for launches whose games all fit the pools at once, where the time is set by
// the latency of the rounds, not by the issue rate (measured when the step still spilled at 96 registers: 65,536 games
// 0.3917 -> 0.3866 ms; 2^20 games 2.076 -> 2.072e10 the other way, profiles/r2_playout.md; the step now takes ~90, and the
// smaller pool still wins for 4 players: 65,536 games 0.636 -> 0.626 ms).
template <int P, int MODE>
static int launch_pool(const PlayoutArgs &a, cudaStream_t stream)
{
    constexpr int NW = SPL_POOL_WARPS;
    if (a.n_games <= (int64_t)sm_count() * SPL_POOL_BLOCKS_PER_SM * 32 * (NW - 2)) return launch_pool_nw<P, MODE, NW - 1>(a, stream);
    return launch_pool_nw<P, MODE, NW>(a, stream);
}

extern "C" {

int spl_abi_version(void) { return 1; }
int spl_state_words(int P) { return (P < 2 || P > 4) ? SPL_E_BADARG : 8 + 4 * P; }
int spl_last_cuda_error(void) { return g_last_cuda_error; }
const char *spl_error_string(int code)
{
    switch (code) {
    case SPL_OK: return "ok";
    case SPL_E_BADARG: return "bad argument";
    case SPL_E_CUDA: return cudaGetErrorString((cudaError_t)g_last_cuda_error);
    case SPL_E_NODEVICE: return "no CUDA device (this engine has no CPU fallback)";
    case SPL_E_RANGE: return "max_steps / trace_len out of range";
    default: return "unknown error";
    }
}

int spl_warmup(void)
{
    if (int rc = check_device()) return rc;
    CK(cudaFree(0));
    sm_count();
    return SPL_OK;
}

int spl_reset(void *state, uint64_t seed, uint64_t first_game, const uint64_t *game_ids, const uint8_t *where, int64_t N, int P,
              uint8_t *deck_lists_out, uint8_t *nobles_out, void *stream)
{
    if (!state || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    DISPATCH_P(P, reset_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(
                      (uint4 *)state, seed, first_game, game_ids, where, N, deck_lists_out, nobles_out));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_inject(void *state, const uint8_t *deck_lists, const uint8_t *nobles, int64_t N, int P, void *stream)
{
    if (!state || !deck_lists || !nobles || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    DISPATCH_P(P, inject_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>((uint4 *)state, deck_lists, nobles, N));
    CK(cudaGetLastError());
    return SPL_OK;
}

// Persistent row kernels: every warp strides over 32-env tiles, so a launch lasts ceil(tiles / resident warps) tile times
// and a tile takes longer the more warps share the SM.  Among the (warps per block) x (blocks per SM) shapes whose tiles
// and tables fit the SM's shared memory, take the one with the smallest rounds x (warps per SM + ROW_LATENCY_WARPS) - the
// second factor models a tile time that is part latency, part issue contention (calibrated on 262,144 and 2^20 envs,
// profiles/r2_kernels.md).  SPL_MASK_SHAPE / SPL_OBS_SHAPE = "<warps per block>x<blocks per SM>" overrides it.
struct RowShape { int warps_per_block, blocks_per_sm; };
constexpr int ROW_LATENCY_WARPS = 8;
static RowShape row_shape(int64_t N, size_t tile_bytes, size_t table_bytes, int regs_per_thread, const char *env_name, int block_warps = 0,
                          int max_blocks_per_sm = 4)
{
    const int warp_regs = 32 * ((regs_per_thread + 7) & ~7);  // registers are allocated per warp, eight per thread at a time
    const auto fits = [&](int wb, int nb) {
        return wb >= 1 && wb * 32 <= TILE_TB_MAX && nb >= 1 && nb <= 8 && nb * wb * warp_regs <= 65536 &&
               (size_t)nb * (wb * tile_bytes + table_bytes + 1024) <= 227u * 1024u;
    };
    if (const char *v = getenv(env_name)) {
        int wb = 0, nb = 0;
        if (sscanf(v, "%dx%d", &wb, &nb) == 2 && fits(wb, nb)) return RowShape{wb, nb};
    }
    const int64_t tiles = (N + 31) / 32, sms = sm_count();
    RowShape best{1, 1};
    int64_t best_cost = INT64_MAX;
    for (int nb = 1; nb <= max_blocks_per_sm; nb++)
        for (int wb = block_warps ? block_warps : 1; fits(wb, nb) && (!block_warps || wb == block_warps); wb++) {
            const int w = wb * nb;
            const int64_t rounds = (tiles + sms * w - 1) / (sms * w), cost = rounds * (w + ROW_LATENCY_WARPS);
            if (cost < best_cost || (cost == best_cost && nb < best.blocks_per_sm)) { best_cost = cost; best = RowShape{wb, nb}; }
        }
    return best;
}
static int launch_mask(const void *state, const int8_t *seat, uint32_t *mask, int64_t N, int P, cudaStream_t stream)
{
    cudaFuncAttributes attr;
    DISPATCH_P(P, CK(cudaFuncGetAttributes(&attr, legal_mask_kernel<PP>)));
    const RowShape sh = row_shape(N, 32 * MASK_ROW_BYTES, TABLES_MASK_BYTES, attr.numRegs, "SPL_MASK_SHAPE");
    const int tb = 32 * sh.warps_per_block;
    const size_t smem = (size_t)tb * MASK_ROW_BYTES;
    const int64_t want = (N + tb - 1) / tb, cap = (int64_t)sm_count() * sh.blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    DISPATCH_P(P, CK(cudaFuncSetAttribute(legal_mask_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
               legal_mask_kernel<PP><<<grid, tb, smem, stream>>>((const uint4 *)state, seat, mask, N));
    CK(cudaGetLastError());
    return SPL_OK;
}
static int launch_observe(const void *state, const int8_t *seat, float *obs, int64_t N, int P, cudaStream_t stream)
{
    cudaFuncAttributes attr;
    DISPATCH_P(P, CK(cudaFuncGetAttributes(&attr, observe_kernel<PP>)));
    // blocks of four warps - one per scheduler - and at most five of them per SM: measured best at 262,144 and 2^20 envs
    // (bigger blocks run their build / store phases in step; 24 warps only add contention once DRAM writes are the limit)
    const RowShape sh = row_shape(N, OBS_TILE_BYTES, sizeof(ObsTables), attr.numRegs, "SPL_OBS_SHAPE", 4, 5);
    const int tb = 32 * sh.warps_per_block;
    const size_t smem = (size_t)sh.warps_per_block * OBS_TILE_BYTES;
    const int64_t want = (N + tb - 1) / tb, cap = (int64_t)sm_count() * sh.blocks_per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    DISPATCH_P(P, CK(cudaFuncSetAttribute(observe_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
               observe_kernel<PP><<<grid, tb, smem, stream>>>((const uint4 *)state, seat, obs, N));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_legal_mask(const void *state, uint32_t *mask, int64_t N, int P, void *stream)
{
    if (!state || !mask || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    return launch_mask(state, nullptr, mask, N, P, (cudaStream_t)stream);
}

int spl_step(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids, const int16_t *action,
             int8_t *reward, uint8_t *done, uint8_t *illegal, int64_t N, int P, uint32_t flags, void *stream)
{
    if (!state || !action || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    DISPATCH_P(P, step_kernel<PP><<<persistent_grid(N, STEP_BLOCKS_PER_SM), ENV_TB, 0, (cudaStream_t)stream>>>(
                      (uint4 *)state, decks, seed, first_game, game_ids, action, reward, done, illegal, N, flags));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_masked_sample(const float *logits, int64_t row_stride, const uint32_t *mask, uint64_t seed, uint64_t first_id,
                      const uint64_t *ids, uint32_t step, int greedy, int16_t *action, float *log_prob, int64_t N, void *stream)
{
    if (!logits || !mask || !action || N < 0 || row_stride < SPL_NUM_ACTIONS) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    const int64_t want = (N + SAMPLE_WARPS - 1) / SAMPLE_WARPS, cap = (int64_t)sm_count() * 8;
    masked_sample_kernel<<<(unsigned)(want < cap ? want : cap), 32 * SAMPLE_WARPS, 0, (cudaStream_t)stream>>>(
        logits, row_stride, mask, seed, first_id, ids, step, greedy, action, log_prob, N);
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_count_legal(const void *state, int32_t *counts, int64_t N, int P, void *stream)
{
    if (!state || !counts || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    DISPATCH_P(P, count_legal_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>((const uint4 *)state, counts, N));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_expand(const void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
               const int64_t *offsets, void *out_state, int16_t *out_action, int8_t *out_reward, int32_t *out_parent, int64_t N,
               int64_t M, int P, void *stream)
{
    if (!state || !offsets || !out_state || N < 0 || M < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0 || M == 0) return SPL_OK;
    DISPATCH_P(P, expand_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(
                      (const uint4 *)state, decks, seed, first_game, game_ids, offsets, (uint4 *)out_state, out_action, out_reward,
                      out_parent, N, M));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_minimax_eval(const void *state, const int8_t *seat, double *value, int64_t N, int P, void *stream)
{
    if (!state || !value || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    DISPATCH_P(P, minimax_eval_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>((const uint4 *)state, seat, value, N));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_minimax2(const void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
                 const int64_t *offsets, double *child_value, int16_t *child_action, int16_t *best_action, double *best_value,
                 int64_t N, int64_t M, int P, void *stream)
{
    if (!state || !offsets || !child_value || !child_action || N < 0 || M < 0) return SPL_E_BADARG;
    if (P != 2) return SPL_E_BADARG;  // the reference agent asserts a 2-player game (minmax.py:16,38)
    if (int rc = check_device()) return rc;
    if (N == 0 || M == 0) return SPL_OK;
    minimax_children_kernel<<<grid_for(M, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>((const uint4 *)state, decks, seed, first_game, game_ids,
                                                                                 offsets, child_value, child_action, N, M);
    CK(cudaGetLastError());
    if (best_action || best_value) {
        minimax_root_kernel<<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(offsets, child_value, child_action, best_action, best_value, N);
        CK(cudaGetLastError());
    }
    return SPL_OK;
}

int spl_undo(void *state, const int8_t *seat, const int16_t *action, const uint8_t *card, const uint8_t *noble,
             const uint8_t *returned, int64_t N, int P, void *stream)
{
    if (!state || !seat || !action || !card || !noble || !returned || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    DISPATCH_P(P, undo_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>((uint4 *)state, seat, action, card, noble, returned, N));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_final_scores(const void *state, int8_t *score2, uint8_t *outcome, uint8_t *end_kind, int64_t N, int P,
                     uint32_t flags, void *stream)
{
    if (!state || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    DISPATCH_P(P, final_scores_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(
                      (const uint4 *)state, score2, outcome, end_kind, N, flags));
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_observe(const void *state, const int8_t *seat, float *obs, int64_t N, int P, void *stream)
{
    if (!state || !obs || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    return launch_observe(state, seat, obs, N, P, (cudaStream_t)stream);
}

static int launch_playout(int mode, const PlayoutArgs &args, int P, cudaStream_t stream)
{
    PlayoutArgs a = args;
    philox_round_keys(a.seed, a.rk);
    if (a.max_steps < 1 || a.max_steps > 4096 || a.trace_len < 0) return SPL_E_RANGE;
    if (!a.trace) a.trace_len = 0;  // the kernel tests trace_len only
    if (a.trace && a.trace_len > 0)
        CK(cudaMemsetAsync(a.trace, 0xFF, (size_t)a.trace_len * (size_t)a.n_games * sizeof(int16_t), stream));
    if (mode == POOL_FRESH) { DISPATCH_P(P, if (int rc = launch_pool<PP, POOL_FRESH>(a, stream)) return rc); }
    else if (mode == POOL_STATES) { DISPATCH_P(P, if (int rc = launch_pool<PP, POOL_STATES>(a, stream)) return rc); }
    else { DISPATCH_P(P, if (int rc = launch_pool<PP, POOL_LISTS>(a, stream)) return rc); }
    CK(cudaGetLastError());
    return SPL_OK;
}

int spl_playout(uint64_t seed, uint64_t first_game, int64_t num_games, int P, uint32_t flags, int max_steps, int8_t *score2,
                int16_t *steps, uint8_t *end_kind, int64_t *counters, int16_t *trace, int trace_len, void *final_state,
                void *stream)
{
    if (num_games < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (num_games == 0) return SPL_OK;
    PlayoutArgs a{(uint4 *)final_state, nullptr, nullptr, seed, first_game, nullptr, num_games, flags, max_steps, score2, steps, end_kind,
                  (unsigned long long *)counters, trace, trace_len};
    return launch_playout(POOL_FRESH, a, P, (cudaStream_t)stream);
}

int spl_playout_from(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids, int64_t N,
                     int P, uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind, int64_t *counters, int16_t *trace,
                     int trace_len, void *stream)
{
    if (!state || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    PlayoutArgs a{(uint4 *)state, decks, nullptr, seed, first_game, game_ids, N, flags, max_steps, score2, steps, end_kind,
                  (unsigned long long *)counters, trace, trace_len};
    return launch_playout(POOL_STATES, a, P, (cudaStream_t)stream);
}

int spl_env_step(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
                 const int8_t *learner, const int16_t *action, int8_t *reward, uint8_t *terminated, uint8_t *illegal, uint32_t *mask, float *obs,
                 int64_t N, int P, uint32_t flags, void *stream)
{
    if (!state || (!learner && !(flags & SPL_F_SELF_PLAY)) || !action || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    const int8_t *seat = (flags & SPL_F_SELF_PLAY) ? nullptr : learner;  // self-play: whoever moves next
    DISPATCH_P(P, env_move_kernel<PP><<<persistent_grid(N, ENV_MOVE_BLOCKS_PER_SM), ENV_TB, 0, (cudaStream_t)stream>>>(
                      (uint4 *)state, decks, seed, first_game, game_ids, learner, action, reward, terminated, illegal, N, flags));
    CK(cudaGetLastError());
    // The two row kernels follow in plain stream order.  (Round 2 ran the mask kernel on a side stream: with the row kernels'
    // present shapes the two grids then fight for shared memory and the persistent blocks that start late finish late -
    // 0.122 ms against 0.111 per 262,144 envs; programmatic dependent launch of one behind the other: no difference,
    // profiles/r2_kernels.md.)
    if (obs)
        if (int rc = launch_observe(state, seat, obs, N, P, (cudaStream_t)stream)) return rc;
    if (mask)  // the LEARNER's mask (get_legal_actions_mask uses my_turn, splendor_env.py:184)
        if (int rc = launch_mask(state, seat, mask, N, P, (cudaStream_t)stream)) return rc;
    return SPL_OK;
}

// ---- host-buffer path: library-owned, grow-only device staging, one set per (device, slot)
namespace {
struct Staging {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return SPL_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CK(cudaMalloc(&p, bytes));
        cap = bytes;
        return SPL_OK;
    }
};
std::mutex g_stage_mu;
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// One pipeline of the host-buffer path: its own staging memory and completion event, so that two calls in flight
// (spl_playout_host_async on alternating slots) share nothing but the GPU.  Slots belong to a DEVICE: a process that drives
// several GPUs gets separate staging memory and events on each (the current device at the time of the call).
constexpr int HOST_SLOTS = 2;
struct HostSlot {
    Staging stage;
    cudaEvent_t ev_call_done = nullptr;
    bool used = false;
};
HostSlot g_slots[SPL_MAX_DEVICES][HOST_SLOTS];

// Everything of a host-buffer playout up to (not including) the wait for its completion, all on the caller's stream:
//   2 H2D copies (deck orders, nobles) -> ONE kernel that builds every game from its deck order and plays it to the end
//   (playout_pool_kernel<P, POOL_LISTS>: no separate inject pass, no state array in HBM) -> 4 D2H copies.
// 10 CUDA API calls per batch (round 1 issued ~35: four chunks on side streams, each with its own inject + playout + events).
// On return the work is queued; `ctr_out` (host, 16 x int64) and the output arrays are written when `stream` gets there.
// Caller holds g_stage_mu.
int playout_host_enqueue(int slot, const uint8_t *deck_lists, const uint8_t *nobles, uint64_t seed, uint64_t first_game, int64_t N,
                         int P, uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind, int64_t *ctr_out,
                         cudaStream_t stream)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= SPL_MAX_DEVICES) return SPL_E_BADARG;
    HostSlot &S = g_slots[dev][slot];
    if (!S.ev_call_done) CK(cudaEventCreateWithFlags(&S.ev_call_done, cudaEventDisableTiming));
    const size_t o_decks = 0, o_nobles = o_decks + align256((size_t)N * SPL_DECK_BYTES), o_score = o_nobles + align256((size_t)N * 5),
                 o_steps = o_score + align256((size_t)N * P), o_kind = o_steps + align256((size_t)N * 2),
                 o_ctr = o_kind + align256((size_t)N), total = o_ctr + 256;
    if (total > S.stage.cap && S.used) CK(cudaEventSynchronize(S.ev_call_done));  // growing frees the old buffer: let its last user finish
    if (int rc = S.stage.ensure(total)) return rc;
    char *b = (char *)S.stage.p;
    // a slot is reused only after the previous call on it has completed (device-side wait: a caller that alternates
    // slots as documented never blocks here; one that does not is serialised instead of corrupted)
    if (S.used) CK(cudaStreamWaitEvent(stream, S.ev_call_done, 0));
    CK(cudaMemsetAsync(b + o_ctr, 0, 128, stream));
    CK(cudaMemcpyAsync(b + o_decks, deck_lists, (size_t)N * SPL_DECK_BYTES, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(b + o_nobles, nobles, (size_t)N * 5, cudaMemcpyHostToDevice, stream));
    PlayoutArgs a{nullptr, (const uint8_t *)(b + o_decks), (const uint8_t *)(b + o_nobles), seed, first_game, nullptr, N, flags, max_steps,
                  (int8_t *)(b + o_score), (int16_t *)(b + o_steps), (uint8_t *)(b + o_kind), (unsigned long long *)(b + o_ctr), nullptr, 0};
    if (int rc = launch_playout(POOL_LISTS, a, P, stream)) return rc;
    if (score2) CK(cudaMemcpyAsync(score2, b + o_score, (size_t)N * P, cudaMemcpyDeviceToHost, stream));
    if (steps) CK(cudaMemcpyAsync(steps, b + o_steps, (size_t)N * 2, cudaMemcpyDeviceToHost, stream));
    if (end_kind) CK(cudaMemcpyAsync(end_kind, b + o_kind, (size_t)N, cudaMemcpyDeviceToHost, stream));
    if (ctr_out) CK(cudaMemcpyAsync(ctr_out, b + o_ctr, 128, cudaMemcpyDeviceToHost, stream));
    CK(cudaEventRecord(S.ev_call_done, stream));
    S.used = true;
    return SPL_OK;
}
}  // namespace

int spl_playout_host(const uint8_t *deck_lists, const uint8_t *nobles, uint64_t seed, uint64_t first_game, int64_t N, int P,
                     uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind, int64_t *counters,
                     void *stream_)
{
    if (!deck_lists || !nobles || N < 0 || P < 2 || P > 4) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    std::lock_guard<std::mutex> lock(g_stage_mu);
    int64_t ctr[16];
    if (int rc = playout_host_enqueue(0, deck_lists, nobles, seed, first_game, N, P, flags, max_steps, score2, steps, end_kind,
                                      ctr, stream))
        return rc;
    CK(cudaStreamSynchronize(stream));
    if (counters)
        for (int i = 0; i < 16; i++) counters[i] += ctr[i];
    return SPL_OK;
}

int spl_playout_host_async(const uint8_t *deck_lists, const uint8_t *nobles, uint64_t seed, uint64_t first_game, int64_t N, int P,
                           uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind, int64_t *counters,
                           int slot, void *stream_)
{
    if (!deck_lists || !nobles || N < 0 || P < 2 || P > 4 || slot < 0 || slot >= HOST_SLOTS) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) {
        if (counters) memset(counters, 0, 128);
        return SPL_OK;
    }
    std::lock_guard<std::mutex> lock(g_stage_mu);
    return playout_host_enqueue(slot, deck_lists, nobles, seed, first_game, N, P, flags, max_steps, score2, steps, end_kind,
                                counters, (cudaStream_t)stream_);
}

}  // extern "C"

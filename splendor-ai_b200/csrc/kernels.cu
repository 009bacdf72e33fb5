// kernels.cu -- CUDA kernels (sm_100a) and the C ABI of include/splendor_b200.h.
//
//   playout_kernel      one thread = one game, state in registers for the whole game: reset -> {terminal test, legal count,
//                       Philox choice, k-th legal action, apply} x steps -> calScore -> counters.  No HBM traffic inside
//                       the loop except the optional trace; the hardware block scheduler refills finished blocks.
//   legal_mask_kernel   persistent; one thread builds one env's 3510-bit row in registers (32/64-bit group patterns) and
//                       flushes finished words into its warp's dense shared-memory tile; lane 0 sends the warp's 32 rows
//                       (14,080 contiguous bytes) to HBM with one bulk copy (cp.async.bulk) while the warp moves on.
//   step_kernel         thread-per-env: decode the ALL_ACTIONS index, membership test, apply, reward / done / illegal.
//   observe_kernel      same scheme for the 265-float observation: rows staged as bytes, widened to float32 on the way
//                       out with coalesced stores, warp by warp (no block barrier).
//   env_move_kernel     the PPO rollout call's first part: learner move + in-kernel random opponents, thread-per-env at
//                       full occupancy; spl_env_step then runs observe_kernel and legal_mask_kernel for the learner's seat.
//   reset / inject / final_scores: thread-per-env.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "engine.cuh"
#include "minimax.cuh"

namespace spl {

__device__ const Tables g_tables = make_tables();
__device__ const uint32_t g_log1p_bits[32] = SPL_LOG1P_BITS;

constexpr int BLOCK = 128;
constexpr int PLAYOUT_BLOCK = 128;  // default threads per block of the playout kernel
#ifndef SPL_PLAYOUT_MINBLOCKS
#define SPL_PLAYOUT_MINBLOCKS 5  // resident 128-thread blocks per SM the playout kernel is compiled for (register cap)
#endif

__device__ __forceinline__ void stage_tables(Tables &sT)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&g_tables);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&sT);
    for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}
static_assert(sizeof(Tables) % 4 == 0, "Tables must be word-copyable");

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

template <int P>
__global__ void __launch_bounds__(BLOCK) inject_kernel(uint4 *state, const uint8_t *deck_lists, const uint8_t *nobles, int64_t N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Env<P> e;
    const uint32_t n = P == 2 ? 4 : (P == 3 ? 5 : 7);
    e.bank = n * (ONES6 & ~(1u << YSH)) | (5u << YSH);
    e.k[0] = 40u | 30u << 8 | 20u << 16; e.k[1] = 0; e.k[2] = 0;
    uint32_t list = 0xFFFFFu;
    for (int k = 0; k < P + 1; k++) list = (list & ~(15u << (4 * k))) | ((uint32_t)(nobles[i * 5 + k] & 15u) << (4 * k));
    e.meta = list | ((uint32_t)(P - 2) << 22) | META_EXPLICIT;
    const DeckSrc src{deck_lists + i * SPL_DECK_BYTES, 0, 0};
#pragma unroll
    for (int t = 0; t < 3; t++) {
        uint32_t w = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) w |= deal(e, t, src) << (8 * c);
        e.d[t] = w;
    }
#pragma unroll
    for (int p = 0; p < P; p++) { e.pg[p] = 0; e.pc[p] = 0; e.pr[p] = 0x00FFFFFFu; e.pi[p] = 0; }
    store_env(e, state, N, i);
}

// ---------------------------------------------------------------------------------------------- playout
struct PlayoutArgs {
    uint4 *state;           // FRESH: optional final-state output; else in/out states
    const uint8_t *decks;   // explicit deck lists (only with states) or nullptr
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

// ONE GAME PER THREAD, one launch = n_games threads in blocks of `blk`.  The hardware block scheduler is the refill
// mechanism: as soon as the slowest game of a block ends, a fresh block takes its place, so every SM always holds as
// many games as its registers allow, and the 32 games of a warp stay in the same phase (reset together, step together,
// score together).  An earlier persistent form (lanes drew their next game themselves) left each lane of a warp in a
// different phase - reset / finalize ran one lane at a time and the instruction cache thrashed - and was 25-40 % slower
// on large sweeps (profiles/r1_playout_ncu.md).
template <int P, bool FRESH>
__global__ void __launch_bounds__(BLOCK, SPL_PLAYOUT_MINBLOCKS) playout_kernel(const __grid_constant__ PlayoutArgs a)  // 92 registers
{
    __shared__ Tables sT;
    __shared__ unsigned long long sctr[16];
    // FRESH: every game's whole deck order, drawn at reset (reset_env_lists); 92-byte pitch = 23 words, odd, so the 32
    // lanes of a warp writing the same entry hit 32 different banks
    constexpr int LIST_PITCH = 92;
    __shared__ __align__(4) uint8_t s_lists[FRESH ? BLOCK * LIST_PITCH : 4];
    if (threadIdx.x < 16) sctr[threadIdx.x] = 0;
    stage_tables(sT);

    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = g < a.n_games;
    uint8_t *const my_lists = s_lists + (FRESH ? threadIdx.x * LIST_PITCH : 0);
    Env<P> e;
    int t = 0, kind = SPL_END_NONE;
    if (live) {
        DeckSrc src{nullptr, a.seed, (!FRESH && a.game_ids) ? a.game_ids[g] : a.first_game + (uint64_t)g, a.rk};
        if (FRESH) {
            reset_env_lists(e, a.seed, src.game, a.rk, my_lists);
            src.lists = my_lists;
        } else {
            load_env(e, a.state, a.n_games, g);
            src.lists = a.decks ? a.decks + g * SPL_DECK_BYTES : nullptr;
        }
        t = (int)turns_total(e);
        const int t_end = t + a.max_steps;
        for (;;) {
            kind = game_ends(e, a.flags);
            if (kind == SPL_END_NONE && t >= t_end) kind = SPL_END_TRUNCATED;
            if (kind != SPL_END_NONE) break;
            Ctx c;
            FamilyCounts f;
            Choice ch;
            const uint32_t r = philox_rk(a.rk, src.game, (uint32_t)t, STREAM_ACTION).x;
            make_ctx(e, sT, c);
            const int total = count_legal(e, sT, c, f);
            select_legal(e, sT, c, f, total, (int)__umulhi(r, (uint32_t)(total ? total : c.nS)), ch);
            if (__builtin_expect(t < a.trace_len, 0)) a.trace[(int64_t)t * a.n_games + g] = (int16_t)ch.idx;  // trace_len is 0 without a trace
            apply_action<P, FRESH>(e, ch, src);
            t++;
        }
    }
    // the warp is back together: calScore, win/tie/loss, counters, outputs
    if (live) {
        int sc[P], oc[P];
        final_scores(e, sc, oc);
        unsigned long long nob = 0, sum = 0;
#pragma unroll
        for (int p = 0; p < P; p++) {
            if (a.score2) a.score2[g * P + p] = (int8_t)sc[p];
            if (oc[p] == 2) atomicAdd(&sctr[SPL_CTR_WINS + p], 1ull);
            if (oc[p] == 1) atomicAdd(&sctr[SPL_CTR_TIES + p], 1ull);
            sum += (unsigned)sc[p];
            nob += __popc(e.pi[p] >> 16 & 0x3FFu);
        }
        atomicAdd(&sctr[SPL_CTR_GAMES], 1ull);
        atomicAdd(&sctr[SPL_CTR_STEPS], (unsigned long long)t);
        atomicAdd(&sctr[SPL_CTR_END_SCORE - 1 + kind], 1ull);
        atomicAdd(&sctr[SPL_CTR_SCORE2_SUM], sum);
        atomicAdd(&sctr[SPL_CTR_NOBLES], nob);
        if (a.steps) a.steps[g] = (int16_t)t;
        if (a.end_kind) a.end_kind[g] = (uint8_t)kind;
        if (a.state) {
            if (FRESH) lists_to_masks(e, my_lists);  // back to the counter-based deck form of the ABI
            store_env(e, a.state, a.n_games, g);
        }
    }
    __syncthreads();
    if (a.counters && threadIdx.x < 16 && sctr[threadIdx.x]) atomicAdd(&a.counters[threadIdx.x], sctr[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------- playout, block pools
// GAMES LIVE IN SHARED MEMORY, WARPS ARE WORKERS.  A block owns a pool of up to CAP = 32 * NW games (packed state words
// in shared memory, 64 B per 2-player game) and plays them in ROUNDS: one round = one move of every live game of the pool.
//   phase A  warp w takes the w-th group of 32 live games (always full except the last group): load the state, terminal
//            context, the legal COUNTS of all four action families, the Philox draw -> (family, rank inside the family);
//            the game gets a slot in the family-sorted list (warp-aggregated shared-memory atomics), and - after a block
//            barrier, when the family totals are known - writes its state, (family, rank, context bits) and the family's
//            count words there;
//   phase B  warp w takes the w-th group of 32 entries of the SORTED list: every lane of the warp (except in the at most
//            three groups that straddle a family boundary) holds a game of the SAME family, so the family-specific
//            selection and the apply run without divergence; the survivors are appended to the next round's live list,
//            finished games are scored and written out;
//   refill   whenever 32 slots are free a spare warp starts 32 new games (reset / load) with all lanes active.
// This removes the two losses the one-game-per-thread kernel measured (profiles/r1q_playout.md): 24 of 32 lanes alive
// (a warp waited for its longest game) and the ~470 instructions per step that every warp executed inside family branches
// at ~8 lanes.  Decks are drawn lazily from the counter-based form (a deal is one Philox call + a bit select, now at full
// lanes inside a reserve / buy group), so nothing but the 16-24 state words per game has to live on chip.
// Results do not depend on where a game sits: every random draw is keyed by (seed, game id, turn / draw number).
#ifndef SPL_POOL_WARPS
#define SPL_POOL_WARPS 10
#endif
#ifndef SPL_POOL_BLOCKS_PER_SM
#define SPL_POOL_BLOCKS_PER_SM 2
#endif
constexpr int POOL_PASS_WORDS = 11;  // header (2) + the collect family's nine group-count words

template <int P, int NW>
struct PoolSmem {
    static constexpr int CAP = 32 * NW, C = 2 + P;
    uint4 st[2][C][CAP];                // [0] live list of the round, [1] family-sorted list; chunk-major like the HBM layout
    uint32_t gid[2][CAP];               // game index inside the launch
    uint32_t t0[2][CAP];                // total turns when the game entered the pool (0 for fresh games)
    uint32_t pw[POOL_PASS_WORDS][CAP];  // phase A -> phase B words, at sorted positions
    Tables T;
    unsigned long long ctr[16];
    int n_alive[2];
    int fam_cnt[2][4];
};

enum PoolMode { POOL_FRESH = 0, POOL_STATES = 1 };

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
__device__ __forceinline__ void pool_finalize(const Env<P> &e, const PlayoutArgs &a, unsigned long long *sctr, int64_t g, int steps, int kind)
{
    int sc[P], oc[P];
    final_scores(e, sc, oc);
    unsigned long long nob = 0, sum = 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        if (a.score2) a.score2[g * P + p] = (int8_t)sc[p];
        if (oc[p] == 2) atomicAdd(&sctr[SPL_CTR_WINS + p], 1ull);
        if (oc[p] == 1) atomicAdd(&sctr[SPL_CTR_TIES + p], 1ull);
        sum += (unsigned)sc[p];
        nob += __popc(e.pi[p] >> 16 & 0x3FFu);
    }
    atomicAdd(&sctr[SPL_CTR_GAMES], 1ull);
    atomicAdd(&sctr[SPL_CTR_STEPS], (unsigned long long)steps);
    atomicAdd(&sctr[SPL_CTR_END_SCORE - 1 + kind], 1ull);
    atomicAdd(&sctr[SPL_CTR_SCORE2_SUM], sum);
    atomicAdd(&sctr[SPL_CTR_NOBLES], nob);
    if (a.steps) a.steps[g] = (int16_t)steps;
    if (a.end_kind) a.end_kind[g] = (uint8_t)kind;
    if (a.state) store_env(e, a.state, a.n_games, g);
}

// slot of this lane in a shared list: lanes with `want` take consecutive slots starting at the counter's old value
__device__ __forceinline__ int pool_append(int *counter, bool want, int lane)
{
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
    int base = 0;
    if (m && lane == __ffs(m) - 1) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, m ? __ffs(m) - 1 : 0);
    return base + __popc(m & ((1u << lane) - 1u));
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
    if (threadIdx.x < 2) S.n_alive[threadIdx.x] = 0;
    if (threadIdx.x < 8) S.fam_cnt[threadIdx.x >> 2][threadIdx.x & 3] = 0;
    stage_tables(S.T);  // ends with a block barrier
    const Tables &T = S.T;

    // the launch's games are split evenly over the blocks; a block starts them as slots of its pool come free
    const int64_t g_hi = a.n_games * (int64_t)(blockIdx.x + 1) / gridDim.x;
    int64_t next = a.n_games * (int64_t)blockIdx.x / gridDim.x;

    for (int round = 0;; round++) {
        const int b = round & 1;
        const int nA = S.n_alive[b];
        const int chunks = (nA + 31) >> 5;
        const int64_t left = g_hi - next;
        int n_new = min(NW - chunks, (CAP - nA) >> 5);
        n_new = (int)min((int64_t)n_new, (left + 31) >> 5);
        if (chunks == 0 && n_new == 0) break;
        const bool stepper = warp < chunks;
        const int p = warp * 32 + lane;
        const bool valid = stepper && p < nA;
        const int64_t g_new = next + (int64_t)(warp - chunks) * 32 + lane;
        const bool loading = !stepper && warp - chunks < n_new && g_new < g_hi;

        Env<P> e;
        uint32_t gid = 0, t0 = 0, w0 = 0, w1 = 0, w2 = 0;
        FamilyCounts f;
        int fam = 4, rank = 0;
        bool fresh_alive = false;
        // ------------------------------------------------------------------ phase A (and the loading of new games)
        if (valid) {
            pool_load<P, NW>(e, S.st[0], p);
            gid = S.gid[0][p]; t0 = S.t0[0][p];
            const uint64_t game = (MODE == POOL_STATES && a.game_ids) ? a.game_ids[gid] : a.first_game + gid;
            const uint32_t r = philox_rk(a.rk, game, turns_total(e), STREAM_ACTION).x;
            Ctx c;
            make_ctx(e, T, c);
            const int total = count_legal(e, T, c, f);
            int kf;
            fam = family_of_k(c, f, total, (int)__umulhi(r, (uint32_t)(total ? total : c.nS)), kf);
            w0 = (uint32_t)kf | (uint32_t)fam << 13 | c.S << 15 | (uint32_t)c.hold << 20 | (c.bank_yellow ? 1u << 24 : 0u);
            w1 = c.h1 | c.h2 << 6 | c.h3 << 12 | c.occ << 18;
            w2 = c.near;
        } else if (loading) {
            gid = (uint32_t)g_new;
            if (MODE == POOL_FRESH) {
                reset_env(e, a.seed, a.first_game + (uint64_t)g_new, a.rk);
                fresh_alive = true;
            } else {
                load_env(e, a.state, a.n_games, g_new);
                t0 = turns_total(e);
                const int kind0 = game_ends(e, a.flags);
                fresh_alive = kind0 == SPL_END_NONE;
                if (!fresh_alive) pool_finalize(e, a, S.ctr, g_new, 0, kind0);  // the game was over when it was handed in
            }
        }
        if (stepper) {  // slot inside the family (all 32 lanes take part; lanes beyond the list carry fam == 4)
#pragma unroll
            for (int F = 0; F < 4; F++) {
                const int slot = pool_append(&S.fam_cnt[b][F], fam == F, lane);
                if (fam == F) rank = slot;
            }
        }
        __syncthreads();  // ---- barrier 1: family totals are final; every read of the live list st[0] is done
        if (threadIdx.x == 0) {
            S.n_alive[b] = 0;  // this round's count has been read by everybody; it is the next round's append counter
#pragma unroll
            for (int F = 0; F < 4; F++) S.fam_cnt[b ^ 1][F] = 0;
        }
        if (valid) {
            int base = 0;
#pragma unroll
            for (int F = 0; F < 3; F++) base += fam > F ? S.fam_cnt[b][F] : 0;
            const int q = base + rank;
            pool_store<P, NW>(e, S.st[1], q);
            S.gid[1][q] = gid; S.t0[1][q] = t0;
            S.pw[0][q] = w0; S.pw[1][q] = w1;
            if (fam == FAM_COLLECT) {
#pragma unroll
                for (int i = 0; i < 9; i++) S.pw[2 + i][q] = f.grp[i];
            } else if (fam == FAM_BUY) {
                S.pw[2][q] = w2;
#pragma unroll
                for (int i = 0; i < 4; i++) S.pw[3 + i][q] = f.buy[i];
            }
        }
        __syncthreads();  // ---- barrier 2: the sorted list is complete
        // ------------------------------------------------------------------ phase B
        bool survive = fresh_alive;
        int kind = SPL_END_NONE, t = 0;
        if (valid) {
            pool_load<P, NW>(e, S.st[1], p);
            gid = S.gid[1][p]; t0 = S.t0[1][p];
            w0 = S.pw[0][p]; w1 = S.pw[1][p];
            fam = (int)(w0 >> 13 & 3u);
            const int kf = (int)(w0 & 0x1FFFu);
            const uint64_t game = (MODE == POOL_STATES && a.game_ids) ? a.game_ids[gid] : a.first_game + gid;
            const DeckSrc src{(MODE == POOL_STATES && a.decks) ? a.decks + (int64_t)gid * SPL_DECK_BYTES : nullptr, a.seed, game, a.rk};
            Ctx c;
            const int seat = seat_of(e.meta);
            c.gems = e.pg[0]; c.cards = e.pc[0]; c.resv = e.pr[0];
#pragma unroll
            for (int q = 1; q < P; q++)
                if (seat == q) { c.gems = e.pg[q]; c.cards = e.pc[q]; c.resv = e.pr[q]; }
            c.bank = e.bank;
            c.S = w0 >> 15 & 31u; c.nS = c.S ? __popc(c.S) : 1;
            c.hold = (int)(w0 >> 20 & 15u); c.bank_yellow = w0 >> 24 & 1u;
            c.h1 = w1 & 63u; c.h2 = w1 >> 6 & 63u; c.h3 = w1 >> 12 & 63u; c.occ = w1 >> 18;
            c.yellow = (int)field(c.gems, YSH);
            c.near = 0;
            Choice ch;
            if (fam == FAM_COLLECT) {
#pragma unroll
                for (int i = 0; i < 9; i++) f.grp[i] = S.pw[2 + i][p];
                select_family(e, T, c, f, FAM_COLLECT, kf, ch);
            } else if (fam == FAM_RESERVE) {
                select_family(e, T, c, f, FAM_RESERVE, kf, ch);
            } else if (fam == FAM_BUY) {
                c.near = S.pw[2][p];
#pragma unroll
                for (int i = 0; i < 4; i++) f.buy[i] = S.pw[3 + i][p];
                select_family(e, T, c, f, FAM_BUY, kf, ch);
            } else {
                select_family(e, T, c, f, FAM_PASS, kf, ch);
            }
            t = (int)turns_total(e);
            if (__builtin_expect(t < a.trace_len, 0)) a.trace[(int64_t)t * a.n_games + gid] = (int16_t)ch.idx;
            apply_action<P, false>(e, ch, src);
            t++;
            kind = game_ends(e, a.flags);
            if (kind == SPL_END_NONE && t - (int)t0 >= a.max_steps) kind = SPL_END_TRUNCATED;
            survive = kind == SPL_END_NONE;
        }
        if (stepper || n_new) {  // warp-uniform: the stepping warps and (when games are being started) the others
            const int q = pool_append(&S.n_alive[b ^ 1], survive, lane);
            if (survive) {
                pool_store<P, NW>(e, S.st[0], q);
                S.gid[0][q] = gid; S.t0[0][q] = t0;
            }
        }
        if (valid && !survive) pool_finalize(e, a, S.ctr, (int64_t)gid, t - (int)t0, kind);
        next = min(g_hi, next + (int64_t)n_new * 32);
        __syncthreads();  // ---- barrier 3: the next round's live list is complete
    }
    if (a.counters && threadIdx.x < 16 && S.ctr[threadIdx.x]) atomicAdd(&a.counters[threadIdx.x], S.ctr[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------- step
template <int P>
__global__ void __launch_bounds__(BLOCK) step_kernel(uint4 *state, const uint8_t *decks, uint64_t seed, uint64_t first_game,
                                                     const uint64_t *game_ids, const int16_t *action, int8_t *reward, uint8_t *done, uint8_t *illegal,
                                                     int64_t N, uint32_t flags)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int a = action[i];
    Env<P> e;
    load_env(e, state, N, i);
    int rew = 0, bad = 0;
    if (a >= 0) {
        Ctx c;
        Choice ch;
        make_ctx(e, sT, c);
        if (decode_and_check(e, sT, c, a, ch)) {
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

// a warp's staged byte rows [0, n) -> float32 observations: lane l writes entries l, l + 32, ... of each row (coalesced
// 128-byte stores); bytes widen to float with the 2^23 trick (one OR, one FADD, exact below 2^23)
__device__ __forceinline__ float byte_to_float(uint32_t b) { return __uint_as_float(0x4B000000u | b) - 8388608.0f; }
__device__ __forceinline__ void obs_tile_store(const uint8_t *wtile, float *out, int n, int lane)
{
    constexpr int FULL = SPL_OBS_SIZE / 32;  // 8 full 32-entry groups, then 9 entries
#pragma unroll 2
    for (int j = 0; j < n; j++) {
        const uint8_t *row = wtile + j * OBS_ROW_BYTES;
        float *dst = out + (int64_t)j * SPL_OBS_SIZE;
        dst[lane] = obs_entry(row, lane, g_log1p_bits);  // the three special entries (turns, variance, log table) are all < 32
#pragma unroll
        for (int k = 1; k < FULL; k++) dst[lane + 32 * k] = byte_to_float(row[lane + 32 * k]);
        if (lane < SPL_OBS_SIZE - 32 * FULL) dst[lane + 32 * FULL] = byte_to_float(row[lane + 32 * FULL]);
    }
}
struct ObsRowSink {
    uint32_t *row;
    const uint32_t *o;
    __device__ __forceinline__ void operator()(int w0, int w1) const
    {
#pragma unroll
        for (int w = w0; w < w1; w++) row[w] = o[w];
    }
};

constexpr int MASK_TB = 224;     // default block sizes (warps work independently in the mask kernel; the tile caps residency)
constexpr int OBS_TB = 128;
constexpr int TILE_TB_MAX = 512; // SPL_MASK_BLOCK / SPL_OBS_BLOCK may raise the block size up to this

template <int P>
__global__ void __launch_bounds__(TILE_TB_MAX) legal_mask_kernel(const uint4 *state, const int8_t *seat, uint32_t *mask, int64_t N)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wtile = dyn_tile + warp * (32 * SPL_MASK_WORDS);  // this warp's 32 dense rows
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32; base < N; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + lane;
        if (lane == 0) bulk_store_wait_read();  // the previous tile has left shared memory
        __syncwarp();
        if (i < N) {
            Env<P> e;
            load_env(e, state, N, i);
            if (seat) e.meta = (e.meta & ~(3u << 20)) | ((uint32_t)seat[i] << 20);  // the mask of this seat, whoever is to move
            Ctx c;
            make_ctx(e, sT, c);
            uint32_t m[MASK_REGS];
#pragma unroll
            for (int w = 0; w < MASK_REGS; w++) m[w] = 0;
            thread_fill_mask(e, sT, c, m, MaskRowSink{reinterpret_cast<uint2 *>(wtile + lane * SPL_MASK_WORDS), m});
        }
        mask_tile_store(wtile, mask + base * SPL_MASK_WORDS, (int)min((int64_t)32, N - base), lane);
    }
    if (lane == 0) bulk_store_wait_read();
}

template <int P>
__global__ void __launch_bounds__(TILE_TB_MAX) observe_kernel(const uint4 *state, const int8_t *seat, float *obs, int64_t N)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *wtile = dyn_tile + warp * (32 * OBS_WORDS);  // this warp's 32 byte rows
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32; base < N; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + lane;
        if (i < N) {
            Env<P> e;
            load_env(e, state, N, i);
            uint32_t o[OBS_WORDS + 1];
#pragma unroll
            for (int w = 0; w <= OBS_WORDS; w++) o[w] = 0;
            thread_fill_obs(e, sT, seat ? (int)seat[i] : seat_of(e.meta), o, ObsRowSink{wtile + lane * OBS_WORDS, o});
        }
        __syncwarp();
        obs_tile_store(reinterpret_cast<const uint8_t *>(wtile), obs + base * SPL_OBS_SIZE, (int)min((int64_t)32, N - base), lane);
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
__global__ void __launch_bounds__(BLOCK) env_move_kernel(uint4 *state, const uint8_t *decks, uint64_t seed, uint64_t first_game,
                                                         const uint64_t *game_ids, const int8_t *learner, const int16_t *action,
                                                         int8_t *reward, uint8_t *terminated, uint8_t *illegal, int64_t N, uint32_t flags)
{
    __shared__ Tables sT;
    stage_tables(sT);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Env<P> e;
    load_env(e, state, N, i);
    const bool self_play = flags & SPL_F_SELF_PLAY;
    const int me = self_play ? seat_of(e.meta) : (int)learner[i];
    const int a = action[i];
    const DeckSrc src{decks ? decks + i * SPL_DECK_BYTES : nullptr, seed, game_ids ? game_ids[i] : first_game + (uint64_t)i};
    int rew = 0, bad = 0;
    bool go = true;
    if (a >= 0) {
        Ctx c;
        Choice ch;
        make_ctx(e, sT, c);
        if (seat_of(e.meta) == me && game_ends(e, flags) == SPL_END_NONE && decode_and_check(e, sT, c, a, ch))
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
static int sm_count()
{
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// tile kernels: block size (a multiple of 32, overridable for experiments), dynamic shared memory above 48 KB, residency
static int tile_block(const char *env_name, int dflt)
{
    const char *v = getenv(env_name);
    const int tb = v ? atoi(v) : dflt;
    return (tb >= 32 && tb <= TILE_TB_MAX && tb % 32 == 0) ? tb : dflt;
}
template <class K>
static cudaError_t allow_smem(K kernel, size_t dyn_bytes)
{
    return dyn_bytes + sizeof(Tables) > 48 * 1024 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_bytes)
                                                  : cudaSuccess;
}
static int resident_blocks(size_t dyn_bytes)  // per SM, by shared memory (227 KB usable, 1 KB reserved per block)
{
    const int n = (int)((227 * 1024) / (dyn_bytes + sizeof(Tables) + 1024));
    return n < 1 ? 1 : n;
}

#define DISPATCH_P(P, ...)             \
    switch (P) {                       \
    case 2: { constexpr int PP = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int PP = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int PP = 4; __VA_ARGS__; } break; \
    default: return SPL_E_BADARG;      \
    }

template <int P, int MODE>
static int launch_pool(const PlayoutArgs &a, cudaStream_t stream)
{
    constexpr int NW = SPL_POOL_WARPS;
    using Sm = PoolSmem<P, NW>;
    auto kernel = playout_pool_kernel<P, MODE, NW>;
    static bool configured[16] = {};  // per device: opt in to > 48 KB of dynamic shared memory once
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 16 && !configured[dev]) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm)));
        configured[dev] = true;
    } else if (dev >= 16) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm)));
    }
    const int64_t cap = (int64_t)sm_count() * SPL_POOL_BLOCKS_PER_SM, want = (a.n_games + 31) / 32;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    kernel<<<grid, 32 * NW, sizeof(Sm), stream>>>(a);
    return SPL_OK;
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

// persistent row kernels: as many blocks as the tile lets reside, each warp strides over 32-env groups
static int launch_mask(const void *state, const int8_t *seat, uint32_t *mask, int64_t N, int P, cudaStream_t stream)
{
    static const int tb = tile_block("SPL_MASK_BLOCK", MASK_TB);
    const size_t smem = (size_t)tb * MASK_ROW_BYTES;
    const int64_t want = (N + tb - 1) / tb, cap = (int64_t)sm_count() * resident_blocks(smem);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    DISPATCH_P(P, CK(allow_smem(legal_mask_kernel<PP>, smem));
               legal_mask_kernel<PP><<<grid, tb, smem, stream>>>((const uint4 *)state, seat, mask, N));
    CK(cudaGetLastError());
    return SPL_OK;
}
static int launch_observe(const void *state, const int8_t *seat, float *obs, int64_t N, int P, cudaStream_t stream)
{
    static const int tb = tile_block("SPL_OBS_BLOCK", OBS_TB);
    const size_t smem = (size_t)tb * OBS_ROW_BYTES;
    const int64_t want = (N + tb - 1) / tb, cap = (int64_t)sm_count() * resident_blocks(smem);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    DISPATCH_P(P, CK(allow_smem(observe_kernel<PP>, smem));
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
    DISPATCH_P(P, step_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(
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

static int launch_playout(bool fresh, const PlayoutArgs &args, int P, cudaStream_t stream)
{
    PlayoutArgs a = args;
    philox_round_keys(a.seed, a.rk);
    if (a.max_steps < 1 || a.max_steps > 4096 || a.trace_len < 0) return SPL_E_RANGE;
    if (!a.trace) a.trace_len = 0;  // the kernel tests trace_len only
    if (a.trace && a.trace_len > 0)
        CK(cudaMemsetAsync(a.trace, 0xFF, (size_t)a.trace_len * (size_t)a.n_games * sizeof(int16_t), stream));
    static int impl = -1;
    if (impl < 0) {  // SPL_PLAYOUT_IMPL=thread: the one-game-per-thread kernel (A/B only)
        const char *s = getenv("SPL_PLAYOUT_IMPL");
        impl = (s && !strcmp(s, "thread")) ? 1 : 0;
    }
    if (impl == 0) {
        if (fresh) { DISPATCH_P(P, if (int rc = launch_pool<PP, POOL_FRESH>(a, stream)) return rc); }
        else { DISPATCH_P(P, if (int rc = launch_pool<PP, POOL_STATES>(a, stream)) return rc); }
        CK(cudaGetLastError());
        return SPL_OK;
    }
    static int blk = 0;
    if (!blk) {  // SPL_PLAYOUT_BLOCK: threads per block (32, 64, 96 or 128) - a tuning knob, see profiles/r1_playout_ncu.md
        const char *s = getenv("SPL_PLAYOUT_BLOCK");
        blk = s ? atoi(s) : PLAYOUT_BLOCK;
        if (blk != 32 && blk != 64 && blk != 96 && blk != 128) blk = PLAYOUT_BLOCK;
    }
    const unsigned grid = (unsigned)((a.n_games + blk - 1) / blk);
    if (fresh) { DISPATCH_P(P, playout_kernel<PP, true><<<grid, blk, 0, stream>>>(a)); }
    else { DISPATCH_P(P, playout_kernel<PP, false><<<grid, blk, 0, stream>>>(a)); }
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
    PlayoutArgs a{(uint4 *)final_state, nullptr, seed, first_game, nullptr, num_games, flags, max_steps, score2, steps, end_kind,
                  (unsigned long long *)counters, trace, trace_len};
    return launch_playout(true, a, P, (cudaStream_t)stream);
}

int spl_playout_from(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids, int64_t N,
                     int P, uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind, int64_t *counters, int16_t *trace,
                     int trace_len, void *stream)
{
    if (!state || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    PlayoutArgs a{(uint4 *)state, decks, seed, first_game, game_ids, N, flags, max_steps, score2, steps, end_kind,
                  (unsigned long long *)counters, trace, trace_len};
    return launch_playout(false, a, P, (cudaStream_t)stream);
}

int spl_env_step(void *state, const uint8_t *decks, uint64_t seed, uint64_t first_game, const uint64_t *game_ids,
                 const int8_t *learner, const int16_t *action, int8_t *reward, uint8_t *terminated, uint8_t *illegal, uint32_t *mask, float *obs,
                 int64_t N, int P, uint32_t flags, void *stream)
{
    if (!state || (!learner && !(flags & SPL_F_SELF_PLAY)) || !action || N < 0) return SPL_E_BADARG;
    if (int rc = check_device()) return rc;
    if (N == 0) return SPL_OK;
    const int8_t *seat = (flags & SPL_F_SELF_PLAY) ? nullptr : learner;  // self-play: whoever moves next
    DISPATCH_P(P, env_move_kernel<PP><<<grid_for(N, BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(
                      (uint4 *)state, decks, seed, first_game, game_ids, learner, action, reward, terminated, illegal, N, flags));
    CK(cudaGetLastError());
    if (obs)
        if (int rc = launch_observe(state, seat, obs, N, P, (cudaStream_t)stream)) return rc;
    if (mask)  // the LEARNER's mask (get_legal_actions_mask uses my_turn, splendor_env.py:184)
        if (int rc = launch_mask(state, seat, mask, N, P, (cudaStream_t)stream)) return rc;
    return SPL_OK;
}

// ---- host-buffer path: library-owned, grow-only device staging
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

// One of the two independent pipelines of the host-buffer path: its own staging memory, chunk streams and events, so
// that two calls in flight (spl_playout_host_async on alternating slots) share nothing but the GPU.
constexpr int HOST_SLOTS = 2, MAX_CHUNKS = 8, HOST_CHUNKS_DEFAULT = 4;
struct HostSlot {
    Staging stage;
    cudaStream_t side[MAX_CHUNKS] = {};
    cudaEvent_t ev_start = nullptr, ev_done[MAX_CHUNKS] = {}, ev_call_done = nullptr;
    bool used = false;
    int init()
    {
        if (ev_start) return SPL_OK;
        CK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_call_done, cudaEventDisableTiming));
        for (int c = 0; c < MAX_CHUNKS; c++) {
            CK(cudaStreamCreateWithFlags(&side[c], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
        }
        return SPL_OK;
    }
};
HostSlot g_slots[HOST_SLOTS];

// Everything of a host-buffer playout up to (not including) the wait for its completion: on return the copies and
// kernels are queued, `ctr_out` (host, 16 x int64) and the output arrays are written when `stream` reaches the end of
// the queued work.  Caller holds g_stage_mu.
int playout_host_enqueue(HostSlot &S, const uint8_t *deck_lists, const uint8_t *nobles, uint64_t seed, uint64_t first_game, int64_t N,
                         int P, uint32_t flags, int max_steps, int8_t *score2, int16_t *steps, uint8_t *end_kind, int64_t *ctr_out,
                         cudaStream_t stream)
{
    if (int rc = S.init()) return rc;
    // The batch is cut into chunks that travel on separate internal streams: the H2D copy of chunk k+1 and the D2H copy
    // of chunk k-1 overlap the kernels of chunk k (and the chunk kernels, latency-bound at this size, overlap each other).
    static int want_chunks = 0;
    if (!want_chunks) {  // SPL_HOST_CHUNKS: 1..8, a tuning knob (profiles/r1q_playout.md)
        const char *s = getenv("SPL_HOST_CHUNKS");
        want_chunks = s ? atoi(s) : HOST_CHUNKS_DEFAULT;
        if (want_chunks < 1 || want_chunks > MAX_CHUNKS) want_chunks = HOST_CHUNKS_DEFAULT;
    }
    const int n_chunks = N >= 16384 ? want_chunks : 1;
    // equal chunks: sizes shrinking towards the end (so that the last kernel's longest game is shorter in expectation)
    // measured 1.5 % slower (profiles/r1q_playout.md)
    const int64_t per = ((N + n_chunks - 1) / n_chunks + 127) / 128 * 128;
    const size_t W = (size_t)(8 + 4 * P) * 4;
    const size_t o_state = 0, o_decks = o_state + align256((size_t)N * W) + 256 * MAX_CHUNKS,
                 o_nobles = o_decks + align256((size_t)N * SPL_DECK_BYTES), o_score = o_nobles + align256((size_t)N * 5),
                 o_steps = o_score + align256((size_t)N * P), o_kind = o_steps + align256((size_t)N * 2),
                 o_ctr = o_kind + align256((size_t)N), total = o_ctr + 256;
    if (total > S.stage.cap && S.used) CK(cudaEventSynchronize(S.ev_call_done));  // growing frees the old buffer: let its last user finish
    if (int rc = S.stage.ensure(total)) return rc;
    char *b = (char *)S.stage.p;
    // a slot is reused only after the previous call on it has completed (device-side wait: a caller that alternates
    // slots as documented never blocks here; one that does not is serialised instead of corrupted)
    if (S.used) CK(cudaStreamWaitEvent(stream, S.ev_call_done, 0));
    CK(cudaMemsetAsync(b + o_ctr, 0, 128, stream));
    CK(cudaEventRecord(S.ev_start, stream));
    size_t state_off = o_state;
    for (int c = 0; c < n_chunks; c++) {
        const int64_t lo = (int64_t)c * per, n = (lo + per <= N ? per : N - lo);
        if (n <= 0) break;
        cudaStream_t s = S.side[c];
        CK(cudaStreamWaitEvent(s, S.ev_start, 0));
        char *st = b + state_off;  // each chunk is its own chunk-major state batch of n envs
        state_off += align256((size_t)n * W);
        CK(cudaMemcpyAsync(b + o_decks + lo * SPL_DECK_BYTES, deck_lists + lo * SPL_DECK_BYTES, (size_t)n * SPL_DECK_BYTES, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + o_nobles + lo * 5, nobles + lo * 5, (size_t)n * 5, cudaMemcpyHostToDevice, s));
        if (int rc = spl_inject(st, (const uint8_t *)(b + o_decks + lo * SPL_DECK_BYTES), (const uint8_t *)(b + o_nobles + lo * 5), n, P, s)) return rc;
        if (int rc = spl_playout_from(st, (const uint8_t *)(b + o_decks + lo * SPL_DECK_BYTES), seed, first_game + (uint64_t)lo, nullptr, n, P,
                                      flags, max_steps, (int8_t *)(b + o_score) + lo * P, (int16_t *)(b + o_steps) + lo,
                                      (uint8_t *)(b + o_kind) + lo, (int64_t *)(b + o_ctr), nullptr, 0, s))
            return rc;
        if (score2) CK(cudaMemcpyAsync(score2 + lo * P, b + o_score + lo * P, (size_t)n * P, cudaMemcpyDeviceToHost, s));
        if (steps) CK(cudaMemcpyAsync(steps + lo, b + o_steps + lo * 2, (size_t)n * 2, cudaMemcpyDeviceToHost, s));
        if (end_kind) CK(cudaMemcpyAsync(end_kind + lo, b + o_kind + lo, (size_t)n, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(S.ev_done[c], s));
        CK(cudaStreamWaitEvent(stream, S.ev_done[c], 0));
    }
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
    if (int rc = playout_host_enqueue(g_slots[0], deck_lists, nobles, seed, first_game, N, P, flags, max_steps, score2, steps, end_kind,
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
    return playout_host_enqueue(g_slots[slot], deck_lists, nobles, seed, first_game, N, P, flags, max_steps, score2, steps, end_kind,
                                counters, (cudaStream_t)stream_);
}

}  // extern "C"

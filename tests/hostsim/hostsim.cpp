// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE.  Compiles the DEVICE rule code (splendor-ai_b200/csrc/engine.cuh)
// with g++ by shimming the handful of CUDA intrinsics it uses, so that `pytest -m "not gpu"` can exercise the exact
// source the kernels are built from (count / select / membership / mask lanes / observation lanes / apply / scoring)
// against the CPU oracle without a GPU.  It is NOT a CPU fallback: nothing under splendor-ai_b200/ can load it, and
// the kernels' glue (launch geometry, warp synchronisation, coalesced stores) is only covered by the `-m gpu` tests.
#define SPL_HOSTSIM 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)  // PTX prmt.b32, default mode
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 15u;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7u))) & 255u;
        if (sel & 8u) byte = (byte & 0x80u) ? 255u : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __uint_as_float(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }
using std::max;
using std::min;

#include "../../splendor-ai_b200/csrc/engine.cuh"

using namespace spl;
static const Tables T = make_tables();
static const DecodeTable DEC = make_decode_table(T);  // the step kernels' table-driven decode
static const uint32_t LOG1P[32] = SPL_LOG1P_BITS;

#define DISPATCH(P, ...)                                   \
    switch (P) {                                           \
    case 2: { constexpr int PP = 2; __VA_ARGS__; } break;  \
    case 3: { constexpr int PP = 3; __VA_ARGS__; } break;  \
    case 4: { constexpr int PP = 4; __VA_ARGS__; } break;  \
    default: return -1;                                    \
    }

template <int P> static void ld(Env<P> &e, const uint32_t *w) { load_env(e, reinterpret_cast<const uint4 *>(w), 1, 0); }
template <int P> static void st(const Env<P> &e, uint32_t *w) { store_env(e, reinterpret_cast<uint4 *>(w), 1, 0); }

template <int P> static int t_reset(uint32_t *w, uint64_t seed, uint64_t game, uint8_t *lists, uint8_t *nobles)
{
    Env<P> e;
    reset_env(e, seed, game);
    st(e, w);
    if (nobles) for (int k = 0; k < 5; k++) { uint32_t id = (e.meta >> (4 * k)) & 15u; nobles[k] = id == 15u ? 0xFF : (uint8_t)id; }
    if (lists) {
        const DeckSrc src{nullptr, seed, game};
        for (int t = 0; t < 3; t++) {
            const int size = t == 0 ? 40 : (t == 1 ? 30 : 20), base = t == 0 ? 0 : (t == 1 ? 40 : 70);
            for (int c = 0; c < 4; c++) lists[base + size - 1 - c] = (uint8_t)byte_of(e.d[t], c);
            for (int d = 4; d < size; d++) lists[base + size - 1 - d] = (uint8_t)deal(e, t, src);
        }
    }
    return 0;
}
template <int P> static int t_inject(uint32_t *w, const uint8_t *lists, const uint8_t *nobles)
{
    Env<P> e;
    const uint32_t n = P == 2 ? 4 : (P == 3 ? 5 : 7);
    e.bank = n * (ONES6 & ~(1u << YSH)) | (5u << YSH);
    e.k[0] = 40u | 30u << 8 | 20u << 16; e.k[1] = 0; e.k[2] = 0;
    uint32_t list = 0xFFFFFu;
    for (int k = 0; k < P + 1; k++) list = (list & ~(15u << (4 * k))) | ((uint32_t)(nobles[k] & 15u) << (4 * k));
    e.meta = list | ((uint32_t)(P - 2) << 22) | META_EXPLICIT;
    const DeckSrc src{lists, 0, 0};
    for (int t = 0; t < 3; t++) {
        uint32_t word = 0;
        for (int c = 0; c < 4; c++) word |= deal(e, t, src) << (8 * c);
        e.d[t] = word;
    }
    for (int p = 0; p < P; p++) { e.pg[p] = 0; e.pc[p] = 0; e.pr[p] = 0x00FFFFFFu; e.pi[p] = 0; }
    st(e, w);
    return 0;
}
template <int P> static int t_legal_select(const uint32_t *w, uint16_t *out)
{
    Env<P> e; ld(e, w);
    Ctx c; FamilyCounts f; Choice ch;
    make_ctx(e, T, c);
    const int total = count_legal(e, T, c, f), n = total ? total : c.nS;
    for (int k = 0; k < n; k++) { select_legal(e, T, c, f, total, k, ch); out[k] = (uint16_t)ch.idx; }
    return n;
}
template <int P> static int t_legal_member(const uint32_t *w, uint16_t *out)
{
    Env<P> e; ld(e, w);
    Ctx c; Choice ch;
    make_ctx(e, T, c);
    int n = 0;
    for (int idx = 0; idx < SPL_NUM_ACTIONS; idx++) {
        // both membership forms - index arithmetic and the table row - must agree on the verdict and on every decoded field
        Choice ch2;
        const bool ok = decode_and_check(e, T, c, idx, ch), ok2 = decode_and_check_row(e, T, c, idx, DEC.row[idx], ch2, T);
        if (ok != ok2 || (ok && (ch.type != ch2.type || ch.slot != ch2.slot || ch.n != ch2.n || ch.C != ch2.C || ch.R != ch2.R ||
                                 ch.info != ch2.info || ch.idx != ch2.idx))) {
            std::fprintf(stderr, "hostsim: decode forms disagree at index %d\n", idx);
            std::abort();
        }
        if (ok) out[n++] = (uint16_t)idx;
    }
    return n;
}
template <int P> static int t_mask(const uint32_t *w, int seat, uint32_t *mask)
{
    Env<P> e; ld(e, w);
    if (seat >= 0) e.meta = (e.meta & ~(3u << 20)) | ((uint32_t)seat << 20);
    Ctx c; make_ctx(e, T, c);
    uint32_t m[MASK_REGS] = {0};
    thread_fill_mask(e, T, c, m);
    std::memcpy(mask, m, SPL_MASK_WORDS * 4);
    return 0;
}
template <int P> static int t_step(uint32_t *w, const uint8_t *lists, uint64_t seed, uint64_t game, int idx, uint32_t flags,
                                   int *reward, int *done)
{
    Env<P> e; ld(e, w);
    Ctx c; Choice ch;
    make_ctx(e, T, c);
    int bad = 0;
    *reward = 0;
    const uint4 row = DEC.row[std::min(std::max(idx, 0), SPL_NUM_ACTIONS - 1)];  // as step_kernel does
    if (decode_and_check_row(e, T, c, idx, row, ch, T)) { const DeckSrc src{lists, seed, game}; *reward = apply_action(e, ch, src); }
    else { bad = 1; e.meta |= META_ILLEGAL; }
    st(e, w);
    *done = game_ends(e, flags);
    return bad;
}
template <int P> static int t_undo(uint32_t *w, int seat, int idx, int card, int noble, const uint8_t *returned)
{
    Env<P> e; ld(e, w);
    uint32_t r = 0;
    for (int c = 0; c < 6; c++) r |= (uint32_t)returned[c] << (5 * field_of_colour_id(c));
    undo_action(e, T, seat, idx, (uint32_t)card, (uint32_t)noble, r);
    st(e, w);
    return 0;
}
template <int P> static int t_observe(const uint32_t *w, int seat, float *o)
{
    Env<P> e; ld(e, w);
    uint32_t row[OBS_WORDS + 1] = {0};
    thread_fill_obs(e, T, seat, row);
    for (int w = 0; w < SPL_OBS_SIZE; w++) o[w] = obs_entry(reinterpret_cast<const uint8_t *>(row), w, LOG1P);
    return 0;
}
template <int P> static int t_scores(const uint32_t *w, uint32_t flags, int *score2, int *outcome, int *kind)
{
    Env<P> e; ld(e, w);
    int sc[P], oc[P];
    final_scores(e, sc, oc);
    for (int p = 0; p < P; p++) { score2[p] = sc[p]; outcome[p] = oc[p]; }
    *kind = game_ends(e, flags);
    return 0;
}
// the body of playout_kernel for one game
template <int P> static int t_playout(uint32_t *w, bool fresh, const uint8_t *lists, uint64_t seed, uint64_t game, uint32_t flags,
                                      int max_steps, int16_t *trace, int trace_len, int *score2, int *kind_out)
{
    Env<P> e;
    DeckSrc src{lists, seed, game};
    uint8_t drawn[92];  // fresh games: the kernel's form - the whole deck order drawn at reset, in-game deals pop the lists
    uint32_t rk[20];
    philox_round_keys(seed, rk);
    if (fresh) { reset_env_lists(e, seed, game, rk, drawn); src.lists = drawn; src.rk = rk; } else ld(e, w);
    int t = 0;
    for (int p = 0; p < P; p++) t += e.pi[p] & 0xFFFFu;
    const int t_end = t + max_steps;
    int kind;
    for (;;) {
        kind = game_ends(e, flags);
        if (kind == SPL_END_NONE && t >= t_end) kind = SPL_END_TRUNCATED;
        if (kind != SPL_END_NONE) break;
        Ctx c; FamilyCounts f; Choice ch;
        make_ctx(e, T, c);
        const int total = count_legal(e, T, c, f);
        const uint32_t r = word_of(philox_rk(rk, game, (uint32_t)t >> 2, STREAM_ACTION), (uint32_t)t & 3u);
        select_legal(e, T, c, f, total, (int)__umulhi(r, (uint32_t)(total ? total : c.nS)), ch);
        if (trace && t < trace_len) trace[t] = (int16_t)ch.idx;
        if (fresh) apply_action<P, true>(e, ch, src); else apply_action(e, ch, src);
        t++;
    }
    int oc[P], sc[P];
    final_scores(e, sc, oc);
    for (int p = 0; p < P; p++) score2[p] = sc[p];
    *kind_out = kind;
    if (fresh) lists_to_masks(e, drawn);
    st(e, w);
    return t;
}

extern "C" {
int sim_reset(uint32_t *w, uint64_t seed, uint64_t game, int P, uint8_t *lists, uint8_t *nobles) { DISPATCH(P, return t_reset<PP>(w, seed, game, lists, nobles)); return -1; }
int sim_inject(uint32_t *w, const uint8_t *lists, const uint8_t *nobles, int P) { DISPATCH(P, return t_inject<PP>(w, lists, nobles)); return -1; }
int sim_legal_select(const uint32_t *w, int P, uint16_t *out) { DISPATCH(P, return t_legal_select<PP>(w, out)); return -1; }
int sim_legal_member(const uint32_t *w, int P, uint16_t *out) { DISPATCH(P, return t_legal_member<PP>(w, out)); return -1; }
int sim_mask(const uint32_t *w, int P, int seat, uint32_t *mask) { DISPATCH(P, return t_mask<PP>(w, seat, mask)); return -1; }
int sim_step(uint32_t *w, const uint8_t *lists, uint64_t seed, uint64_t game, int P, int idx, uint32_t flags, int *reward, int *done)
{ DISPATCH(P, return t_step<PP>(w, lists, seed, game, idx, flags, reward, done)); return -1; }
int sim_undo(uint32_t *w, int P, int seat, int idx, int card, int noble, const uint8_t *returned) { DISPATCH(P, return t_undo<PP>(w, seat, idx, card, noble, returned)); return -1; }
int sim_observe(const uint32_t *w, int P, int seat, float *o) { DISPATCH(P, return t_observe<PP>(w, seat, o)); return -1; }
int sim_scores(const uint32_t *w, int P, uint32_t flags, int *score2, int *outcome, int *kind) { DISPATCH(P, return t_scores<PP>(w, flags, score2, outcome, kind)); return -1; }
int sim_playout(uint32_t *w, int fresh, const uint8_t *lists, uint64_t seed, uint64_t game, int P, uint32_t flags, int max_steps,
                int16_t *trace, int trace_len, int *score2, int *kind)
{ DISPATCH(P, return t_playout<PP>(w, fresh != 0, lists, seed, game, flags, max_steps, trace, trace_len, score2, kind)); return -1; }
void sim_philox(uint64_t seed, uint64_t game, uint32_t index, uint32_t stream, uint32_t *out)
{ uint4 r = philox(seed, game, index, stream); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
void sim_philox_rk(uint64_t seed, uint64_t game, uint32_t index, uint32_t stream, uint32_t *out)
{ uint32_t rk[20]; philox_round_keys(seed, rk); uint4 r = philox_rk(rk, game, index, stream); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
int sim_tables_bytes(void) { return (int)sizeof(Tables); }
int sim_small_div_mismatches(void)  // the whole documented domain of small_div against integer division
{
    int bad = 0;
    for (int d = 1; d <= 64; d++)
        for (int k = 0; k < 8192; k++) bad += small_div(T, k, d) != k / d;
    return bad;
}
}

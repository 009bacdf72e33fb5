// engine.cuh -- device-side Splendor rules for sm_100a: packed state words, SWAR helpers, legal-action
// count / select / membership, action application and its inverse, terminal test and scoring, and the per-thread
// builders of one legal-mask row and one observation row.
//
// Rule semantics follow the reference (paths relative to /root/reference/src/splendor/):
//   getLegalActions      splendor/splendor_model.py:404-576   generate_return_combos :329-367
//   resources_sufficient splendor/splendor_model.py:371-394   noble_visit            :397-402
//   generateSuccessor    splendor/splendor_model.py:222-295   GameRule.update template.py:47-53
//   gameEnds / calScore  splendor/splendor_model.py:299-323   LimitRoundsGameRule splendor/utils.py:14-25
//   generatePredecessor splendor/splendor_model.py:156-220  features          splendor/features.py:470-480
//   ALL_ACTIONS order    splendor/gym/envs/actions.py:109-237
// but the formulation is different on purpose: the reference materialises a list of dicts; here every count is a
// closed form over a handful of 32-bit words, so that one thread can play one game entirely out of registers.
//
// Word format ("g6"): six 5-bit fields in the reference's COLOURS order (splendor_utils.py:124-131)
//   field 0 black, 1 red, 2 yellow, 3 green, 4 blue, 5 white; every value is <= 15, bit 4 of each field is a guard
//   bit, which makes per-field saturating subtraction and >= tests a three-instruction SWAR operation.
#pragma once
#include <stdint.h>
#ifndef SPL_HOSTSIM  // tests/hostsim compiles this header with g++ and its own shims (test only)
#include <cuda_runtime.h>
#endif

#include "../../include/spl_tables.h"
#include "../../include/splendor_b200.h"

// SPL_CHECK: bounds / invariant assertions, compiled in only with -DSPL_DEBUG (compute-sanitizer is not available on
// the GPU pool, so index safety is checked with a debug build run over tools/touch_all_kernels.py and the parity tests)
#if defined(SPL_DEBUG) && !defined(SPL_HOSTSIM)
#include <assert.h>
#define SPL_CHECK(cond) assert(cond)
#elif defined(SPL_DEBUG)
#include <cassert>
#define SPL_CHECK(cond) assert(cond)
#else
#define SPL_CHECK(cond) ((void)0)
#endif

namespace spl {

constexpr uint32_t ONES6 = 0x02108421u;        // 1 in each field
constexpr uint32_t H6 = ONES6 << 4;            // guard bit of each field
constexpr uint32_t NORMAL_FIELDS = 0x3Bu;      // fields 0,1,3,4,5 (everything but yellow)
constexpr int YSH = 10;                        // bit offset of the yellow field
constexpr uint32_t NONE8 = 0xFFu;

__host__ __device__ constexpr int field_of_normal(int c) { return c + (c >= 2 ? 1 : 0); }  // NORMAL_COLORS idx -> field
__host__ __device__ constexpr int field_of_colour_id(int c) { return c == 5 ? 2 : (c >= 2 ? c + 1 : c); }

// ------------------------------------------------------------------------------------------------ static tables
enum ActType { T_PASS = 0, T_RESERVE = 1, T_SAME = 2, T_DIFF = 3, T_BUY_RESERVE = 4, T_BUY_AVAILABLE = 5 };
constexpr int BASE_RESERVE = 6, BASE_SAME = 510, BASE_DIFF = 1140, BASE_BUY_RESERVE = 3420, BASE_BUY_AVAILABLE = 3438;

struct alignas(16) Tables {
    // Four parts, each starting on a 16-byte boundary, ordered so that a kernel stages only a PREFIX (plus, for the observation
    // kernel, the last part): [card, noble] everyone | decode / apply (step) | legality patterns (mask) | counting and
    // selecting (playouts, expansion) | observation card vectors.  kernels.cu: TABLES_*_BYTES.
    uint2 card[128];         // .x = cost (g6, yellow field 0) ; .y = 5*colour_field | points << 5 | tier << 8;
                             // entries 90..127 (an empty slot's 0xFF & 127 lands on 127): colour shift = yellow, never buyable
    uint32_t noble[16];      // requirement (g6); entry 15 = "no noble" (unsatisfiable)
    alignas(16) uint32_t combo_c[25];  // collect_diff combo ci: collected gems (g6).  ci 0..4 L=1, 5..14 L=2, 15..24 L=3
    uint16_t combo_base[25]; // ALL_ACTIONS index of (combo ci, noble slot 0, variant 0)
    uint16_t combo_off[25];  // first entry of the combo's variants in ret_diff
    uint8_t combo_mask[25];  // 6-bit field mask of the combo's colours
    uint8_t pad[3];
    uint32_t ret_diff[380];  // returned gems (g6) of variant v: 6 / 15 / 20 variants per combo of length 1 / 2 / 3
    uint32_t ret_same[105];  // [c*21 + v], c = NORMAL_COLORS index
    // feasibility LUTs: which return variants of a group the agent can pay, from its held-colour masks (no scanning)
    alignas(16) uint8_t pext_diff[25][64];  // 6-bit colour-field mask -> bits over the combo's OTHER colours, in variant order
    uint8_t pext_same[5][64];   // same for collect_same colour c (5 other colours)
    uint16_t pairs5[32];        // held>=1 bits over 5 others -> combinations(others, 2) bits (collect_same variants 1..10)
    uint16_t cwr2_pair[2][16];  // [k==4][x]: k others held>=1 -> bits of the i<j entries of combinations_with_replacement(k,2)
    uint16_t cwr2_dbl[2][16];   // [k==4][y]: held>=2 -> bits of the (i,i) entries
    uint16_t cwr3[512];         // k=3: x | y<<3 | z<<6 (held >=1/>=2/>=3) -> bits of combinations_with_replacement(3,3)
    // return-multiset generating functions (see collect_groups below): (1 + x + .. + x^lvl)^n and 1 / (1 + x + .. + x^lvl)
    // as base-256 digit strings modulo 2^32
    alignas(16) uint32_t poly_pow[3][8];  // [lvl-1][n], n = 0..6
    uint32_t poly_inv[16];      // [gems held of the colour] (level = min(held, 3))
    uint32_t recip[68];         // ceil(2^31 / d), d = 1..64 (small_div)
    uint2 buyrow[256];          // card[] for count_legal's affordability test: guard bits already set in the cost (.x | H6),
                                // and indexed by the raw slot byte (0xFF = empty slot -> filler), so no masking of the id
    uint8_t kth[256][8];        // position of the j-th (0-based) set bit of a byte (select_bit_lut)
    uint4 cardvec[128];         // the observation's 13-byte card vector (vectorize_card, features.py:377-413) as bytes 0..12:
                                // one-hot over black, blue, green, red, white, yellow | cost black, blue, green, red, white |
                                // deck_id | points; all zero for the filler entries (an empty slot contributes zeros)
};

namespace detail {
struct CardRow { int colour, points, c[5]; };
struct NobleRow { int c[5]; };
constexpr CardRow CARD_ROWS[SPL_NUM_CARDS] = {
#define X(col, pts, c0, c1, c2, c3, c4) {col, pts, {c0, c1, c2, c3, c4}},
    SPL_CARD_ROWS
#undef X
};
constexpr NobleRow NOBLE_ROWS[SPL_NUM_NOBLES] = {
#define X(c0, c1, c2, c3, c4) {{c0, c1, c2, c3, c4}},
    SPL_NOBLE_ROWS
#undef X
};
__host__ __device__ constexpr uint32_t g6_from_normal(const int *c)
{
    uint32_t w = 0;
    for (int i = 0; i < 5; i++) w |= (uint32_t)c[i] << (5 * field_of_normal(i));
    return w;
}
}  // namespace detail

// itertools.combinations / combinations_with_replacement orders of gym/envs/actions.py:145-199, unrolled at compile time
__host__ __device__ constexpr Tables make_tables()
{
    Tables T{};
    for (int i = 0; i < 128; i++) T.card[i] = uint2{0u, (uint32_t)YSH};  // filler: a free card of "colour yellow", which count_legal treats as capped
    for (int i = 0; i < SPL_NUM_CARDS; i++) {
        const detail::CardRow &r = detail::CARD_ROWS[i];
        int tier = i < 40 ? 0 : (i < 70 ? 1 : 2);
        T.card[i].x = detail::g6_from_normal(r.c);
        T.card[i].y = (uint32_t)(5 * field_of_colour_id(r.colour)) | ((uint32_t)r.points << 5) | ((uint32_t)tier << 8);
    }
    for (int i = 0; i < 16; i++) T.noble[i] = 0x1EF7BDEFu & ~(31u << YSH);  // 15 of everything: never satisfied
    for (int i = 0; i < SPL_NUM_NOBLES; i++) T.noble[i] = detail::g6_from_normal(detail::NOBLE_ROWS[i].c);

    // collect_same variants (actions.py:145-174): none | pairs of the 5 other colours | one of | two of
    for (int c = 0; c < 5; c++) {
        int others[5] = {0, 0, 0, 0, 0}, k = 0;
        for (int f = 0; f < 6; f++)
            if (f != field_of_normal(c)) others[k++] = f;
        uint32_t *row = T.ret_same + c * 21;
        int v = 0;
        row[v++] = 0;
        for (int i = 0; i < 5; i++)
            for (int j = i + 1; j < 5; j++) row[v++] = (1u << (5 * others[i])) + (1u << (5 * others[j]));
        for (int i = 0; i < 5; i++) row[v++] = 1u << (5 * others[i]);
        for (int i = 0; i < 5; i++) row[v++] = 2u << (5 * others[i]);
    }
    // collect_diff (actions.py:177-199)
    int ci = 0, off = 0, base = BASE_DIFF;
    for (int L = 1; L <= 3; L++) {
        const int V = L == 1 ? 6 : (L == 2 ? 15 : 20);
        for (int a = 0; a < 5; a++)
            for (int b = (L >= 2 ? a + 1 : 4); b < 5; b++)
                for (int c = (L >= 3 ? b + 1 : 4); c < 5; c++) {
                    uint32_t fmask = 1u << field_of_normal(a);
                    if (L >= 2) fmask |= 1u << field_of_normal(b);
                    if (L >= 3) fmask |= 1u << field_of_normal(c);
                    uint32_t cw = 0;
                    int others[6] = {0, 0, 0, 0, 0, 0}, k = 0;
                    for (int f = 0; f < 6; f++) {
                        if (fmask >> f & 1) cw |= 1u << (5 * f);
                        else others[k++] = f;
                    }
                    T.combo_c[ci] = cw;
                    T.combo_mask[ci] = (uint8_t)fmask;
                    T.combo_base[ci] = (uint16_t)base;
                    T.combo_off[ci] = (uint16_t)off;
                    int v = 0;
                    T.ret_diff[off + v++] = 0;
                    for (int r = 1; r <= L; r++)
                        for (int x = 0; x < k; x++)
                            for (int y = (r >= 2 ? x : k - 1); y < k; y++)
                                for (int z = (r >= 3 ? y : k - 1); z < k; z++) {
                                    uint32_t w = 1u << (5 * others[x]);
                                    if (r >= 2) w += 1u << (5 * others[y]);
                                    if (r >= 3) w += 1u << (5 * others[z]);
                                    T.ret_diff[off + v++] = w;
                                }
                    ci++;
                    off += V;
                    base += 6 * V;
                }
    }
    // ---- feasibility LUTs
    for (int ci2 = 0; ci2 < 25; ci2++)
        for (int h = 0; h < 64; h++) {
            int k = 0, out = 0;
            for (int f = 0; f < 6; f++) {
                if (T.combo_mask[ci2] >> f & 1) continue;
                if (h >> f & 1) out |= 1 << k;
                k++;
            }
            T.pext_diff[ci2][h] = (uint8_t)out;
        }
    for (int c = 0; c < 5; c++)
        for (int h = 0; h < 64; h++) {
            int k = 0, out = 0;
            for (int f = 0; f < 6; f++) {
                if (f == field_of_normal(c)) continue;
                if (h >> f & 1) out |= 1 << k;
                k++;
            }
            T.pext_same[c][h] = (uint8_t)out;
        }
    for (int x = 0; x < 32; x++) {
        int bits = 0, r = 0;
        for (int i = 0; i < 5; i++)
            for (int j = i + 1; j < 5; j++, r++)
                if ((x >> i & 1) && (x >> j & 1)) bits |= 1 << r;
        T.pairs5[x] = (uint16_t)bits;
    }
    for (int kk = 3; kk <= 4; kk++)
        for (int x = 0; x < 16; x++) {
            int pb = 0, db = 0, r = 0;
            for (int i = 0; i < kk; i++)
                for (int j = i; j < kk; j++, r++) {
                    if (i == j) { if (x >> i & 1) db |= 1 << r; }
                    else if ((x >> i & 1) && (x >> j & 1)) pb |= 1 << r;
                }
            T.cwr2_pair[kk - 3][x] = (uint16_t)pb;
            T.cwr2_dbl[kk - 3][x] = (uint16_t)db;
        }
    for (int v = 0; v < 512; v++) {
        const int x = v & 7, y = (v >> 3) & 7, z = v >> 6;
        int bits = 0, r = 0;
        for (int i = 0; i < 3; i++)
            for (int j = i; j < 3; j++)
                for (int l = j; l < 3; l++, r++) {
                    int need[3] = {0, 0, 0};
                    need[i]++; need[j]++; need[l]++;
                    bool ok = true;
                    for (int q = 0; q < 3; q++) {
                        const int have = (z >> q & 1) ? 3 : ((y >> q & 1) ? 2 : ((x >> q & 1) ? 1 : 0));
                        if (need[q] > have) ok = false;
                    }
                    if (ok) bits |= 1 << r;
                }
        T.cwr3[v] = (uint16_t)bits;
    }
    const uint32_t unit[4] = {0x00000001u, 0x00000101u, 0x00010101u, 0x01010101u};
    for (int l = 1; l <= 3; l++) {
        uint32_t p = 1;
        for (int n = 0; n < 8; n++) { T.poly_pow[l - 1][n] = p; p *= unit[l]; }
    }
    for (int h = 0; h < 16; h++) {  // Newton iteration for the inverse of an odd number modulo 2^32
        const uint32_t u = unit[h < 3 ? h : 3];
        uint32_t x = u;
        for (int it = 0; it < 5; it++) x *= 2u - u * x;
        T.poly_inv[h] = x;
    }
    for (int i = 0; i < 256; i++) T.buyrow[i] = uint2{T.card[i & 127].x | H6, T.card[i & 127].y & 31u};  // .y: the colour's bit offset only
    for (int b = 0; b < 256; b++) {
        int j = 0;
        for (int k = 0; k < 8; k++) T.kth[b][k] = 0;
        for (int pos = 0; pos < 8; pos++)
            if (b >> pos & 1) T.kth[b][j++] = (uint8_t)pos;
    }
    for (int i = 0; i < 128; i++) T.cardvec[i] = uint4{0u, 0u, 0u, 0u};
    for (int i = 0; i < SPL_NUM_CARDS; i++) {
        const detail::CardRow &r = detail::CARD_ROWS[i];
        const int alpha = r.colour == 0 ? 0 : (r.colour == 1 ? 3 : (r.colour == 2 ? 2 : (r.colour == 3 ? 1 : 4)));  // colour id -> alphabetical rank
        const uint32_t tier = i < 40 ? 0u : (i < 70 ? 1u : 2u);
        uint8_t b[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        b[alpha] = 1;
        b[6] = (uint8_t)r.c[0]; b[7] = (uint8_t)r.c[3]; b[8] = (uint8_t)r.c[2]; b[9] = (uint8_t)r.c[1]; b[10] = (uint8_t)r.c[4];
        b[11] = (uint8_t)tier; b[12] = (uint8_t)r.points;
        T.cardvec[i] = uint4{(uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24,
                             (uint32_t)b[4] | (uint32_t)b[5] << 8 | (uint32_t)b[6] << 16 | (uint32_t)b[7] << 24,
                             (uint32_t)b[8] | (uint32_t)b[9] << 8 | (uint32_t)b[10] << 16 | (uint32_t)b[11] << 24, (uint32_t)b[12]};
    }
    for (int d = 1; d <= 64; d++) T.recip[d] = (uint32_t)(((1ull << 31) + (uint64_t)d - 1) / (uint64_t)d);
    return T;
}

// ------------------------------------------------------------------------------------------------ SWAR helpers
__device__ __forceinline__ uint32_t satsub6(uint32_t a, uint32_t b)  // per field max(a-b,0); fields < 16
{
    uint32_t d = (a | H6) - b;
    uint32_t m = d & H6;
    m -= m >> 4;
    return d & m;
}
__device__ __forceinline__ uint32_t satsub6g(uint32_t ag, uint32_t b)  // satsub6 for an `a` that already carries the guard bits
{
    uint32_t d = ag - b;
    uint32_t m = d & H6;
    m -= m >> 4;
    return d & m;
}
__device__ __forceinline__ uint32_t hsum6(uint32_t x)  // sum of the six fields (each <= 15)
{
    x = (x & 0x7FFFu) + (x >> 15);
    return (x & 31u) + ((x >> 5) & 31u) + (x >> 10);
}
__device__ __forceinline__ uint32_t hsum6_small(uint32_t x)  // same, valid when the total is <= 31
{
    return ((x * ONES6) >> 25) & 31u;
}
__device__ __forceinline__ uint32_t compress6(uint32_t h)  // guard bits (subset of H6) -> 6-bit field mask
{
    return ((((h >> 4) & 0x00108421u) * 0x00111110u) >> 20 & 31u) | ((h >> 24) & 32u);
}
__device__ __forceinline__ uint32_t ge_mask(uint32_t x, uint32_t t)  // fields with value >= t (1 <= t <= 15)
{
    return compress6((x + (16u - t) * ONES6) & H6);
}
__device__ __forceinline__ uint32_t field(uint32_t w, int sh) { return (w >> sh) & 31u; }
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xFFu; }
__device__ __forceinline__ uint32_t set_byte(uint32_t w, int i, uint32_t v)
{
    return (w & ~(0xFFu << (8 * i))) | (v << (8 * i));
}
__device__ __forceinline__ int select_bit(uint32_t m, int j)  // position of the j-th (0-based) set bit; m has > j bits
{
    int pos = 0, c;
    c = __popc(m & 0xFFFFu); if (j >= c) { j -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (j >= c) { j -= c; pos += 8;  m >>= 8; }
    c = __popc(m & 0xFu);    if (j >= c) { j -= c; pos += 4;  m >>= 4; }
    c = __popc(m & 0x3u);    if (j >= c) { j -= c; pos += 2;  m >>= 2; }
    if (j >= (int)(m & 1u)) pos += 1;
    return pos;
}
__device__ __forceinline__ int select_bit8(uint32_t m, int j)  // same for masks of <= 8 bits
{
    int pos = 0, c;
    c = __popc(m & 0xFu); if (j >= c) { j -= c; pos += 4; m >>= 4; }
    c = __popc(m & 0x3u); if (j >= c) { j -= c; pos += 2; m >>= 2; }
    if (j >= (int)(m & 1u)) pos += 1;
    return pos;
}
// The same two with a table of the j-th set bit of a BYTE: SWAR byte popcounts, their running sums by one multiply, the byte
// that holds the j-th bit by a guard-bit compare (as in the group search of select_family), then one lookup - ~21
// instructions and ONE POPC (a quarter-rate instruction) instead of ~30 and five; 3 instead of 13 for masks of <= 8 bits.
__device__ __forceinline__ int select_bit8_lut(const Tables &T, uint32_t m, int j)
{
    SPL_CHECK(m < 256u && j >= 0 && j < __popc(m));
    return (int)T.kth[m][j];
}
__device__ __forceinline__ int select_bit_lut(const Tables &T, uint32_t m, int j)
{
    SPL_CHECK(j >= 0 && j < __popc(m));
    uint32_t b = m - ((m >> 1) & 0x55555555u);
    b = (b & 0x33333333u) + ((b >> 2) & 0x33333333u);
    b = (b + (b >> 4)) & 0x0F0F0F0Fu;                  // bits set in each byte
    const uint32_t pre = b * 0x01010101u;              // byte k: bits set in bytes 0 .. k
    const uint32_t ahead = ((((uint32_t)j * 0x01010101u) | 0x80808080u) - pre) & 0x80808080u;  // bytes that end before bit j
    const int bi = __popc(ahead);
    const uint32_t before = ((pre << 8) >> (8 * bi)) & 0xFFu;
    return 8 * bi + (int)T.kth[(m >> (8 * bi)) & 0xFFu][(uint32_t)j - before];
}
// floor(k/d) for 0 <= k < 8192, 1 <= d <= 64: ((2k+1) * ceil(2^31/d)) >> 32.  (2k+1)/(2d) is never an integer and lies at
// least 1/(2d) below the next one, while the rounded-up reciprocal adds less than 2^14 / 2^32: exact (checked
// exhaustively in tests/test_device_rules_hostsim.py).
__device__ __forceinline__ int small_div(const Tables &T, int k, int d)
{
    SPL_CHECK(k >= 0 && k < 8192 && d >= 1 && d <= 64);
    return (int)__umulhi(2u * (uint32_t)k + 1u, T.recip[d]);
}
__device__ __forceinline__ uint32_t occupied4(uint32_t dealt_word)  // 4-bit mask of non-empty slots (ids < 128, empty 0xFF)
{
    return ((((~dealt_word) & 0x80808080u) >> 7) * 0x00204081u) >> 21 & 15u;
}

// Philox4x32-10 (Salmon et al. 2011).  counter = (game lo, game hi, index, stream), key = seed.
__device__ __forceinline__ uint4 philox_inline(uint64_t seed, uint64_t game, uint32_t index, uint32_t stream)
{
    uint32_t c0 = (uint32_t)game, c1 = (uint32_t)(game >> 32), c2 = index, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// Out-of-line copy for the rare call sites (every deal, the reset): inlining all of them made the playout kernel
// overflow the instruction cache (profiles/r1a: no_instruction stalls).  The per-step action draw uses the inline
// form so that its 10-round dependent chain overlaps the branch-free counting code.
__device__ __noinline__ uint4 philox(uint64_t seed, uint64_t game, uint32_t index, uint32_t stream)
{
    return philox_inline(seed, game, index, stream);
}
// Same generator with the key schedule precomputed (rk[2r] = key.x + r * 0x9E3779B9, rk[2r+1] = key.y + r * 0xBB67AE85).
// The playout kernel passes the 20 words as kernel parameters, so each round is two wide multiplies and two
// three-input XORs with a constant-bank operand: no key arithmetic, no key registers.
__host__ __device__ inline void philox_round_keys(uint64_t seed, uint32_t *rk)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
__device__ __forceinline__ uint4 philox_rk(const uint32_t *__restrict__ rk, uint64_t game, uint32_t index, uint32_t stream)
{
    uint32_t c0 = (uint32_t)game, c1 = (uint32_t)(game >> 32), c2 = index, c3 = stream;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r]; c1 = (uint32_t)p1;
        c2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1]; c3 = (uint32_t)p0;
    }
    return make_uint4(c0, c1, c2, c3);
}
enum Stream { STREAM_ACTION = 0, STREAM_DECK0 = 1, STREAM_NOBLES = 4, STREAM_ENV = 5 };
// The playouts' action draw for move t of a game: word t & 3 of the Philox block (seed, game, t >> 2, STREAM_ACTION) - one
// block serves four consecutive moves (the env kernels' in-kernel opponents draw word 0 of block (seed, game, total turns)).
__host__ __device__ inline uint32_t word_of(const uint4 &v, uint32_t i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// ------------------------------------------------------------------------------------------------ state in registers
constexpr uint32_t META_EXPLICIT = 1u << 24, META_ILLEGAL = 1u << 26;

template <int P>
struct Env {
    uint32_t bank, d[3], k[3], meta;          // board words 0..7
    uint32_t pg[P], pc[P], pr[P], pi[P];      // per player: gems, bought cards, reserved ids + score, turns|nobles|passed
};

template <int P>
__device__ __forceinline__ void load_env(Env<P> &e, const uint4 *__restrict__ st, int64_t N, int64_t i)
{
    uint4 a = st[i], b = st[N + i];
    e.bank = a.x; e.d[0] = a.y; e.d[1] = a.z; e.d[2] = a.w;
    e.k[0] = b.x; e.k[1] = b.y; e.k[2] = b.z; e.meta = b.w;
#pragma unroll
    for (int p = 0; p < P; p++) {
        uint4 q = st[(int64_t)(2 + p) * N + i];
        e.pg[p] = q.x; e.pc[p] = q.y; e.pr[p] = q.z; e.pi[p] = q.w;
    }
}
template <int P>
__device__ __forceinline__ void store_env(const Env<P> &e, uint4 *__restrict__ st, int64_t N, int64_t i)
{
    st[i] = make_uint4(e.bank, e.d[0], e.d[1], e.d[2]);
    st[N + i] = make_uint4(e.k[0], e.k[1], e.k[2], e.meta);
#pragma unroll
    for (int p = 0; p < P; p++) st[(int64_t)(2 + p) * N + i] = make_uint4(e.pg[p], e.pc[p], e.pr[p], e.pi[p]);
}
__device__ __forceinline__ int seat_of(uint32_t meta) { return (meta >> 20) & 3; }

// per-game context a deal needs
struct DeckSrc {
    const uint8_t *lists;  // explicit deck lists of THIS env (90 bytes) or nullptr
    uint64_t seed, game;
    const uint32_t *rk;    // optional: the seed's Philox key schedule (philox_round_keys) in constant / parameter space
    const Tables *tab;     // optional: the staged tables (for the table form of the bit select)
};

// BoardState.deal (splendor_model.py:95-99).  Counter-based decks: the d-th card dealt from a tier is the
// floor(u*c)-th smallest id still in the deck (c cards left, u = Philox word of (seed, game, d, tier)) - a lazy
// Fisher-Yates draw whose order is a function of (seed, game) only, never of the actions played.
// The deck form is a property of the ENV (META_EXPLICIT, set by spl_inject), not of the batch: after a partial spl_reset of
// an injected batch the reset envs are counter-based while the others still pop their lists.
template <int P, bool LISTS = false>  // LISTS: the caller guarantees src.lists (no counter-based code is generated)
__device__ __forceinline__ uint32_t deal(Env<P> &e, int tier, const DeckSrc &src)
{
    if (LISTS || (src.lists && (e.meta & META_EXPLICIT))) {  // explicit lists: pop from the end
        uint32_t cnt = byte_of(e.k[0], tier);
        if (cnt == 0) return NONE8;
        cnt--;
        e.k[0] = set_byte(e.k[0], tier, cnt);
        const int base = tier == 0 ? 0 : (tier == 1 ? 40 : 70);
        SPL_CHECK(cnt < (tier == 0 ? 40u : (tier == 1 ? 30u : 20u)));
        return src.lists[base + cnt];
    }
    uint32_t lo, hi = 0;
    int size, base;
    if (tier == 0) { lo = e.k[0]; hi = (e.k[2] >> 20) & 0xFFu; size = 40; base = 0; }
    else if (tier == 1) { lo = e.k[1]; size = 30; base = 40; }
    else { lo = e.k[2] & 0xFFFFFu; size = 20; base = 70; }
    const int clo = __popc(lo), c = clo + __popc(hi);
    if (c == 0) return NONE8;
    const uint32_t r = src.rk ? philox_rk(src.rk, src.game, (uint32_t)(size - c), STREAM_DECK0 + tier).x
                              : philox(src.seed, src.game, (uint32_t)(size - c), STREAM_DECK0 + tier).x;
    const int j = (int)__umulhi(r, (uint32_t)c);
    // one select for both halves of the 40-bit first tier: no second code path for the eight high bits
    const bool up = j >= clo;
    const int pos = (src.tab ? select_bit_lut(*src.tab, up ? hi : lo, up ? j - clo : j) : select_bit(up ? hi : lo, up ? j - clo : j)) + (up ? 32 : 0);
    const uint32_t bit = 1u << (pos & 31);
    e.k[0] &= ~((tier == 0 && !up) ? bit : 0u);
    e.k[1] &= ~(tier == 1 ? bit : 0u);
    e.k[2] &= ~(tier == 2 ? bit : (up ? bit << 20 : 0u));
    return (uint32_t)(base + pos);
}

// random.sample(NOBLES, P+1): partial Fisher-Yates over the 10 ids held as nibbles of a 64-bit word; returns the board's
// noble list (5 nibbles, 0xF = none)
template <int P>
__device__ __forceinline__ uint32_t sample_nobles(uint64_t seed, uint64_t game)
{
    uint64_t ids = 0x9876543210ull;
    const uint4 r0 = philox(seed, game, 0, STREAM_NOBLES), r1 = philox(seed, game, 1, STREAM_NOBLES);
    const uint32_t rr[5] = {r0.x, r0.y, r0.z, r0.w, r1.x};
    uint32_t list = 0xFFFFFu;
#pragma unroll
    for (int i = 0; i < P + 1; i++) {
        const int j = i + (int)__umulhi(rr[i], (uint32_t)(10 - i));
        const uint64_t vi = (ids >> (4 * i)) & 15, vj = (ids >> (4 * j)) & 15;
        ids = (ids & ~(15ull << (4 * i)) & ~(15ull << (4 * j))) | (vj << (4 * i)) | (vi << (4 * j));
        list = (list & ~(15u << (4 * i))) | ((uint32_t)vj << (4 * i));
    }
    return list;
}

// SplendorState.__init__ (splendor_model.py:53-93) with Philox nobles and lazily drawn decks
template <int P>
__device__ __forceinline__ void reset_env(Env<P> &e, uint64_t seed, uint64_t game, const uint32_t *rk = nullptr)
{
    const uint32_t n = P == 2 ? 4 : (P == 3 ? 5 : 7);
    e.bank = n * (ONES6 & ~(1u << YSH)) | (5u << YSH);
    e.k[0] = 0xFFFFFFFFu; e.k[1] = 0x3FFFFFFFu; e.k[2] = 0x0FFFFFFFu;
    e.meta = sample_nobles<P>(seed, game) | ((uint32_t)(P - 2) << 22);
    DeckSrc src{nullptr, seed, game, rk};
    e.d[0] = e.d[1] = e.d[2] = 0;
#pragma unroll 1
    for (int i = 0; i < 12; i++) {  // one copy of deal() in the binary: the reset runs once per game, the step loop 90 times
        const int t = i >> 2;
        const uint32_t card = deal(e, t, src) << (8 * (i & 3));
        if (t == 0) e.d[0] |= card; else if (t == 1) e.d[1] |= card; else e.d[2] |= card;
    }
#pragma unroll
    for (int p = 0; p < P; p++) { e.pg[p] = 0; e.pc[p] = 0; e.pr[p] = 0x00FFFFFFu; e.pi[p] = 0; }
}

// The same initial state with the WHOLE deck order drawn up front into `lists` (90 bytes, the explicit-list format of
// spl_inject: a tier's next card is the last entry of its list).  The d-th card of a tier is the same card deal()
// would draw lazily - the order is a function of (seed, game) only - so a game played from here is identical, but its
// in-game deals are one byte load instead of a Philox call and a bit select.  The playout kernel keeps the lists in
// shared memory; the draws run here with all 32 lanes of the warp active instead of ~7 lanes inside the step loop.
// On return the env is in explicit-list form (deck counts in k[0]); lists_to_masks() converts it back.
template <int P>
__device__ __forceinline__ void reset_env_lists(Env<P> &e, uint64_t seed, uint64_t game, const uint32_t *rk, uint8_t *lists)
{
    const uint32_t n = P == 2 ? 4 : (P == 3 ? 5 : 7);
    e.bank = n * (ONES6 & ~(1u << YSH)) | (5u << YSH);
    e.k[0] = 0xFFFFFFFFu; e.k[1] = 0x3FFFFFFFu; e.k[2] = 0x0FFFFFFFu;
    e.meta = sample_nobles<P>(seed, game) | ((uint32_t)(P - 2) << 22);
    const DeckSrc philox_src{nullptr, seed, game, rk};
#pragma unroll 1
    for (int t = 0; t < 3; t++) {
        const int size = t == 0 ? 40 : (t == 1 ? 30 : 20), base = t == 0 ? 0 : (t == 1 ? 40 : 70);
#pragma unroll 1
        for (int d = 0; d < size; d++) lists[base + size - 1 - d] = (uint8_t)deal(e, t, philox_src);
    }
    e.k[0] = 40u | 30u << 8 | 20u << 16; e.k[1] = 0; e.k[2] = 0;
    const DeckSrc list_src{lists, seed, game, rk};
    e.d[0] = e.d[1] = e.d[2] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) e.d[i >> 2] |= deal<P, true>(e, i >> 2, list_src) << (8 * (i & 3));
#pragma unroll
    for (int p = 0; p < P; p++) { e.pg[p] = 0; e.pc[p] = 0; e.pr[p] = 0x00FFFFFFu; e.pi[p] = 0; }
}
// explicit-list deck counts (k[0] bytes) + lists -> the counter-based form: one bit per card still in its deck
template <int P>
__device__ __forceinline__ void lists_to_masks(Env<P> &e, const uint8_t *lists)
{
    const uint32_t counts = e.k[0];
    uint32_t k0 = 0, k1 = 0, k2 = 0;
    for (int t = 0; t < 3; t++) {
        const int base = t == 0 ? 0 : (t == 1 ? 40 : 70), cnt = (int)byte_of(counts, t);
        for (int i = 0; i < cnt; i++) {
            const int pos = (int)lists[base + i] - base;
            if (t == 0 && pos < 32) k0 |= 1u << pos;
            else if (t == 0) k2 |= 1u << (20 + pos - 32);
            else if (t == 1) k1 |= 1u << pos;
            else k2 |= 1u << pos;
        }
    }
    e.k[0] = k0; e.k[1] = k1; e.k[2] = k2;
}

// ------------------------------------------------------------------------------------------------ legal actions
struct Choice {
    int idx;        // ALL_ACTIONS index
    int type;       // ActType
    int slot;       // reserve / buy_available: board position 0..11; buy_reserve: reserved index 0..2
    int n;          // noble slot 0 = none, k+1 = board noble at list position k
    uint32_t C, R;  // gems collected / returned (g6); for buys R is the payment
    uint32_t info;  // buys: Tables::card[].y of the card
};

// Everything getLegalActions derives from the state before enumerating, as a few words.
struct Ctx {
    uint32_t gems, cards, resv, bank;
    uint32_t S, near;      // S: nobles already earned (bit = list position); near: bit 5*colour_field_shift+k
    uint32_t h1, h2, h3;   // field masks: gems held >= 1 / 2 / 3
    uint32_t A1, A4;       // field masks (normal colours): bank >= 1 / >= 4
    uint32_t A1g, A4g;     // the same as guard bits (bit 5f+4), which is how the SWAR compare produces them
    uint32_t occ;          // 12-bit mask of non-empty board slots
    int nS, hold, yellow, n_resv, nA, Lmin, Lmax;
    bool bank_yellow;
};

// MOVER0: the caller keeps the player words ROTATED so that the seat to move is always entry 0 (the playout kernel does:
// a rotation is a handful of register moves per move, the per-seat selects it replaces are ~4 P compare/select pairs)
template <int P, bool MOVER0 = false>
__device__ __forceinline__ void make_ctx(const Env<P> &e, const Tables &T, Ctx &c)
{
    const int seat = MOVER0 ? 0 : seat_of(e.meta);
    c.gems = e.pg[0]; c.cards = e.pc[0]; c.resv = e.pr[0];
#pragma unroll
    for (int p = 1; p < P; p++)
        if (seat == p) { c.gems = e.pg[p]; c.cards = e.pc[p]; c.resv = e.pr[p]; }
    c.bank = e.bank;
    // noble_visit over the board list (:430-435) + which nobles are exactly one card of one colour away (:543-554)
    c.S = 0; c.near = 0;
#pragma unroll
    for (int k = 0; k < P + 1; k++) {  // the board list never holds more than P + 1 nobles (:90-93)
        const uint32_t d = satsub6(T.noble[(e.meta >> (4 * k)) & 15u], c.cards);
        c.S |= (d == 0 ? 1u : 0u) << k;
        c.near |= ((d & (d - 1)) == 0 ? d & ONES6 : 0u) << k;  // exactly one gem of one colour short (branch-free)
    }
    c.nS = c.S ? __popc(c.S) : 1;
    c.hold = (int)hsum6(c.gems);
    c.yellow = (int)field(c.gems, YSH);
    c.h1 = ge_mask(c.gems, 1); c.h2 = ge_mask(c.gems, 2); c.h3 = ge_mask(c.gems, 3);
    c.A1g = (c.bank + 15u * ONES6) & (H6 & ~(16u << YSH)); c.A4g = (c.bank + 12u * ONES6) & (H6 & ~(16u << YSH));
    c.A1 = compress6(c.A1g); c.A4 = compress6(c.A4g);  // dead code in kernels that only count and select
    c.bank_yellow = field(c.bank, YSH) > 0;
    c.n_resv = __popc(~c.resv & 0x00808080u);
    c.occ = occupied4(e.d[0]) | occupied4(e.d[1]) << 4 | occupied4(e.d[2]) << 8;
    c.nA = __popc(c.A1g);
    c.Lmax = min(c.nA, 3);
    c.Lmin = c.hold <= 7 ? c.Lmax : (c.hold == 8 ? min(c.nA, 2) : min(c.nA, 1));
}

// Family totals, per-group collect counts and per-slot buy counts of one decision.
struct FamilyCounts {
    int res, same, diff, bres, bav;  // actions per family (noble multiplicity included)
    uint32_t grp[9];                 // return variants per collect group, one byte each, in ALL_ACTIONS order:
                                     // bytes 0..4 the collect_same colours, bytes 8..32 the collect_diff combos 0..24
    uint32_t buy[4];                 // buy actions per card, one byte each, in ALL_ACTIONS order: the 3 reserved cards | board rows 0, 1, 2
};

__device__ __forceinline__ int reserve_variants(const Ctx &c)
{
    return (!c.bank_yellow || c.hold < 10) ? 1 : __popc(c.h1 & NORMAL_FIELDS);
}
__device__ __forceinline__ int combo_len(int ci) { return ci < 5 ? 1 : (ci < 15 ? 2 : 3); }

// Bitmask (bit v = variant v) of the return variants of a collect group that the agent can pay when it must return `ex`
// gems (<= 0: only "return nothing").  Table lookups only: pext of the held-colour masks onto the group's other colours, then
// the combinations_with_replacement patterns (actions.py:145-199).  collect_same colour: none | pairs 1..10 | one of 11..15 |
// two of 16..20.  Callers never branch on the case "how many gems must go back": it is turned into AND-masks (ExSel) and
// every group becomes a short branch-free sequence of lookups, shifts and AND/OR.  (A first form with nested selects
// compiled to one branch region per group, executed at ~9 active lanes: 41 % of the mask kernel's instructions.)
struct ExSel { uint32_t is0, m1, m2, m3; };  // excess <= 0 -> bit 0 ("return nothing") ; == 1 / 2 / 3 -> all-ones masks
__device__ __forceinline__ ExSel make_exsel(int ex)
{
    return ExSel{ex <= 0 ? 1u : 0u, ex == 1 ? ~0u : 0u, ex == 2 ? ~0u : 0u, ex == 3 ? ~0u : 0u};
}
template <int L>
__device__ __forceinline__ uint32_t diff_variant_bits_sel(const Tables &T, const Ctx &c, int ci, const ExSel &s)
{
    constexpr int kk = 6 - L;
    const uint32_t x = T.pext_diff[ci][c.h1];
    uint32_t r = s.is0 | ((x << 1) & s.m1);
    if (L >= 2) {
        const uint32_t y = T.pext_diff[ci][c.h2];
        const uint32_t two = (uint32_t)(T.cwr2_pair[L == 2][x & 15u] | T.cwr2_dbl[L == 2][y & 15u]);
        r |= (two << (1 + kk)) & s.m2;
        if (L == 3) {
            const uint32_t z = T.pext_diff[ci][c.h3];
            r |= ((uint32_t)T.cwr3[x | y << 3 | z << 6] << 10) & s.m3;
        }
    }
    return r;
}
__device__ __forceinline__ uint32_t same_variant_bits_sel(const Tables &T, const Ctx &c, int col, const ExSel &s)
{
    const uint32_t x = T.pext_same[col][c.h1], y = T.pext_same[col][c.h2];
    return s.is0 | ((x << 11) & s.m1) | (((uint32_t)T.pairs5[x] << 1 | y << 16) & s.m2);
}

// the same for ONE group whose size is only known at run time (select_family): all lookups unconditional, the case picked
// by masks; entries that do not apply to the size (pairs for L = 1, triples for L < 3) are masked out because the excess
// cannot reach them (hold <= 10)
__device__ __forceinline__ uint32_t diff_variant_bits_rt(const Tables &T, const Ctx &c, int ci, int L, int ex)
{
    SPL_CHECK(ci >= 0 && ci < 25 && L == combo_len(ci) && ex <= L);
    const ExSel s = make_exsel(ex);
    const uint32_t x = T.pext_diff[ci][c.h1], y = T.pext_diff[ci][c.h2], z = T.pext_diff[ci][c.h3];
    const uint32_t two = (uint32_t)(T.cwr2_pair[L == 2][x & 15u] | T.cwr2_dbl[L == 2][y & 15u]);
    const uint32_t three = (uint32_t)T.cwr3[(x | y << 3 | z << 6) & 511u];
    return s.is0 | ((x << 1) & s.m1) | ((two << (7 - L)) & s.m2) | ((three << 10) & s.m3);
}

// resources_sufficient (:371-394): returns true and the payment if the card is affordable and below the 7-card cap
__device__ __forceinline__ bool can_buy(const Ctx &c, uint2 cw, uint32_t &pay)
{
    const uint32_t need = satsub6(cw.x, c.cards);
    const uint32_t shortfall = satsub6(need, c.gems);
    const uint32_t wild = hsum6(shortfall);
    pay = (need - shortfall) + (wild << YSH);
    return field(c.cards, cw.y & 31u) != 7u && wild <= (uint32_t)c.yellow;
}

// ---- collect_same / collect_diff return variants as generating functions
// The return multisets of e gems out of a colour set O, with at most min(held_c, 3) of colour c, are counted by the
// coefficient of x^e in  prod_{c in O} (1 + x + .. + x^lvl_c),  lvl_c = min(held_c, 3)  (generate_return_combos
// :329-367 builds exactly these multisets).  Only e <= 3 matters, and every coefficient up to x^3 is < 256 (at most
// C(8,3) = 56), so a polynomial mod x^4 is kept as four base-256 digits of one 32-bit word: the product of two
// polynomials is ONE integer multiply (digits never carry; the x^4.. terms fall off the top), and removing a colour
// is a multiply by the modular inverse of its factor (the factors are odd numbers: 1/(1+x) = 0xFF00FF01 ...).
// So with G = the product over all six colours, a collect group that takes the colours {a,b,c} has
//     variants = digit_e( G * I_a * I_b * I_c ),   e = gems over the limit of 10 (digit 0 is 1: "return nothing"),
// and an unavailable colour (bank empty, or < 4 for collect_same) is folded in as I = 0, which zeroes the whole group.
// 45 multiplies and 23 byte permutes give all 30 groups of a decision at once, packed one byte per group in
// ALL_ACTIONS order; select_legal finds the k-th action with two SWAR prefix-sum steps instead of walking the groups.
__device__ __forceinline__ uint32_t pick2(uint32_t x, uint32_t y, uint32_t sel) { return __byte_perm(x, y, sel); }
__device__ __forceinline__ uint32_t pack4(uint32_t xy, uint32_t zw) { return __byte_perm(xy, zw, 0x5410u); }

__device__ __forceinline__ void collect_groups(const Tables &T, const Ctx &c, FamilyCounts &f)
{
    const int t1 = __popc(c.h1), t2 = __popc(c.h2), t3 = __popc(c.h3);
    SPL_CHECK(t1 >= t2 && t2 >= t3 && t1 <= 6 && c.hold <= 10);
    const uint32_t G = T.poly_pow[0][t1 - t2] * T.poly_pow[1][t2 - t3] * T.poly_pow[2][t3];
    uint32_t Js[5], Jd[5];  // inverse factor of normal colour n, zero when collect_same / collect_diff cannot take it
#pragma unroll
    for (int n = 0; n < 5; n++) {
        const int fl = field_of_normal(n);
        const uint32_t inv = T.poly_inv[field(c.gems, 5 * fl) & 15u];
        Js[n] = inv * (c.A4g >> (5 * fl + 4) & 1u);  // a multiply instead of a compare + select: the FMA pipe has room, the ALU pipe has not
        Jd[n] = inv * (c.A1g >> (5 * fl + 4) & 1u);
    }
    // digit selectors: a group taking L gems returns e_L = max(hold + L - 10, 0) <= 3; collect_same returns e_2
    const uint32_t e1 = (uint32_t)max(c.hold - 9, 0), e2 = (uint32_t)max(c.hold - 8, 0), e3 = (uint32_t)max(c.hold - 7, 0);
    const uint32_t s11 = e1 * 17u + 0x40u, s12 = e1 + e2 * 16u + 0x40u, s22 = e2 * 17u + 0x40u, s23 = e2 + e3 * 16u + 0x40u,
                   s33 = e3 * 17u + 0x40u;
    // collect_same (:475-496): colour n excluded from the returnable colours
    {
        const uint32_t p0 = G * Js[0], p1 = G * Js[1], p2 = G * Js[2], p3 = G * Js[3], p4 = G * Js[4];
        f.grp[0] = pack4(pick2(p0, p1, s22), pick2(p2, p3, s22));
        f.grp[1] = pick2(p4, 0u, e2 + 0x4440u);
    }
    // collect_diff (:437-473): the sizes outside [Lmin, Lmax] are zeroed through G (sizes above Lmax contain an
    // unavailable colour anyway)
    {
        const uint32_t G1 = c.Lmin <= 1 ? G : 0u, G2 = c.Lmin <= 2 ? G : 0u, G3 = G;
        uint32_t a[5], b[10], d[10];
#pragma unroll
        for (int n = 0; n < 5; n++) a[n] = G1 * Jd[n];
        uint32_t pr[10];
        {
            int i = 0;
#pragma unroll
            for (int x = 0; x < 5; x++)
#pragma unroll
                for (int y = x + 1; y < 5; y++) pr[i++] = Jd[x] * Jd[y];  // 01 02 03 04 12 13 14 23 24 34
        }
#pragma unroll
        for (int i = 0; i < 10; i++) b[i] = G2 * pr[i];
        const uint32_t g0 = G3 * Jd[0], g1 = G3 * Jd[1], g2 = G3 * Jd[2];
        d[0] = g0 * pr[4]; d[1] = g0 * pr[5]; d[2] = g0 * pr[6];   // 012 013 014
        d[3] = g0 * pr[7]; d[4] = g0 * pr[8]; d[5] = g0 * pr[9];   // 023 024 034
        d[6] = g1 * pr[7]; d[7] = g1 * pr[8]; d[8] = g1 * pr[9];   // 123 124 134
        d[9] = g2 * pr[9];                                         // 234
        f.grp[2] = pack4(pick2(a[0], a[1], s11), pick2(a[2], a[3], s11));
        f.grp[3] = pack4(pick2(a[4], b[0], s12), pick2(b[1], b[2], s22));
        f.grp[4] = pack4(pick2(b[3], b[4], s22), pick2(b[5], b[6], s22));
        f.grp[5] = pack4(pick2(b[7], b[8], s22), pick2(b[9], d[0], s23));
        f.grp[6] = pack4(pick2(d[1], d[2], s33), pick2(d[3], d[4], s33));
        f.grp[7] = pack4(pick2(d[5], d[6], s33), pick2(d[7], d[8], s33));
        f.grp[8] = pick2(d[9], 0u, e3 + 0x4440u);
    }
    // totals: bytes are <= 15, so whole words can be added before the horizontal byte sum
    f.same = (int)((((f.grp[0] + f.grp[1]) * 0x01010101u) >> 24) * (uint32_t)c.nS);
    const uint32_t dsum = f.grp[2] + f.grp[3] + f.grp[4] + f.grp[5] + f.grp[6] + f.grp[7] + f.grp[8];  // bytes <= 70
    f.diff = (int)(((dsum * 0x01010101u) >> 24) * (uint32_t)c.nS);
}

template <int P>
__device__ __forceinline__ int count_legal(const Env<P> &e, const Tables &T, const Ctx &c, FamilyCounts &f)
{
    f.res = c.n_resv < 3 ? __popc(c.occ) * c.nS * reserve_variants(c) : 0;
    collect_groups(T, c, f);
    // buy (:522-568): 12 board slots + 3 reserved cards, branch-free; empty slots (0xFF) index an unaffordable filler
    const uint32_t eff = (c.gems & ~(31u << YSH)) + c.cards;    // gems + bought cards per colour (<= 14)
    // actions per buyable card of each colour = nobles to choose from after the purchase (at least 1), as a 5-bit field
    // per colour; 0 where 7 cards of the colour are owned (the cap) and for the filler's "yellow"
    uint32_t mult = 0;
#pragma unroll
    for (int n = 0; n < 5; n++) {
        const int sh = 5 * field_of_normal(n);
        mult |= (uint32_t)max(__popc(c.S | ((c.near >> sh) & 31u)), 1) << sh;
    }
    mult &= ~((((c.cards + 9u * ONES6) & H6) >> 4) * 31u);
    // wilds needed = sum of max(cost - cards - gems, 0) <= 14; x * (ONES6 << 2) leaves that sum in the top five bits (what
    // would lie above falls off), so "sum <= yellow" is one compare of the product against ((yellow + 1) << 27) - 1
    const uint32_t wild_max = (((uint32_t)c.yellow + 1u) << 27) - 1u;
    uint32_t bw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int s = 0; s < 15; s++) {
        const uint32_t w = s < 12 ? e.d[s >> 2] : c.resv;
        const uint2 cw = T.buyrow[__byte_perm(w, 0u, 0x4440u + (s & 3))];  // byte s & 3 of w
        const uint32_t sh = cw.y;
        const uint32_t cnt = satsub6g(cw.x, eff) * (ONES6 << 2) <= wild_max ? (mult >> sh) & 31u : 0u;
        bw[s < 12 ? 1 + (s >> 2) : 0] += cnt << (8 * (s & 3));
    }
    f.buy[0] = bw[0]; f.buy[1] = bw[1]; f.buy[2] = bw[2]; f.buy[3] = bw[3];
    f.bres = (int)((bw[0] * 0x01010101u) >> 24);
    f.bav = (int)(((bw[1] + bw[2] + bw[3]) * 0x01010101u) >> 24);  // bytes <= 5, 6 nobles slots at most
    return f.res + f.same + f.diff + f.bres + f.bav;
}

// The four action families in ascending ALL_ACTIONS order (pass | reserve | collect_same + collect_diff | buy_reserve +
// buy_available).  A random choice k among `total` legal actions falls into one family; kf is its rank inside it.
enum Family { FAM_PASS = 0, FAM_RESERVE = 1, FAM_COLLECT = 2, FAM_BUY = 3 };
__device__ __forceinline__ int family_of_k(const Ctx &c, const FamilyCounts &f, int total, int k, int &kf)
{
    SPL_CHECK(k >= 0 && k < (total ? total : c.nS));
    kf = k;
    if (total == 0) return FAM_PASS;  // pass (:570-574) is legal only when nothing else is; k < nS picks the noble
    if (k < f.res) return FAM_RESERVE;
    kf = k - f.res;
    if (kf < f.same + f.diff) return FAM_COLLECT;
    kf -= f.same + f.diff;
    return FAM_BUY;
}

// The kf-th legal action (ascending ALL_ACTIONS order) of family `fam`.  Reads of c / f per family:
//   pass     -
//   reserve  c.nS c.S c.occ c.bank_yellow c.hold c.h1
//   collect  c.nS c.S c.hold c.h1 c.h2 c.h3, f.grp
//   buy      c.S c.near c.gems c.cards c.yellow c.resv, f.buy, e.d
// so a caller that knows the family at compile time (the playout kernel's second phase, where a warp holds games of one
// family) only has to provide those.
template <int P>
__device__ __forceinline__ void select_family(const Env<P> &e, const Tables &T, const Ctx &c, const FamilyCounts &f, int fam, int k,
                                              Choice &ch)
{
    ch.C = 0; ch.R = 0; ch.info = 0; ch.slot = 0;
    // Every family ends the same way - pick the pj-th set bit of a mask (board slot / return variant), pick the ni-th
    // noble of a set, index = base + n * nmul + pos * pmul - so the families only prepare those operands and the picks
    // run ONCE after the branches have rejoined.
    uint32_t pm = 1u, Sb = c.S;
    int pj = 0, ni = 0, base = 0, nmul = 1, pmul = 0;
    const uint32_t *row = T.ret_same;  // collect families: returned gems of variant `pos`; others: not read
    if (fam == FAM_PASS) {  // pass (:570-574): index = noble slot
        ch.type = T_PASS;
        ni = k;
    } else if (fam == FAM_RESERVE) {  // reserve: index 6 + p*42 + n*7 + v
        const int nv = reserve_variants(c), per = c.nS * nv;
        const int q = small_div(T, k, per), rem = k - q * per;
        ni = small_div(T, rem, nv);
        const int vi = rem - ni * nv;
        ch.type = T_RESERVE;
        pm = c.occ; pj = q;
        int v = 0;
        if (c.bank_yellow) {
            v = 1; ch.C = 1u << YSH;
            if (c.hold >= 10) {
                const int fl = select_bit8_lut(T, c.h1 & NORMAL_FIELDS, vi);
                v = 2 + fl - (fl > 2 ? 1 : 0);
                ch.R = 1u << (5 * fl);
            }
        }
        base = BASE_RESERVE + v; nmul = 7; pmul = 42;
    } else if (fam == FAM_COLLECT) {
        // collect_same (510 + c*126 + n*21 + v) and collect_diff (combo_base + n*V + v).  The 30 groups (colours, then
        // colour combinations) lie in index order in f.grp, one byte of return-variant count each; a group holds
        // variants * nS consecutive legal actions.  Find the group of action k in two prefix-sum steps - the word, then
        // the byte inside it (a multiply by 0x01010101 turns four counts into their four running sums) - then the noble
        // slot and the vi-th payable return variant of the group's table row.
        const uint32_t q = (uint32_t)small_div(T, k, c.nS);
        uint32_t W = f.grp[0], before = 0, run = 0;
        int wi = 0;
#pragma unroll
        for (int w = 0; w < 9; w++) {
            if (w > 0 && q >= run) { W = f.grp[w]; before = run; wi = w; }
            run += (f.grp[w] * 0x01010101u) >> 24;
        }
        SPL_CHECK(q < run && q - before < 64u);
        const uint32_t sums = W * 0x01010101u;                                    // running sums of the word's bytes (<= 60)
        const uint32_t ahead = (((q - before) * 0x01010101u | 0x80808080u) - sums) & 0x80808080u;  // bytes wholly before q
        const int b = __popc(ahead), g = 4 * wi + b;
        SPL_CHECK(b < 4 && g != 5 && g != 6 && g != 7 && g <= 32);
        before += ((sums << 8) >> (8 * b)) & 255u;
        const int nv = (int)((W >> (8 * b)) & 255u);
        k -= (int)before * c.nS;
        if (g < 8) {
            const int col = g, e2 = c.hold - 8;
            row = T.ret_same + col * 21;
            base = BASE_SAME + col * 126; nmul = 21;
            pm = same_variant_bits_sel(T, c, col, make_exsel(e2));
            ch.type = T_SAME; ch.slot = col;
            ch.C = 2u << (5 * field_of_normal(col));
        } else {
            const int ci = g - 8, L = combo_len(ci);
            nmul = L == 1 ? 6 : (L == 2 ? 15 : 20);
            row = T.ret_diff + T.combo_off[ci];
            base = T.combo_base[ci];
            pm = diff_variant_bits_rt(T, c, ci, L, c.hold + L - 10);
            ch.type = T_DIFF; ch.slot = ci;
            ch.C = T.combo_c[ci];
        }
        ni = small_div(T, k, nv);
        pj = k - ni * nv;  // the pj-th payable return variant, ascending
        pmul = 1;
        SPL_CHECK(nv >= 1 && nv == __popc(pm) && pj >= 0 && pj < nv && ni < c.nS);
    } else {
        // buys: reserved cards first (3420 + r*6 + n), then the board (3438 + p*6 + n).  The card of action k by the same
        // two prefix-sum steps as the collect groups (word, then byte); what is left of k is the noble choice
        uint32_t W = f.buy[0], before = 0, run = 0;
        int wi = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            if (w > 0 && (uint32_t)k >= run) { W = f.buy[w]; before = run; wi = w; }
            run += (f.buy[w] * 0x01010101u) >> 24;
        }
        SPL_CHECK((uint32_t)k < run);
        const uint32_t sums = W * 0x01010101u;
        const uint32_t ahead = ((((uint32_t)k - before) * 0x01010101u | 0x80808080u) - sums) & 0x80808080u;
        const int b = __popc(ahead);
        ni = k - (int)(before + (((sums << 8) >> (8 * b)) & 255u));
        const int s = wi == 0 ? 12 + b : 4 * (wi - 1) + b;
        SPL_CHECK(b < 4 && s < 15 && ni >= 0 && ni < 6);
        const uint32_t w = s < 4 ? e.d[0] : (s < 8 ? e.d[1] : (s < 12 ? e.d[2] : c.resv));
        const uint2 cw = T.card[byte_of(w, s & 3) & 127u];
        Sb = c.S | ((c.near >> (cw.y & 31u)) & 31u);
        can_buy(c, cw, ch.R);
        ch.type = s < 12 ? T_BUY_AVAILABLE : T_BUY_RESERVE;
        ch.slot = s < 12 ? s : s - 12;
        ch.info = cw.y;
        base = s < 12 ? BASE_BUY_AVAILABLE + s * 6 : BASE_BUY_RESERVE + (s - 12) * 6;
    }
    // ---- the shared tail
    SPL_CHECK(pj >= 0 && pj < __popc(pm));
    const int pos = select_bit_lut(T, pm, pj);
    ch.n = Sb ? select_bit8_lut(T, Sb, ni) + 1 : 0;
    ch.idx = base + ch.n * nmul + pos * pmul;
    if (ch.type == T_RESERVE) ch.slot = pos;
    if (pmul == 1) { SPL_CHECK(pos < nmul); ch.R = row[pos]; }
}

// k-th legal action in ascending ALL_ACTIONS order (k < total; total == 0 means only passes are legal, k < nS)
template <int P>
__device__ __forceinline__ void select_legal(const Env<P> &e, const Tables &T, const Ctx &c, const FamilyCounts &f,
                                             int total, int k, Choice &ch)
{
    int kf;
    const int fam = family_of_k(c, f, total, k, kf);
    select_family(e, T, c, f, fam, kf, ch);
}

// Is ALL_ACTIONS[idx] legal for the seat to move?  Decodes it into ch either way (build_action, gym/envs/utils.py:43-145).
// This is the membership form of the same rules, written independently of count/select so the two cross-check.
// Tcount: the tables the (rare) pass check counts the legal actions with - a kernel that stages only the decode / apply
// part of the tables into shared memory passes the complete copy in global memory here.
template <int P>
__device__ __forceinline__ bool decode_and_check(const Env<P> &e, const Tables &T, const Ctx &c, int idx, Choice &ch, const Tables &Tcount)
{
    ch.idx = idx; ch.C = 0; ch.R = 0; ch.info = 0; ch.slot = 0; ch.n = 0; ch.type = T_PASS;
    if (idx < 0 || idx >= SPL_NUM_ACTIONS) return false;
    bool ok;
    uint32_t S = c.S;
    if (idx < BASE_RESERVE) {
        FamilyCounts f;
        ch.n = idx;
        ok = count_legal(e, Tcount, c, f) == 0;
    } else if (idx < BASE_SAME) {
        const int r = idx - BASE_RESERVE, p = r / 42, v = r % 7;
        ch.type = T_RESERVE; ch.slot = p; ch.n = (r % 42) / 7;
        int expect_lo, expect_hi;
        if (!c.bank_yellow) expect_lo = expect_hi = 0;
        else if (c.hold < 10) expect_lo = expect_hi = 1;
        else { expect_lo = 2; expect_hi = 6; }
        ok = c.n_resv < 3 && (c.occ >> p & 1) && v >= expect_lo && v <= expect_hi;
        if (v >= 1) ch.C = 1u << YSH;
        if (v >= 2) {
            const int fl = field_of_normal(v - 2);
            ch.R = 1u << (5 * fl);
            ok = ok && (c.h1 >> fl & 1);
        }
    } else if (idx < BASE_DIFF) {
        const int r = idx - BASE_SAME, col = r / 126, v = r % 21, fl = field_of_normal(col);
        ch.type = T_SAME; ch.slot = col; ch.n = (r % 126) / 21;
        ch.C = 2u << (5 * fl);
        ch.R = T.ret_same[col * 21 + v];
        ok = (c.A4 >> fl & 1) && (int)hsum6(ch.R) == max(c.hold - 8, 0) && satsub6(ch.R, c.gems) == 0;
    } else if (idx < BASE_BUY_RESERVE) {
        const int r = idx - BASE_DIFF;
        int ci, rem, V;
        if (r < 180) { ci = r / 36; rem = r % 36; V = 6; }
        else if (r < 1080) { ci = 5 + (r - 180) / 90; rem = (r - 180) % 90; V = 15; }
        else { ci = 15 + (r - 1080) / 120; rem = (r - 1080) % 120; V = 20; }
        const int L = combo_len(ci), v = rem % V;
        const uint32_t m = T.combo_mask[ci];
        ch.type = T_DIFF; ch.slot = ci; ch.n = rem / V;
        ch.C = T.combo_c[ci];
        ch.R = T.ret_diff[T.combo_off[ci] + v];
        ok = !(m & ~c.A1) && L >= c.Lmin && (int)hsum6(ch.R) == max(c.hold + L - 10, 0) && satsub6(ch.R, c.gems) == 0;
    } else {
        const bool avail = idx >= BASE_BUY_AVAILABLE;
        const int r = idx - (avail ? BASE_BUY_AVAILABLE : BASE_BUY_RESERVE), s = r / 6;
        ch.type = avail ? T_BUY_AVAILABLE : T_BUY_RESERVE; ch.slot = s; ch.n = r % 6;
        uint32_t w = e.d[0];
        if ((s >> 2) == 1) w = e.d[1];
        if ((s >> 2) == 2) w = e.d[2];
        const uint32_t id = avail ? byte_of(w, s & 3) : byte_of(c.resv, s);
        ok = id != NONE8;
        const uint2 cw = T.card[id & 127u];
        ok = ok && can_buy(c, cw, ch.R);
        ch.info = cw.y;
        S |= (c.near >> (cw.y & 31u)) & 31u;
    }
    const bool noble_ok = S ? (ch.n >= 1 && (S >> (ch.n - 1) & 1)) : ch.n == 0;
    return ok && noble_ok;
}

template <int P>
__device__ __forceinline__ bool decode_and_check(const Env<P> &e, const Tables &T, const Ctx &c, int idx, Choice &ch)
{
    return decode_and_check(e, T, c, idx, ch, T);
}

// ---- the same membership test with the index decoded by TABLE (the step kernels): every family's divisions and remainders
// of the index above become one 16-byte row read, identical code for all 32 lanes of a warp, and collect_same /
// collect_diff share one check.  Row of index i:  .x = gems collected, .y = gems returned (both g6; zero for buys),
// .z = type | slot << 3 | noble slot << 11 | gems taken << 15 (2 for collect_same, the combo's length for collect_diff)
//      | the length tested against Lmin << 18 (7 = always passes, for collect_same) | reserve variant << 21,
// .w = colour-field mask that must be available (collect: the colours taken; reserve: the colour handed back, if any).
struct DecodeTable { uint4 row[SPL_NUM_ACTIONS]; };
__host__ __device__ constexpr DecodeTable make_decode_table(const Tables &T)
{
    DecodeTable D{};
    for (int idx = 0; idx < SPL_NUM_ACTIONS; idx++) {
        uint32_t C = 0, R = 0, type = T_PASS, slot = 0, n = 0, take = 0, lchk = 7, v = 0, need = 0;
        if (idx < BASE_RESERVE) {
            n = (uint32_t)idx;
        } else if (idx < BASE_SAME) {
            const int r = idx - BASE_RESERVE;
            type = T_RESERVE; slot = (uint32_t)(r / 42); n = (uint32_t)((r % 42) / 7); v = (uint32_t)(r % 7);
            if (v >= 1) C = 1u << YSH;
            if (v >= 2) { const int fl = field_of_normal((int)v - 2); R = 1u << (5 * fl); need = 1u << fl; }
        } else if (idx < BASE_DIFF) {
            const int r = idx - BASE_SAME, col = r / 126, fl = field_of_normal(col);
            type = T_SAME; slot = (uint32_t)col; n = (uint32_t)((r % 126) / 21); take = 2;
            C = 2u << (5 * fl); R = T.ret_same[col * 21 + r % 21]; need = 1u << fl;
        } else if (idx < BASE_BUY_RESERVE) {
            const int r = idx - BASE_DIFF;
            int ci = 0, rem = 0, V = 0;
            if (r < 180) { ci = r / 36; rem = r % 36; V = 6; }
            else if (r < 1080) { ci = 5 + (r - 180) / 90; rem = (r - 180) % 90; V = 15; }
            else { ci = 15 + (r - 1080) / 120; rem = (r - 1080) % 120; V = 20; }
            const int L = ci < 5 ? 1 : (ci < 15 ? 2 : 3);
            type = T_DIFF; slot = (uint32_t)ci; n = (uint32_t)(rem / V); take = (uint32_t)L; lchk = (uint32_t)L;
            C = T.combo_c[ci]; R = T.ret_diff[T.combo_off[ci] + rem % V]; need = T.combo_mask[ci];
        } else {
            const bool avail = idx >= BASE_BUY_AVAILABLE;
            const int r = idx - (avail ? BASE_BUY_AVAILABLE : BASE_BUY_RESERVE);
            type = avail ? T_BUY_AVAILABLE : T_BUY_RESERVE; slot = (uint32_t)(r / 6); n = (uint32_t)(r % 6);
        }
        D.row[idx] = uint4{C, R, type | slot << 3 | n << 11 | take << 15 | lchk << 18 | v << 21, need};
    }
    return D;
}
// `row` = DecodeTable::row[idx] for a valid idx (anything for an invalid one); the caller may have loaded it early
template <int P>
__device__ __forceinline__ bool decode_and_check_row(const Env<P> &e, const Tables &T, const Ctx &c, int idx, const uint4 &row, Choice &ch,
                                                     const Tables &Tcount)
{
    ch.idx = idx; ch.C = 0; ch.R = 0; ch.info = 0; ch.slot = 0; ch.n = 0; ch.type = T_PASS;
    if (idx < 0 || idx >= SPL_NUM_ACTIONS) return false;
    ch.C = row.x; ch.R = row.y;
    ch.type = (int)(row.z & 7u); ch.slot = (int)((row.z >> 3) & 255u); ch.n = (int)((row.z >> 11) & 15u);
    bool ok;
    uint32_t S = c.S;
    if (ch.type == T_PASS) {
        FamilyCounts f;
        ok = count_legal(e, Tcount, c, f) == 0;
    } else if (ch.type == T_RESERVE) {
        const int v = (int)((row.z >> 21) & 7u);
        int expect_lo, expect_hi;
        if (!c.bank_yellow) expect_lo = expect_hi = 0;
        else if (c.hold < 10) expect_lo = expect_hi = 1;
        else { expect_lo = 2; expect_hi = 6; }
        ok = c.n_resv < 3 && (c.occ >> ch.slot & 1) && v >= expect_lo && v <= expect_hi && !(row.w & ~c.h1);
    } else if (ch.type <= T_DIFF) {  // collect_same (two of a colour with >= 4 in the bank) or collect_diff (one each)
        const uint32_t avail = ch.type == T_SAME ? c.A4 : c.A1;
        const int take = (int)((row.z >> 15) & 7u), lchk = (int)((row.z >> 18) & 7u);
        ok = !(row.w & ~avail) && lchk >= c.Lmin && (int)hsum6(ch.R) == max(c.hold + take - 10, 0) && satsub6(ch.R, c.gems) == 0;
    } else {
        const bool avail = ch.type == T_BUY_AVAILABLE;
        const int s = ch.slot;
        uint32_t w = e.d[0];
        if ((s >> 2) == 1) w = e.d[1];
        if ((s >> 2) == 2) w = e.d[2];
        const uint32_t id = avail ? byte_of(w, s & 3) : byte_of(c.resv, s);
        ok = id != NONE8;
        const uint2 cw = T.card[id & 127u];
        ok = ok && can_buy(c, cw, ch.R);
        ch.info = cw.y;
        S |= (c.near >> (cw.y & 31u)) & 31u;
    }
    const bool noble_ok = S ? (ch.n >= 1 && (S >> (ch.n - 1) & 1)) : ch.n == 0;
    return ok && noble_ok;
}

// ------------------------------------------------------------------------------------------------ successor
// generateSuccessor + GameRule.update.  Returns the acting agent's score delta.
template <int P, bool LISTS = false, bool MOVER0 = false>
__device__ __forceinline__ int apply_action(Env<P> &e, const Choice &ch, const DeckSrc &src)
{
    const int seat = seat_of(e.meta), slot = MOVER0 ? 0 : seat;  // slot: where the mover's words are
    uint32_t g = e.pg[0], cd = e.pc[0], rv = e.pr[0], pi = e.pi[0];
#pragma unroll
    for (int p = 1; p < P; p++)
        if (slot == p) { g = e.pg[p]; cd = e.pc[p]; rv = e.pr[p]; pi = e.pi[p]; }

    SPL_CHECK(satsub6(ch.R, g + ch.C) == 0 && satsub6(ch.C, e.bank + ch.R) == 0);  // nobody pays or takes gems that are not there
    g = g + ch.C - ch.R;
    e.bank = e.bank - ch.C + ch.R;
    SPL_CHECK(!(g & H6) && !(e.bank & H6) && hsum6(g) <= 10);
    int score = 0;
    // one test, not two short-circuit branches: the reserving and the buying lanes of a warp must enter the deal together
    if ((1u << ch.type) & (1u << T_RESERVE | 1u << T_BUY_AVAILABLE)) {
        const int tier = ch.slot >> 2, col = ch.slot & 3;
        uint32_t w = e.d[0];
        if (tier == 1) w = e.d[1];
        if (tier == 2) w = e.d[2];
        const uint32_t id = byte_of(w, col);
        w = set_byte(w, col, deal<P, LISTS>(e, tier, src));
        if (tier == 0) e.d[0] = w; else if (tier == 1) e.d[1] = w; else e.d[2] = w;
        if (ch.type == T_RESERVE) rv = set_byte(rv, __popc(~rv & 0x00808080u), id);
    } else if (ch.type == T_BUY_RESERVE) {
        const uint32_t list = rv & 0x00FFFFFFu, sh = 8u * ch.slot;
        rv = (rv & 0xFF000000u) | (list & ((1u << sh) - 1u)) | ((list >> (sh + 8u)) << sh) | 0x00FF0000u;
    }
    if (ch.type >= T_BUY_RESERVE) {
        cd += 1u << (ch.info & 31u);
        score += (int)((ch.info >> 5) & 7u);
    }
    if (ch.n > 0) {  // the noble leaves the board list; later nobles shift down (:280-289)
        const uint32_t list = e.meta & 0xFFFFFu, sh = 4u * (ch.n - 1);
        const uint32_t id = (list >> sh) & 15u;
        e.meta = (e.meta & ~0xFFFFFu) | (list & ((1u << sh) - 1u)) | ((list >> (sh + 4u)) << sh) | 0xF0000u;
        pi |= 1u << (16 + id);
        score += 3;
    }
    rv += (uint32_t)score << 24;
    pi = ((pi + 1u) & ~(1u << 26)) | (ch.type == T_PASS ? 1u << 26 : 0u);
#pragma unroll
    for (int p = 0; p < P; p++)
        if (slot == p) { e.pg[p] = g; e.pc[p] = cd; e.pr[p] = rv; e.pi[p] = pi; }
    const int next = seat + 1 == P ? 0 : seat + 1;
    e.meta = (e.meta & ~(3u << 20)) | ((uint32_t)next << 20);
    return score;
}

// player words one place to the left (entry 1 becomes entry 0): after a move of a MOVER0 state the next mover is in front
template <int P>
__device__ __forceinline__ void rotate_players_left(Env<P> &e)
{
    const uint32_t g = e.pg[0], c = e.pc[0], r = e.pr[0], i = e.pi[0];
#pragma unroll
    for (int p = 0; p + 1 < P; p++) { e.pg[p] = e.pg[p + 1]; e.pc[p] = e.pc[p + 1]; e.pr[p] = e.pr[p + 1]; e.pi[p] = e.pi[p + 1]; }
    e.pg[P - 1] = g; e.pc[P - 1] = c; e.pr[P - 1] = r; e.pi[P - 1] = i;
}
// seat-order state <-> MOVER0 state (entry 0 = the seat to move, entry k = seat (seat + k) mod P)
template <int P>
__device__ __forceinline__ void to_mover_first(Env<P> &e)
{
    for (int k = seat_of(e.meta); k > 0; k--) rotate_players_left(e);
}
template <int P>
__device__ __forceinline__ void to_seat_order(Env<P> &e)
{
    for (int k = (P - seat_of(e.meta)) % P; k > 0; k--) rotate_players_left(e);
}

// generatePredecessor (splendor_model.py:156-220): undo ALL_ACTIONS[idx] for `seat`.  `card` = the card that was bought
// or reserved, `noble` = the noble id that was claimed (NONE8 if none), `returned` = the action's returned_gems (g6; for
// buys the payment).  The reference's quirks are kept: the noble and an un-bought reserved card are APPENDED to their
// lists, the card that had been dealt into the slot goes back on top of its deck, `passed` comes from the undone action.
// The reference keeps the seat to move in the rule object; the packed state carries it, so it is set back to `seat`.
template <int P>
__device__ __forceinline__ void undo_action(Env<P> &e, const Tables &T, int seat, int idx, uint32_t card, uint32_t noble, uint32_t returned)
{
    uint32_t g = e.pg[0], cd = e.pc[0], rv = e.pr[0], pi = e.pi[0];
#pragma unroll
    for (int p = 1; p < P; p++)
        if (seat == p) { g = e.pg[p]; cd = e.pc[p]; rv = e.pr[p]; pi = e.pi[p]; }
    int type = T_PASS, slot = 0, score = 0;
    uint32_t C = 0, R = 0;
    if (idx >= BASE_BUY_AVAILABLE) { type = T_BUY_AVAILABLE; slot = (idx - BASE_BUY_AVAILABLE) / 6; R = returned; }
    else if (idx >= BASE_BUY_RESERVE) { type = T_BUY_RESERVE; R = returned; }
    else if (idx >= BASE_DIFF) {
        const int r = idx - BASE_DIFF;
        int ci, rem, V;
        if (r < 180) { ci = r / 36; rem = r % 36; V = 6; }
        else if (r < 1080) { ci = 5 + (r - 180) / 90; rem = (r - 180) % 90; V = 15; }
        else { ci = 15 + (r - 1080) / 120; rem = (r - 1080) % 120; V = 20; }
        type = T_DIFF; C = T.combo_c[ci]; R = T.ret_diff[T.combo_off[ci] + rem % V];
    } else if (idx >= BASE_SAME) {
        const int r = idx - BASE_SAME, col = r / 126;
        type = T_SAME; C = 2u << (5 * field_of_normal(col)); R = T.ret_same[col * 21 + r % 21];
    } else if (idx >= BASE_RESERVE) {
        const int r = idx - BASE_RESERVE, v = r % 7;
        type = T_RESERVE; slot = r / 42;
        if (v >= 1) C = 1u << YSH;
        if (v >= 2) R = 1u << (5 * field_of_normal(v - 2));
    }
    g = g - C + R;
    e.bank = e.bank + C - R;
    if (type == T_RESERVE || type == T_BUY_AVAILABLE) {
        const int tier = slot >> 2, col = slot & 3;
        uint32_t w = tier == 0 ? e.d[0] : (tier == 1 ? e.d[1] : e.d[2]);
        const uint32_t cur = byte_of(w, col);
        if (cur != NONE8) {  // back on top of its deck
            if (e.meta & META_EXPLICIT) e.k[0] += 1u << (8 * tier);
            else {
                const uint32_t pos = cur - (tier == 0 ? 0u : (tier == 1 ? 40u : 70u));
                if (tier == 0 && pos < 32) e.k[0] |= 1u << pos;
                else if (tier == 0) e.k[2] |= 1u << (20 + pos - 32);
                else if (tier == 1) e.k[1] |= 1u << pos;
                else e.k[2] |= 1u << pos;
            }
        }
        w = set_byte(w, col, card);
        if (tier == 0) e.d[0] = w; else if (tier == 1) e.d[1] = w; else e.d[2] = w;
        if (type == T_RESERVE) {  // agent.cards["yellow"].remove(card)
            const uint32_t list = rv & 0x00FFFFFFu;
            const uint32_t sh = byte_of(list, 0) == card ? 0u : (byte_of(list, 1) == card ? 8u : 16u);
            rv = (rv & 0xFF000000u) | (list & ((1u << sh) - 1u)) | ((list >> (sh + 8u)) << sh) | 0x00FF0000u;
        }
    } else if (type == T_BUY_RESERVE) {
        rv = set_byte(rv, __popc(~rv & 0x00808080u), card);  // appended, not re-inserted at its old index
    }
    if (type >= T_BUY_RESERVE) {
        const uint32_t info = T.card[card & 127u].y;
        cd -= 1u << (info & 31u);
        score -= (int)((info >> 5) & 7u);
    }
    if (noble != NONE8) {
        int count = 0;  // board.nobles.append(noble): first free nibble
#pragma unroll
        for (int k = 0; k < 5; k++) count += ((e.meta >> (4 * k)) & 15u) != 15u;
        e.meta = (e.meta & ~(15u << (4 * count))) | (noble << (4 * count));
        pi &= ~(1u << (16 + noble));
        score -= 3;
    }
    rv += (uint32_t)score << 24;
    pi = ((pi - 1u) & ~(1u << 26)) | (type == T_PASS ? 1u << 26 : 0u);
    e.meta = (e.meta & ~(3u << 20)) | ((uint32_t)seat << 20);  // the undone agent is to move again (inverse of apply_action)
#pragma unroll
    for (int p = 0; p < P; p++)
        if (seat == p) { e.pg[p] = g; e.pc[p] = cd; e.pr[p] = rv; e.pi[p] = pi; }
}

template <int P>
__device__ __forceinline__ int game_ends(const Env<P> &e, uint32_t flags)
{
    bool all100 = true, all_passed = true, any15 = false;
#pragma unroll
    for (int p = 0; p < P; p++) {
        all100 = all100 && (e.pi[p] & 0xFFFFu) == 100u;
        all_passed = all_passed && (e.pi[p] >> 26 & 1u);
        any15 = any15 || (e.pr[p] >> 24) >= 15u;
    }
    if ((flags & SPL_F_LIMIT_ROUNDS) && all100) return SPL_END_ROUNDCAP;
    if (any15 && seat_of(e.meta) == 0) return SPL_END_SCORE;
    return all_passed ? SPL_END_DEADLOCK : SPL_END_NONE;
}

// calScore x 2 for every seat (+1 = the half point, :308-323) and win(2)/tie(1)/loss(0)
template <int P>
__device__ __forceinline__ void final_scores(const Env<P> &e, int (&score2)[P], int (&outcome)[P])
{
    int sc[P], bought[P], mx = 0, min_cards = 1 << 20, victors = 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        sc[p] = (int)(e.pr[p] >> 24);
        bought[p] = (int)hsum6(e.pc[p]);
        mx = max(mx, sc[p]);
        min_cards = min(min_cards, bought[p]);
    }
#pragma unroll
    for (int p = 0; p < P; p++) victors += sc[p] == mx;
    int best = -1, nbest = 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        score2[p] = 2 * sc[p] + ((victors > 1 && sc[p] == mx && bought[p] == min_cards) ? 1 : 0);
        best = max(best, score2[p]);
    }
#pragma unroll
    for (int p = 0; p < P; p++) nbest += score2[p] == best;
#pragma unroll
    for (int p = 0; p < P; p++) outcome[p] = score2[p] == best ? (nbest > 1 ? 1 : 2) : 0;
}

// ------------------------------------------------------------------------------------------------ mask + observation
// The legal-action MASK (3510 bits) and the OBSERVATION (265 floats) are built by ONE THREAD PER ENV, in registers, and
// handed to the kernel as a finished row (kernels.cu stages the rows in shared memory and moves them to HBM in bulk).
// A first version used one warp per env; it spent ~500-800 warp-instructions per env with most lanes idle and reached only
// 6-7 % of HBM bandwidth (profiles/r1_kernels.md), so the rule work moved to one thread and only the data movement is shared.
constexpr int MASK_REGS = 112; // thread_fill_mask's row: 110 words + 2 spare for or_bits' overhang; every bit offset is a
                               // compile-time constant once its loops are unrolled, so the array lives in registers
constexpr int OBS_ROW_BYTES = 276;  // observation staged as BYTES: 265 entries + 4 spill bytes, 69-word (odd) row stride
// Every observation entry is an integer in 0..255 except three, which thread_fill_obs stashes and obs_entry decodes:
//   [1]  turns            16 bits in row[267..268]
//   [6]  np.var numerator 16 bits in row[265..266] (5*sum(x^2) - sum(x)^2 <= 4900), divided by 25 on the way out
//   [19..23] log(1+bp)    the table index bp in the byte, looked up on the way out

// OR a bit pattern into the thread's private mask words at a bit offset (a compile-time constant at every call site)
__device__ __forceinline__ void or_bits64(uint32_t *m, int start, uint64_t pat)
{
    const int w = start >> 5, sh = start & 31;
    SPL_CHECK(start >= 0 && w + 2 < MASK_REGS);
    const uint64_t lo = pat << sh;
    m[w] |= (uint32_t)lo;
    m[w + 1] |= (uint32_t)(lo >> 32);
    m[w + 2] |= sh ? (uint32_t)(pat >> (64 - sh)) : 0u;
}
__device__ __forceinline__ void or_bits32(uint32_t *m, int start, uint32_t pat)
{
    const int w = start >> 5, sh = start & 31;
    SPL_CHECK(start >= 0 && w + 1 < MASK_REGS);
    const uint64_t lo = (uint64_t)pat << sh;
    m[w] |= (uint32_t)lo;
    m[w + 1] |= (uint32_t)(lo >> 32);
}
// variants x three noble slots: bit (n*V + v) for every v in `variants` and n in the low 3 bits of `slots`
__device__ __forceinline__ uint64_t times_nobles(uint32_t variants, int V, uint32_t slots)
{
    const uint64_t v = variants;
    return ((slots & 1u) ? v : 0ull) | ((slots & 2u) ? v << V : 0ull) | ((slots & 4u) ? v << (2 * V) : 0ull);
}
// getLegalActions + create_legal_actions_mask (splendor_model.py:404-576, gym/envs/utils.py:148-164) for the seat to move,
// OR-ed into the zeroed row m.  Noble slots: NP has bit n set for every noble slot n the action may carry.
// `done(w0, w1)` is called as soon as words [w0, w1) (even bounds) are final, in ascending order and covering 0..110 with
// [0, 2) last: a kernel stores them right away, so only a quarter of the row is ever live in registers.
struct MaskKeep { __device__ __forceinline__ void operator()(int, int) const {} };
// NO_S: the caller guarantees c.S == 0 (no noble is waiting: true for almost every decision - a noble is claimed with the
// action that earns it, c.S != 0 needs two nobles earned at once).  Then every collect / reserve action carries noble slot 0
// only, the (noble x variant) patterns are the variant bits themselves, and the 64-bit shift / select chains of
// times_nobles fold away at compile time (they were 14 % of the mask kernel's instructions, profiles/r2_kernels.md).
template <int P, class Done = MaskKeep, bool NO_S = false>
__device__ __forceinline__ void thread_fill_mask(const Env<P> &e, const Tables &T, const Ctx &c, uint32_t *m, Done done = Done())
{
    SPL_CHECK(!NO_S || c.S == 0);
    const uint32_t NP = NO_S ? 1u : (c.S ? c.S << 1 : 1u);
    bool any = false;
    // reserve (:498-520): the same 42-bit (noble x variant) pattern for every occupied slot
    if (c.n_resv < 3 && c.occ) {
        uint32_t vr = 1u;
        if (c.bank_yellow) {
            vr = 2u;
            if (c.hold >= 10) {
                vr = 0;
#pragma unroll
                for (int n = 0; n < 5; n++) vr |= (c.h1 >> field_of_normal(n) & 1u) << (2 + n);
            }
        }
        const uint64_t pat = times_nobles(vr, 7, NP) | times_nobles(vr, 7, NP >> 3) << 21;
#pragma unroll
        for (int p = 0; p < 12; p++) or_bits64(m, BASE_RESERVE + 42 * p, (c.occ >> p & 1u) ? pat : 0ull);
        any = true;
    }
    // collect_same (:475-496) and collect_diff (:437-473): every group computes its pattern unconditionally (zero when
    // the group is not available) - uniform control flow beats skipping work when 32 different games share a warp
    const uint32_t np_lo = NP & 7u, np_hi = NP >> 3;
    uint32_t any_bits = 0;
    {
        const ExSel s2 = make_exsel(c.hold - 8);
#pragma unroll
        for (int col = 0; col < 5; col++) {
            const uint32_t vs = same_variant_bits_sel(T, c, col, s2) & ((c.A4 >> field_of_normal(col) & 1u) ? ~0u : 0u);
            or_bits64(m, BASE_SAME + 126 * col, times_nobles(vs, 21, np_lo));
            if (np_hi) or_bits64(m, BASE_SAME + 126 * col + 63, times_nobles(vs, 21, np_hi));  // nobles at list position >= 2: rare
            any_bits |= vs;
        }
    }
    done(2, 34);  // reserve and collect_same end at bit 1139
    const ExSel sd1 = make_exsel(c.hold + 1 - 10), sd2 = make_exsel(c.hold + 2 - 10), sd3 = make_exsel(c.hold + 3 - 10);
#pragma unroll
    for (int ci = 0; ci < 25; ci++) {
        const int L = ci < 5 ? 1 : (ci < 15 ? 2 : 3), V = L == 1 ? 6 : (L == 2 ? 15 : 20);
        const uint32_t cm = T.combo_mask[ci];
        const bool valid = !(cm & ~c.A1) && L >= c.Lmin;
        const uint32_t vd_all = L == 1 ? diff_variant_bits_sel<1>(T, c, ci, sd1)
                                       : (L == 2 ? diff_variant_bits_sel<2>(T, c, ci, sd2) : diff_variant_bits_sel<3>(T, c, ci, sd3));
        const uint32_t vd = valid ? vd_all : 0u;
        const int base = BASE_DIFF + (L == 1 ? 36 * ci : (L == 2 ? 180 + 90 * (ci - 5) : 1080 + 120 * (ci - 15)));
        if (L == 1) or_bits64(m, base, times_nobles(vd, V, np_lo) | times_nobles(vd, V, np_hi) << 18);
        else {
            or_bits64(m, base, times_nobles(vd, V, np_lo));
            if (np_hi) or_bits64(m, base + 3 * V, times_nobles(vd, V, np_hi));
        }
        any_bits |= vd;
        if (ci == 14) done(34, 68);  // lengths 1 and 2 end at bit 2219
        if (ci == 19) done(68, 88);  // five of the ten triples: bit 2819
    }
    any = any || any_bits;
    // buy (:522-568): 6 noble bits per card; board slots 0-4 / 5-9 / 10-11 and the reserved cards as four 32-bit patterns
    {
        const uint32_t eff = (c.gems & ~(31u << YSH)) + c.cards;
        const uint32_t cap = ((c.cards + 9u * ONES6) & H6) | (16u << YSH);
        uint32_t b0 = 0, b1 = 0, b2 = 0, br = 0;
#pragma unroll
        for (int s = 0; s < 15; s++) {
            const uint32_t w = s < 12 ? e.d[s >> 2] : c.resv;
            const uint2 cw = T.card[(w >> (8 * (s & 3))) & 127u];
            const uint32_t sh = cw.y & 31u;
            const bool ok = hsum6_small(satsub6(cw.x, eff)) <= (uint32_t)c.yellow && !((cap >> sh) & 16u);
            const uint32_t Sb = c.S | ((c.near >> sh) & 31u);
            const uint32_t np = ok ? (Sb ? Sb << 1 : 1u) : 0u;
            if (s < 5) b0 |= np << (6 * s);
            else if (s < 10) b1 |= np << (6 * (s - 5));
            else if (s < 12) b2 |= np << (6 * (s - 10));
            else br |= np << (6 * (s - 12));
        }
        or_bits32(m, BASE_BUY_AVAILABLE, b0);
        or_bits32(m, BASE_BUY_AVAILABLE + 30, b1);
        or_bits32(m, BASE_BUY_AVAILABLE + 60, b2);
        or_bits32(m, BASE_BUY_RESERVE, br);
        any = any || (b0 | b1 | b2 | br);
    }
    done(88, 110);
    if (!any) m[0] |= NP;  // pass, only when nothing else is legal (:570-574)
    done(0, 2);
}

// insert d into the ascending 4-tuple held in the bytes of t (keeps the four smallest)
__device__ __forceinline__ uint32_t insert_sorted4(uint32_t t, uint32_t d)
{
    const uint32_t a0 = t & 0xFFu, a1 = (t >> 8) & 0xFFu, a2 = (t >> 16) & 0xFFu, a3 = t >> 24;
    const uint32_t t0 = min(a0, d), r0 = max(a0, d), t1 = min(a1, r0), r1 = max(a1, r0), t2 = min(a2, r1), r2 = max(a2, r1);
    return t0 | t1 << 8 | t2 << 16 | min(a3, r2) << 24;
}

// float32 value of observation entry w from a staged byte row
__device__ __forceinline__ float obs_entry(const uint8_t *row, int w, const uint32_t *log1p_bits)
{
    const uint32_t b = row[w];
    float f = (float)b;
    if (w == 1) f = (float)(row[267] | row[268] << 8);
    // np.var of five small integers = (5*sum(x^2) - sum(x)^2) / 25; numerator < 2^24, IEEE division: exact rounding
    if (w == 6) f = __fdiv_rn((float)(row[265] | row[266] << 8), 25.0f);
    if (w >= 19 && w < 24) f = __uint_as_float(log1p_bits[b & 31u]);
    return f;
}

// features.extract_metrics_with_cards(state, seat).astype(float32) (features.py:214-310, :377-480) into the zeroed staged
// row o (bytes packed four to a word, assembled in registers)
// staged observation words: OBS_ROW_BYTES / 4 registers per thread; every byte offset below is a compile-time constant
constexpr int OBS_WORDS = OBS_ROW_BYTES / 4;
__device__ __forceinline__ void put_byte(uint32_t *o, int idx, uint32_t v) { o[idx >> 2] |= v << (8 * (idx & 3)); }
__device__ __forceinline__ void put_bytes64(uint32_t *o, int idx, uint64_t pat)  // up to 8 bytes starting at byte idx
{
    const int w = idx >> 2, sh = 8 * (idx & 3);
    const uint64_t lo = pat << sh;
    o[w] |= (uint32_t)lo;
    o[w + 1] |= (uint32_t)(lo >> 32);
    o[w + 2] |= sh ? (uint32_t)(pat >> (64 - sh)) : 0u;
}

// `done(w0, w1)` is called when words [w0, w1) are final: the card vectors as each card is finished, the metrics last
// (TT: Tables, or any struct with its card / noble / cardvec members - the observation kernel stages only those)
template <int P, class Done = MaskKeep, class TT = Tables>
__device__ __forceinline__ void thread_fill_obs(const Env<P> &e, const TT &T, int seat, uint32_t *o, Done done = Done())
{
    uint32_t gems = e.pg[0], cards = e.pc[0], resv = e.pr[0], pinfo = e.pi[0];
#pragma unroll
    for (int p = 1; p < P; p++)
        if (seat == p) { gems = e.pg[p]; cards = e.pc[p]; resv = e.pr[p]; pinfo = e.pi[p]; }
    const uint32_t bp = (gems & ~(31u << YSH)) + cards;  // buying power per normal colour (features.py:65-74), <= 14
    // per colour, the four cheapest candidate cards (board + own reserve) by missing gems; absent ones count 14
    // (MAX_CARD_GEMS_DIST_3, features.py:257-274) - exactly what "sorted distances, padded" needs for the nobles
    uint32_t best[6] = {0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu, 0x0E0E0E0Eu};
#pragma unroll
    for (int s = 0; s < 15; s++) {
        const uint32_t w = s < 12 ? e.d[s >> 2] : resv;
        const uint32_t id = (w >> (8 * (s & 3))) & 0xFFu;
        const bool present = id != NONE8;
        const uint2 cw = T.card[id & 127u];  // an empty slot reads the all-zero filler and contributes nothing
        const uint32_t miss = satsub6(cw.x, bp), sum = hsum6_small(miss), sh = cw.y & 31u;
        uint32_t mx = 0;
#pragma unroll
        for (int f = 0; f < 6; f++) mx = max(mx, field(miss, 5 * f));
        const uint32_t turns = sum == 0 ? 0u : max((sum + 2u) / 3u, mx);  // features.py:124-135
        put_byte(o, (s < 12 ? 30 : 54 - 12) + s, sum);
        put_byte(o, (s < 12 ? 42 : 57 - 12) + s, turns);
        {  // insert into this card's colour tuple only: select, insert, write back
            const uint32_t t = insert_sorted4(sh == 0u ? best[0] : (sh == 5u ? best[1] : (sh == 15u ? best[3] : (sh == 20u ? best[4] : best[5]))), sum);
            best[0] = (present && sh == 0u) ? t : best[0]; best[1] = (present && sh == 5u) ? t : best[1];
            best[3] = (present && sh == 15u) ? t : best[3]; best[4] = (present && sh == 20u) ? t : best[4];
            best[5] = (present && sh == 25u) ? t : best[5];
        }
        // vectorize_card (features.py:377-413): one-hot over sklearn's alphabetical categories black, blue, green, red,
        // white, yellow (bytes 0..5); the cost in alphabetical order black, blue, green, red, white (6..10); deck_id; points
        const uint4 vec = T.cardvec[id & 127u];  // a table row per card id (the 13 bytes are a pure function of the card)
        const uint64_t head = (uint64_t)vec.x | (uint64_t)vec.y << 32, tail = (uint64_t)vec.z | (uint64_t)vec.w << 32;
        put_bytes64(o, 70 + 13 * s, head);      // bytes 0..7: one-hot, cost black, cost blue
        put_bytes64(o, 70 + 13 * s + 8, tail);  // bytes 8..12: cost green, red, white, deck_id, points
        done(s == 0 ? 18 : (70 + 13 * s) >> 2, (70 + 13 * (s + 1)) >> 2);  // later cards start at byte 70 + 13 (s + 1)
    }
    // distance to each board noble (features.py:257-274)
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t id = (e.meta >> (4 * k)) & 15u;
        const uint32_t missing = id == 15u ? 0u : satsub6(T.noble[id], cards);
        uint32_t dist = 0;
#pragma unroll
        for (int f = 0; f < 6; f++) {
            if (f == 2) continue;
            const uint32_t cnt = field(missing, 5 * f), t = best[f];
            dist += (cnt >= 1 ? t & 0xFFu : 0u) + (cnt >= 2 ? (t >> 8) & 0xFFu : 0u) + (cnt >= 3 ? (t >> 16) & 0xFFu : 0u) +
                    (cnt >= 4 ? t >> 24 : 0u);
        }
        put_byte(o, 60 + k, dist);
        put_byte(o, 65 + k, hsum6_small(missing));
    }
    // the scalar metrics (features.py:284-299)
    const uint32_t score = resv >> 24;
    put_byte(o, 0, 1);
    put_byte(o, 267, pinfo & 0xFFu); put_byte(o, 268, (pinfo >> 8) & 0xFFu);
    put_byte(o, 2, score);
    put_byte(o, 3, score >= 15u ? 1u : 0u);
    put_byte(o, 4, hsum6(cards));
    put_byte(o, 5, (uint32_t)__popc(~resv & 0x00808080u));
    int s1 = 0, s2 = 0;
#pragma unroll
    for (int n = 0; n < 5; n++) {
        const int sh = 5 * field_of_normal(n);
        const int b = (int)field(bp, sh);
        s1 += b; s2 += b * b;
        put_byte(o, 9 + n, field(cards, sh));
        put_byte(o, 14 + n, (uint32_t)b);
        put_byte(o, 19 + n, (uint32_t)b);
    }
    const uint32_t var_num = (uint32_t)(5 * s2 - s1 * s1);
    put_byte(o, 265, var_num & 0xFFu); put_byte(o, 266, (var_num >> 8) & 0xFFu);
    put_byte(o, 7, field(gems, YSH));
    put_byte(o, 8, hsum6(gems));
#pragma unroll
    for (int p = 0; p < P; p++) {  // features.py:276-282, its i-1 indexing quirk kept
        const uint32_t sc = e.pr[p] >> 24;
        if (p < 3) put_byte(o, 24 + p, p < seat ? sc : 0u);
        if (p >= 1) put_byte(o, 27 + p - 1, p > seat ? sc : 0u);
    }
    done(0, 18);
    done((70 + 13 * 15) >> 2, OBS_WORDS);
}

}  // namespace spl

/* oracle/splendor_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference's game-rule path (roeey777/Splendor-AI, all citations relative to
 * /root/reference/src/splendor/):
 *   state + dealing          splendor/splendor_model.py:52-138
 *   getLegalActions          splendor/splendor_model.py:404-576  (noble_visit :397, resources_sufficient :371,
 *                                                                generate_return_combos :329)
 *   generateSuccessor/update splendor/splendor_model.py:222-295, template.py:47-53
 *   gameEnds / LimitRounds   splendor/splendor_model.py:299-305, splendor/utils.py:14-25
 *   calScore                 splendor/splendor_model.py:308-323
 *   ALL_ACTIONS index        splendor/gym/envs/actions.py:109-237, gym/envs/utils.py:148-181
 *   observation vector       splendor/features.py:214-310, :377-480
 *   RandomAgent              agents/generic/random.py:21-27
 *   MiniMaxAgent             agents/our_agents/minmax.py:32-136 (float64 evaluation, depth-2 search)
 *
 * PARITY PIN: the reference ships no tests or golden vectors.  This oracle is pinned against
 * the reference itself, executed in the build container: oracle/gen_golden.py drives the unmodified Python engine
 * and records per-step states, legal ALL_ACTIONS index sets, rewards, terminal flags, final scores and observation
 * vectors into tests/golden/; tests/test_oracle_golden.py replays them through this file bit-exactly.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (splendor-ai_b200/) never calls it and has no CPU fallback.
 *
 * Style: deliberately close to the reference's control flow (enumerate actions family by family, materialise the
 * list, map each to its ALL_ACTIONS index), NOT the count/select formulation the CUDA kernels use, so that the two
 * are independent derivations of the same rules.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include "../include/spl_tables.h"

#define NONE 0xFF

/* ------------------------------------------------------------------ tables */
typedef struct { uint8_t colour, points, cost[5]; } card_t;
static const card_t CARD[SPL_NUM_CARDS] = {
#define X(col, pts, c0, c1, c2, c3, c4) {col, pts, {c0, c1, c2, c3, c4}},
    SPL_CARD_ROWS
#undef X
};
static const uint8_t NOBLE_COST[SPL_NUM_NOBLES][5] = {
#define X(c0, c1, c2, c3, c4) {c0, c1, c2, c3, c4},
    SPL_NOBLE_ROWS
#undef X
};
static const uint32_t LOG1P_BITS[32] = SPL_LOG1P_BITS;
static const int TIER_BASE[3] = {0, 40, 70};
static const int TIER_SIZE[3] = {40, 30, 20};
static int card_tier(int id) { return id < 40 ? 0 : (id < 70 ? 1 : 2); }

/* ALL_GEMS_COLORS order (actions.py:25): black, red, yellow, green, blue, white -> our colour ids */
static const int ALLG[6] = {0, 1, 5, 2, 3, 4};

/* ------------------------------------------------------------------ state */
typedef struct {
    uint8_t gems[6];      /* black red green blue white yellow */
    uint8_t cards[5];     /* bought cards per colour (len(agent.cards[colour])) */
    uint8_t resv[3];      /* agent.cards["yellow"]: reserved card ids in list order */
    uint8_t n_resv;
    uint8_t score;
    uint8_t passed;
    uint8_t n_nobles;     /* len(agent.nobles) */
    uint8_t pad;
    uint16_t nobles_mask; /* which noble ids the agent holds */
    uint16_t turns;       /* len(agent.agent_trace.action_reward) */
} orc_player;

typedef struct {
    uint8_t P, cur;
    uint8_t bank[6];
    uint8_t dealt[12];    /* tier*4+col, NONE = empty slot */
    uint8_t nobles[5];    /* board noble ids in LIST order (deletion shifts later ones down) */
    uint8_t n_nobles;
    uint8_t deck[3][40];  /* reference list order: deal() pops from the END (splendor_model.py:95-99) */
    uint8_t deck_n[3];
    uint8_t pad;
    orc_player pl[4];
} orc_state;

int orc_sizeof_state(void) { return (int)sizeof(orc_state); }

static int deal(orc_state *s, int tier)
{
    if (s->deck_n[tier] == 0) return NONE;
    return s->deck[tier][--s->deck_n[tier]];
}

/* splendor_model.py:69-93.  deck_lists = the three shuffled lists (40+30+20 card ids, reference list order),
 * nobles = the P+1 sampled noble ids in list order. */
void orc_init(orc_state *s, int P, const uint8_t *deck_lists, const uint8_t *nobles)
{
    memset(s, 0, sizeof *s);
    s->P = (uint8_t)P;
    int n = (P == 2) ? 4 : (P == 3 ? 5 : 7);
    for (int c = 0; c < 5; c++) s->bank[c] = (uint8_t)n;
    s->bank[5] = 5;
    for (int i = 0; i < 5; i++) s->nobles[i] = NONE;
    s->n_nobles = (uint8_t)(P + 1);
    for (int i = 0; i < P + 1; i++) s->nobles[i] = nobles[i];
    for (int t = 0; t < 3; t++) {
        memcpy(s->deck[t], deck_lists + TIER_BASE[t], (size_t)TIER_SIZE[t]);
        s->deck_n[t] = (uint8_t)TIER_SIZE[t];
    }
    for (int t = 0; t < 3; t++)
        for (int j = 0; j < 4; j++) s->dealt[t * 4 + j] = (uint8_t)deal(s, t);
    for (int p = 0; p < 4; p++)
        for (int i = 0; i < 3; i++) s->pl[p].resv[i] = NONE;
}

/* ------------------------------------------------------------------ ALL_ACTIONS closed-form index */
static int popc(unsigned x) { return __builtin_popcount(x); }

/* rank of a sorted position tuple within itertools.combinations_with_replacement(range(k), r) */
static int cwr_rank(int k, int r, const int *pos)
{
    int rank = 0, q[3];
    if (r == 1) {
        for (q[0] = 0; q[0] < k; q[0]++, rank++)
            if (q[0] == pos[0]) return rank;
    } else if (r == 2) {
        for (q[0] = 0; q[0] < k; q[0]++)
            for (q[1] = q[0]; q[1] < k; q[1]++, rank++)
                if (q[0] == pos[0] && q[1] == pos[1]) return rank;
    } else {
        for (q[0] = 0; q[0] < k; q[0]++)
            for (q[1] = q[0]; q[1] < k; q[1]++)
                for (q[2] = q[1]; q[2] < k; q[2]++, rank++)
                    if (q[0] == pos[0] && q[1] == pos[1] && q[2] == pos[2]) return rank;
    }
    return -1;
}
static int cwr_count(int k, int r) { return r == 0 ? 1 : (r == 1 ? k : (r == 2 ? k * (k + 1) / 2 : k * (k + 1) * (k + 2) / 6)); }

/* rank of a colour set (bitmask over colour ids 0..4) within itertools.combinations(NORMAL_COLORS, L) */
static int comb_rank(unsigned mask, int L)
{
    int rank = 0;
    int a, b, c;
    if (L == 1) {
        for (a = 0; a < 5; a++, rank++) if (mask == (1u << a)) return rank;
    } else if (L == 2) {
        for (a = 0; a < 5; a++) for (b = a + 1; b < 5; b++, rank++) if (mask == ((1u << a) | (1u << b))) return rank;
    } else {
        for (a = 0; a < 5; a++) for (b = a + 1; b < 5; b++) for (c = b + 1; c < 5; c++, rank++)
            if (mask == ((1u << a) | (1u << b) | (1u << c))) return rank;
    }
    return -1;
}

/* positions (in ALL_GEMS order minus the collected colours) of a returned multiset ret[6] (counts per colour id);
 * returns r = number of gems returned, pos[] sorted ascending. */
static int returned_positions(unsigned collected_mask, const uint8_t *ret, int *pos, int *k_out)
{
    int k = 0, r = 0;
    for (int i = 0; i < 6; i++) {
        int c = ALLG[i];
        if (collected_mask >> c & 1) continue;
        for (int j = 0; j < ret[c]; j++) pos[r++] = k;
        k++;
    }
    *k_out = k;
    return r;
}

/* actions.py:177-199 */
static int idx_collect_diff(unsigned combo_mask, int n, const uint8_t *ret)
{
    static const int BASE[3] = {0, 180, 1080}, V[3] = {6, 15, 20};
    int L = popc(combo_mask), pos[3], k;
    int r = returned_positions(combo_mask, ret, pos, &k);
    int v = 0;
    if (r > 0) {
        v = 1;
        for (int rr = 1; rr < r; rr++) v += cwr_count(k, rr);
        v += cwr_rank(k, r, pos);
    }
    return 1140 + BASE[L - 1] + comb_rank(combo_mask, L) * 6 * V[L - 1] + n * V[L - 1] + v;
}

/* actions.py:145-174 */
static int idx_collect_same(int c, int n, const uint8_t *ret)
{
    int pos[3], k, v = 0;
    int r = returned_positions(1u << c, ret, pos, &k);
    if (r == 1) v = 11 + pos[0];
    else if (r == 2 && pos[0] == pos[1]) v = 16 + pos[0];
    else if (r == 2) { /* rank within combinations(range(5), 2) */
        int rank = 0;
        for (int a = 0; a < 5; a++) for (int b = a + 1; b < 5; b++, rank++) if (a == pos[0] && b == pos[1]) v = 1 + rank;
    }
    return 510 + c * 126 + n * 21 + v;
}

/* actions.py:115-142 */
static int idx_reserve(int p, int n, int took_yellow, const uint8_t *ret)
{
    int v = took_yellow ? 1 : 0;
    for (int c = 0; c < 5; c++) if (ret[c]) v = 2 + c;
    return 6 + p * 42 + n * 7 + v;
}

/* ------------------------------------------------------------------ rules */
/* splendor_model.py:397-402 */
static int noble_visit(const uint8_t *cards, int noble_id)
{
    for (int c = 0; c < 5; c++) if (cards[c] < NOBLE_COST[noble_id][c]) return 0;
    return 1;
}

/* splendor_model.py:371-394 : returns 1 and fills pay[6] if affordable */
static int resources_sufficient(const orc_player *a, const uint8_t *cost, uint8_t *pay)
{
    int wild = a->gems[5];
    memset(pay, 0, 6);
    for (int c = 0; c < 5; c++) {
        if (cost[c] == 0) continue; /* the reference iterates only over colours present in the cost dict */
        int available = a->gems[c] + a->cards[c];
        int shortfall = cost[c] - available; if (shortfall < 0) shortfall = 0;
        wild -= shortfall;
        if (wild < 0) return 0;
        int gem_cost = cost[c] - a->cards[c]; if (gem_cost < 0) gem_cost = 0;
        int gem_short = gem_cost - a->gems[c]; if (gem_short < 0) gem_short = 0;
        pay[c] = (uint8_t)(gem_cost - gem_short);
        pay[5] = (uint8_t)(pay[5] + gem_short);
    }
    return 1;
}

/* all multisets of size r (0..3) over the 6 colour ids, as count vectors; returns how many (1, 6, 21, 56) */
static int multisets(int r, uint8_t out[56][6])
{
    int n = 0;
    if (r == 0) { memset(out[0], 0, 6); return 1; }
    for (int x = 0; x < 6; x++)
        for (int y = (r >= 2 ? x : 5); y < 6; y++)
            for (int z = (r >= 3 ? y : 5); z < 6; z++) {
                memset(out[n], 0, 6);
                out[n][x]++;
                if (r >= 2) out[n][y]++;
                if (r >= 3) out[n][z]++;
                n++;
            }
    return n;
}

/* splendor_model.py:329-367 : enumerate all return multisets; out[i][6] counts. returns number of combos. */
static int return_combos(const orc_player *a, unsigned collected_mask, int n_collected, uint8_t out[64][6])
{
    int hold = 0;
    for (int c = 0; c < 6; c++) hold += a->gems[c];
    int total = hold + n_collected;
    if (total <= 10) { memset(out[0], 0, 6); return 1; }
    int num_return = total - 10, n = 0, avail = 0;
    for (int c = 0; c < 6; c++) if (!(collected_mask >> c & 1)) avail += a->gems[c];
    if (avail < num_return) return 0;
    /* every distinct multiset of num_return gems the agent holds, in colours it did not just collect */
    uint8_t all[56][6];
    int m = multisets(num_return, all);
    for (int i = 0; i < m; i++) {
        int ok = 1;
        for (int c = 0; c < 6; c++)
            if (all[i][c] > a->gems[c] || (all[i][c] && (collected_mask >> c & 1))) ok = 0;
        if (ok) memcpy(out[n++], all[i], 6);
    }
    return n;
}

static int cmp_u16(const void *x, const void *y) { return (int)*(const uint16_t *)x - (int)*(const uint16_t *)y; }

/* getLegalActions (splendor_model.py:404-576) -> ascending ALL_ACTIONS indices. out must hold >= 1024 entries. */
int orc_legal(const orc_state *s, int seat, uint16_t *out)
{
    const orc_player *a = &s->pl[seat];
    int n_out = 0;
    int nslot[6], n_nslot = 0; /* noble slots to emit for non-buy actions (0 = None, k+1 = board position k) */
    for (int k = 0; k < s->n_nobles; k++) if (noble_visit(a->cards, s->nobles[k])) nslot[n_nslot++] = k + 1;
    if (n_nslot == 0) nslot[n_nslot++] = 0;

    uint8_t combos[64][6];
    int hold = 0;
    for (int c = 0; c < 6; c++) hold += a->gems[c];

    /* collect_diff :437-473 */
    int avail[5], n_avail = 0;
    for (int c = 0; c < 5; c++) if (s->bank[c] > 0) avail[n_avail++] = c;
    int min_len = hold <= 7 ? (n_avail < 3 ? n_avail : 3) : (hold == 8 ? (n_avail < 2 ? n_avail : 2) : (n_avail < 1 ? n_avail : 1));
    int max_len = n_avail < 3 ? n_avail : 3;
    for (int len = min_len; len <= max_len; len++) {
        if (len == 0) continue; /* "collected_gems == {}" is skipped :456 */
        for (unsigned sub = 1; sub < (1u << n_avail); sub++) {
            if (popc(sub) != len) continue;
            unsigned cm = 0;
            for (int i = 0; i < n_avail; i++) if (sub >> i & 1) cm |= 1u << avail[i];
            int nc = return_combos(a, cm, len, combos);
            for (int r = 0; r < nc; r++)
                for (int j = 0; j < n_nslot; j++) out[n_out++] = (uint16_t)idx_collect_diff(cm, nslot[j], combos[r]);
        }
    }
    /* collect_same :475-496 */
    for (int c = 0; c < 5; c++) {
        if (s->bank[c] < 4) continue;
        int nc = return_combos(a, 1u << c, 2, combos);
        for (int r = 0; r < nc; r++)
            for (int j = 0; j < n_nslot; j++) out[n_out++] = (uint16_t)idx_collect_same(c, nslot[j], combos[r]);
    }
    /* reserve :498-520 */
    if (a->n_resv < 3) {
        int yellow = s->bank[5] > 0;
        int nc = return_combos(a, yellow ? (1u << 5) : 0, yellow ? 1 : 0, combos);
        for (int r = 0; r < nc; r++)
            for (int p = 0; p < 12; p++) {
                if (s->dealt[p] == NONE) continue;
                for (int j = 0; j < n_nslot; j++) out[n_out++] = (uint16_t)idx_reserve(p, nslot[j], yellow, combos[r]);
            }
    }
    /* buy :522-568 : 12 dealt slots then the reserved list */
    for (int row = 0; row < 4; row++)
        for (int col = 0; col < (row < 3 ? 4 : a->n_resv); col++) {
            int id = row < 3 ? s->dealt[row * 4 + col] : a->resv[col];
            if (id == NONE) continue;
            const card_t *cd = &CARD[id];
            if (a->cards[cd->colour] == 7) continue;
            uint8_t pay[6];
            if (!resources_sufficient(a, cd->cost, pay)) continue;
            uint8_t after[5];
            memcpy(after, a->cards, 5);
            after[cd->colour]++;
            int emitted = 0;
            for (int k = 0; k < s->n_nobles; k++)
                if (noble_visit(after, s->nobles[k])) {
                    out[n_out++] = (uint16_t)((row < 3 ? 3438 + (row * 4 + col) * 6 : 3420 + col * 6) + k + 1);
                    emitted = 1;
                }
            if (!emitted) out[n_out++] = (uint16_t)(row < 3 ? 3438 + (row * 4 + col) * 6 : 3420 + col * 6);
        }
    /* pass :570-574 */
    if (n_out == 0)
        for (int j = 0; j < n_nslot; j++) out[n_out++] = (uint16_t)nslot[j];
    qsort(out, (size_t)n_out, sizeof(uint16_t), cmp_u16);
    return n_out;
}

/* 3510-bit mask (110 x u32), bit i of word i/32 = ALL_ACTIONS index i legal (gym/envs/utils.py:148-164) */
int orc_legal_mask(const orc_state *s, int seat, uint32_t *mask)
{
    uint16_t idx[1024];
    int n = orc_legal(s, seat, idx);
    memset(mask, 0, SPL_MASK_WORDS * 4);
    for (int i = 0; i < n; i++) mask[idx[i] >> 5] |= 1u << (idx[i] & 31);
    return n;
}

/* decoded action (the fields of the reference's action dict that generateSuccessor reads) */
typedef struct { int type; int p; int n; uint8_t collected[6], returned[6]; } action_t;
enum { T_PASS, T_RESERVE, T_SAME, T_DIFF, T_BUY_RESERVE, T_BUY_AVAILABLE };

/* build_action (gym/envs/utils.py:43-145) restated by brute force: find the unique (family, fields) whose closed-form
 * index equals idx.  Returns 0 on success. */
static int decode_action(const orc_state *s, int seat, int idx, action_t *A)
{
    const orc_player *a = &s->pl[seat];
    memset(A, 0, sizeof *A);
    if (idx < 0 || idx >= SPL_NUM_ACTIONS) return -1;
    if (idx < 6) { A->type = T_PASS; A->n = idx; return 0; }
    if (idx < 510) {
        int r = idx - 6; A->type = T_RESERVE; A->p = r / 42; A->n = (r % 42) / 7; int v = r % 7;
        if (v >= 1) A->collected[5] = 1;
        if (v >= 2) A->returned[v - 2] = 1;
        return 0;
    }
    if (idx < 1140) {
        int r = idx - 510, c = r / 126; A->type = T_SAME; A->p = c; A->n = (r % 126) / 21;
        A->collected[c] = 2;
        uint8_t ret[6];
        /* search the <=21 variants */
        for (int x = 0; x < 6; x++) for (int y = x; y < 6; y++) for (int cntmode = 0; cntmode < 3; cntmode++) {
            memset(ret, 0, 6);
            if (cntmode == 0) { if (x || y) continue; }
            else if (cntmode == 1) { if (x != y) continue; ret[x] = 1; }
            else { ret[x]++; ret[y]++; }
            if (ret[c]) continue;
            if (idx_collect_same(c, A->n, ret) == idx) { memcpy(A->returned, ret, 6); return 0; }
        }
        return -1;
    }
    if (idx < 3420) {
        static const int BASE[3] = {0, 180, 1080}, V[3] = {6, 15, 20};
        int r = idx - 1140, L = r < 180 ? 1 : (r < 1080 ? 2 : 3);
        r -= BASE[L - 1];
        int rank = r / (6 * V[L - 1]), n = (r % (6 * V[L - 1])) / V[L - 1];
        A->type = T_DIFF; A->n = n;
        for (unsigned cm = 1; cm < 32; cm++) {
            if (popc(cm) != L || comb_rank(cm, L) != rank) continue;
            A->p = (int)cm;
            for (int c = 0; c < 5; c++) A->collected[c] = (uint8_t)(cm >> c & 1);
            for (int nr = 0; nr <= L; nr++) {
                uint8_t all[56][6];
                int m = multisets(nr, all);
                for (int i = 0; i < m; i++) {
                    int bad = 0;
                    for (int c = 0; c < 6; c++) if (all[i][c] && (cm >> c & 1)) bad = 1;
                    if (!bad && idx_collect_diff(cm, n, all[i]) == idx) { memcpy(A->returned, all[i], 6); return 0; }
                }
            }
        }
        return -1;
    }
    if (idx < 3438) { A->type = T_BUY_RESERVE; A->p = (idx - 3420) / 6; A->n = (idx - 3420) % 6; }
    else { A->type = T_BUY_AVAILABLE; A->p = (idx - 3438) / 6; A->n = (idx - 3438) % 6; }
    (void)a;
    return 0;
}

/* generateSuccessor + GameRule.update (splendor_model.py:222-295, template.py:47-53).
 * Returns the acting agent's score delta (the env's reward, splendor_env.py:142-161), or -1 if idx is not in the
 * legal set (mirrors the ValueError at splendor_env.py:139-153); the state is untouched in that case. */
int orc_step(orc_state *s, int idx)
{
    uint16_t legal[1024];
    int seat = s->cur, n = orc_legal(s, seat, legal), found = 0;
    for (int i = 0; i < n; i++) if (legal[i] == idx) found = 1;
    if (!found) return -1;
    action_t A;
    if (decode_action(s, seat, idx, &A)) return -1;
    orc_player *a = &s->pl[seat];
    int score = 0, card = NONE;

    if (A.type == T_SAME || A.type == T_DIFF || A.type == T_RESERVE) {
        for (int c = 0; c < 6; c++) { s->bank[c] -= A.collected[c]; a->gems[c] += A.collected[c]; }
        for (int c = 0; c < 6; c++) { a->gems[c] -= A.returned[c]; s->bank[c] += A.returned[c]; }
        if (A.type == T_RESERVE) {
            card = s->dealt[A.p];
            s->dealt[A.p] = (uint8_t)deal(s, card_tier(card));
            a->resv[a->n_resv++] = (uint8_t)card;
        }
    } else if (A.type == T_BUY_AVAILABLE || A.type == T_BUY_RESERVE) {
        card = A.type == T_BUY_AVAILABLE ? s->dealt[A.p] : a->resv[A.p];
        uint8_t pay[6];
        resources_sufficient(a, CARD[card].cost, pay);
        for (int c = 0; c < 6; c++) { a->gems[c] -= pay[c]; s->bank[c] += pay[c]; }
        if (A.type == T_BUY_AVAILABLE) s->dealt[A.p] = (uint8_t)deal(s, card_tier(card));
        else {
            for (int i = A.p; i < 2; i++) a->resv[i] = a->resv[i + 1];
            a->resv[2] = NONE;
            a->n_resv--;
        }
        a->cards[CARD[card].colour]++;
        score += CARD[card].points;
    }
    if (A.n > 0) {
        int k = A.n - 1, id = s->nobles[k];
        for (int i = k; i < 4; i++) s->nobles[i] = s->nobles[i + 1];
        s->nobles[4] = NONE;
        s->n_nobles--;
        a->nobles_mask |= (uint16_t)(1u << id);
        a->n_nobles++;
        score += 3;
    }
    a->turns++;
    a->score = (uint8_t)(a->score + score);
    a->passed = (A.type == T_PASS);
    s->cur = (uint8_t)((s->cur + 1) % s->P);
    return score;
}

/* gameEnds (splendor_model.py:299-305); limit_rounds!=0 adds LimitRoundsGameRule (splendor/utils.py:14-25).
 * returns 0 = not ended, 1 = score>=15 at round end, 2 = deadlock (all passed), 3 = 100-round cap */
int orc_game_ends(const orc_state *s, int limit_rounds)
{
    if (limit_rounds) {
        int all = 1;
        for (int i = 0; i < s->P; i++) if (s->pl[i].turns != 100) all = 0;
        if (all) return 3;
    }
    int deadlock = 0;
    for (int i = 0; i < s->P; i++) {
        deadlock += s->pl[i].passed ? 1 : 0;
        if (s->pl[i].score >= 15 && s->cur == 0) return 1;
    }
    return deadlock == s->P ? 2 : 0;
}

/* calScore x 2 (splendor_model.py:308-323) */
int orc_cal_score2(const orc_state *s, int agent)
{
    int max_score = 0, n_victors = 0, min_cards = 1 << 30, mine = 0;
    for (int i = 0; i < s->P; i++) {
        int bought = 0;
        for (int c = 0; c < 5; c++) bought += s->pl[i].cards[c];
        if (bought < min_cards) min_cards = bought; /* minimum over ALL agents, not only victors (:319) */
        if (i == agent) mine = bought;
        if (s->pl[i].score > max_score) max_score = s->pl[i].score;
    }
    for (int i = 0; i < s->P; i++) if (s->pl[i].score == max_score) n_victors++;
    if (n_victors > 1 && s->pl[agent].score == max_score && mine == min_cards) return 2 * s->pl[agent].score + 1;
    return 2 * s->pl[agent].score;
}

/* ------------------------------------------------------------------ observation (features.py) */
static float f32_from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

static void card_distance(const uint8_t *bp, int id, int *gems_out, int *turns_out)
{
    int sum = 0, mx = 0;
    for (int c = 0; c < 5; c++) {
        int miss = CARD[id].cost[c] - bp[c];
        if (miss > 0) { sum += miss; if (miss > mx) mx = miss; }
    }
    int t = (sum + 2) / 3;           /* np.ceil(sum/3) features.py:124-135 */
    *gems_out = sum;
    *turns_out = sum == 0 ? 0 : (t > mx ? t : mx);
}

static void vectorize_card(int id, float *o) /* features.py:377-413; 13 floats */
{
    static const int ALPHA[5] = {0, 3, 2, 1, 4}; /* our colour id -> index among black,blue,green,red,white(,yellow) */
    for (int i = 0; i < 13; i++) o[i] = 0.f;
    if (id == NONE) return;
    o[ALPHA[CARD[id].colour]] = 1.f;
    for (int c = 0; c < 5; c++) o[6 + ALPHA[c]] = (float)CARD[id].cost[c];
    o[11] = (float)card_tier(id);
    o[12] = (float)CARD[id].points;
}

static int cmp_int(const void *x, const void *y) { return *(const int *)x - *(const int *)y; }

/* extract_metrics_with_cards(...).astype(float32) (features.py:470-480, splendor_env.py:206) */
void orc_observe(const orc_state *s, int seat, float *o)
{
    const orc_player *a = &s->pl[seat];
    uint8_t bp[5];
    int owned = 0, total_gems = 0;
    for (int c = 0; c < 5; c++) { bp[c] = (uint8_t)(a->gems[c] + a->cards[c]); owned += a->cards[c]; }
    for (int c = 0; c < 6; c++) total_gems += a->gems[c];
    for (int i = 0; i < SPL_OBS_SIZE; i++) o[i] = 0.f;
    o[0] = 1.f; o[1] = (float)a->turns; o[2] = (float)a->score; o[3] = a->score >= 15 ? 1.f : 0.f;
    o[4] = (float)owned; o[5] = (float)a->n_resv;
    { /* np.var of the 5 buying powers, float64 arithmetic in numpy's order, then float32 */
        double m = 0, v = 0;
        for (int c = 0; c < 5; c++) m += (double)bp[c];
        m /= 5.0;
        for (int c = 0; c < 5; c++) { double d = (double)bp[c] - m; v += d * d; }
        o[6] = (float)(v / 5.0);
    }
    o[7] = (float)a->gems[5]; o[8] = (float)total_gems;
    for (int c = 0; c < 5; c++) {
        o[9 + c] = (float)a->cards[c];
        o[14 + c] = (float)bp[c];
        o[19 + c] = f32_from_bits(LOG1P_BITS[bp[c]]);
    }
    for (int i = 0; i < s->P; i++) { /* features.py:276-282, including its i-1 indexing quirk */
        if (i < seat) o[24 + i] = (float)s->pl[i].score;
        else if (i > seat) o[27 + i - 1] = (float)s->pl[i].score;
    }
    int per_colour[5][16], n_per[5] = {0, 0, 0, 0, 0};
    for (int p = 0; p < 12; p++) {
        int id = s->dealt[p];
        if (id == NONE) continue;
        int g, t;
        card_distance(bp, id, &g, &t);
        o[30 + p] = (float)g; o[42 + p] = (float)t;
        per_colour[CARD[id].colour][n_per[CARD[id].colour]++] = g;
    }
    for (int i = 0; i < a->n_resv; i++) {
        int id = a->resv[i], g, t;
        card_distance(bp, id, &g, &t);
        o[54 + i] = (float)g; o[57 + i] = (float)t;
        per_colour[CARD[id].colour][n_per[CARD[id].colour]++] = g;
    }
    for (int c = 0; c < 5; c++) qsort(per_colour[c], (size_t)n_per[c], sizeof(int), cmp_int);
    for (int k = 0; k < s->n_nobles; k++) {
        int dist_cards = 0, dist_gems = 0;
        for (int c = 0; c < 5; c++) {
            int count = NOBLE_COST[s->nobles[k]][c] - a->cards[c];
            if (count <= 0) continue;
            dist_cards += count;
            for (int j = 0; j < count && j < n_per[c]; j++) dist_gems += per_colour[c][j];
            if (count > n_per[c]) dist_gems += 14 * (count - n_per[c]);
        }
        o[60 + k] = (float)dist_gems; o[65 + k] = (float)dist_cards;
    }
    for (int p = 0; p < 12; p++) vectorize_card(s->dealt[p], o + 70 + 13 * p);
    for (int i = 0; i < 3; i++) vectorize_card(i < a->n_resv ? a->resv[i] : NONE, o + 70 + 13 * (12 + i));
}

/* ------------------------------------------------------------------ synthetic games: Philox4x32-10 */
/* These restate THIS REPO's synthetic-input definition (DESIGN.md "Synthetic games"), not reference code: the
 * reference draws decks with Python's random.shuffle, which parity tests bypass by injecting deck orders. */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out)
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_philox(uint64_t seed, uint64_t game, uint32_t index, uint32_t stream, uint32_t *out)
{
    philox4x32_10((uint32_t)game, (uint32_t)(game >> 32), index, stream, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}
static uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

enum { STREAM_ACTION = 0, STREAM_DECK0 = 1, STREAM_NOBLES = 4, STREAM_ENV = 5 };

/* deck order + nobles of synthetic game `game`: deck_lists in REFERENCE LIST ORDER (first dealt card = last element) */
void orc_synth_game(uint64_t seed, uint64_t game, int P, uint8_t *deck_lists, uint8_t *nobles)
{
    uint32_t r[8];
    uint8_t ids[10];
    for (int i = 0; i < 10; i++) ids[i] = (uint8_t)i;
    orc_philox(seed, game, 0, STREAM_NOBLES, r);
    orc_philox(seed, game, 1, STREAM_NOBLES, r + 4);
    for (int i = 0; i < P + 1; i++) {
        int j = i + (int)mulhi(r[i], (uint32_t)(10 - i));
        uint8_t t = ids[i]; ids[i] = ids[j]; ids[j] = t;
        nobles[i] = ids[i];
    }
    for (int t = 0; t < 3; t++) {
        int n = TIER_SIZE[t];
        uint8_t remaining[40];
        for (int i = 0; i < n; i++) remaining[i] = (uint8_t)i; /* ascending positions still in the deck */
        for (int d = 0; d < n; d++) {
            uint32_t w[4];
            orc_philox(seed, game, (uint32_t)d, (uint32_t)(STREAM_DECK0 + t), w);
            int c = n - d, j = (int)mulhi(w[0], (uint32_t)c);
            int pos = remaining[j];
            memmove(remaining + j, remaining + j + 1, (size_t)(c - 1 - j));
            deck_lists[TIER_BASE[t] + (n - 1 - d)] = (uint8_t)(TIER_BASE[t] + pos); /* d-th deal pops from the end */
        }
    }
}

/* One random playout of synthetic game `game` (Game.Run with RandomAgents, game.py:104-213, restated for uniform
 * choice over ascending ALL_ACTIONS indices).  trace (optional) receives the action index per step, -1 terminated.
 * returns number of steps; *end_kind as orc_game_ends, 4 = truncated at max_steps. */
int orc_playout(uint64_t seed, uint64_t game, int P, int limit_rounds, int max_steps,
                orc_state *final_state, int16_t *trace, int *end_kind)
{
    uint8_t decks[90], nobles[5];
    orc_state s;
    uint16_t legal[1024];
    orc_synth_game(seed, game, P, decks, nobles);
    orc_init(&s, P, decks, nobles);
    int t = 0, e = 0;
    while (t < max_steps && !(e = orc_game_ends(&s, limit_rounds))) {
        int n = orc_legal(&s, s.cur, legal);
        uint32_t w[4];
        orc_philox(seed, game, (uint32_t)t >> 2, STREAM_ACTION, w);  /* one Philox block serves four consecutive moves */
        int a = legal[mulhi(w[t & 3], (uint32_t)n)];
        if (trace) trace[t] = (int16_t)a;
        orc_step(&s, a);
        t++;
    }
    if (!e) e = orc_game_ends(&s, limit_rounds) ? orc_game_ends(&s, limit_rounds) : 4;
    if (trace) for (int i = t; i < max_steps; i++) trace[i] = -1;
    if (final_state) *final_state = s;
    if (end_kind) *end_kind = e;
    return t;
}

/* counters layout shared with the CUDA engine (include/splendor_b200.h SPL_CTR_*) */
static void tally(const orc_state *s, int steps, int end_kind, int64_t *ctr, int8_t *score2_out)
{
    int sc[4], mx = -1, nmax = 0;
    for (int i = 0; i < s->P; i++) { sc[i] = orc_cal_score2(s, i); if (sc[i] > mx) mx = sc[i]; }
    for (int i = 0; i < s->P; i++) if (sc[i] == mx) nmax++;
    ctr[0] += 1; ctr[1] += steps;
    for (int i = 0; i < s->P; i++) {
        if (sc[i] == mx) ctr[(nmax > 1 ? 6 : 2) + i] += 1; /* win = unique max, tie = shared max (general_game_runner.py:416-430) */
        ctr[14] += sc[i];
        ctr[15] += s->pl[i].n_nobles;
        if (score2_out) score2_out[i] = (int8_t)sc[i];
    }
    ctr[9 + end_kind] += 1; /* 10 score, 11 deadlock, 12 round cap, 13 truncated */
}

/* Batch of synthetic playouts on `nthreads` host threads (pthreads; games handed out 16 at a time from an atomic
 * cursor).  out_score2 [n][P] int8, out_steps [n] int32 (both optional), counters int64[16]. */
typedef struct {
    uint64_t seed, first_game;
    int n, P, limit_rounds, max_steps;
    int8_t *out_score2;
    int32_t *out_steps;
    int *cursor;
    int64_t local[16];
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (;;) {
        int lo = __atomic_fetch_add(j->cursor, 16, __ATOMIC_RELAXED);
        if (lo >= j->n) break;
        int hi = lo + 16 < j->n ? lo + 16 : j->n;
        for (int i = lo; i < hi; i++) {
            orc_state fin;
            int kind, steps = orc_playout(j->seed, j->first_game + (uint64_t)i, j->P, j->limit_rounds, j->max_steps, &fin, NULL, &kind);
            tally(&fin, steps, kind, j->local, j->out_score2 ? j->out_score2 + (size_t)i * j->P : NULL);
            if (j->out_steps) j->out_steps[i] = steps;
        }
    }
    return NULL;
}

void orc_playout_many(uint64_t seed, uint64_t first_game, int n, int P, int limit_rounds, int max_steps,
                      int8_t *out_score2, int32_t *out_steps, int64_t *counters, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    job_t jobs[256];
    int cursor = 0;
    for (int t = 0; t < nthreads; t++) {
        job_t j = {seed, first_game, n, P, limit_rounds, max_steps, out_score2, out_steps, &cursor, {0}};
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    memset(counters, 0, 16 * sizeof(int64_t));
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        for (int k = 0; k < 16; k++) counters[k] += jobs[t].local[k];
    }
}

/* test hook: the decoded fields of ALL_ACTIONS[idx] (checked against tests/golden/tables.json for all 3510) */
int orc_describe_action(int idx, int *type, int *p, int *n, uint8_t *collected, uint8_t *returned)
{
    orc_state dummy;
    action_t A;
    memset(&dummy, 0, sizeof dummy);
    if (decode_action(&dummy, 0, idx, &A)) return -1;
    *type = A.type; *p = A.p; *n = A.n;
    memcpy(collected, A.collected, 6);
    memcpy(returned, A.returned, 6);
    return 0;
}

/* generatePredecessor (splendor_model.py:156-220): undo ALL_ACTIONS[idx] for `seat`.  card = the card that was bought or
 * reserved (NONE for other actions), noble = the noble id that was claimed (NONE if none), returned[6] = the action's
 * returned_gems (for buys: the payment the successor computed).  Quirks kept: the undone noble and an un-bought
 * reserved card are APPENDED to their lists (:201,:210), the card that had been dealt into the slot goes back on top of
 * its deck (:180,:196), `passed` is set from the undone action (:219), and the seat to move is not touched (that lives
 * in the rule object). */
int orc_undo(orc_state *s, int seat, int idx, int card, int noble, const uint8_t *returned)
{
    action_t A;
    if (decode_action(s, seat, idx, &A)) return -1;
    orc_player *a = &s->pl[seat];
    int score = 0;
    if (A.type == T_SAME || A.type == T_DIFF || A.type == T_RESERVE) {
        for (int c = 0; c < 6; c++) { s->bank[c] += A.collected[c]; a->gems[c] -= A.collected[c]; }
        for (int c = 0; c < 6; c++) { a->gems[c] += A.returned[c]; s->bank[c] -= A.returned[c]; }
        if (A.type == T_RESERVE) {
            int cur = s->dealt[A.p];
            if (cur != NONE) { int t = card_tier(cur); s->deck[t][s->deck_n[t]++] = (uint8_t)cur; }
            int k = -1;
            for (int i = 0; i < a->n_resv; i++) if (a->resv[i] == card) k = i;
            if (k < 0) return -1;
            for (int i = k; i < 2; i++) a->resv[i] = a->resv[i + 1];
            a->resv[2] = NONE;
            a->n_resv--;
            s->dealt[A.p] = (uint8_t)card;
        }
    } else if (A.type == T_BUY_AVAILABLE || A.type == T_BUY_RESERVE) {
        for (int c = 0; c < 6; c++) { a->gems[c] += returned[c]; s->bank[c] -= returned[c]; }
        if (A.type == T_BUY_AVAILABLE) {
            int cur = s->dealt[A.p];
            if (cur != NONE) { int t = card_tier(cur); s->deck[t][s->deck_n[t]++] = (uint8_t)cur; }
            s->dealt[A.p] = (uint8_t)card;
        } else {
            a->resv[a->n_resv++] = (uint8_t)card;
        }
        a->cards[CARD[card].colour]--;
        score -= CARD[card].points;
    }
    if (noble != NONE) {
        s->nobles[s->n_nobles++] = (uint8_t)noble;
        a->nobles_mask &= (uint16_t)~(1u << noble);
        a->n_nobles--;
        score -= 3;
    }
    a->turns--;
    a->score = (uint8_t)(a->score + score);
    a->passed = (A.type == T_PASS);
    return 0;
}

/* ------------------------------------------------------------------ MiniMaxAgent (agents/our_agents/minmax.py) */
static const uint16_t COST_ORDER[SPL_NUM_CARDS] = SPL_CARD_COST_ORDER;

/* _evaluation_function (:90-136) from agent `me`'s point of view.  float64 with the reference's operation order:
 * Python evaluates a * b * c left to right, `reward -= x` one term at a time over board cards (tier-major, slots in
 * order) then the agent's reserved cards, each card's cost.items() in dict order; np.var(list(gems.values())) =
 * mean of squared deviations from sum/6 over black, red, yellow, green, blue, white.  Build with -ffp-contract=off. */
double orc_minimax_eval(const orc_state *s, int me)
{
    const orc_player *a = &s->pl[me];
    int max_score = 0;
    for (int i = 0; i < s->P; i++) if (s->pl[i].score > max_score) max_score = s->pl[i].score;
    if (max_score >= 15) {
        double r = (double)(99999 + max_score);
        return max_score > a->score ? -r : r;
    }
    int gem_sum = 0;
    for (int c = 0; c < 6; c++) gem_sum += a->gems[c];
    const double gems_factor = gem_sum >= 8 ? -0.7 : 0.1;
    double x[6], mean = 0.0, ss = 0.0;
    for (int k = 0; k < 6; k++) { x[k] = (double)a->gems[ALLG[k]]; mean += x[k]; }
    mean = mean / 6.0;
    for (int k = 0; k < 6; k++) { double d = x[k] - mean; ss += d * d; }
    const double gems_var = ss / 6.0;

    double reward = 0.0;
    for (int slot = 0; slot < 12 + a->n_resv; slot++) {
        const int id = slot < 12 ? s->dealt[slot] : a->resv[slot - 12];
        if (id == NONE) continue;
        const int col = CARD[id].colour;
        int relevant = 0;
        for (int k = 0; k < s->n_nobles; k++) {
            const uint8_t *nc = NOBLE_COST[s->nobles[k]];
            if (nc[col] > 0 && a->cards[col] < nc[col]) relevant++;
        }
        const double weight = (double)(CARD[id].points + 1) + (double)relevant * 0.5;
        for (unsigned o = COST_ORDER[id]; (o & 7u) != 7u; o >>= 3) {
            const int c = (int)(o & 7u);
            const int diff = (int)CARD[id].cost[c] - ((int)a->gems[c] + (int)a->cards[c]);
            double t = (double)diff * 0.1;
            t = t * weight;
            reward = reward - (t < 0 ? -t : t);
        }
    }
    double r = reward + (double)(a->score * 2);
    r = r + 6.0 * 0.7; /* len(agent_state.cards) is the number of colour keys: always 6 */
    r = r + (double)gem_sum * gems_factor;
    r = r + gems_var * -0.2;
    return r;
}

/* _select_action_recursion (:43-88) at DEPTH 2 for the seat to move, without the alpha-beta cut (same root values for
 * the chosen action): values[i] = min over the opponent's replies of the evaluation, for root action
 * actions[i] (ascending ALL_ACTIONS order).  No terminal test between plies, as in the reference.  Returns the count. */
int orc_minimax2(const orc_state *s, uint16_t *actions, double *values)
{
    const int me = s->cur;
    const int n = orc_legal(s, me, actions);
    for (int i = 0; i < n; i++) {
        orc_state c1 = *s;
        orc_step(&c1, actions[i]);
        uint16_t reply[1024];
        const int m = orc_legal(&c1, c1.cur, reply);
        double best = 1.0 / 0.0;
        for (int j = 0; j < m; j++) {
            orc_state c2 = c1;
            orc_step(&c2, reply[j]);
            const double v = orc_minimax_eval(&c2, me);
            if (v < best) best = v;
        }
        values[i] = best;
    }
    return n;
}

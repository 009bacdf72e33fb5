// minimax.cuh -- the reference's search agent on the packed state (device code, included by kernels.cu only).
//
//   MiniMaxAgent._evaluation_function   agents/our_agents/minmax.py:90-136
//   _select_action_recursion, DEPTH 2   agents/our_agents/minmax.py:43-88
//
// The evaluation is float64 in the reference (Python floats + one np.var) and is reproduced BIT-EXACTLY: same operation
// order, round-to-nearest intrinsics so that nothing is contracted into an FMA.  `reward -= abs(diff * 0.1 * weight)` runs
// over the board cards (tier-major, slot order), then the agent's reserved cards, and per card over cost.items() in the
// dict order of the reference's CARDS table (SPL_CARD_COST_ORDER, generated from it).
#pragma once
#include "engine.cuh"

namespace spl {

__device__ const uint16_t g_cost_order[SPL_NUM_CARDS] = SPL_CARD_COST_ORDER;

// the `reward -= abs(...)` terms of the cards whose ids are the low `n` bytes of w (:113-130)
__device__ __noinline__ double minimax_card_terms(double reward, uint32_t w, int n, uint32_t meta, uint32_t gems, uint32_t cards, const Tables &T)
{
    for (int b = 0; b < n; b++, w >>= 8) {
        const uint32_t id = w & 0xFFu;
        if (id >= (uint32_t)SPL_NUM_CARDS) continue;  // empty slot
        const uint2 cw = T.card[id];
        const uint32_t sh = cw.y & 31u, have_col = field(cards, sh);
        int relevant = 0;  // nobles this card's colour still counts towards (:114-121)
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint32_t nid = (meta >> (4 * k)) & 15u;
            const uint32_t need = nid == 15u ? 0u : field(T.noble[nid], sh);
            relevant += (need > 0u && have_col < need) ? 1 : 0;
        }
        const double weight = __dadd_rn((double)(((cw.y >> 5) & 7u) + 1u), __dmul_rn((double)relevant, 0.5));
        for (uint32_t o = g_cost_order[id]; (o & 7u) != 7u; o >>= 3) {  // cost.items() in dict order (:123-130)
            const int fsh = 5 * field_of_normal((int)(o & 7u));
            const int diff = (int)field(cw.x, fsh) - (int)(field(gems, fsh) + field(cards, fsh));
            reward = __dsub_rn(reward, fabs(__dmul_rn(__dmul_rn((double)diff, 0.1), weight)));
        }
    }
    return reward;
}

template <int P>
__device__ __forceinline__ double minimax_eval(const Env<P> &e, const Tables &T, int me)
{
    uint32_t gems = e.pg[0], cards = e.pc[0], resv = e.pr[0];
    uint32_t max_score = 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        max_score = max(max_score, e.pr[p] >> 24);
        if (p && me == p) { gems = e.pg[p]; cards = e.pc[p]; resv = e.pr[p]; }
    }
    const uint32_t score = resv >> 24;
    if (max_score >= 15u) {  // :103-108
        const double r = (double)(99999u + max_score);
        return max_score > score ? -r : r;
    }
    const int gem_sum = (int)hsum6(gems);
    const double gems_factor = gem_sum >= 8 ? -0.7 : 0.1;  // :109-110
    // np.var(list(gems.values())): mean of squared deviations from sum / 6, black red yellow green blue white = fields 0..5
    const double mean = __ddiv_rn((double)gem_sum, 6.0);
    double ss = 0.0;
#pragma unroll
    for (int f = 0; f < 6; f++) {
        const double d = __dsub_rn((double)field(gems, 5 * f), mean);
        ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    const double gems_var = __ddiv_rn(ss, 6.0);

    // board.dealt_list() + agent_state.cards["yellow"] (:113): four words of card ids, named one by one so that the state
    // stays in registers (an indexed walk over e.d[] puts it in local memory)
    double reward = 0.0;
    reward = minimax_card_terms(reward, e.d[0], 4, e.meta, gems, cards, T);
    reward = minimax_card_terms(reward, e.d[1], 4, e.meta, gems, cards, T);
    reward = minimax_card_terms(reward, e.d[2], 4, e.meta, gems, cards, T);
    reward = minimax_card_terms(reward, resv, 3, e.meta, gems, cards, T);
    double r = __dadd_rn(reward, (double)(2u * score));                       // + score * score_factor
    r = __dadd_rn(r, __dmul_rn(6.0, 0.7));                                    // + len(agent_state.cards) * cards_factor: 6 colour keys
    r = __dadd_rn(r, __dmul_rn((double)gem_sum, gems_factor));                // + sum(gems) * gems_factor
    return __dadd_rn(r, __dmul_rn(gems_var, -0.2));                           // + gems_var * gems_var_factor
}

// rank of an action's type string in the reference's sort (minmax.py:59): buy_available < buy_reserve < collect_diff <
// collect_same < pass < reserve
__device__ __forceinline__ int minimax_type_rank(int idx)
{
    return idx >= BASE_BUY_AVAILABLE ? 0 : (idx >= BASE_BUY_RESERVE ? 1 : (idx >= BASE_DIFF ? 2 : (idx >= BASE_SAME ? 3 : (idx >= BASE_RESERVE ? 5 : 4))));
}

}  // namespace spl

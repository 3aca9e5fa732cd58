// msw_avoid.cu -- avoidability.analyze_avoidability (minesweeper/avoidability.py:145-394) for a
// whole batch of boards in one launch (the evaluator calls it per env
// and per step from a Python loop, eval.py:381-398).
//
// One warp per board, working set in shared memory.  The reference's answer is a set of logical
// consequences, so it does not depend on iteration order, and the kernel uses the formulation
// that parallelises:
//   * frontier F = hidden & dilate8(revealed), constraint cells = revealed & ~mine & dilate8(F),
//     targets = adjacent_counts: bitboard algebra on the lane-distributed words (msw_common.cuh);
//   * frontier components (two variables are connected when they share a constraint, :212-237):
//     min-label propagation through the constraint cells; the label of a component is its smallest
//     cell, which is also the order in which the reference discovers components;
//   * unit propagation + pairwise subset rule (:268-329) as a monotone fixed point: every round
//     derives assignments from ONE snapshot of the assignment map (a constraint pair evaluated on
//     two different snapshots can make the subset rule unsound) and commits them afterwards.  A
//     subset pair (a, b) needs rem(a) inside rem(b), so b lies in the 5x5 window of a;
//   * only if the rules found no safe cell (:343-351): exact feasibility of "cell = mine" for every
//     free variable of every component (:353-375).  Lane l answers queries l, l+32, ... with its
//     own depth-first search over the component's variables in row-major order; the bound checks
//     are the reference's (_ConstraintSolver._assign, :98-126), the constraint counters live in
//     per-lane local memory indexed by cell.  The search is exact, hence order-independent, but
//     exponential in the worst case exactly like the reference's; a per-lane step budget turns a
//     runaway search into flag bit 3 instead of a hung GPU.
// Output is the array form documented in include/msw_b200.h (msw_avoidability).
#include "../../include/msw_b200.h"
#include "msw_common.cuh"
#include "msw_error.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace msw {

struct AvoidParams {
    const uint32_t *mines, *revealed, *flags;   // [n][wpb]; flags nullable
    const int4 *meta;                            // [n]: .x = first_click_done
    uint32_t *safe_bits;                         // [n][wpb]
    int16_t *comp_of_cell, *comp_size;           // [n][HW]
    uint8_t *out_flags;                          // [n]
    long long n;
    int H, W, HW, wpb;
    unsigned budget;
};

struct AvoidSmem {
    int label[MSW_MAX_CELLS];        // frontier cells: smallest cell of the component; others: -1
    int size[MSW_MAX_CELLS];
    int16_t vcell[MSW_MAX_CELLS];    // free variables of the current component, ascending
    int16_t ccell[MSW_MAX_CELLS];    // its constraint cells
    uint8_t front[MSW_MAX_CELLS], cons[MSW_MAX_CELLS], assign[MSW_MAX_CELLS], pend[MSW_MAX_CELLS];
    int8_t target[MSW_MAX_CELLS], need0[MSW_MAX_CELLS], unk0[MSW_MAX_CELLS];
    uint32_t safe[32];
};

__global__ void __launch_bounds__(32) avoid_kernel(const __grid_constant__ AvoidParams p)
{
    __shared__ AvoidSmem S;
    const int lane = threadIdx.x;
    const int H = p.H, W = p.W, HW = p.HW, wpb = p.wpb;
    const long long b = blockIdx.x;
    if (b >= p.n) return;

    // outputs default to "nothing on the frontier"
    for (int c = lane; c < HW; c += 32) {
        p.comp_of_cell[b * HW + c] = -1;
        p.comp_size[b * HW + c] = 0;
    }
    const bool fcd = p.meta[b].x != 0;
    if (!fcd) {                                        // avoidability.py:152-159: avoidable, nothing else
        if (lane < wpb) p.safe_bits[b * wpb + lane] = 0u;
        if (lane == 0) p.out_flags[b] = 1;
        return;
    }

    Geo g;
    g.valid = g.notcol0 = g.notlast = 0u;
    for (int j = 0; j < 32; ++j) {
        const int cell = lane * 32 + j;
        if (cell < HW) {
            g.valid |= 1u << j;
            if (cell % W != 0) g.notcol0 |= 1u << j;
            if (cell % W != W - 1) g.notlast |= 1u << j;
        }
    }
    uint32_t M = 0, R = 0, Fl = 0;
    if (lane < wpb) {
        M = p.mines[b * wpb + lane];
        R = p.revealed[b * wpb + lane] & g.valid;
        if (p.flags) Fl = p.flags[b * wpb + lane];
    }
    const uint32_t hidden = ~R & ~Fl & g.valid;                          // :162
    const uint32_t F = hidden & dilate8<0>(R, lane, W, g);               // :164-172
    const bool has_frontier = __any_sync(FULL, F != 0u);
    if (!has_frontier) {                                                 // :175-186
        if (lane < wpb) p.safe_bits[b * wpb + lane] = 0u;
        if (lane == 0) p.out_flags[b] = 4;
        return;
    }
    const uint32_t Cn = R & ~M & dilate8<0>(F, lane, W, g);              // :197-210
    const Planes pl = count_planes<0>(M, lane, W, g);
    for (int it = 0; it < wpb; ++it) {
        const int cell = it * 32 + lane;
        const uint32_t wf = __shfl_sync(FULL, F, it), wc = __shfl_sync(FULL, Cn, it);
        const uint32_t c0 = __shfl_sync(FULL, pl.c0, it), c1 = __shfl_sync(FULL, pl.c1, it);
        const uint32_t c2 = __shfl_sync(FULL, pl.c2, it), c3 = __shfl_sync(FULL, pl.c3, it);
        if (cell < HW) {
            const bool f = (wf >> lane) & 1u;
            S.front[cell] = f;
            S.cons[cell] = (wc >> lane) & 1u;
            S.target[cell] = (int8_t)(((c0 >> lane) & 1u) | (((c1 >> lane) & 1u) << 1) | (((c2 >> lane) & 1u) << 2) |
                                      (((c3 >> lane) & 1u) << 3));
            S.assign[cell] = 0;
            S.pend[cell] = 0;
            S.label[cell] = f ? cell : -1;
            S.size[cell] = 0;
        }
    }
    S.safe[lane] = 0u;
    __syncwarp();

    // frontier neighbours of cell a (bounds-checked); fn(neighbour cell)
    auto for_nbrs = [&](int a, auto &&fn) {
        const int r = a / W, c = a % W;
        for (int dr = -1; dr <= 1; ++dr) {
            const int rr = r + dr;
            if (rr < 0 || rr >= H) continue;
            for (int dc = -1; dc <= 1; ++dc) {
                const int cc = c + dc;
                if ((dr | dc) == 0 || cc < 0 || cc >= W) continue;
                fn(rr * W + cc);
            }
        }
    };

    // ---- components: min-label propagation through the constraint cells
    for (bool changed = true; changed;) {
        bool ch = false;
        for (int a = lane; a < HW; a += 32) {
            if (!S.cons[a]) continue;
            int m = 0x7fffffff;
            for_nbrs(a, [&](int x) { if (S.front[x]) m = min(m, S.label[x]); });
            for_nbrs(a, [&](int x) {
                if (S.front[x] && S.label[x] > m) { atomicMin(&S.label[x], m); ch = true; }
            });
        }
        __syncwarp();
        changed = __any_sync(FULL, ch);
    }
    for (int a = lane; a < HW; a += 32)
        if (S.front[a]) atomicAdd(&S.size[S.label[a]], 1);
    __syncwarp();
    for (int a = lane; a < HW; a += 32) {
        p.comp_of_cell[b * HW + a] = (int16_t)S.label[a];
        p.comp_size[b * HW + a] = (int16_t)S.size[a];
    }

    // ---- unit + subset rules to the fixed point; rem(a) = unassigned frontier neighbours of a,
    // tgt(a) = target - neighbours known to be mines, both read from the snapshot S.assign
    auto remaining = [&](int a, int (&rem)[8], int &tgt) {
        int k = 0, t = S.target[a];
        for_nbrs(a, [&](int x) {
            if (!S.front[x]) return;
            const int v = S.assign[x];
            if (v == 0) rem[k++] = x;
            else if (v == 2) --t;
        });
        tgt = t;
        return k;
    };
    bool any_safe = false;
    for (bool changed = true; changed;) {
        for (int a = lane; a < HW; a += 32) {
            if (!S.cons[a]) continue;
            int ra[8], ta;
            const int na = remaining(a, ra, ta);
            if (na == 0) continue;
            if (ta >= 0 && ta <= na) {                                   // unit propagation (:271-287)
                if (ta == 0) for (int k = 0; k < na; ++k) S.pend[ra[k]] = 1;
                else if (ta == na) for (int k = 0; k < na; ++k) S.pend[ra[k]] = 2;
            }
            const int r0 = a / W, c0 = a % W;                            // subset rule (:291-329), a = subset side
            for (int br = -2; br <= 2; ++br) {
                const int rb = r0 + br;
                if (rb < 0 || rb >= H) continue;
                for (int bc = -2; bc <= 2; ++bc) {
                    const int cb = c0 + bc;
                    if ((br | bc) == 0 || cb < 0 || cb >= W) continue;
                    const int bb = rb * W + cb;
                    if (!S.cons[bb]) continue;
                    bool subset = true;                                  // every cell of rem(a) is a neighbour of b
                    for (int k = 0; k < na && subset; ++k) {
                        const int dr = ra[k] / W - rb, dc = ra[k] % W - cb;
                        subset = dr >= -1 && dr <= 1 && dc >= -1 && dc <= 1;
                    }
                    if (!subset) continue;
                    int rb8[8], tb;
                    const int nb = remaining(bb, rb8, tb);
                    int diff[8], nd = 0;
                    for (int k = 0; k < nb; ++k) {
                        const int dr = rb8[k] / W - r0, dc = rb8[k] % W - c0;
                        if (!(dr >= -1 && dr <= 1 && dc >= -1 && dc <= 1)) diff[nd++] = rb8[k];
                    }
                    if (nd == 0) continue;
                    if (ta == tb) for (int k = 0; k < nd; ++k) S.pend[diff[k]] = 1;
                    else if (tb - ta == nd) for (int k = 0; k < nd; ++k) S.pend[diff[k]] = 2;
                }
            }
        }
        __syncwarp();
        bool ch = false;
        for (int a = lane; a < HW; a += 32)
            if (S.pend[a] && !S.assign[a]) {
                S.assign[a] = S.pend[a];
                ch = true;
                if (S.pend[a] == 1) any_safe = true;
            }
        __syncwarp();
        changed = __any_sync(FULL, ch);
    }
    any_safe = __any_sync(FULL, any_safe);

    uint8_t out_flag = 2 | 4;
    if (any_safe) {                                                      // :343-351
        for (int a = lane; a < HW; a += 32)
            if (S.assign[a] == 1) atomicOr(&S.safe[a >> 5], 1u << (a & 31));
        out_flag |= 1;
    } else {
        // ---- exact search (:353-375).  need0 / unk0: the constraints reduced by the mines the rules found
        for (int a = lane; a < HW; a += 32) {
            if (!S.cons[a]) { S.unk0[a] = 0; continue; }
            int rem[8], t;
            S.unk0[a] = (int8_t)remaining(a, rem, t);
            S.need0[a] = (int8_t)t;
        }
        __syncwarp();
        int8_t need[MSW_MAX_CELLS], unk[MSW_MAX_CELLS], val[MSW_MAX_CELLS];   // per-lane search state (local memory)
        unsigned budget = p.budget;
        bool over = false, found = false;
        // assign `value` to variable x: update the counters of its constraints, undo on a violated bound
        auto try_assign = [&](int x, int value) {
            const int r = x / W, c = x % W;
            bool ok = true;
            int done = 0;
            for (int q = 0; q < 9 && ok; ++q) {
                const int rr = r + q / 3 - 1, cc = c + q % 3 - 1;
                done = q + 1;
                if (q == 4 || rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
                const int a = rr * W + cc;
                if (!S.cons[a]) continue;
                unk[a] -= 1;
                need[a] -= (int8_t)value;
                ok = need[a] >= 0 && need[a] <= unk[a];
            }
            if (!ok)
                for (int q = done - 1; q >= 0; --q) {
                    const int rr = r + q / 3 - 1, cc = c + q % 3 - 1;
                    if (q == 4 || rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
                    const int a = rr * W + cc;
                    if (!S.cons[a]) continue;
                    unk[a] += 1;
                    need[a] += (int8_t)value;
                }
            return ok;
        };
        auto undo = [&](int x, int value) {
            for_nbrs(x, [&](int a) {
                if (S.cons[a]) { unk[a] += 1; need[a] += (int8_t)value; }
            });
        };
        for (int lab = 0; lab < HW; ++lab) {
            if (S.size[lab] == 0) continue;                              // warp-uniform
            // compact the component: free variables and constraint cells, ascending
            int K = 0, Kc = 0;
            for (int it = 0; it < wpb; ++it) {
                const int cell = it * 32 + lane;
                const bool in = cell < HW;
                const bool isv = in && S.front[cell] && S.label[cell] == lab && S.assign[cell] == 0;
                bool isc = false;
                if (in && S.cons[cell] && S.unk0[cell] > 0)
                    for_nbrs(cell, [&](int x) { if (S.front[x] && S.label[x] == lab) isc = true; });
                const unsigned mv = __ballot_sync(FULL, isv), mc = __ballot_sync(FULL, isc);
                const unsigned lt = (1u << lane) - 1u;
                if (isv) S.vcell[K + __popc(mv & lt)] = (int16_t)cell;
                if (isc) S.ccell[Kc + __popc(mc & lt)] = (int16_t)cell;
                K += __popc(mv);
                Kc += __popc(mc);
            }
            __syncwarp();
            if (K == 0 || Kc == 0) continue;                             // :355-360
            for (int qi = lane; qi < K && !over; qi += 32) {
                const int q = S.vcell[qi];
                for (int k = 0; k < Kc; ++k) {
                    const int a = S.ccell[k];
                    need[a] = S.need0[a];
                    unk[a] = S.unk0[a];
                }
                bool feasible = false;
                if (try_assign(q, 1)) {
                    int pos = 0;
                    bool forward = true;
                    while (true) {
                        if (budget == 0u) { over = true; feasible = true; break; }
                        --budget;
                        if (forward) {
                            if (pos == K) { feasible = true; break; }
                            const int x = S.vcell[pos];
                            if (x == q) { ++pos; continue; }
                            if (try_assign(x, 0)) { val[x] = 0; ++pos; }
                            else if (try_assign(x, 1)) { val[x] = 1; ++pos; }
                            else { forward = false; --pos; }
                        } else {
                            if (pos < 0) break;
                            const int x = S.vcell[pos];
                            if (x == q) { --pos; continue; }
                            const int v = val[x];
                            undo(x, v);
                            if (v == 0 && try_assign(x, 1)) { val[x] = 1; forward = true; ++pos; }
                            else --pos;
                        }
                    }
                }
                if (!feasible) {
                    atomicOr(&S.safe[q >> 5], 1u << (q & 31));
                    found = true;
                }
            }
            __syncwarp();
        }
        if (__any_sync(FULL, found)) out_flag |= 1;
        if (__any_sync(FULL, over)) out_flag |= 8;
    }
    __syncwarp();
    if (lane < wpb) p.safe_bits[b * wpb + lane] = S.safe[lane];
    if (lane == 0) p.out_flags[b] = out_flag;
}

}  // namespace msw

extern "C" int msw_avoidability(const msw_env_desc *desc, const msw_state *st, int64_t n, uint32_t *safe_bits,
                                int16_t *comp_of_cell, int16_t *comp_size, uint8_t *flags, uint32_t search_budget,
                                void *stream)
{
    using namespace msw;
    if (!desc || !st || !st->mines || !st->revealed || !st->meta) return fail(MSW_ERR_NULL, "msw_avoidability: NULL state");
    if (!safe_bits || !comp_of_cell || !comp_size || !flags) return fail(MSW_ERR_NULL, "msw_avoidability: NULL output");
    const int wpb = msw_words_per_board(desc->H, desc->W);
    if (wpb == 0) return fail(MSW_ERR_BAD_SHAPE, "msw_avoidability: unsupported board %dx%d", desc->H, desc->W);
    if (n < 0 || n > 0x7fffffffLL) return fail(MSW_ERR_BAD_SHAPE, "msw_avoidability: n=%lld", (long long)n);
    if (n == 0) return MSW_OK;
    AvoidParams p;
    p.mines = st->mines; p.revealed = st->revealed; p.flags = st->flags;
    p.meta = reinterpret_cast<const int4 *>(st->meta);
    p.safe_bits = safe_bits; p.comp_of_cell = comp_of_cell; p.comp_size = comp_size; p.out_flags = flags;
    p.n = n; p.H = desc->H; p.W = desc->W; p.HW = desc->H * desc->W; p.wpb = wpb;
    p.budget = search_budget ? search_budget : (1u << 23);
    avoid_kernel<<<(unsigned)n, 32, 0, (cudaStream_t)stream>>>(p);
    MSW_CUDA_TRY(cudaGetLastError());
    return MSW_OK;
}

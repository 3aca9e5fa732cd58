/*
 * msw_oracle_avoid.c -- CPU restatement of avoidability.analyze_avoidability
 * (minesweeper/avoidability.py:145-394 with its _ConstraintSolver, :43-142).
 *
 * TEST INFRASTRUCTURE ONLY (see msw_oracle.h).  Follows the reference step by step and in the
 * reference's own iteration orders -- frontier variables in row-major order (:174, :189-195),
 * one constraint per revealed non-mine cell with frontier neighbours (:197-210), components by
 * shared constraints discovered in variable order (:212-237), unit propagation to exhaustion then
 * the first productive subset pair (:268-329), and, only when the rules found no safe cell, the
 * exact search per component with variables branched by descending degree (:55, :345-375) --
 * deliberately NOT the fixed-point / spatial-order formulation of the CUDA kernel.
 * Pinned against tests/golden/avoidability.npz, recorded from the live reference.
 */
#include "msw_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
    int nv;
    int vars[8];
    int target;
} cons_t;

typedef struct {
    int num_vars, num_cons;
    int (*cvars)[8];
    int *cnv, *targets;
    int (*v2c)[8];
    int *v2cn;
    int *order;
    /* search state */
    int *assignment;       /* -1 none */
    int *assigned_sum, *unknown_count;
} solver_t;

/* _ConstraintSolver._assign (:98-126); returns 0 on contradiction (all changes undone).  The list of
 * touched constraints is exactly var's constraint list, so _revert needs no separate log. */
static int sv_assign(solver_t *s, int var, int value)
{
    if (s->assignment[var] >= 0) return s->assignment[var] == value ? 2 : 0;   /* 2: nothing to revert */
    for (int k = 0; k < s->v2cn[var]; ++k) {
        const int c = s->v2c[var][k];
        s->assigned_sum[c] += value == 1;
        s->unknown_count[c] -= 1;
        if (s->assigned_sum[c] > s->targets[c] || s->assigned_sum[c] + s->unknown_count[c] < s->targets[c]) {
            for (int j = k; j >= 0; --j) {
                const int cj = s->v2c[var][j];
                s->assigned_sum[cj] -= value == 1;
                s->unknown_count[cj] += 1;
            }
            return 0;
        }
    }
    s->assignment[var] = value;
    return 1;
}

static void sv_revert(solver_t *s, int var)                                      /* :128-139 */
{
    const int value = s->assignment[var];
    s->assignment[var] = -1;
    for (int k = s->v2cn[var] - 1; k >= 0; --k) {
        const int c = s->v2c[var][k];
        s->assigned_sum[c] -= value == 1;
        s->unknown_count[c] += 1;
    }
}

static int sv_dfs(solver_t *s, int pos)                                          /* :70-83 */
{
    if (pos == s->num_vars) return 1;
    const int var = s->order[pos];
    if (s->assignment[var] >= 0) return sv_dfs(s, pos + 1);
    for (int value = 0; value <= 1; ++value) {
        if (!sv_assign(s, var, value)) continue;
        if (sv_dfs(s, pos + 1)) {
            /* the reference returns with the assignment in place and rebuilds its arrays per call;
             * here the arrays are reused, so unwind before reporting success */
            sv_revert(s, var);
            return 1;
        }
        sv_revert(s, var);
    }
    return 0;
}

/* is_feasible({var: 1}) (:57-90) */
static int sv_feasible_with_mine(solver_t *s, int var)
{
    for (int v = 0; v < s->num_vars; ++v) s->assignment[v] = -1;
    for (int c = 0; c < s->num_cons; ++c) {
        s->assigned_sum[c] = 0;
        s->unknown_count[c] = s->cnv[c];
    }
    if (!sv_assign(s, var, 1)) return 0;
    return sv_dfs(s, 0);
}

static const int DR[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, DC[8] = {-1, 0, 1, -1, 1, -1, 0, 1};

static void avoid_one(int H, int W, const uint8_t *mine, const uint8_t *rev, const uint8_t *flg,
                      const uint8_t *counts, int first_click_done, uint8_t *safe, int16_t *comp_of_cell,
                      int16_t *comp_size, uint8_t *flags_out)
{
    const int HW = H * W;
    memset(safe, 0, (size_t)HW);
    for (int i = 0; i < HW; ++i) { comp_of_cell[i] = -1; comp_size[i] = 0; }
    if (!first_click_done) { *flags_out = 1; return; }                           /* :152-159, avoidable=True */

    int *var_of_cell = malloc(sizeof(int) * (size_t)HW), *cell_of_var = malloc(sizeof(int) * (size_t)HW);
    int nv = 0;
    for (int a = 0; a < HW; ++a) {                                               /* :164-174 */
        var_of_cell[a] = -1;
        if (rev[a] || flg[a]) continue;
        for (int d = 0; d < 8; ++d) {
            const int r = a / W + DR[d], c = a % W + DC[d];
            if (r >= 0 && r < H && c >= 0 && c < W && rev[r * W + c]) {
                var_of_cell[a] = nv; cell_of_var[nv++] = a;
                break;
            }
        }
    }
    if (nv == 0) {                                                               /* :176-186, avoidable=False */
        *flags_out = 4;
        free(var_of_cell); free(cell_of_var);
        return;
    }

    cons_t *cons = malloc(sizeof(cons_t) * (size_t)HW);
    int nc = 0;
    for (int a = 0; a < HW; ++a) {                                               /* :197-210 */
        if (!rev[a] || mine[a]) continue;
        cons_t k; k.nv = 0; k.target = counts[a];
        for (int d = 0; d < 8; ++d) {
            const int r = a / W + DR[d], c = a % W + DC[d];
            if (r >= 0 && r < H && c >= 0 && c < W && var_of_cell[r * W + c] >= 0) k.vars[k.nv++] = var_of_cell[r * W + c];
        }
        if (k.nv) cons[nc++] = k;
    }

    /* components over the full frontier (:212-237) */
    uint8_t *adj = calloc((size_t)nv * (size_t)nv, 1);
    for (int c = 0; c < nc; ++c)
        for (int i = 0; i < cons[c].nv; ++i)
            for (int j = i + 1; j < cons[c].nv; ++j) {
                adj[(size_t)cons[c].vars[i] * nv + cons[c].vars[j]] = 1;
                adj[(size_t)cons[c].vars[j] * nv + cons[c].vars[i]] = 1;
            }
    int *comp_id = malloc(sizeof(int) * (size_t)nv), *stack = malloc(sizeof(int) * (size_t)nv);
    int *csize = calloc((size_t)nv, sizeof(int)), *cfirst = malloc(sizeof(int) * (size_t)nv);
    int ncomp = 0;
    for (int v = 0; v < nv; ++v) comp_id[v] = -1;
    for (int v = 0; v < nv; ++v) {
        if (comp_id[v] >= 0) continue;
        int sp = 0;
        stack[sp++] = v; comp_id[v] = ncomp; cfirst[ncomp] = v;
        while (sp) {
            const int cur = stack[--sp];
            csize[ncomp]++;
            for (int o = 0; o < nv; ++o)
                if (adj[(size_t)cur * nv + o] && comp_id[o] < 0) { comp_id[o] = ncomp; stack[sp++] = o; }
        }
        ncomp++;
    }
    for (int v = 0; v < nv; ++v) comp_of_cell[cell_of_var[v]] = (int16_t)cell_of_var[cfirst[comp_id[v]]];
    for (int k = 0; k < ncomp; ++k) comp_size[cell_of_var[cfirst[k]]] = (int16_t)csize[k];

    /* unit + subset rules (:268-329) */
    int *assign = malloc(sizeof(int) * (size_t)nv);                              /* -1 / 0 / 1 */
    for (int v = 0; v < nv; ++v) assign[v] = -1;
    int n_safe = 0, changed = 1;
    int rem_a[8], rem_b[8];
#define REMAINING(K, REM, NREM, TGT)                                         \
    do {                                                                     \
        NREM = 0; TGT = (K).target;                                          \
        for (int q_ = 0; q_ < (K).nv; ++q_) {                                \
            const int v_ = (K).vars[q_];                                     \
            if (assign[v_] < 0) REM[NREM++] = v_;                            \
            else if (assign[v_] == 1) TGT -= 1;                              \
        }                                                                    \
    } while (0)
    while (changed) {
        changed = 0;
        for (int c = 0; c < nc; ++c) {                                           /* unit propagation */
            int nr, tg;
            REMAINING(cons[c], rem_a, nr, tg);
            if (tg < 0 || tg > nr) continue;
            if (tg == 0) {
                for (int q = 0; q < nr; ++q) if (assign[rem_a[q]] < 0) { assign[rem_a[q]] = 0; n_safe++; changed = 1; }
            } else if (tg == nr) {
                for (int q = 0; q < nr; ++q) if (assign[rem_a[q]] < 0) { assign[rem_a[q]] = 1; changed = 1; }
            }
        }
        if (changed) continue;
        for (int i = 0; i < nc && !changed; ++i) {                               /* subset rule */
            int na, ta;
            REMAINING(cons[i], rem_a, na, ta);
            if (!na) continue;
            for (int j = 0; j < nc; ++j) {
                if (i == j) continue;
                int nb, tb;
                REMAINING(cons[j], rem_b, nb, tb);
                if (!nb) continue;
                int subset = 1;
                for (int x = 0; x < na && subset; ++x) {
                    int found = 0;
                    for (int y = 0; y < nb; ++y) found |= rem_b[y] == rem_a[x];
                    subset = found;
                }
                if (!subset) continue;
                int diff[8], nd = 0;
                for (int y = 0; y < nb; ++y) {
                    int in_a = 0;
                    for (int x = 0; x < na; ++x) in_a |= rem_a[x] == rem_b[y];
                    if (!in_a) diff[nd++] = rem_b[y];
                }
                if (!nd) continue;
                if (ta == tb) {
                    for (int q = 0; q < nd; ++q) if (assign[diff[q]] < 0) { assign[diff[q]] = 0; n_safe++; changed = 1; }
                    if (changed) break;
                } else if (tb - ta == nd) {
                    for (int q = 0; q < nd; ++q) if (assign[diff[q]] < 0) { assign[diff[q]] = 1; changed = 1; }
                    if (changed) break;
                }
            }
        }
    }

    uint8_t fl = 2 | 4;                 /* bit1: has frontier, bit2: first click done */
    if (n_safe) {                                                                /* :343-351 */
        for (int v = 0; v < nv; ++v) if (assign[v] == 0) safe[cell_of_var[v]] = 1;
        fl |= 1;
    } else {
        /* exact search per component on the constraints that still have free variables (:331-375) */
        solver_t s;
        s.cvars = malloc(sizeof(int[8]) * (size_t)nc);
        s.cnv = malloc(sizeof(int) * (size_t)nc); s.targets = malloc(sizeof(int) * (size_t)nc);
        s.v2c = malloc(sizeof(int[8]) * (size_t)nv); s.v2cn = malloc(sizeof(int) * (size_t)nv);
        s.order = malloc(sizeof(int) * (size_t)nv); s.assignment = malloc(sizeof(int) * (size_t)nv);
        s.assigned_sum = malloc(sizeof(int) * (size_t)nc); s.unknown_count = malloc(sizeof(int) * (size_t)nc);
        int *local_of = malloc(sizeof(int) * (size_t)nv), *free_vars = malloc(sizeof(int) * (size_t)nv);
        int any = 0;
        for (int k = 0; k < ncomp; ++k) {
            int nf = 0;
            for (int v = 0; v < nv; ++v) {                                       /* free_vars in comp_vars order is */
                local_of[v] = -1;                                               /* irrelevant to the answer: use index order */
                if (comp_id[v] == k && assign[v] < 0) { local_of[v] = nf; free_vars[nf++] = v; }
            }
            if (!nf) continue;
            s.num_vars = nf; s.num_cons = 0;
            for (int v = 0; v < nf; ++v) s.v2cn[v] = 0;
            for (int c = 0; c < nc; ++c) {
                int nr, tg;
                REMAINING(cons[c], rem_a, nr, tg);
                if (!nr || comp_id[rem_a[0]] != k) continue;
                const int ci = s.num_cons++;
                s.cnv[ci] = nr; s.targets[ci] = tg;
                for (int q = 0; q < nr; ++q) {
                    const int lv = local_of[rem_a[q]];
                    s.cvars[ci][q] = lv;
                    s.v2c[lv][s.v2cn[lv]++] = ci;
                }
            }
            if (!s.num_cons) continue;
            /* order = sorted(range(num_vars), key=degree, reverse=True): stable, so ties keep index order (:55) */
            int pos = 0;
            for (int deg = 8; deg >= 0; --deg)
                for (int v = 0; v < nf; ++v) if (s.v2cn[v] == deg) s.order[pos++] = v;
            for (int v = 0; v < nf; ++v)
                if (!sv_feasible_with_mine(&s, v)) { safe[cell_of_var[free_vars[v]]] = 1; any = 1; }
        }
        if (any) fl |= 1;
        free(s.cvars); free(s.cnv); free(s.targets); free(s.v2c); free(s.v2cn); free(s.order);
        free(s.assignment); free(s.assigned_sum); free(s.unknown_count); free(local_of); free(free_vars);
    }
#undef REMAINING
    *flags_out = fl;
    free(var_of_cell); free(cell_of_var); free(cons); free(adj); free(comp_id); free(stack); free(csize);
    free(cfirst); free(assign);
}

void orc_avoidability(const orc_cfg *cfg, int64_t n, const orc_state *st, uint8_t *safe, int16_t *comp_of_cell,
                      int16_t *comp_size, uint8_t *flags)
{
    const int HW = cfg->H * cfg->W;
    for (int64_t i = 0; i < n; ++i)
        avoid_one(cfg->H, cfg->W, st->mine + i * HW, st->revealed + i * HW, st->flags + i * HW, st->counts + i * HW,
                  st->first_click_done[i], safe + i * HW, comp_of_cell + i * HW, comp_size + i * HW, flags + i);
}

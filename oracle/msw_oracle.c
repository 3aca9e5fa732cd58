/*
 * msw_oracle.c -- CPU restatement of the minesweeper-ppo rollout hot path.
 * TEST INFRASTRUCTURE ONLY; see msw_oracle.h for scope, citations and the
 * parity pin (tests/golden/, generated from the live reference).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared).
 */
#include "msw_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_VERSION 1
#define ORC_MAX_CELLS 1024

int orc_version(void) { return ORC_VERSION; }

/* ------------------------------------------------------------------------ */
/* Minimal pthread parallel-for (no OpenMP: the image's default CC has no    */
/* libgomp).  Envs are independent (env.py:491-505 touches only envs[i]).    */
/* ------------------------------------------------------------------------ */
typedef void (*orc_range_fn)(void *ctx, int64_t lo, int64_t hi);
typedef struct { orc_range_fn fn; void *ctx; int64_t lo, hi; } orc_job;

static void *orc_job_main(void *p)
{
    orc_job *j = (orc_job *)p;
    j->fn(j->ctx, j->lo, j->hi);
    return 0;
}

static void orc_parallel_for(int64_t n, int nthreads, orc_range_fn fn, void *ctx)
{
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1 || n < 2 * (int64_t)nthreads) { fn(ctx, 0, n); return; }
    pthread_t tid[256];
    orc_job job[256];
    int started[256];
    const int64_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        job[t].fn = fn; job[t].ctx = ctx;
        job[t].lo = t * chunk;
        job[t].hi = job[t].lo + chunk < n ? job[t].lo + chunk : n;
        started[t] = 0;
        if (job[t].lo >= job[t].hi) continue;
        if (t == nthreads - 1 || pthread_create(&tid[t], 0, orc_job_main, &job[t]) != 0)
            orc_job_main(&job[t]);          /* last chunk (or a failed spawn) runs inline */
        else
            started[t] = 1;
    }
    for (int t = 0; t < nthreads; ++t)
        if (started[t]) pthread_join(tid[t], 0);
}

/* ------------------------------------------------------------------------ */
/* Counter-based RNG (Philox4x32-10, Salmon et al. 2011).  Not part of the   */
/* reference -- its layouts come from NumPy's PCG64 Generator.choice stream  */
/* (env.py:49, 309), which is outside the parity contract.      */
/* ------------------------------------------------------------------------ */
void orc_philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t c[4], uint32_t out[4])
{
    uint32_t x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * x0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * x2;
        uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ k0;
        uint32_t y1 = (uint32_t)p1;
        uint32_t y2 = (uint32_t)(p0 >> 32) ^ x3 ^ k1;
        uint32_t y3 = (uint32_t)p0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

/*
 * First-click-safe placement, env.py:280-312.  The forbidden set and the
 * "relax to the clicked cell only" fallback follow the reference; the choice
 * of the subset uses the counter-based sampler specified in DESIGN.md:
 * 16-bit Lemire draws over all cells in Philox stream order, rejecting
 * forbidden/already-chosen cells, choosing the complement when more than half
 * the allowed cells are mines.  Sequential rejection sampling without replacement is exactly
 * uniform over mine_count-subsets of the allowed cells.
 */
void orc_place_mines(const orc_cfg *cfg, int64_t env_id, uint32_t episode,
                     int r0, int c0, uint8_t *mine)
{
    const int H = cfg->H, W = cfg->W, HW = H * W, M = cfg->mine_count;
    uint8_t forbidden[ORC_MAX_CELLS];
    uint8_t chosen[ORC_MAX_CELLS];
    memset(forbidden, 0, (size_t)HW);
    memset(chosen, 0, (size_t)HW);

    if (cfg->safe_nbhd) {                                   /* env.py:288-299 */
        for (int dr = -1; dr <= 1; ++dr)
            for (int dc = -1; dc <= 1; ++dc) {
                int r = r0 + dr, c = c0 + dc;
                if (r >= 0 && r < H && c >= 0 && c < W) forbidden[r * W + c] = 1;
            }
    }
    forbidden[r0 * W + c0] = 1;                             /* env.py:300 */
    int allowed = 0;
    for (int i = 0; i < HW; ++i) allowed += !forbidden[i];
    if (allowed < M) {                                      /* env.py:303-307 */
        memset(forbidden, 0, (size_t)HW);
        forbidden[r0 * W + c0] = 1;
        allowed = HW - 1;
    }
    if (M > allowed) {           /* rng.choice would raise; the product rejects */
        memset(mine, 0, (size_t)HW);   /* this config at construction time.      */
        return;
    }

    const int complement = (2 * M > allowed);
    const int K = complement ? allowed - M : M;
    const uint32_t thresh = 65536u % (uint32_t)HW;
    const uint32_t k0 = (uint32_t)cfg->seed, k1 = (uint32_t)(cfg->seed >> 32);
    uint32_t ctr[4] = { (uint32_t)(uint64_t)env_id, (uint32_t)((uint64_t)env_id >> 32), episode, 0 };
    /* Draw p is 16-bit slot (p%256)/32 of Philox block 32*(p/256) + p%32 ("board
     * sampler"): 32 consecutive draws come from 32 consecutive counter blocks. */
    uint32_t w[32][4];
    uint32_t have_batch = 0xFFFFFFFFu;
    uint32_t p = 0;
    int cnt = 0;
    while (cnt < K) {
        const uint32_t batch = p >> 8, slot = (p & 255u) >> 5, ln = p & 31u;
        ++p;
        if (batch != have_batch) {
            for (uint32_t l = 0; l < 32; ++l) {
                ctr[3] = batch * 32u + l;
                orc_philox4x32_10(k0, k1, ctr, w[l]);
            }
            have_batch = batch;
        }
        uint32_t x = (w[ln][slot >> 1] >> (16u * (slot & 1u))) & 0xFFFFu;
        uint32_t m = x * (uint32_t)HW;
        if ((m & 0xFFFFu) < thresh) continue;
        uint32_t d = m >> 16;
        if (forbidden[d] || chosen[d]) continue;
        chosen[d] = 1;
        ++cnt;
    }
    for (int i = 0; i < HW; ++i)
        mine[i] = complement ? (uint8_t)(!forbidden[i] && !chosen[i]) : chosen[i];
}

/* env.py:314-335: 8-neighbour sum, defined for every cell, self excluded. */
void orc_adjacent_counts(int H, int W, const uint8_t *mine, uint8_t *counts)
{
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            int s = 0;
            for (int dr = -1; dr <= 1; ++dr)
                for (int dc = -1; dc <= 1; ++dc) {
                    if (!dr && !dc) continue;
                    int rr = r + dr, cc = c + dc;
                    if (rr >= 0 && rr < H && cc >= 0 && cc < W) s += mine[rr * W + cc] != 0;
                }
            counts[r * W + c] = (uint8_t)s;
        }
}

/* env_numba.py:16-77: array-queue BFS with a `queued` map. */
int orc_flood_fill(int H, int W, uint8_t *revealed, const uint8_t *flags,
                   const uint8_t *mine, const uint8_t *counts, int r, int c)
{
    const int start = r * W + c;
    if (revealed[start] || flags[start]) return 0;          /* :25-26 */
    if (mine[start]) return 0;                              /* :28-29 */

    int16_t queue[ORC_MAX_CELLS];
    uint8_t queued[ORC_MAX_CELLS];
    memset(queued, 0, (size_t)(H * W));
    int head = 0, tail = 0, newly = 0;
    queue[tail++] = (int16_t)start;
    queued[start] = 1;
    while (head < tail) {
        int cell = queue[head++];
        if (revealed[cell] || flags[cell]) continue;        /* :50-51 */
        if (mine[cell]) continue;                           /* :52-53 */
        revealed[cell] = 1;
        ++newly;
        if (counts[cell] != 0) continue;                    /* :58 */
        int rr = cell / W, cc = cell % W;
        for (int dr = -1; dr <= 1; ++dr)
            for (int dc = -1; dc <= 1; ++dc) {
                if (!dr && !dc) continue;
                int nr = rr + dr, nc = cc + dc;
                if (nr < 0 || nr >= H || nc < 0 || nc >= W) continue;
                int nb = nr * W + nc;
                if (queued[nb]) continue;
                if (revealed[nb] || flags[nb] || mine[nb]) continue;
                queue[tail++] = (int16_t)nb;
                queued[nb] = 1;
            }
    }
    return newly;
}

/* env.py:172-192 (obs), :194-196 (mask), train_rl.py:205-212 (labels/valid). */
void orc_encode(int H, int W, const uint8_t *revealed, const uint8_t *flags,
                const uint8_t *mine, const uint8_t *counts, int first_click_done,
                float *obs, uint8_t *mask, float *labels, uint8_t *valid)
{
    const int HW = H * W;
    memset(obs, 0, sizeof(float) * 10u * (size_t)HW);
    for (int i = 0; i < HW; ++i) {
        if (revealed[i]) {
            obs[i] = 1.0f;                                          /* ch 0 */
            if (first_click_done) obs[(1 + counts[i]) * HW + i] = 1.0f;  /* env.py:181-185 */
        }
        mask[i] = (uint8_t)!revealed[i];
    }
    if (labels) {
        for (int i = 0; i < HW; ++i)
            labels[i] = (first_click_done && mine[i]) ? 1.0f : 0.0f;
    }
    if (valid) {
        for (int i = 0; i < HW; ++i)
            valid[i] = (uint8_t)(first_click_done && !revealed[i] && !flags[i]);
    }
}

static void reset_one(int HW, orc_state *st, int64_t i)    /* env.py:87-95 */
{
    memset(st->mine + i * HW, 0, (size_t)HW);
    memset(st->counts + i * HW, 0, (size_t)HW);
    memset(st->revealed + i * HW, 0, (size_t)HW);
    memset(st->flags + i * HW, 0, (size_t)HW);
    st->first_click_done[i] = 0;
    st->step_count[i] = 0;
    st->last_new_reveals[i] = 0;
    st->episode_idx[i] += 1u;
}

typedef struct {
    const orc_cfg *cfg; orc_state *st;
    float *obs; uint8_t *mask; float *labels; uint8_t *valid;
} reset_ctx;

static void reset_range(void *p, int64_t lo, int64_t hi)
{
    reset_ctx *x = (reset_ctx *)p;
    const int H = x->cfg->H, W = x->cfg->W, HW = H * W;
    orc_state *st = x->st;
    for (int64_t i = lo; i < hi; ++i) {
        reset_one(HW, st, i);
        orc_encode(H, W, st->revealed + i * HW, st->flags + i * HW, st->mine + i * HW,
                   st->counts + i * HW, 0, x->obs + i * 10 * HW, x->mask + i * HW,
                   x->labels ? x->labels + i * HW : 0, x->valid ? x->valid + i * HW : 0);
    }
}

void orc_vec_reset(const orc_cfg *cfg, int64_t n, orc_state *st,
                   float *obs, uint8_t *mask, float *labels, uint8_t *valid,
                   int nthreads)
{
    reset_ctx x = { cfg, st, obs, mask, labels, valid };
    orc_parallel_for(n, nthreads, reset_range, &x);
}

static void encode_range(void *p, int64_t lo, int64_t hi)
{
    reset_ctx *x = (reset_ctx *)p;
    const int H = x->cfg->H, W = x->cfg->W, HW = H * W;
    orc_state *st = x->st;
    for (int64_t i = lo; i < hi; ++i)
        orc_encode(H, W, st->revealed + i * HW, st->flags + i * HW, st->mine + i * HW, st->counts + i * HW,
                   st->first_click_done[i], x->obs + i * 10 * HW, x->mask + i * HW,
                   x->labels ? x->labels + i * HW : 0, x->valid ? x->valid + i * HW : 0);
}

void orc_vec_encode(const orc_cfg *cfg, int64_t n, orc_state *st, float *obs, uint8_t *mask, float *labels,
                    uint8_t *valid, int nthreads)
{
    reset_ctx x = { cfg, st, obs, mask, labels, valid };
    orc_parallel_for(n, nthreads, encode_range, &x);
}

typedef struct {
    const orc_cfg *cfg; int64_t env_id_base; const int64_t *actions;
    const uint8_t *inject_mine, *inject_sel; orc_state *st; orc_step_out *out;
} step_ctx;

static void step_range(void *p, int64_t lo, int64_t hi)
{
    step_ctx *x = (step_ctx *)p;
    const orc_cfg *cfg = x->cfg;
    orc_state *st = x->st;
    orc_step_out *out = x->out;
    const int H = cfg->H, W = cfg->W, HW = H * W;
    const int total_safe = HW - cfg->mine_count;                /* env.py:116 */
    for (int64_t i = lo; i < hi; ++i) {
        uint8_t *mine = st->mine + i * HW, *rev = st->revealed + i * HW;
        uint8_t *flg = st->flags + i * HW, *cnt = st->counts + i * HW;

        int64_t a = x->actions[i] % HW;                         /* env.py:106 */
        if (a < 0) a += HW;                                     /* Python modulo */
        const int cell = (int)a, r = cell / W, c = cell % W;

        double reward = 0.0;
        int done = 0, outcome = 0, newly = 0;
        st->last_new_reveals[i] = 0;                            /* env.py:113 */

        if (!rev[cell]) {                                       /* env.py:118 */
            if (!st->first_click_done[i]) {                     /* env.py:119-122 */
                if (x->inject_sel && x->inject_sel[i])
                    memcpy(mine, x->inject_mine + i * HW, (size_t)HW);
                else
                    orc_place_mines(cfg, x->env_id_base + i, st->episode_idx[i], r, c, mine);
                orc_adjacent_counts(H, W, mine, cnt);
                st->first_click_done[i] = 1;
            }
            if (mine[cell]) {                                   /* env.py:124-128 */
                rev[cell] = 1;
                done = 1;
                outcome = 2;
                reward += cfg->loss_reward;
            } else {                                            /* env.py:129-137 */
                newly = orc_flood_fill(H, W, rev, flg, mine, cnt, r, c);
                st->last_new_reveals[i] = newly;
                int total = 0;
                for (int k = 0; k < HW; ++k) total += rev[k];
                if (total >= total_safe) {
                    done = 1;
                    outcome = 1;
                    reward += cfg->win_reward;
                }
            }
        }
        reward -= cfg->step_penalty;                            /* env.py:142 */
        st->step_count[i] += 1;                                 /* env.py:143 */

        /* aux of the pre-reset state, env.py:163-170 via :494 */
        int total = 0;
        for (int k = 0; k < HW; ++k) total += rev[k];
        out->reward[i] = (float)reward;                         /* env.py:501 (f64 -> f32 store) */
        out->done[i] = (uint8_t)done;
        out->outcome[i] = (int8_t)(done ? outcome : 0);         /* env.py:493 */
        out->new_reveals[i] = newly;
        out->step[i] = st->step_count[i];
        out->revealed_count[i] = total;

        if (done) reset_one(HW, st, i);                         /* env.py:497-498 */

        orc_encode(H, W, rev, flg, mine, cnt, st->first_click_done[i],
                   out->obs + i * 10 * HW, out->mask + i * HW,
                   out->mine_labels ? out->mine_labels + i * HW : 0,
                   out->mine_valid ? out->mine_valid + i * HW : 0);
    }
}

void orc_vec_step(const orc_cfg *cfg, int64_t n, int64_t env_id_base,
                  const int64_t *actions, const uint8_t *inject_mine,
                  const uint8_t *inject_sel, orc_state *st, orc_step_out *out,
                  int nthreads)
{
    step_ctx x = { cfg, env_id_base, actions, inject_mine, inject_sel, st, out };
    orc_parallel_for(n, nthreads, step_range, &x);
}

/* ---- late start (env.py:416-466) ---------------------------------------------------------- */
typedef struct { uint32_t k0, k1, c[3], idx, w[4]; } late_rng;

static uint32_t late_next(late_rng *r)
{
    if ((r->idx & 3u) == 0u) {
        uint32_t ctr[4] = { r->c[0], r->c[1], r->c[2], r->idx >> 2 };
        orc_philox4x32_10(r->k0, r->k1, ctr, r->w);
    }
    return r->w[r->idx++ & 3u];
}

static uint32_t late_below(late_rng *r, uint32_t range)
{
    const uint32_t thresh = (0u - range) % range;
    for (;;) {
        const uint64_t m = (uint64_t)late_next(r) * range;
        if ((uint32_t)m >= thresh) return (uint32_t)(m >> 32);
    }
}

/* MinesweeperEnv.step (env.py:103-152) on one env of the state arrays; returns done. */
static int late_click(const orc_cfg *cfg, int64_t env_id, orc_state *st, int64_t i, int cell)
{
    const int H = cfg->H, W = cfg->W, HW = H * W;
    uint8_t *mine = st->mine + i * HW, *rev = st->revealed + i * HW;
    uint8_t *flg = st->flags + i * HW, *cnt = st->counts + i * HW;
    int done = 0;
    st->last_new_reveals[i] = 0;
    if (!rev[cell]) {
        if (!st->first_click_done[i]) {
            orc_place_mines(cfg, env_id, st->episode_idx[i], cell / W, cell % W, mine);
            orc_adjacent_counts(H, W, mine, cnt);
            st->first_click_done[i] = 1;
        }
        if (mine[cell]) {
            rev[cell] = 1;
            done = 1;
        } else {
            st->last_new_reveals[i] = orc_flood_fill(H, W, rev, flg, mine, cnt, cell / W, cell % W);
            int total = 0;
            for (int k = 0; k < HW; ++k) total += rev[k];
            if (total >= HW - cfg->mine_count) done = 1;
        }
    }
    st->step_count[i] += 1;
    return done;
}

void orc_late_start(const orc_cfg *cfg, int64_t n, int64_t env_id_base, orc_state *st,
                    const uint8_t *sel, uint64_t late_seed, float prob, int min_hidden,
                    int max_hidden, int max_attempts, int max_extra_steps)
{
    const int HW = cfg->H * cfg->W, safe_total = HW - cfg->mine_count;
    const double p24d = (double)prob * 16777216.0;
    const uint32_t prob24 = p24d >= 16777216.0 ? 16777216u : (uint32_t)p24d;
    if (!(prob > 0.0f)) return;                                         /* env.py:421-423 */
    for (int64_t i = 0; i < n; ++i) {
        if (sel && !sel[i]) continue;
        const uint64_t id = (uint64_t)(env_id_base + i);
        late_rng r = { (uint32_t)late_seed, (uint32_t)(late_seed >> 32),
                       { (uint32_t)id, (uint32_t)(id >> 32), st->episode_idx[i] }, 0, {0, 0, 0, 0} };
        if ((late_next(&r) >> 8) >= prob24) continue;
        uint8_t *mine = st->mine + i * HW, *rev = st->revealed + i * HW, *flg = st->flags + i * HW;
        int success = 0;
        for (int attempt = 0; attempt < max_attempts && !success; ++attempt) {
            if (st->first_click_done[i]) reset_one(HW, st, i);          /* env.py:437-438 */
            int done = late_click(cfg, env_id_base + i, st, i, (int)late_below(&r, (uint32_t)HW));
            if (done) continue;                                         /* env.py:443-444 */
            int target = min_hidden + (int)late_below(&r, (uint32_t)(max_hidden - min_hidden + 1));
            if (target < 1) target = 1;
            if (target > safe_total) target = safe_total;               /* env.py:447 */
            for (int e = 0; e < max_extra_steps; ++e) {                 /* env.py:449-458 */
                int revealed = 0, ncand = 0;
                for (int k = 0; k < HW; ++k) revealed += rev[k];
                if (safe_total - revealed <= target) { success = 1; break; }
                for (int k = 0; k < HW; ++k) ncand += (!mine[k] && !rev[k] && !flg[k]);
                if (ncand == 0) break;
                int pick = (int)late_below(&r, (uint32_t)ncand), cell = -1;
                for (int k = 0; k < HW; ++k)
                    if (!mine[k] && !rev[k] && !flg[k] && pick-- == 0) { cell = k; break; }
                done = late_click(cfg, env_id_base + i, st, i, cell);
                if (done) break;
            }
            if (!success && !done) {                                    /* env.py:460-462 */
                int revealed = 0;
                for (int k = 0; k < HW; ++k) revealed += rev[k];
                if (safe_total - revealed <= target) success = 1;
            }
        }
        if (!success) reset_one(HW, st, i);                             /* env.py:465-466 */
    }
}

/* ---- rules.analyze_forced_modules (rules.py:206-259), all pairs as the reference ------------ */
void orc_forced_subset(const orc_cfg *cfg, int64_t n, const orc_state *st, uint8_t *out)
{
    const int H = cfg->H, W = cfg->W, HW = H * W;
    static const int DR[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, DC[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    int16_t (*sets)[8] = malloc(sizeof(int16_t[8]) * (size_t)HW);
    int *cnt = malloc(sizeof(int) * (size_t)HW), *mines = malloc(sizeof(int) * (size_t)HW);
    int *keys = malloc(sizeof(int) * (size_t)HW);
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *rev = st->revealed + i * HW, *mine = st->mine + i * HW, *counts = st->counts + i * HW;
        uint8_t *o = out + i * HW;
        memset(o, 0, (size_t)HW);
        int nk = 0;
        for (int a = 0; a < HW; ++a) {                              /* :222-235 */
            if (!(rev[a] && counts[a] > 0)) continue;
            int k = 0, m = 0;
            for (int d = 0; d < 8; ++d) {
                const int r = a / W + DR[d], c = a % W + DC[d];
                if (r < 0 || r >= H || c < 0 || c >= W) continue;
                if (!rev[r * W + c]) {
                    sets[a][k++] = (int16_t)(r * W + c);
                    m += mine[r * W + c] != 0;
                }
            }
            if (k == 0) continue;
            cnt[a] = k; mines[a] = m;
            keys[nk++] = a;
        }
        for (int x = 0; x < nk; ++x)                                /* :237-257, both directions */
            for (int y = 0; y < nk; ++y) {
                if (x == y) continue;
                const int a = keys[x], b = keys[y];
                if (mines[a] != mines[b]) continue;
                int subset = 1;
                for (int k = 0; k < cnt[a] && subset; ++k) {
                    int found = 0;
                    for (int l = 0; l < cnt[b]; ++l) found |= sets[b][l] == sets[a][k];
                    subset = found;
                }
                if (!subset) continue;
                for (int l = 0; l < cnt[b]; ++l) {
                    int in_a = 0;
                    for (int k = 0; k < cnt[a]; ++k) in_a |= sets[a][k] == sets[b][l];
                    if (!in_a) o[sets[b][l]] = 1;
                }
            }
    }
    free(sets); free(cnt); free(mines); free(keys);
}

/*
 * buffers.py:78-94.  torch evaluates, per t from T-1 down to 0 (all fp32,
 * Python scalars rounded to fp32 when they meet an fp32 tensor):
 *   nnt   = 1.0 - float(done[t])
 *   delta = (rewards[t] + (gamma * next_value) * nnt) - values[t]
 *   last  = delta + ((gamma*lam) * nnt) * last
 * Built with -ffp-contract=off so no FMA is formed.
 */
void orc_gae(int64_t T, int64_t N, const float *rewards, const float *values,
             const uint8_t *dones, const float *last_values, float gamma_f32,
             float gamma_lam_f32, float *adv, float *ret)
{
    for (int64_t i = 0; i < N; ++i) {
        float last = 0.0f;
        for (int64_t t = T - 1; t >= 0; --t) {
            const int64_t k = t * N + i;
            const float next_value = (t == T - 1) ? last_values[i] : values[k + N];
            const float nnt = 1.0f - (dones[k] ? 1.0f : 0.0f);
            volatile float gv = gamma_f32 * next_value;
            volatile float gvn = gv * nnt;
            volatile float s = rewards[k] + gvn;
            volatile float delta = s - values[k];
            volatile float cl = gamma_lam_f32 * nnt;
            volatile float cla = cl * last;
            last = delta + cla;
            adv[k] = last;
            ret[k] = last + values[k];
        }
    }
}

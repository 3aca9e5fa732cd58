/*
 * msw_oracle.h -- CPU restatement of the minesweeper-ppo rollout hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product path (minesweeper_ppo_b200/) never links or
 * calls it and fails loudly when the CUDA library is missing.
 *
 * The reference (/root/reference, yakvrz/minesweeper-ppo) is pure Python with
 * one numba kernel; there is nothing to compile from it, so this is a plain-C
 * restatement ("port") that follows, function by function:
 *   minesweeper/env.py:87-152      reset / step
 *   minesweeper/env.py:163-196     aux / obs planes / action mask
 *   minesweeper/env.py:280-335     first-click-safe placement / adjacency counts
 *   minesweeper/env.py:406-414, 468-511   vector reset / step with auto-reset
 *   minesweeper/env_numba.py:16-77 array-queue BFS flood fill
 *   minesweeper/buffers.py:78-94   GAE / returns
 *   train_rl.py:203-219            auxiliary mine labels / valid map
 *
 * Parity pin: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), so the oracle is pinned against outputs of the
 * reference itself, generated in the authoring container by
 * tests/golden/make_golden.py and committed under tests/golden/.
 *
 * Data layout is the reference's own (one byte per cell, row-major [H][W]),
 * deliberately NOT the bitboard layout of the CUDA path, so the two
 * implementations share no representation.
 */
#ifndef MSW_ORACLE_H
#define MSW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_cfg {
    int32_t H, W;
    int32_t mine_count;
    int32_t safe_nbhd;      /* EnvConfig.guarantee_safe_neighborhood */
    double win_reward;      /* EnvConfig.win_reward  (env.py:28) */
    double loss_reward;     /* EnvConfig.loss_reward (env.py:29) */
    double step_penalty;    /* EnvConfig.step_penalty (env.py:30) */
    uint64_t seed;          /* key of the counter-based board sampler */
} orc_cfg;

/* Per-env state, structure-of-arrays over n envs, one byte per cell. */
typedef struct orc_state {
    uint8_t *mine;          /* [n][H*W]  env.mine_mask        */
    uint8_t *revealed;      /* [n][H*W]  env.revealed         */
    uint8_t *flags;         /* [n][H*W]  env.flags            */
    uint8_t *counts;        /* [n][H*W]  env.adjacent_counts  */
    int32_t *first_click_done; /* [n] */
    int32_t *step_count;       /* [n] */
    int32_t *last_new_reveals; /* [n] */
    uint32_t *episode_idx;     /* [n] sampler counter, bumped on every reset */
} orc_state;

typedef struct orc_step_out {
    float   *obs;           /* [n][10][H][W] f32  (env.py:172-192) */
    uint8_t *mask;          /* [n][H*W] bool      (env.py:194-196) */
    float   *reward;        /* [n] */
    uint8_t *done;          /* [n] */
    int8_t  *outcome;       /* [n] 0 none, 1 win, 2 loss (env.py:145, 493) */
    int32_t *new_reveals;   /* [n] aux.last_new_reveals, pre-reset */
    int32_t *step;          /* [n] aux.step, pre-reset */
    int32_t *revealed_count;/* [n] popcount(revealed), pre-reset */
    float   *mine_labels;   /* nullable [n][H][W] f32 (train_rl.py:205-212) */
    uint8_t *mine_valid;    /* nullable [n][H][W] bool */
} orc_step_out;

int orc_version(void);

/* Philox4x32-10 block: key (k0,k1), counter c[4] -> out[4]. */
void orc_philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t c[4], uint32_t out[4]);

/* Board sampler shared (by specification) with the CUDA path: uniform
 * mine_count-subset of the allowed cells, keyed by (seed, env_id, episode). */
void orc_place_mines(const orc_cfg *cfg, int64_t env_id, uint32_t episode,
                     int r0, int c0, uint8_t *mine /* [H*W] out */);

/* env.py:314-335 */
void orc_adjacent_counts(int H, int W, const uint8_t *mine, uint8_t *counts);

/* env_numba.py:16-77; returns number of newly revealed cells. */
int orc_flood_fill(int H, int W, uint8_t *revealed, const uint8_t *flags,
                   const uint8_t *mine, const uint8_t *counts, int r, int c);

/* env.py:172-196 and train_rl.py:205-212 for one env. */
void orc_encode(int H, int W, const uint8_t *revealed, const uint8_t *flags,
                const uint8_t *mine, const uint8_t *counts, int first_click_done,
                float *obs, uint8_t *mask, float *labels, uint8_t *valid);

/* env.py:468-477: zero all state, bump episode counters, emit obs/mask
 * (and labels/valid when non-null). */
void orc_vec_reset(const orc_cfg *cfg, int64_t n, orc_state *st,
                   float *obs, uint8_t *mask, float *labels, uint8_t *valid,
                   int nthreads);

/* Observation of the current state of every env (env.py:172-196). */
void orc_vec_encode(const orc_cfg *cfg, int64_t n, orc_state *st, float *obs, uint8_t *mask,
                    float *labels, uint8_t *valid, int nthreads);

/* env.py:479-511.  inject_sel[i] != 0 => if env i places mines in this step
 * it takes inject_mine[i][H*W] (bytes) instead of sampling. */
void orc_vec_step(const orc_cfg *cfg, int64_t n, int64_t env_id_base,
                  const int64_t *actions, const uint8_t *inject_mine,
                  const uint8_t *inject_sel, orc_state *st, orc_step_out *out,
                  int nthreads);

/* env.py:416-466 for the envs with sel[i] != 0 (all when sel == NULL), which must be
 * freshly reset.  Randomness: draw k of env e is word k%4 of Philox block k/4, key =
 * late_seed, counter = (env id lo, env id hi, episode index at entry, block) -- the
 * stream specified in DESIGN.md and used by the CUDA path (the reference's shared
 * sequential NumPy generator cannot be reproduced in parallel). */
void orc_late_start(const orc_cfg *cfg, int64_t n, int64_t env_id_base, orc_state *st,
                    const uint8_t *sel, uint64_t late_seed, float prob, int min_hidden,
                    int max_hidden, int max_attempts, int max_extra_steps);

/* rules.analyze_forced_modules (rules.py:206-259): out[i][cell] = 1 for the cells of
 * "subset_reveal" of env i.  Restated with the reference's all-pairs loop. */
void orc_forced_subset(const orc_cfg *cfg, int64_t n, const orc_state *st, uint8_t *out);

/* avoidability.analyze_avoidability (avoidability.py:145-394) for every env (msw_oracle_avoid.c).
 * safe[i][cell] = 1 for forced_safe_cells; comp_of_cell[i][cell] = smallest cell index of the frontier
 * component the cell belongs to (-1 off the frontier); comp_size[i][cell] = size of the component whose
 * smallest cell is `cell` (0 elsewhere), so the non-zero entries in index order are component_sizes;
 * flags[i]: bit 0 avoidable, bit 1 frontier non-empty, bit 2 first_click_done.  The chosen-cell fields
 * of the result follow from these (chosen_is_forced_safe = safe[chosen] on the frontier,
 * chosen_component_size = comp_size[comp_of_cell[chosen]]). */
void orc_avoidability(const orc_cfg *cfg, int64_t n, const orc_state *st, uint8_t *safe,
                      int16_t *comp_of_cell, int16_t *comp_size, uint8_t *flags);

/* buffers.py:78-94 in IEEE fp32 without contraction. */
void orc_gae(int64_t T, int64_t N, const float *rewards, const float *values,
             const uint8_t *dones, const float *last_values, float gamma_f32,
             float gamma_lam_f32, float *adv, float *ret);

#ifdef __cplusplus
}
#endif
#endif

/*
 * TEST INFRASTRUCTURE — scalar CPU restatement of the batched MuZero search behind muax.MuZero.act.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load
 * this library; the product (muax_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the algorithm lives in the third-party package `mctx` (unpinned in the reference's
 * setup.py:13-20; newest release at the reference's commit date is 0.0.5) plus `jax.random`; neither is
 * installable in this image and the reference has no tests or golden vectors on this path
 * (SURVEY.md §4, §8c).  This file restates the published algorithm one tree at a
 * time; it is cross-checked against an independently written batched NumPy restatement
 * (oracle/np_mctx.py) and pinned only by the threefry known-answer vectors.
 *
 * Reference call sites restated here:
 *   muax/model.py:222-243   _plan (root inference -> policy -> (plan_output, root.value))
 *   muax/model.py:251-263   _root_inference
 *   muax/model.py:265-282   _recurrent_inference
 *   muax/nn.py:37-44        min_max_normalize
 *   muax/nn.py:59-115       Representation / Prediction / Dynamic (as declarative MLP stacks)
 *   muax/utils.py:70-102    _inv_scaling / support_to_scalar (via include/mz_math.h)
 *   muax/policy.py:13-47    MuZeroPolicy / GumbelMuZeroPolicy kwargs and defaults
 *   mctx (Appendix A.1-A.7) tree, search/simulate/expand/backward, action selection, qtransforms,
 *                           sequential halving, threefry split/uniform/gumbel/categorical
 *
 * Scalar math (exp/log/expm1/...) comes from include/mz_math.h so that this checker and the CUDA
 * kernels agree bit-for-bit.  Build: see oracle/Makefile (-ffp-contract=off is mandatory).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>
#include "../include/mz_math.h"

#define MZO_MAX_LAYERS 8
#define MZO_MAX_ACTIONS 32

typedef struct {
  int32_t n_layers;
  int32_t in_dim[MZO_MAX_LAYERS];
  int32_t out_dim[MZO_MAX_LAYERS];
  int64_t w_off[MZO_MAX_LAYERS]; /* float offsets into the weight blob; W is [in][out] row-major */
  int64_t b_off[MZO_MAX_LAYERS];
} mzo_stack;

typedef struct {
  int32_t batch, num_actions, embed_dim, obs_dim, support_size;
  int32_t policy;      /* 0 muzero, 1 gumbel */
  int32_t qtransform;  /* 0 by_parent_and_siblings, 1 completed_by_mix_value */
  int32_t prng_mode;   /* 0 legacy threefry layout, 1 partitionable */
  int32_t num_simulations, max_depth; /* max_depth <= 0 means None (= num_simulations) */
  int32_t max_considered;
  int32_t activation;  /* 0 elu, 1 relu */
  int32_t repr_minmax, dyn_minmax;
  int32_t global_batch, batch_offset;
  int32_t noise_injected; /* 1: `noise` holds dirichlet noise (muzero) / root gumbel (gumbel policy) */
  float temperature, dirichlet_fraction, dirichlet_alpha, pb_c_init, pb_c_base;
  float gumbel_scale, discount, value_scale, maxvisit_init;
  mzo_stack repr, pred_v, pred_pi, dyn_ns, dyn_r;
} mzo_config;

typedef struct { /* optional tree dump, mctx field names (Appendix A.1); any pointer may be NULL */
  int32_t *node_visits, *parents, *action_from_parent, *children_index, *children_visits;
  float *raw_values, *node_values, *children_prior_logits, *children_values, *children_rewards,
      *children_discounts, *embeddings;
  float *root_noise;      /* [B,A] the dirichlet noise / gumbel actually used */
  int32_t *sim_depth;     /* [B,num_sim] selected path length per simulation */
} mzo_tree_out;

/* ------------------------------------------------------------------ threefry / jax.random (A.7) */

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t *o0, uint32_t *o1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl32(x1, R[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  *o0 = x0;
  *o1 = x1;
}

/* m-th 32-bit word of random_bits(key, n) */
static uint32_t bits_word(const uint32_t key[2], int64_t n, int64_t m, int mode) {
  uint32_t y0, y1;
  if (mode == 0) {
    int64_t half = (n + (n & 1)) / 2;
    int64_t i = m < half ? m : m - half;
    int64_t hi = (half + i < n) ? half + i : 0;
    threefry2x32(key[0], key[1], (uint32_t)i, (uint32_t)hi, &y0, &y1);
    return m < half ? y0 : y1;
  }
  threefry2x32(key[0], key[1], 0u, (uint32_t)m, &y0, &y1);
  return y0 ^ y1;
}

/* j-th key of split(key, num) */
static void split_key(const uint32_t key[2], int64_t num, int64_t j, int mode, uint32_t out[2]) {
  if (mode == 0) {
    out[0] = bits_word(key, 2 * num, 2 * j, 0);
    out[1] = bits_word(key, 2 * num, 2 * j + 1, 0);
  } else {
    threefry2x32(key[0], key[1], 0u, (uint32_t)j, &out[0], &out[1]);
  }
}

static void split2(const uint32_t key[2], int mode, uint32_t a[2], uint32_t b[2]) {
  if (mode == 0) {
    uint32_t p0, p1, q0, q1;
    threefry2x32(key[0], key[1], 0u, 2u, &p0, &p1);
    threefry2x32(key[0], key[1], 1u, 3u, &q0, &q1);
    a[0] = p0; a[1] = q0; b[0] = p1; b[1] = q1;
  } else {
    threefry2x32(key[0], key[1], 0u, 0u, &a[0], &a[1]);
    threefry2x32(key[0], key[1], 0u, 1u, &b[0], &b[1]);
  }
}

static void random_bits_vec(const uint32_t key[2], int n, int mode, uint32_t *out) {
  if (mode == 0) {
    int half = (n + (n & 1)) / 2;
    for (int i = 0; i < half; ++i) {
      uint32_t y0, y1;
      int hi = (half + i < n) ? half + i : 0;
      threefry2x32(key[0], key[1], (uint32_t)i, (uint32_t)hi, &y0, &y1);
      out[i] = y0;
      if (half + i < n) out[half + i] = y1;
    }
  } else {
    for (int i = 0; i < n; ++i) out[i] = bits_word(key, n, i, 1);
  }
}

/* ------------------------------------------------------------------ nets (muax/nn.py:37-115) */

static float act_fn(float x, int kind) { return kind == 0 ? mz_elu(x) : (x > 0.0f ? x : 0.0f); }

/* y = x @ W + b with a fixed accumulation order: acc = 0; acc = fma(x_k, W_kj, acc), k ascending; + b_j.
 * `onehot` >= 0 appends one_hot(onehot, n_onehot) to x (muax/nn.py:105-108): its only nonzero term is
 * 1 * W[in_x + onehot][j]. */
static void dense(const float *W, const float *b, int in_x, int n_onehot, int onehot, int out, const float *x,
                  float *y) {
  (void)n_onehot;
  for (int j = 0; j < out; ++j) {
    float acc = 0.0f;
    for (int k = 0; k < in_x; ++k) acc = fmaf(x[k], W[(int64_t)k * out + j], acc);
    if (onehot >= 0) acc = acc + W[(int64_t)(in_x + onehot) * out + j];
    y[j] = acc + b[j];
  }
}

static void stack_forward(const mzo_stack *s, const float *w, int act, const float *x, int in_x, int n_onehot,
                          int onehot, float *out, float *tmp0, float *tmp1) {
  const float *cur = x;
  float *bufs[2] = {tmp0, tmp1};
  for (int l = 0; l < s->n_layers; ++l) {
    int last = (l == s->n_layers - 1);
    float *dst = last ? out : bufs[l & 1];
    if (l == 0)
      dense(w + s->w_off[l], w + s->b_off[l], in_x, n_onehot, onehot, s->out_dim[l], cur, dst);
    else
      dense(w + s->w_off[l], w + s->b_off[l], s->in_dim[l], 0, -1, s->out_dim[l], cur, dst);
    if (!last)
      for (int j = 0; j < s->out_dim[l]; ++j) dst[j] = act_fn(dst[j], act);
    cur = dst;
  }
}

static void min_max_normalize(float *s, int n) { /* muax/nn.py:37-44 */
  float lo = s[0], hi = s[0];
  for (int i = 1; i < n; ++i) {
    lo = mz_fmin(lo, s[i]);
    hi = mz_fmax(hi, s[i]);
  }
  float scale = hi - lo;
  if (scale < 1e-5f) scale = scale + 1e-5f;
  for (int i = 0; i < n; ++i) s[i] = (s[i] - lo) / scale;
}

static void softmax(const float *x, int n, float *p) { /* jax.nn.softmax: exp(x - max) / sum, left to right */
  float m = x[0];
  for (int i = 1; i < n; ++i) m = mz_fmax(m, x[i]);
  float s = 0.0f;
  for (int i = 0; i < n; ++i) {
    p[i] = mz_expf(x[i] - m);
    s = s + p[i];
  }
  for (int i = 0; i < n; ++i) p[i] = p[i] / s;
}

static float support_from_probs(const float *probs, int S) { /* muax/utils.py:94-102 */
  int F = 2 * S + 1;
  float x = 0.0f;
  for (int i = 0; i < F; ++i) x = x + (float)(i - S) * probs[i];
  return mz_inv_scaling(x);
}

static float support_to_scalar(const float *logits, int S, float *tmp) { /* model.py:260,273-274 */
  softmax(logits, 2 * S + 1, tmp);
  return support_from_probs(tmp, S);
}

/* ------------------------------------------------------------------ tree for ONE env (A.1) */

typedef struct {
  int A, E, N;
  int32_t *node_visits, *parents, *action_from_parent, *children_index, *children_visits;
  float *raw_values, *node_values, *children_prior_logits, *children_values, *children_rewards,
      *children_discounts, *embeddings;
} tree_t;

static void qvalues(const tree_t *t, int node, float *q) {
  for (int a = 0; a < t->A; ++a) {
    int i = node * t->A + a;
    q[a] = t->children_rewards[i] + t->children_discounts[i] * t->children_values[i];
  }
}

static void qtransform_by_parent_and_siblings(const tree_t *t, int node, float *out) { /* A.6 */
  int A = t->A;
  float q[MZO_MAX_ACTIONS];
  qvalues(t, node, q);
  const int32_t *vc = t->children_visits + node * A;
  float nv = t->node_values[node];
  float lo = nv, hi = nv;
  for (int a = 0; a < A; ++a) {
    float safe = vc[a] > 0 ? q[a] : nv;
    lo = mz_fmin(lo, safe);
    hi = mz_fmax(hi, safe);
  }
  float denom = mz_fmax(hi - lo, 1e-8f);
  for (int a = 0; a < A; ++a) {
    float completed = vc[a] > 0 ? q[a] : lo;
    out[a] = (completed - lo) / denom;
  }
}

static void qtransform_completed_by_mix_value(const tree_t *t, int node, float value_scale, float maxvisit_init,
                                              float *out) { /* A.6 */
  int A = t->A;
  float q[MZO_MAX_ACTIONS], p[MZO_MAX_ACTIONS];
  qvalues(t, node, q);
  const int32_t *vc = t->children_visits + node * A;
  float raw = t->raw_values[node];
  softmax(t->children_prior_logits + node * A, A, p);
  int32_t sum_vc = 0, max_vc = 0;
  float sum_p = 0.0f;
  for (int a = 0; a < A; ++a) {
    p[a] = mz_fmax(MZ_F32_TINY, p[a]);
    sum_vc += vc[a];
    if (vc[a] > max_vc) max_vc = vc[a];
    sum_p = sum_p + (vc[a] > 0 ? p[a] : 0.0f);
  }
  float weighted_q = 0.0f;
  for (int a = 0; a < A; ++a) {
    float term = vc[a] > 0 ? (p[a] * q[a]) / sum_p : 0.0f;
    weighted_q = weighted_q + term;
  }
  float mixed = (raw + (float)sum_vc * weighted_q) / (float)(sum_vc + 1);
  float lo = 0.0f, hi = 0.0f;
  for (int a = 0; a < A; ++a) {
    out[a] = vc[a] > 0 ? q[a] : mixed;
    lo = a == 0 ? out[a] : mz_fmin(lo, out[a]);
    hi = a == 0 ? out[a] : mz_fmax(hi, out[a]);
  }
  float denom = mz_fmax(hi - lo, 1e-8f);
  float visit_scale = maxvisit_init + (float)max_vc;
  for (int a = 0; a < A; ++a) out[a] = (visit_scale * value_scale) * ((out[a] - lo) / denom);
}

static void qtransform(const mzo_config *c, const tree_t *t, int node, float *out) {
  if (c->qtransform == 0)
    qtransform_by_parent_and_siblings(t, node, out);
  else
    qtransform_completed_by_mix_value(t, node, c->value_scale, c->maxvisit_init, out);
}

static int masked_argmax(const float *x, const uint8_t *invalid, int A) { /* first max; all invalid -> 0 */
  int best = 0;
  float bv = 0.0f;
  for (int a = 0; a < A; ++a) {
    float v = (invalid && invalid[a]) ? -mz_inf() : x[a];
    if (a == 0 || v > bv) {
      bv = v;
      best = a;
    }
  }
  return best;
}

static float pb_c_of(float node_visit, float pb_c_init, float pb_c_base) {
  return pb_c_init + mz_logf(((node_visit + pb_c_base) + 1.0f) / pb_c_base);
}

static int muzero_action_selection(const mzo_config *c, const tree_t *t, const uint32_t key[2], int node, int depth,
                                   const uint8_t *root_invalid) { /* A.5 */
  int A = t->A;
  const int32_t *vc = t->children_visits + node * A;
  float node_visit = (float)t->node_visits[node];
  float pb_c = pb_c_of(node_visit, c->pb_c_init, c->pb_c_base);
  float probs[MZO_MAX_ACTIONS], value_score[MZO_MAX_ACTIONS], to_argmax[MZO_MAX_ACTIONS];
  uint32_t bits[MZO_MAX_ACTIONS];
  softmax(t->children_prior_logits + node * A, A, probs);
  qtransform(c, t, node, value_score);
  random_bits_vec(key, A, c->prng_mode, bits);
  float sq = sqrtf(node_visit);
  for (int a = 0; a < A; ++a) {
    float policy_score = ((sq * pb_c) * probs[a]) / (float)(vc[a] + 1);
    float noise = 1e-7f * mz_fmax(0.0f, mz_bits_to_unit(bits[a]));
    to_argmax[a] = (value_score[a] + policy_score) + noise;
  }
  return masked_argmax(to_argmax, depth == 0 ? root_invalid : NULL, A);
}

/* seq_halving.get_sequence_of_considered_visits (A.4) */
static void considered_visits_sequence(int m, int n, int32_t *seq) {
  if (m <= 1) {
    for (int i = 0; i < n; ++i) seq[i] = i;
    return;
  }
  int log2max = (int)ceil(log2((double)m));
  int32_t visits[MZO_MAX_ACTIONS];
  for (int i = 0; i < m; ++i) visits[i] = 0;
  int k = m, len = 0;
  while (len < n) {
    int extra = n / (log2max * k);
    if (extra < 1) extra = 1;
    for (int e = 0; e < extra; ++e) {
      for (int i = 0; i < k && len < n; ++i) seq[len++] = visits[i];
      for (int i = 0; i < k; ++i) visits[i] += 1;
    }
    k = k / 2 > 2 ? k / 2 : 2;
  }
}

static void score_considered(int considered_visit, const float *gumbel, const float *logits, const float *q,
                             const int32_t *visits, int A, float *out) { /* A.4 */
  float m = logits[0];
  for (int a = 1; a < A; ++a) m = mz_fmax(m, logits[a]);
  for (int a = 0; a < A; ++a) {
    float l = logits[a] - m;
    float s = mz_fmax(-1e9f, (gumbel[a] + l) + q[a]);
    out[a] = visits[a] == considered_visit ? s : -mz_inf();
  }
}

static int gumbel_root_action_selection(const mzo_config *c, const tree_t *t, const float *root_gumbel,
                                        const uint8_t *root_invalid, const int32_t *table) {
  int A = t->A;
  const int32_t *vc = t->children_visits;
  float q[MZO_MAX_ACTIONS], to_argmax[MZO_MAX_ACTIONS];
  qtransform(c, t, 0, q);
  int num_valid = 0, sim_index = 0;
  for (int a = 0; a < A; ++a) {
    num_valid += (root_invalid && root_invalid[a]) ? 0 : 1;
    sim_index += vc[a];
  }
  int num_considered = c->max_considered < num_valid ? c->max_considered : num_valid;
  int considered_visit = table[num_considered * c->num_simulations + sim_index];
  score_considered(considered_visit, root_gumbel, t->children_prior_logits, q, vc, A, to_argmax);
  return masked_argmax(to_argmax, root_invalid, A);
}

static int gumbel_interior_action_selection(const mzo_config *c, const tree_t *t, int node) {
  int A = t->A;
  const int32_t *vc = t->children_visits + node * A;
  float q[MZO_MAX_ACTIONS], x[MZO_MAX_ACTIONS], p[MZO_MAX_ACTIONS];
  qtransform(c, t, node, q);
  int32_t sum_vc = 0;
  for (int a = 0; a < A; ++a) {
    x[a] = t->children_prior_logits[node * A + a] + q[a];
    sum_vc += vc[a];
  }
  softmax(x, A, p);
  for (int a = 0; a < A; ++a) x[a] = p[a] - (float)vc[a] / (float)(1 + sum_vc);
  return masked_argmax(x, NULL, A);
}

/* recurrent_fn = muax/model.py:265-282 on one row */
typedef struct {
  float *sa, *t0, *t1, *r_logits, *v_logits, *probs;
} scratch_t;

static void recurrent_inference(const mzo_config *c, const float *w, const float *emb, int action, scratch_t *s,
                                float *reward, float *value, float *prior_logits, float *next_emb) {
  int E = c->embed_dim, A = c->num_actions;
  stack_forward(&c->dyn_r, w, c->activation, emb, E, A, action, s->r_logits, s->t0, s->t1);
  stack_forward(&c->dyn_ns, w, c->activation, emb, E, A, action, next_emb, s->t0, s->t1);
  if (c->dyn_minmax) min_max_normalize(next_emb, E);
  stack_forward(&c->pred_v, w, c->activation, next_emb, E, 0, -1, s->v_logits, s->t0, s->t1);
  stack_forward(&c->pred_pi, w, c->activation, next_emb, E, 0, -1, prior_logits, s->t0, s->t1);
  *reward = support_to_scalar(s->r_logits, c->support_size, s->probs);
  *value = support_to_scalar(s->v_logits, c->support_size, s->probs);
}

static void root_inference(const mzo_config *c, const float *w, const float *obs, scratch_t *s, float *value,
                           float *prior_logits, float *emb) { /* muax/model.py:251-263 */
  stack_forward(&c->repr, w, c->activation, obs, c->obs_dim, 0, -1, emb, s->t0, s->t1);
  if (c->repr_minmax) min_max_normalize(emb, c->embed_dim);
  stack_forward(&c->pred_v, w, c->activation, emb, c->embed_dim, 0, -1, s->v_logits, s->t0, s->t1);
  stack_forward(&c->pred_pi, w, c->activation, emb, c->embed_dim, 0, -1, prior_logits, s->t0, s->t1);
  *value = support_to_scalar(s->v_logits, c->support_size, s->probs);
}

/* Production-mode Dirichlet(alpha) draw.  jax.random.dirichlet is NOT bit-reproducible off-XLA
 * (SURVEY.md §7 hard part 2), so the framework defines its own counter-based sampler; this is its
 * restatement.  Stream: threefry(key, c0 = global_row * A + a, c1 = draw index).
 *   draw 0            : boost uniform (alpha < 1 only)
 *   draws 2i+1, 2i+2  : Marsaglia polar pair (x0,x1 of 2i+1) and the acceptance uniform (x0 of 2i+2), i >= 0
 * Marsaglia-Tsang: d = alpha' - 1/3, c = 1/sqrt(9 d); accept when log(u) < x^2/2 + d - d v + d log v. */
static float unit_open(uint32_t bits) { return mz_bits_to_unit(bits) + 5.9604645e-8f; } /* (0, 1] */

static float gamma_draw(const uint32_t key[2], uint32_t idx, float alpha) {
  float boost = 1.0f;
  uint32_t y0, y1;
  if (alpha < 1.0f) {
    threefry2x32(key[0], key[1], idx, 0u, &y0, &y1);
    boost = mz_expf(mz_logf(unit_open(y0)) / alpha);
    alpha = alpha + 1.0f;
  }
  float d = alpha - 0.333333343f;
  float cc = 1.0f / sqrtf(9.0f * d);
  for (uint32_t it = 0; it < 64; ++it) {
    threefry2x32(key[0], key[1], idx, 2 * it + 1, &y0, &y1);
    float v1 = 2.0f * mz_bits_to_unit(y0) - 1.0f;
    float v2 = 2.0f * mz_bits_to_unit(y1) - 1.0f;
    float s = v1 * v1 + v2 * v2;
    if (s >= 1.0f || s == 0.0f) continue;
    float x = v1 * sqrtf((-2.0f * mz_logf(s)) / s);
    float v = 1.0f + cc * x;
    if (v <= 0.0f) continue;
    v = (v * v) * v;
    threefry2x32(key[0], key[1], idx, 2 * it + 2, &y0, &y1);
    float u = unit_open(y0);
    float rhs = ((0.5f * (x * x) + d) - d * v) + d * mz_logf(v);
    if (mz_logf(u) < rhs) return (d * v) * boost;
  }
  return d * boost;
}

static void dirichlet_row(const uint32_t key[2], int64_t global_row, int A, float alpha, float *out) {
  float s = 0.0f;
  for (int a = 0; a < A; ++a) {
    out[a] = gamma_draw(key, (uint32_t)(global_row * A + a), alpha);
    s = s + out[a];
  }
  for (int a = 0; a < A; ++a) out[a] = s > 0.0f ? out[a] / s : 1.0f / (float)A;
}

static void mask_invalid_actions(float *logits, const uint8_t *invalid, int A) { /* A.2 */
  float m = logits[0];
  for (int a = 1; a < A; ++a) m = mz_fmax(m, logits[a]);
  for (int a = 0; a < A; ++a) logits[a] = invalid[a] ? -MZ_F32_MAX : logits[a] - m;
}

/* ------------------------------------------------------------------ one env, whole act */

static void search_one(const mzo_config *c, const float *w, int b, const float *obs, const float *root_logits_in,
                       const float *root_value_in, const float *root_emb_in, const uint8_t *invalid_all,
                       const float *noise_all, const uint32_t dirichlet_or_gumbel_key[2],
                       const uint32_t *simulate_keys, const uint32_t final_key[2], const int32_t *table,
                       int32_t *action_out, float *weights_out, float *root_value_out, const mzo_tree_out *dump) {
  const int A = c->num_actions, E = c->embed_dim, NS = c->num_simulations, N = NS + 1;
  const int F = 2 * c->support_size + 1;
  const int64_t gb = (int64_t)c->batch_offset + b;
  const int max_depth = c->max_depth > 0 ? c->max_depth : NS;
  const uint8_t *invalid = invalid_all ? invalid_all + (int64_t)b * A : NULL;

  int maxw = E + A > F ? E + A : F;
  if (c->obs_dim > maxw) maxw = c->obs_dim;
  const mzo_stack *stacks[5] = {&c->repr, &c->pred_v, &c->pred_pi, &c->dyn_ns, &c->dyn_r};
  for (int s = 0; s < 5; ++s)
    for (int l = 0; l < stacks[s]->n_layers; ++l)
      if (stacks[s]->out_dim[l] > maxw) maxw = stacks[s]->out_dim[l];

  float *fbuf = (float *)calloc((size_t)6 * maxw + (size_t)N * (3 + 4 * A + E) + 8 * A, sizeof(float));
  int32_t *ibuf = (int32_t *)calloc((size_t)N * (3 + 2 * A), sizeof(int32_t));
  scratch_t sc;
  float *fp = fbuf;
  sc.sa = fp; fp += maxw;
  sc.t0 = fp; fp += maxw;
  sc.t1 = fp; fp += maxw;
  sc.r_logits = fp; fp += maxw;
  sc.v_logits = fp; fp += maxw;
  sc.probs = fp; fp += maxw;
  tree_t t;
  t.A = A; t.E = E; t.N = N;
  t.raw_values = fp; fp += N;
  t.node_values = fp; fp += N;
  t.children_prior_logits = fp; fp += N * A;
  t.children_values = fp; fp += N * A;
  t.children_rewards = fp; fp += N * A;
  t.children_discounts = fp; fp += N * A;
  t.embeddings = fp; fp += (size_t)N * E;
  float *root_logits = fp; fp += A;
  float *noise = fp; fp += A;
  float *tmpA = fp; fp += A;
  float *tmpB = fp; fp += A;
  int32_t *ip = ibuf;
  t.node_visits = ip; ip += N;
  t.parents = ip; ip += N;
  t.action_from_parent = ip; ip += N;
  t.children_index = ip; ip += N * A;
  t.children_visits = ip; ip += N * A;
  for (int i = 0; i < N; ++i) t.parents[i] = t.action_from_parent[i] = -1;
  for (int i = 0; i < N * A; ++i) t.children_index[i] = -1;

  /* ---- root (muax/model.py:251-263, or supplied by the caller) */
  float root_value;
  float *root_emb = t.embeddings;
  if (obs) {
    root_inference(c, w, obs + (int64_t)b * c->obs_dim, &sc, &root_value, root_logits, root_emb);
  } else {
    root_value = root_value_in[b];
    memcpy(root_logits, root_logits_in + (int64_t)b * A, sizeof(float) * A);
    memcpy(root_emb, root_emb_in + (int64_t)b * E, sizeof(float) * E);
  }
  root_value_out[b] = root_value; /* raw network value: muax/model.py:243 */

  /* ---- policy prologue (A.2 / A.4) */
  if (c->policy == 0) {
    softmax(root_logits, A, tmpA);
    if (c->noise_injected)
      memcpy(noise, noise_all + (int64_t)b * A, sizeof(float) * A);
    else
      dirichlet_row(dirichlet_or_gumbel_key, gb, A, c->dirichlet_alpha, noise);
    float one_minus = 1.0f - c->dirichlet_fraction;
    for (int a = 0; a < A; ++a) {
      float noisy = one_minus * tmpA[a] + c->dirichlet_fraction * noise[a];
      root_logits[a] = mz_logf(mz_fmax(noisy, MZ_F32_TINY));
    }
    if (invalid) mask_invalid_actions(root_logits, invalid, A);
  } else {
    if (invalid) mask_invalid_actions(root_logits, invalid, A);
    if (c->noise_injected)
      memcpy(noise, noise_all + (int64_t)b * A, sizeof(float) * A);
    else
      for (int a = 0; a < A; ++a) {
        uint32_t bits = bits_word(dirichlet_or_gumbel_key, (int64_t)c->global_batch * A, gb * A + a, c->prng_mode);
        noise[a] = c->gumbel_scale * mz_bits_to_gumbel(bits);
      }
  }

  /* ---- instantiate_tree_from_root (A.3) */
  memcpy(t.children_prior_logits, root_logits, sizeof(float) * A);
  t.raw_values[0] = t.node_values[0] = root_value;
  t.node_visits[0] = 1;

  /* ---- simulations */
  for (int sim = 0; sim < NS; ++sim) {
    uint32_t key[2], sel[2], nk[2];
    split_key(simulate_keys + 2 * sim, c->global_batch, gb, c->prng_mode, key);
    /* simulate */
    int node = 0, action = 0, depth = 0, next;
    for (;;) {
      split2(key, c->prng_mode, nk, sel);
      key[0] = nk[0]; key[1] = nk[1];
      if (c->policy == 0)
        action = muzero_action_selection(c, &t, sel, node, depth, invalid);
      else if (depth == 0)
        action = gumbel_root_action_selection(c, &t, noise, invalid, table);
      else
        action = gumbel_interior_action_selection(c, &t, node);
      next = t.children_index[node * A + action];
      depth += 1;
      if (next == -1 || depth >= max_depth) break;
      node = next;
    }
    if (dump && dump->sim_depth) dump->sim_depth[(int64_t)b * NS + sim] = depth;
    int parent = node;
    if (next == -1) next = sim + 1;
    /* expand */
    float reward, value;
    recurrent_inference(c, w, t.embeddings + (size_t)parent * E, action, &sc, &reward, &value, tmpA, sc.sa);
    /* tmpA = prior logits of the new node, sc.sa = next embedding (width >= E) */
    t.node_visits[next] += 1;
    memcpy(t.children_prior_logits + next * A, tmpA, sizeof(float) * A);
    t.raw_values[next] = t.node_values[next] = value;
    memcpy(t.embeddings + (size_t)next * E, sc.sa, sizeof(float) * E);
    t.children_index[parent * A + action] = next;
    t.children_rewards[parent * A + action] = reward;
    t.children_discounts[parent * A + action] = c->discount;
    t.parents[next] = parent;
    t.action_from_parent[next] = action;
    /* backward */
    int index = next;
    float G = t.node_values[index];
    while (index != 0) {
      int p = t.parents[index];
      int a = t.action_from_parent[index];
      float count = (float)t.node_visits[p];
      G = t.children_rewards[p * A + a] + t.children_discounts[p * A + a] * G;
      t.node_values[p] = (t.node_values[p] * count + G) / (count + 1.0f);
      t.node_visits[p] += 1;
      t.children_values[p * A + a] = t.node_values[index];
      t.children_visits[p * A + a] += 1;
      index = p;
    }
  }

  /* ---- policy epilogue */
  float *wout = weights_out + (int64_t)b * A;
  if (c->policy == 0) {
    float total = 0.0f;
    for (int a = 0; a < A; ++a) total = total + (float)t.children_visits[a];
    for (int a = 0; a < A; ++a)
      wout[a] = total > 0.0f ? (float)t.children_visits[a] / mz_fmax(total, 1.0f) : 1.0f / (float)A;
    float m = 0.0f;
    for (int a = 0; a < A; ++a) {
      tmpA[a] = mz_logf(mz_fmax(wout[a], MZ_F32_TINY));
      m = a == 0 ? tmpA[a] : mz_fmax(m, tmpA[a]);
    }
    float temp = mz_fmax(MZ_F32_TINY, c->temperature);
    for (int a = 0; a < A; ++a) {
      uint32_t bits = bits_word(final_key, (int64_t)c->global_batch * A, gb * A + a, c->prng_mode);
      tmpA[a] = mz_bits_to_gumbel(bits) + (tmpA[a] - m) / temp;
    }
    action_out[b] = masked_argmax(tmpA, NULL, A);
  } else {
    int32_t cv = 0;
    for (int a = 0; a < A; ++a)
      if (t.children_visits[a] > cv) cv = t.children_visits[a];
    qtransform(c, &t, 0, tmpB);
    score_considered(cv, noise, root_logits, tmpB, t.children_visits, A, tmpA);
    action_out[b] = masked_argmax(tmpA, invalid, A);
    for (int a = 0; a < A; ++a) tmpA[a] = root_logits[a] + tmpB[a];
    if (invalid) mask_invalid_actions(tmpA, invalid, A);
    softmax(tmpA, A, wout);
  }

  if (dump) {
#define DUMP(field, per) \
  if (dump->field) memcpy(dump->field + (int64_t)b * (per), t.field, sizeof(*t.field) * (size_t)(per))
    DUMP(node_visits, N);
    DUMP(parents, N);
    DUMP(action_from_parent, N);
    DUMP(children_index, N * A);
    DUMP(children_visits, N * A);
    DUMP(raw_values, N);
    DUMP(node_values, N);
    DUMP(children_prior_logits, N * A);
    DUMP(children_values, N * A);
    DUMP(children_rewards, N * A);
    DUMP(children_discounts, N * A);
    DUMP(embeddings, (size_t)N * E);
#undef DUMP
    if (dump->root_noise) memcpy(dump->root_noise + (int64_t)b * A, noise, sizeof(float) * A);
  }
  free(fbuf);
  free(ibuf);
}

/* ------------------------------------------------------------------ exported entry points */

int mzo_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

typedef struct { /* trees are independent: threads pull rows from a shared counter */
  const mzo_config *c;
  const float *weights, *obs, *root_logits, *root_value, *root_emb;
  const uint8_t *invalid;
  const float *noise;
  uint32_t aux_key[2];
  const uint32_t *simulate_keys;
  uint32_t final_key[2];
  const int32_t *table;
  int32_t *action_out;
  float *action_weights_out, *root_value_out;
  const mzo_tree_out *dump;
  atomic_int next_row;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (;;) {
    int b0 = atomic_fetch_add(&j->next_row, 8);
    if (b0 >= j->c->batch) break;
    int b1 = b0 + 8 < j->c->batch ? b0 + 8 : j->c->batch;
    for (int b = b0; b < b1; ++b)
      search_one(j->c, j->weights, b, j->obs, j->root_logits, j->root_value, j->root_emb, j->invalid, j->noise,
                 j->aux_key, j->simulate_keys, j->final_key, j->table, j->action_out, j->action_weights_out,
                 j->root_value_out, j->dump);
  }
  return NULL;
}

int mzo_search(const mzo_config *c, const float *weights, const float *obs, const float *root_logits,
               const float *root_value, const float *root_emb, const uint8_t *invalid, const float *noise,
               uint32_t key0, uint32_t key1, int32_t *action_out, float *action_weights_out, float *root_value_out,
               const mzo_tree_out *dump, int nthreads) {
  if (c->num_actions < 1 || c->num_actions > MZO_MAX_ACTIONS) return 1;
  if (!obs && !(root_logits && root_value && root_emb)) return 2;
  if (c->noise_injected && !noise) return 3;
  const int NS = c->num_simulations;
  uint32_t rng[2] = {key0, key1}, aux_key[2], search_key[2], final_key[2];
  if (c->policy == 0) { /* rng_key, dirichlet_key, search_key = split(rng_key, 3) */
    split_key(rng, 3, 0, c->prng_mode, final_key);
    split_key(rng, 3, 1, c->prng_mode, aux_key);
    split_key(rng, 3, 2, c->prng_mode, search_key);
  } else { /* rng_key, gumbel_key = split(rng_key); search runs on the new rng_key */
    split2(rng, c->prng_mode, search_key, aux_key);
    final_key[0] = final_key[1] = 0;
  }
  uint32_t *simulate_keys = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (size_t)(NS > 0 ? NS : 1));
  uint32_t cur[2] = {search_key[0], search_key[1]};
  for (int sim = 0; sim < NS; ++sim) { /* rng_key, simulate_key, expand_key = split(rng_key, 3) */
    uint32_t nxt[2];
    split_key(cur, 3, 0, c->prng_mode, nxt);
    split_key(cur, 3, 1, c->prng_mode, simulate_keys + 2 * sim);
    cur[0] = nxt[0]; cur[1] = nxt[1];
  }
  int32_t *table = NULL;
  if (c->policy == 1) {
    int M = c->max_considered;
    if (M < 0 || M > MZO_MAX_ACTIONS) { free(simulate_keys); return 4; }
    table = (int32_t *)calloc((size_t)(M + 1) * (NS > 0 ? NS : 1), sizeof(int32_t));
    for (int m = 0; m <= M; ++m) considered_visits_sequence(m, NS, table + (size_t)m * NS);
  }
  job_t job = {c, weights, obs, root_logits, root_value, root_emb, invalid, noise, {aux_key[0], aux_key[1]},
               simulate_keys, {final_key[0], final_key[1]}, table, action_out, action_weights_out, root_value_out,
               dump, 0};
  if (nthreads <= 0) nthreads = mzo_max_threads();
  if (nthreads > c->batch) nthreads = c->batch > 0 ? c->batch : 1;
  if (nthreads <= 1) {
    worker(&job);
  } else {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int i = 0; i < nthreads - 1; ++i)
      if (pthread_create(&th[started], NULL, worker, &job) == 0) ++started;
    worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
  }
  free(simulate_keys);
  free(table);
  return 0;
}


/* elementwise math for the NumPy restatement and for the mz_math accuracy tests */
void mzo_expf_v(const float *x, float *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = mz_expf(x[i]); }
void mzo_logf_v(const float *x, float *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = mz_logf(x[i]); }
void mzo_expm1f_v(const float *x, float *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = mz_expm1f(x[i]); }
void mzo_inv_scaling_v(const float *x, float *y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = mz_inv_scaling(x[i]); }
void mzo_fmaf_v(const float *a, const float *b, const float *c, float *y, int64_t n) {
  for (int64_t i = 0; i < n; ++i) y[i] = fmaf(a[i], b[i], c[i]);
}
void mzo_threefry_v(uint32_t k0, uint32_t k1, const uint32_t *c0, const uint32_t *c1, uint32_t *o0, uint32_t *o1,
                    int64_t n) {
  for (int64_t i = 0; i < n; ++i) threefry2x32(k0, k1, c0[i], c1[i], &o0[i], &o1[i]);
}
void mzo_dirichlet(uint32_t k0, uint32_t k1, int64_t row0, int64_t rows, int A, float alpha, float *out) {
  uint32_t key[2] = {k0, k1};
  for (int64_t r = 0; r < rows; ++r) dirichlet_row(key, row0 + r, A, alpha, out + r * A);
}
void mzo_considered_visits(int m, int n, int32_t *seq) { considered_visits_sequence(m, n, seq); }
void mzo_support_from_probs(const float *probs, int rows, int S, float *out) {
  for (int r = 0; r < rows; ++r) out[r] = support_from_probs(probs + (int64_t)r * (2 * S + 1), S);
}
void mzo_min_max_normalize(float *s, int rows, int n) {
  for (int r = 0; r < rows; ++r) min_max_normalize(s + (int64_t)r * n, n);
}
void mzo_pb_c(const int32_t *visits, int n, float pb_c_init, float pb_c_base, float *out) {
  for (int i = 0; i < n; ++i) out[i] = pb_c_of((float)visits[i], pb_c_init, pb_c_base);
}

/* Exhaustive check over all 2^32 float bit patterns: branch-free mz_expf / mz_expm1f == early-return forms.
 * Returns the number of mismatching inputs (NaN results compare equal when both are NaN). */
static int same_bits(float a, float b) { return (a != a && b != b) || mz_f2u(a) == mz_f2u(b); }
int64_t mzo_check_branch_free(uint32_t lo, uint32_t hi_inclusive, uint32_t *first_bad) {
  int64_t bad = 0;
  uint64_t u = lo;
  for (;; ++u) {
    float x = mz_u2f((uint32_t)u);
    if (!same_bits(mz_expf(x), mz_expf_ref(x)) || !same_bits(mz_expm1f(x), mz_expm1f_ref(x))) {
      if (bad == 0 && first_bad) *first_bad = (uint32_t)u;
      ++bad;
    }
    if (u == hi_inclusive) break;
  }
  return bad;
}

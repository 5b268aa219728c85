// mz_lane.cuh — "lane" engine: the fused search with lane == tree.
//
// A CTA owns 32 trees for the whole act; trees, weights, activations and the tie-break noise rows live in shared
// memory; ONE search launch per act (plus the noise-table pre-pass for the MuZero policy).
//
//   * tree phases (select, expand + backup, begin, finish): ONE THREAD PER TREE, plain scalar code with loops over
//     the A children — no shuffles, no redundant lanes.  Each warp owns 32/nwarps of the CTA's trees, so the
//     data-dependent walk only couples a handful of trees.
//   * recurrent_fn: lane == tree for every warp; the output units of a layer are dealt to the warps in blocks of
//     4, weights are read as warp-uniform LDS.128 broadcasts straight from the (row-padded) haiku layout, the
//     input activation is one conflict-free LDS per k.  Accumulation is sequential in k per unit: the canonical
//     order of the CPU checkers, so results stay bit-identical.
//   * categorical heads: exps are split over the warps, the two left-to-right sums (softmax denominator and the
//     support expectation) run on one warp per head.
// Profiling history (profiles/): the group engine spent 3x the ideal instruction count at IPC 0.15 per warp on
// shuffles, redundant lanes and per-lane ELU/exp branches; this layout removes all three.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "mz_device.cuh"
#include "mz_fused.cuh"
#include "mz_group.cuh"

namespace mz {

constexpr int kLT = 32;          // trees per CTA (== lanes)
constexpr int kLMaxLayers = 4;

struct LLayer {   // one dense layer in the padded blob: W'[K + extra][out4] then b'[out4]
  int32_t K, extra, out, out4, off, act;
};

struct LPackDesc {
  LLayer l;
  PackSrc src;
  int32_t in_x;  // real input rows before the one-hot rows
};

struct LaneNet {
  LLayer repr[kLMaxLayers], pred_v[kLMaxLayers], pred_pi[kLMaxLayers], dyn_ns[kLMaxLayers], dyn_r[kLMaxLayers];
  int32_t n_repr, n_pred, n_dyn;
  int32_t obs_dim, E, A, S, F, activation, repr_minmax, dyn_minmax;
  int32_t Hmax;  // widest hidden layer
  int32_t packed_floats;
};

struct LaneArgs {
  LaneNet net;
  const float* packed;
  Tree out;
  SearchParams p;
  const float* obs;
  const uint8_t* invalid;
  const float* noise;
  const float* noise_table;   // [B][NS][kGNoiseFloats] or null
  const uint32_t* cont_keys;  // [B][NS][2]
  int32_t K;
  int32_t* action_out;
  float* weights_out;
  float* root_value_out;
  int32_t B, N, dump_tree;
  int32_t walkers;  // lane2: warps that walk trees (select / expand / backup); 0 = all of them
  // warp engine: action / action_weights / root_value are also stored at (pointer + peer_delta[i]) — the same slots of
  // the peer GPUs' gather buffers (NVLink peer stores; muax_b200/sharded.py)
  int32_t n_peers;
  int64_t peer_delta[7];
};

__global__ void lane_pack_kernel(const float* __restrict__ raw, float* __restrict__ packed, LPackDesc d) {
  const int rows = d.l.K + d.l.extra;
  const int total = (rows + 1) * d.l.out4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / d.l.out4, j = i - r * d.l.out4;
    float v = 0.0f;
    if (j < d.src.out) {
      if (r == rows) {
        v = raw[d.src.b_off + j];
      } else {
        int k_src = -1;
        if (r < d.l.K) {
          if (r < d.in_x) k_src = r;
        } else {
          k_src = d.in_x + (r - d.l.K);
        }
        if (k_src >= 0 && k_src < d.src.in) v = raw[d.src.w_off + (int64_t)k_src * d.src.out + j];
      }
    }
    packed[d.l.off + i] = v;
  }
}

// ---------------------------------------------------------------------------------------- per-tree block

struct LaneLayout {
  int node_visits, parents, afp, cindex, cvisits, raw, values, logits, probs, cvalues, rewards, emb, root_noise,
      root_invalid, stride;
};

__host__ __device__ inline LaneLayout lane_layout(int N, int A, int E) {
  LaneLayout L;
  int o = 0;
  auto seg = [&](int n) { const int at = o; o += round_up(n, 4); return at; };
  L.node_visits = seg(N); L.parents = seg(N); L.afp = seg(N);
  L.cindex = seg(N * A); L.cvisits = seg(N * A);
  L.raw = seg(N); L.values = seg(N);
  L.logits = seg(N * A); L.probs = seg(N * A); L.cvalues = seg(N * A); L.rewards = seg(N * A);
  L.emb = seg(N * E);
  L.root_noise = seg(A); L.root_invalid = seg((A + 3) / 4);
  while (o % 32 != 4) o += 4;  // the owner lanes of a warp hit distinct banks on same-offset accesses
  L.stride = o;
  return L;
}

struct LTree {  // one thread's view of its own tree
  int32_t *node_visits, *parents, *afp, *cindex, *cvisits;
  float *raw, *values, *logits, *probs, *cvalues, *rewards, *emb, *root_noise;
  uint8_t* root_invalid;
};

__device__ __forceinline__ LTree lane_tree(float* blk, const LaneLayout& L) {
  LTree t;
  int32_t* ib = reinterpret_cast<int32_t*>(blk);
  t.node_visits = ib + L.node_visits; t.parents = ib + L.parents; t.afp = ib + L.afp;
  t.cindex = ib + L.cindex; t.cvisits = ib + L.cvisits;
  t.raw = blk + L.raw; t.values = blk + L.values; t.logits = blk + L.logits; t.probs = blk + L.probs;
  t.cvalues = blk + L.cvalues; t.rewards = blk + L.rewards; t.emb = blk + L.emb; t.root_noise = blk + L.root_noise;
  t.root_invalid = reinterpret_cast<uint8_t*>(blk + L.root_invalid);
  return t;
}

// ---------------------------------------------------------------------------------------- scalar tree code (one thread)

struct QT {  // qtransform of one node, prepared once per level (Appendix A.6)
  int kind;
  float lo, denom, mixed, mul;
};

__device__ __forceinline__ float lt_q(const LTree& t, int i, float gamma) {
  return MZ_ADD(t.rewards[i], MZ_MUL(gamma, t.cvalues[i]));
}

__device__ __forceinline__ QT lt_qt_prepare(const LTree& t, const SearchParams& p, int node, int A, float gamma) {
  QT qt;
  qt.kind = p.qtransform;
  const int base = node * A;
  if (p.qtransform == MZ_QTRANSFORM_BY_PARENT_AND_SIBLINGS) {
    const float nv = t.values[node];
    float lo = nv, hi = nv;
    for (int a = 0; a < A; ++a)
      if (t.cvisits[base + a] > 0) {
        const float q = lt_q(t, base + a, gamma);
        lo = fminf(lo, q);
        hi = fmaxf(hi, q);
      }
    qt.lo = lo;
    qt.denom = fmaxf(MZ_SUB(hi, lo), 1e-8f);
    qt.mixed = lo;
    qt.mul = 1.0f;
    return qt;
  }
  int sum_vc = 0, max_vc = 0;
  float sum_p = 0.0f;
  for (int a = 0; a < A; ++a) {
    const int vc = t.cvisits[base + a];
    sum_vc += vc;
    max_vc = max(max_vc, vc);
    sum_p = MZ_ADD(sum_p, vc > 0 ? fmaxf(MZ_F32_TINY, t.probs[base + a]) : 0.0f);
  }
  float weighted_q = 0.0f;
  for (int a = 0; a < A; ++a) {
    const bool vis = t.cvisits[base + a] > 0;
    const float term =
        vis ? MZ_DIV(MZ_MUL(fmaxf(MZ_F32_TINY, t.probs[base + a]), lt_q(t, base + a, gamma)), sum_p) : 0.0f;
    weighted_q = MZ_ADD(weighted_q, term);
  }
  const float mixed = MZ_DIV(MZ_ADD(t.raw[node], MZ_MUL((float)sum_vc, weighted_q)), (float)(sum_vc + 1));
  float lo = 0.0f, hi = 0.0f;
  for (int a = 0; a < A; ++a) {
    const float c = t.cvisits[base + a] > 0 ? lt_q(t, base + a, gamma) : mixed;
    lo = a == 0 ? c : fminf(lo, c);
    hi = a == 0 ? c : fmaxf(hi, c);
  }
  qt.lo = lo;
  qt.denom = fmaxf(MZ_SUB(hi, lo), 1e-8f);
  qt.mixed = mixed;
  qt.mul = MZ_MUL(MZ_ADD(p.maxvisit_init, (float)max_vc), p.value_scale);
  return qt;
}

__device__ __forceinline__ float lt_qt_value(const QT& qt, bool visited, float q) {
  const float c = visited ? q : qt.mixed;
  const float n = MZ_DIV(MZ_SUB(c, qt.lo), qt.denom);
  return qt.kind == MZ_QTRANSFORM_BY_PARENT_AND_SIBLINGS ? n : MZ_MUL(qt.mul, n);
}

struct LaneAux {
  const float* noise_row;  // this tree's [K][A] row in shared memory, or null
  int K;
  const uint32_t* cont;    // global: carried key after K levels for (tree, sim)
  const float* pbc;
};

// One level of simulate for MuZero (A.5), Gumbel root / interior (A.4).
__device__ __forceinline__ int lt_select_action(const LTree& t, const SearchParams& p, int node, int depth, int A,
                                                float gamma, const float* pbc, const float* noise_row,
                                                uint32_t s0, uint32_t s1) {
  const int base = node * A;
  const QT qt = lt_qt_prepare(t, p, node, A, gamma);
  int best = 0;
  float bestv = 0.0f;
  if (p.policy == MZ_POLICY_MUZERO) {
    const float explore = pbc[t.node_visits[node]];
    for (int a = 0; a < A; ++a) {
      const int vc = t.cvisits[base + a];
      const float vs = lt_qt_value(qt, vc > 0, lt_q(t, base + a, gamma));
      const float ps = MZ_DIV(MZ_MUL(explore, t.probs[base + a]), (float)(vc + 1));
      const float nz = noise_row != nullptr ? noise_row[a] : tie_break_noise(bits_word(s0, s1, A, a, p.prng_mode));
      float s = MZ_ADD(MZ_ADD(vs, ps), nz);
      if (depth == 0 && t.root_invalid[a] != 0) s = -mz_inf();
      if (a == 0 || s > bestv) {
        bestv = s;
        best = a;
      }
    }
    return best;
  }
  if (depth == 0) {
    int num_valid = 0, sim_index = 0;
    float lmax = -mz_inf();
    for (int a = 0; a < A; ++a) {
      num_valid += t.root_invalid[a] != 0 ? 0 : 1;
      sim_index += t.cvisits[a];
      lmax = fmaxf(lmax, t.logits[a]);
    }
    const int num_considered = min(p.max_considered, num_valid);
    const int cv = p.considered_table[num_considered * p.num_simulations + sim_index];
    for (int a = 0; a < A; ++a) {
      const int vc = t.cvisits[a];
      const float q = lt_qt_value(qt, vc > 0, lt_q(t, a, gamma));
      float s = fmaxf(-1e9f, MZ_ADD(MZ_ADD(t.root_noise[a], MZ_SUB(t.logits[a], lmax)), q));
      if (vc != cv || t.root_invalid[a] != 0) s = -mz_inf();
      if (a == 0 || s > bestv) {
        bestv = s;
        best = a;
      }
    }
    return best;
  }
  float xmax = -mz_inf();
  int sum_vc = 0;
  for (int a = 0; a < A; ++a) {
    const int vc = t.cvisits[base + a];
    sum_vc += vc;
    xmax = fmaxf(xmax, MZ_ADD(t.logits[base + a], lt_qt_value(qt, vc > 0, lt_q(t, base + a, gamma))));
  }
  float sum = 0.0f;
  for (int a = 0; a < A; ++a) {
    const float x = MZ_ADD(t.logits[base + a], lt_qt_value(qt, t.cvisits[base + a] > 0, lt_q(t, base + a, gamma)));
    sum = MZ_ADD(sum, mz_expf(MZ_SUB(x, xmax)));
  }
  for (int a = 0; a < A; ++a) {
    const int vc = t.cvisits[base + a];
    const float x = MZ_ADD(t.logits[base + a], lt_qt_value(qt, vc > 0, lt_q(t, base + a, gamma)));
    const float prob = MZ_DIV(mz_expf(MZ_SUB(x, xmax)), sum);
    const float s = MZ_SUB(prob, MZ_DIV((float)vc, (float)(1 + sum_vc)));
    if (a == 0 || s > bestv) {
      bestv = s;
      best = a;
    }
  }
  return best;
}

// jax.random.split on one thread: (n, s) = split(k).
__device__ __forceinline__ void lt_split2(uint32_t k0, uint32_t k1, int mode, uint32_t& n0, uint32_t& n1, uint32_t& s0,
                                          uint32_t& s1) {
  if (mode == MZ_PRNG_THREEFRY_LEGACY) {
    uint32_t p0, p1, q0, q1;
    threefry2x32(k0, k1, 0u, 2u, p0, p1);
    threefry2x32(k0, k1, 1u, 3u, q0, q1);
    n0 = p0; n1 = q0; s0 = p1; s1 = q1;
  } else {
    threefry2x32(k0, k1, 0u, 0u, n0, n1);
    threefry2x32(k0, k1, 0u, 1u, s0, s1);
  }
}

__device__ __forceinline__ void lt_simulate(const LTree& t, const SearchParams& p, int sim, int A, float gamma,
                                            const LaneAux& aux, int& parent, int& action, int& next, int& depth_out) {
  const bool need_rng = p.policy == MZ_POLICY_MUZERO;
  const bool table = aux.noise_row != nullptr;
  uint32_t k0 = 0, k1 = 0;
  if (need_rng && !table)
    split_key(p.sim_keys[2 * sim], p.sim_keys[2 * sim + 1], (uint32_t)p.global_batch, (uint32_t)p.batch_offset,
              p.prng_mode, k0, k1);
  const int max_depth = p.max_depth > 0 ? p.max_depth : p.num_simulations;
  int node = 0, depth = 0;
  for (;;) {
    uint32_t s0 = 0, s1 = 0;
    const float* row = nullptr;
    if (need_rng) {
      if (table && depth < aux.K) {
        row = aux.noise_row + depth * A;
      } else {
        if (table && depth == aux.K) {
          k0 = __ldg(aux.cont);
          k1 = __ldg(aux.cont + 1);
        }
        lt_split2(k0, k1, p.prng_mode, k0, k1, s0, s1);
      }
    }
    action = lt_select_action(t, p, node, depth, A, gamma, aux.pbc, row, s0, s1);
    next = t.cindex[node * A + action];
    ++depth;
    if (next == kUnvisited || depth >= max_depth) break;
    node = next;
  }
  parent = node;
  depth_out = depth;
  if (next == kUnvisited) next = sim + 1;
}

// expand (A.3) + backward for one tree; `plogits` / `nemb` are this tree's columns of [.][32] activation buffers.
__device__ __forceinline__ void lt_expand_backup(const LTree& t, int A, int E, int parent, int action, int next,
                                                 float reward, float gamma, float value, const float* plogits,
                                                 const float* nemb, int col_stride) {
  float mx = -mz_inf();
  for (int a = 0; a < A; ++a) mx = fmaxf(mx, plogits[a * col_stride]);
  float sum = 0.0f;
  for (int a = 0; a < A; ++a) sum = MZ_ADD(sum, mz_expf(MZ_SUB(plogits[a * col_stride], mx)));
  for (int a = 0; a < A; ++a) {
    const float lg = plogits[a * col_stride];
    t.logits[next * A + a] = lg;
    t.probs[next * A + a] = MZ_DIV(mz_expf(MZ_SUB(lg, mx)), sum);
  }
  for (int e = 0; e < E; ++e) t.emb[next * E + e] = nemb[e * col_stride];
  t.node_visits[next] += 1;
  t.raw[next] = value;
  t.values[next] = value;
  const int edge = parent * A + action;
  t.cindex[edge] = next;
  t.rewards[edge] = reward;
  t.parents[next] = parent;
  t.afp[next] = action;
  int index = next;
  float G_ = value, child_value = value;
  while (index != 0) {
    const int pn = t.parents[index];
    const int e2 = pn * A + t.afp[index];
    const int ci = t.node_visits[pn];
    const float count = (float)ci;
    G_ = MZ_ADD(t.rewards[e2], MZ_MUL(gamma, G_));
    const float pv = MZ_DIV(MZ_ADD(MZ_MUL(t.values[pn], count), G_), MZ_ADD(count, 1.0f));
    t.values[pn] = pv;
    t.node_visits[pn] = ci + 1;
    t.cvalues[e2] = child_value;
    t.cvisits[e2] += 1;
    child_value = pv;
    index = pn;
  }
}

// Policy prologue (A.2 / A.4) + node 0.
__device__ __forceinline__ void lt_begin(const LTree& t, const SearchParams& p, int A, int E, long gb,
                                         const float* rlogits, int col_stride, float root_value, const float* remb,
                                         const uint8_t* invalid, const float* noise) {
  float mx = -mz_inf();
  for (int a = 0; a < A; ++a) mx = fmaxf(mx, rlogits[a * col_stride]);
  if (p.policy == MZ_POLICY_MUZERO) {
    float sum = 0.0f;
    for (int a = 0; a < A; ++a) sum = MZ_ADD(sum, mz_expf(MZ_SUB(rlogits[a * col_stride], mx)));
    float gsum = 0.0f;
    if (noise == nullptr)
      for (int a = 0; a < A; ++a) {
        const float g = gamma_draw(p.aux_key0, p.aux_key1, (uint32_t)(gb * A + a), p.dirichlet_alpha);
        t.root_noise[a] = g;
        gsum = MZ_ADD(gsum, g);
      }
    float lmax = -mz_inf();
    for (int a = 0; a < A; ++a) {
      const float prob = MZ_DIV(mz_expf(MZ_SUB(rlogits[a * col_stride], mx)), sum);
      float nz;
      if (noise != nullptr)
        nz = noise[a];
      else
        nz = gsum > 0.0f ? MZ_DIV(t.root_noise[a], gsum) : MZ_DIV(1.0f, (float)A);
      t.root_noise[a] = nz;
      const float noisy = MZ_ADD(MZ_MUL(MZ_SUB(1.0f, p.dirichlet_fraction), prob), MZ_MUL(p.dirichlet_fraction, nz));
      const float lg = mz_logf(fmaxf(noisy, MZ_F32_TINY));
      t.logits[a] = lg;
      lmax = fmaxf(lmax, lg);
    }
    if (invalid != nullptr)
      for (int a = 0; a < A; ++a) t.logits[a] = invalid[a] != 0 ? -MZ_F32_MAX : MZ_SUB(t.logits[a], lmax);
  } else {
    for (int a = 0; a < A; ++a) {
      float lg = rlogits[a * col_stride];
      if (invalid != nullptr) lg = invalid[a] != 0 ? -MZ_F32_MAX : MZ_SUB(lg, mx);
      t.logits[a] = lg;
      float nz;
      if (noise != nullptr) {
        nz = noise[a];
      } else {
        const uint32_t bits = bits_word(p.aux_key0, p.aux_key1, (uint32_t)p.global_batch * (uint32_t)A,
                                        (uint32_t)(gb * A + a), p.prng_mode);
        nz = MZ_MUL(p.gumbel_scale, mz_bits_to_gumbel(bits));
      }
      t.root_noise[a] = nz;
    }
  }
  float m2 = -mz_inf();
  for (int a = 0; a < A; ++a) {
    t.root_invalid[a] = (invalid != nullptr && invalid[a] != 0) ? 1 : 0;
    m2 = fmaxf(m2, t.logits[a]);
  }
  float s2 = 0.0f;
  for (int a = 0; a < A; ++a) s2 = MZ_ADD(s2, mz_expf(MZ_SUB(t.logits[a], m2)));
  for (int a = 0; a < A; ++a) t.probs[a] = MZ_DIV(mz_expf(MZ_SUB(t.logits[a], m2)), s2);
  for (int e = 0; e < E; ++e) t.emb[e] = remb[e * col_stride];
  t.raw[0] = root_value;
  t.values[0] = root_value;
  t.node_visits[0] = 1;
}

// Policy epilogue: writes action_weights, returns the action.
__device__ __forceinline__ int lt_finish(const LTree& t, const SearchParams& p, int A, float gamma, long gb,
                                         bool has_invalid, float* weights_out) {
  int best = 0;
  float bestv = 0.0f;
  if (p.policy == MZ_POLICY_MUZERO) {
    float total = 0.0f;
    for (int a = 0; a < A; ++a) total = MZ_ADD(total, (float)t.cvisits[a]);
    float lmax = -mz_inf();
    for (int a = 0; a < A; ++a) {
      const float w = total > 0.0f ? MZ_DIV((float)t.cvisits[a], fmaxf(total, 1.0f)) : MZ_DIV(1.0f, (float)A);
      weights_out[a] = w;
      lmax = fmaxf(lmax, mz_logf(fmaxf(w, MZ_F32_TINY)));
    }
    const float temp = fmaxf(MZ_F32_TINY, p.temperature);
    for (int a = 0; a < A; ++a) {
      const float w = total > 0.0f ? MZ_DIV((float)t.cvisits[a], fmaxf(total, 1.0f)) : MZ_DIV(1.0f, (float)A);
      const float l = MZ_DIV(MZ_SUB(mz_logf(fmaxf(w, MZ_F32_TINY)), lmax), temp);
      const uint32_t bits = bits_word(p.final_key0, p.final_key1, (uint32_t)p.global_batch * (uint32_t)A,
                                      (uint32_t)(gb * A + a), p.prng_mode);
      const float s = MZ_ADD(mz_bits_to_gumbel(bits), l);
      if (a == 0 || s > bestv) {
        bestv = s;
        best = a;
      }
    }
    return best;
  }
  const QT qt = lt_qt_prepare(t, p, 0, A, gamma);
  int cv = 0;
  float lmax = -mz_inf(), xmax = -mz_inf();
  for (int a = 0; a < A; ++a) {
    cv = max(cv, t.cvisits[a]);
    lmax = fmaxf(lmax, t.logits[a]);
    xmax = fmaxf(xmax, MZ_ADD(t.logits[a], lt_qt_value(qt, t.cvisits[a] > 0, lt_q(t, a, gamma))));
  }
  for (int a = 0; a < A; ++a) {
    const int vc = t.cvisits[a];
    const float q = lt_qt_value(qt, vc > 0, lt_q(t, a, gamma));
    float s = fmaxf(-1e9f, MZ_ADD(MZ_ADD(t.root_noise[a], MZ_SUB(t.logits[a], lmax)), q));
    if (vc != cv || t.root_invalid[a] != 0) s = -mz_inf();
    if (a == 0 || s > bestv) {
      bestv = s;
      best = a;
    }
  }
  // action_weights = softmax(mask_invalid(logits + completed_q))
  float m3 = -mz_inf();
  for (int a = 0; a < A; ++a) {
    float x = MZ_ADD(t.logits[a], lt_qt_value(qt, t.cvisits[a] > 0, lt_q(t, a, gamma)));
    if (has_invalid) x = t.root_invalid[a] != 0 ? -MZ_F32_MAX : MZ_SUB(x, xmax);
    m3 = fmaxf(m3, x);
  }
  float sum = 0.0f;
  for (int a = 0; a < A; ++a) {
    float x = MZ_ADD(t.logits[a], lt_qt_value(qt, t.cvisits[a] > 0, lt_q(t, a, gamma)));
    if (has_invalid) x = t.root_invalid[a] != 0 ? -MZ_F32_MAX : MZ_SUB(x, xmax);
    sum = MZ_ADD(sum, mz_expf(MZ_SUB(x, m3)));
  }
  for (int a = 0; a < A; ++a) {
    float x = MZ_ADD(t.logits[a], lt_qt_value(qt, t.cvisits[a] > 0, lt_q(t, a, gamma)));
    if (has_invalid) x = t.root_invalid[a] != 0 ? -MZ_F32_MAX : MZ_SUB(x, xmax);
    weights_out[a] = MZ_DIV(mz_expf(MZ_SUB(x, m3)), sum);
  }
  return best;
}

// ---------------------------------------------------------------------------------------- CTA-cooperative MLP, lane == tree

// One block of 4 output units of one dense layer for this lane's tree.
__device__ __forceinline__ void lane_block(const float* __restrict__ w, const LLayer& L, int j0, const float* in,
                                           int onehot, int act_kind, float* out, int lane) {
  const float* wj = w + L.off + j0;
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 4
  for (int k = 0; k < L.K; ++k) {
    const float x = in[k * kLT + lane];
    const float4 wv = *reinterpret_cast<const float4*>(wj + k * L.out4);
    a0 = MZ_FMA(x, wv.x, a0);
    a1 = MZ_FMA(x, wv.y, a1);
    a2 = MZ_FMA(x, wv.z, a2);
    a3 = MZ_FMA(x, wv.w, a3);
  }
  if (onehot >= 0) {
    const float4 wv = *reinterpret_cast<const float4*>(wj + (L.K + onehot) * L.out4);
    a0 = MZ_ADD(a0, wv.x); a1 = MZ_ADD(a1, wv.y); a2 = MZ_ADD(a2, wv.z); a3 = MZ_ADD(a3, wv.w);
  }
  const float4 bv = *reinterpret_cast<const float4*>(wj + (L.K + L.extra) * L.out4);
  a0 = MZ_ADD(a0, bv.x); a1 = MZ_ADD(a1, bv.y); a2 = MZ_ADD(a2, bv.z); a3 = MZ_ADD(a3, bv.w);
  if (L.act) {
    a0 = activate(a0, act_kind); a1 = activate(a1, act_kind); a2 = activate(a2, act_kind); a3 = activate(a3, act_kind);
  }
  if (j0 + 0 < L.out) out[(j0 + 0) * kLT + lane] = a0;
  if (j0 + 1 < L.out) out[(j0 + 1) * kLT + lane] = a1;
  if (j0 + 2 < L.out) out[(j0 + 2) * kLT + lane] = a2;
  if (j0 + 3 < L.out) out[(j0 + 3) * kLT + lane] = a3;
}

// Layer i of one or two heads: the 4-unit blocks of both heads are dealt round-robin to the warps.
__device__ __forceinline__ void lane_dense_pair(const float* w, const LLayer* l0, const LLayer* l1, const float* in0,
                                                const float* in1, float* out0, float* out1, int onehot, int act_kind,
                                                int lane, int warp, int nwarps) {
  const int nb0 = l0->out4 >> 2;
  const int nb1 = l1 != nullptr ? l1->out4 >> 2 : 0;
  for (int g = warp; g < nb0 + nb1; g += nwarps) {
    if (g < nb0)
      lane_block(w, *l0, g << 2, in0, onehot, act_kind, out0, lane);
    else
      lane_block(w, *l1, (g - nb0) << 2, in1, onehot, act_kind, out1, lane);
  }
}

struct LaneBufs {
  float *in, *h0a, *h0b, *h1a, *h1b;
};

// A module = one or two hk.Sequential heads on the same input.  Ends with a CTA barrier.
__device__ __forceinline__ void lane_module(const float* w, const LLayer* s0, const LLayer* s1, int n,
                                            const LaneBufs& b, float* out0, float* out1, int onehot, int act_kind,
                                            int lane, int warp, int nwarps) {
  for (int i = 0; i < n; ++i) {
    const bool last = i + 1 == n;
    const float* in0 = i == 0 ? b.in : ((i & 1) ? b.h0a : b.h0b);
    const float* in1 = i == 0 ? b.in : ((i & 1) ? b.h1a : b.h1b);
    float* o0 = last ? out0 : ((i & 1) ? b.h0b : b.h0a);
    float* o1 = last ? out1 : ((i & 1) ? b.h1b : b.h1a);
    lane_dense_pair(w, s0 + i, s1 != nullptr ? s1 + i : nullptr, in0, in1, o0, o1, i == 0 ? onehot : -1, act_kind,
                    lane, warp, nwarps);
    __syncthreads();
  }
}

// min_max_normalize (muax/nn.py:37-44): raw [n][32] -> dst [n][32]; every warp recomputes the range of its lane's
// tree, rows are dealt to the warps.  No barrier inside.
__device__ __forceinline__ void lane_min_max(const float* raw, float* dst, int n, bool enabled, int lane, int warp,
                                             int nwarps) {
  float lo = mz_inf(), hi = -mz_inf();
  if (enabled)
    for (int k = 0; k < n; ++k) {
      const float v = raw[k * kLT + lane];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
  float scale = MZ_SUB(hi, lo);
  if (scale < 1e-5f) scale = MZ_ADD(scale, 1e-5f);
  for (int k = warp; k < n; k += nwarps) {
    const float v = raw[k * kLT + lane];
    dst[k * kLT + lane] = enabled ? MZ_DIV(MZ_SUB(v, lo), scale) : v;
  }
}

// exps of a categorical head: e[j] = exp(logit[j] - max), j dealt over `parts` warps (this warp is part `part`).
__device__ __forceinline__ void lane_head_exps(const float* logits, float* e, int F, int part, int parts, int lane) {
  float mx = -mz_inf();
  for (int j = 0; j < F; ++j) mx = fmaxf(mx, logits[j * kLT + lane]);
  for (int j = part; j < F; j += parts) e[j * kLT + lane] = mz_expf(MZ_SUB(logits[j * kLT + lane], mx));
}

// support_to_scalar from the exps (muax/utils.py:94-102): both sums left to right.
__device__ __forceinline__ float lane_head_scalar(const float* e, int F, int S, int lane) {
  float s = 0.0f;
  for (int j = 0; j < F; ++j) s = MZ_ADD(s, e[j * kLT + lane]);
  float x = 0.0f;
  for (int j = 0; j < F; ++j) x = MZ_ADD(x, MZ_MUL((float)(j - S), MZ_DIV(e[j * kLT + lane], s)));
  return mz_inv_scaling(x);
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------------------- the kernel

__global__ void __launch_bounds__(256) lane_search_kernel(LaneArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  const LaneNet& net = a.net;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int N = a.N, A = net.A, E = net.E, F = net.F, NS = a.p.num_simulations;
  const int row0 = blockIdx.x * kLT;
  const float gamma = a.p.discount;

  // ---- shared memory carve-up
  float* w = smem;
  int off = round_up(net.packed_floats, 4);
  float* pbc = smem + off; off += round_up(NS + 2, 4);
  const int inW = max(max(E, net.obs_dim), 1);
  LaneBufs bufs;
  bufs.in = smem + off; off += inW * kLT;
  bufs.h0a = smem + off; off += net.Hmax * kLT;
  bufs.h0b = smem + off; off += net.Hmax * kLT;
  bufs.h1a = smem + off; off += net.Hmax * kLT;
  bufs.h1b = smem + off; off += net.Hmax * kLT;
  float* bufNs = smem + off; off += E * kLT;
  float* bufR = smem + off; off += F * kLT;
  float* bufV = smem + off; off += F * kLT;
  float* bufP = smem + off; off += A * kLT;
  float* bufEr = smem + off; off += F * kLT;
  float* bufEv = smem + off; off += F * kLT;
  float* nzbuf = smem + off; off += 2 * kLT * kGNoiseFloats;
  float* sc_reward = smem + off; off += kLT;
  float* sc_value = smem + off; off += kLT;
  int32_t* sc_action = reinterpret_cast<int32_t*>(smem + off); off += kLT;
  const LaneLayout L = lane_layout(N, A, E);
  float* blocks = smem + off;

  // ---- prologue: TMA bulk copy of the weights, pb_c table, tree init, observations
  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_expect_tx(&wbar, (uint32_t)(round_up(net.packed_floats, 4) * 4));
    tma_bulk_g2s(w, a.packed, (uint32_t)(round_up(net.packed_floats, 4) * 4), &wbar);
  }
  for (int n = tid; n < NS + 2; n += blockDim.x) pbc[n] = pbc_explore((float)n, a.p.pb_c_init, a.p.pb_c_base);
  for (int i = tid; i < kLT * L.stride; i += blockDim.x) {
    const int o = i % L.stride;
    const bool minus1 = (o >= L.parents && o < L.cindex + round_up(N * A, 4)) ;  // parents, afp, cindex
    reinterpret_cast<int32_t*>(blocks)[i] = minus1 ? -1 : 0;
  }
  for (int i = tid; i < kLT * net.obs_dim; i += blockDim.x) {
    const int tr = i / net.obs_dim, k = i - tr * net.obs_dim;
    const int b = min(row0 + tr, a.B - 1);
    bufs.in[k * kLT + tr] = a.obs[(size_t)b * net.obs_dim + k];
  }
  __syncthreads();  // thread 0 initialised the mbarrier: it must exist before any other thread polls it
  mbar_wait(&wbar, 0);
  __syncthreads();

  // ---- thread -> tree ownership for the tree phases
  const int tpw = kLT / nwarps;                 // trees per warp
  const bool owner = lane < tpw;
  const int ti = warp * tpw + lane;             // owned tree (valid when owner)
  const bool live = owner && row0 + ti < a.B;
  const int b = min(row0 + (owner ? ti : 0), a.B - 1);
  const LTree t = lane_tree(blocks + (size_t)(owner ? ti : 0) * L.stride, L);
  SearchParams p = a.p;
  p.batch_offset += b;

  // ---- root inference (muax/model.py:251-263)
  lane_module(w, net.repr, nullptr, net.n_repr, bufs, bufNs, nullptr, -1, net.activation, lane, warp, nwarps);
  lane_min_max(bufNs, bufs.in, E, net.repr_minmax != 0, lane, warp, nwarps);
  __syncthreads();
  lane_module(w, net.pred_v, net.pred_pi, net.n_pred, bufs, bufV, bufP, -1, net.activation, lane, warp, nwarps);
  if (warp < 2) lane_head_exps(bufV, bufEv, F, warp, min(2, nwarps), lane);
  __syncthreads();
  if (warp == 0) sc_value[lane] = lane_head_scalar(bufEv, F, net.S, lane);
  __syncthreads();
  if (owner) {
    const float rv = sc_value[ti];
    if (live && a.root_value_out != nullptr) a.root_value_out[b] = rv;  // raw network value (model.py:243)
    const size_t ba = (size_t)b * A;
    lt_begin(t, p, A, E, (long)p.batch_offset, bufP + ti, kLT, rv, bufs.in + ti,
             a.invalid != nullptr ? a.invalid + ba : nullptr, a.noise != nullptr ? a.noise + ba : nullptr);
  }
  // first noise rows
  const bool use_table = a.noise_table != nullptr && NS > 0;
  auto prefetch_noise = [&](int sim, int which) {
    // kLT rows of kGNoiseFloats floats -> 16-byte chunks dealt to the threads
    const int chunks = kLT * (kGNoiseFloats / 4);
    for (int c = tid; c < chunks; c += blockDim.x) {
      const int tr = c / (kGNoiseFloats / 4), q = c - tr * (kGNoiseFloats / 4);
      const int bb = min(row0 + tr, a.B - 1);
      cp_async16(nzbuf + ((size_t)which * kLT + tr) * kGNoiseFloats + q * 4,
                 a.noise_table + ((size_t)bb * NS + sim) * kGNoiseFloats + q * 4);
    }
  };
  if (use_table) {
    prefetch_noise(0, 0);
    cp_async_wait_all();
  }
  __syncthreads();

  // ---- simulations
  for (int sim = 0; sim < NS; ++sim) {
    int parent = 0, action = 0, next = 0, depth = 0;
    if (owner) {
      LaneAux aux;
      aux.pbc = pbc;
      aux.K = a.K;
      aux.noise_row = use_table ? nzbuf + ((size_t)(sim & 1) * kLT + ti) * kGNoiseFloats : nullptr;
      aux.cont = a.cont_keys != nullptr ? a.cont_keys + ((size_t)b * NS + sim) * 2 : nullptr;
      lt_simulate(t, p, sim, A, gamma, aux, parent, action, next, depth);
      sc_action[ti] = action;
      if (live) a.out.sim_depth[(size_t)b * NS + sim] = depth;
      for (int e = 0; e < E; ++e) bufs.in[e * kLT + ti] = t.emb[parent * E + e];
    }
    __syncthreads();
    // recurrent_fn (muax/model.py:265-282)
    lane_module(w, net.dyn_ns, net.dyn_r, net.n_dyn, bufs, bufNs, bufR, sc_action[lane], net.activation, lane, warp,
                nwarps);
    lane_min_max(bufNs, bufs.in, E, net.dyn_minmax != 0, lane, warp, nwarps);
    if (use_table && sim + 1 < NS) prefetch_noise(sim + 1, (sim + 1) & 1);
    __syncthreads();
    lane_module(w, net.pred_v, net.pred_pi, net.n_pred, bufs, bufV, bufP, -1, net.activation, lane, warp, nwarps);
    {
      const int half = max(nwarps / 2, 1);
      if (warp < half)
        lane_head_exps(bufR, bufEr, F, warp, half, lane);
      else
        lane_head_exps(bufV, bufEv, F, warp - half, nwarps - half, lane);
    }
    __syncthreads();
    if (warp == 0) sc_reward[lane] = lane_head_scalar(bufEr, F, net.S, lane);
    if (warp == nwarps / 2) sc_value[lane] = lane_head_scalar(bufEv, F, net.S, lane);
    cp_async_wait_all();
    __syncthreads();
    if (owner)
      lt_expand_backup(t, A, E, parent, action, next, sc_reward[ti], gamma, sc_value[ti], bufP + ti, bufs.in + ti,
                       kLT);
    // the next select of a tree runs on the thread that just backed it up: no barrier needed here
  }

  // ---- policy epilogue
  if (owner) {
    // surplus trees (beyond the batch) sink their weights into their own root_noise row (read before it is
    // overwritten inside lt_finish) instead of HBM
    float* wdst = live ? a.weights_out + (size_t)b * A : t.root_noise;
    const int action = lt_finish(t, p, A, gamma, (long)p.batch_offset, a.invalid != nullptr, wdst);
    if (live) a.action_out[b] = action;
  }
  __syncthreads();

  // ---- dump the trees to the global SoA arrays (mctx layout)
  if (a.dump_tree) {
    const Tree& o = a.out;
    const int live_trees = min(kLT, a.B - row0);
    for (int i = tid; i < live_trees * N; i += blockDim.x) {
      const int tr = i / N, n = i - tr * N;
      const LTree s = lane_tree(blocks + (size_t)tr * L.stride, L);
      const size_t g = (size_t)(row0 + tr) * o.N + n;
      o.node_visits[g] = s.node_visits[n];
      o.parents[g] = s.parents[n];
      o.action_from_parent[g] = s.afp[n];
      o.raw_values[g] = s.raw[n];
      o.node_values[g] = s.values[n];
    }
    for (int i = tid; i < live_trees * N * A; i += blockDim.x) {
      const int tr = i / (N * A), k = i - tr * (N * A);
      const LTree s = lane_tree(blocks + (size_t)tr * L.stride, L);
      const size_t g = (size_t)(row0 + tr) * o.N * A + k;
      const int ci = s.cindex[k];
      o.children_index[g] = ci;
      o.children_visits[g] = s.cvisits[k];
      o.children_prior_logits[g] = s.logits[k];
      o.children_prior_probs[g] = s.probs[k];
      o.children_values[g] = s.cvalues[k];
      o.children_rewards[g] = s.rewards[k];
      o.children_discounts[g] = ci >= 0 ? gamma : 0.0f;
    }
    for (int i = tid; i < live_trees * N * E; i += blockDim.x) {
      const int tr = i / (N * E), k = i - tr * (N * E);
      o.embeddings[(size_t)(row0 + tr) * o.N * E + k] = (blocks + (size_t)tr * L.stride + L.emb)[k];
    }
    for (int i = tid; i < live_trees * A; i += blockDim.x) {
      const int tr = i / A, x = i - tr * A;
      const LTree s = lane_tree(blocks + (size_t)tr * L.stride, L);
      o.root_noise[(size_t)(row0 + tr) * A + x] = s.root_noise[x];
      o.root_invalid[(size_t)(row0 + tr) * A + x] = s.root_invalid[x];
    }
  }
}

// ---------------------------------------------------------------------------------------- host side

struct LaneState {
  bool available = false;
  LaneNet net{};
  std::vector<LPackDesc> descs;
  float* packed = nullptr;
  float* noise_table = nullptr;
  uint32_t* cont_keys = nullptr;
  size_t noise_capacity = 0;
  int max_smem = 0, num_sms = 0, warps = 4;
  std::string why;
};

inline bool lane_plan_stack(const mz_stack& s, int in_x, int extra, LLayer* out, std::vector<LPackDesc>& descs, int& off,
                            int& hmax, std::string* why) {
  if (s.n_layers < 1 || s.n_layers > kLMaxLayers) {
    *why = "stack depth unsupported by the lane engine";
    return false;
  }
  for (int i = 0; i < s.n_layers; ++i) {
    LPackDesc d{};
    d.src = PackSrc{s.w_off[i], s.b_off[i], s.in_dim[i], s.out_dim[i]};
    d.in_x = i == 0 ? in_x : s.in_dim[i];
    d.l.K = d.in_x;
    d.l.extra = i == 0 ? extra : 0;
    d.l.out = s.out_dim[i];
    d.l.out4 = round_up(s.out_dim[i], 4);
    d.l.act = i + 1 < s.n_layers;
    d.l.off = off;
    off += (d.l.K + d.l.extra + 1) * d.l.out4;
    if (i + 1 < s.n_layers) hmax = std::max(hmax, (int)s.out_dim[i]);
    out[i] = d.l;
    descs.push_back(d);
  }
  return true;
}

inline size_t lane_smem_bytes(const LaneNet& g, int N, int NS) {
  const LaneLayout L = lane_layout(N, g.A, g.E);
  size_t f = (size_t)round_up(g.packed_floats, 4) + round_up(NS + 2, 4);
  f += (size_t)kLT * (std::max(std::max(g.E, g.obs_dim), 1) + 4 * g.Hmax + g.E + 4 * g.F + g.A);
  f += (size_t)2 * kLT * kGNoiseFloats + 3 * kLT;
  f += (size_t)kLT * L.stride;
  return f * 4;
}

inline int lane_init(LaneState& st, const Net& net, int device, std::string* err) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    *err = "cudaGetDeviceProperties failed";
    return 1;
  }
  st.max_smem = (int)prop.sharedMemPerBlockOptin;
  st.num_sms = prop.multiProcessorCount;
  st.available = false;
  LaneNet& g = st.net;
  g = LaneNet{};
  if (net.obs_dim <= 0) { st.why = "no Representation in the library (obs_dim = 0)"; return 0; }
  if (net.pred_v.n_layers != net.pred_pi.n_layers || net.dyn_ns.n_layers != net.dyn_r.n_layers) {
    st.why = "heads of a module differ in depth";
    return 0;
  }
  if (kGNoiseFloats / net.num_actions < 1) { st.why = "too many actions for the noise row"; return 0; }
  g.obs_dim = net.obs_dim; g.E = net.embed_dim; g.A = net.num_actions; g.S = net.support_size;
  g.F = 2 * net.support_size + 1;
  g.activation = net.activation; g.repr_minmax = net.repr_minmax; g.dyn_minmax = net.dyn_minmax;
  int off = 0, hmax = 1;
  st.descs.clear();
  if (!lane_plan_stack(net.repr, net.obs_dim, 0, g.repr, st.descs, off, hmax, &st.why) ||
      !lane_plan_stack(net.pred_v, net.embed_dim, 0, g.pred_v, st.descs, off, hmax, &st.why) ||
      !lane_plan_stack(net.pred_pi, net.embed_dim, 0, g.pred_pi, st.descs, off, hmax, &st.why) ||
      !lane_plan_stack(net.dyn_ns, net.embed_dim, net.num_actions, g.dyn_ns, st.descs, off, hmax, &st.why) ||
      !lane_plan_stack(net.dyn_r, net.embed_dim, net.num_actions, g.dyn_r, st.descs, off, hmax, &st.why))
    return 0;
  g.n_repr = net.repr.n_layers; g.n_pred = net.pred_v.n_layers; g.n_dyn = net.dyn_ns.n_layers;
  g.Hmax = hmax;
  g.packed_floats = off;
  if (cudaMalloc((void**)&st.packed, (size_t)round_up(off, 4) * 4 + 16) != cudaSuccess) {
    *err = "cudaMalloc(lane weights) failed";
    return 1;
  }
  const cudaError_t e = cudaFuncSetAttribute(lane_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             st.max_smem - 1024);
  if (e != cudaSuccess) {
    *err = std::string("lane engine: cudaFuncSetAttribute failed: ") + cudaGetErrorString(e);
    return 1;
  }
  if (const char* wv = getenv("MZ_LANE_WARPS")) st.warps = atoi(wv);
  if (st.warps != 2 && st.warps != 4 && st.warps != 8) st.warps = 4;
  st.available = true;
  return 0;
}

inline void lane_destroy(LaneState& st) {
  if (st.packed) cudaFree(st.packed);
  if (st.noise_table) cudaFree(st.noise_table);
  if (st.cont_keys) cudaFree(st.cont_keys);
  st.packed = nullptr;
  st.noise_table = nullptr;
  st.cont_keys = nullptr;
}

inline int lane_pack(LaneState& st, const float* raw, cudaStream_t stream, int64_t* launches) {
  if (!st.available) return 0;
  for (const LPackDesc& d : st.descs) {
    const int total = (d.l.K + d.l.extra + 1) * d.l.out4;
    lane_pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(raw, st.packed, d);
    *launches += 1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

inline bool lane_supported(const LaneState& st, const SearchParams& p) {
  return st.available &&
         lane_smem_bytes(st.net, p.num_simulations + 1, p.num_simulations) + 1024 <= (size_t)st.max_smem;
}

inline int lane_launch(LaneState& st, const Tree& out, const SearchParams& p, const float* obs, const uint8_t* invalid,
                       const float* noise, int32_t* action_out, float* weights_out, float* root_value_out,
                       cudaStream_t stream, int64_t* launches, std::string* err) {
  const int B = out.B, NS = p.num_simulations, N = NS + 1, A = st.net.A;
  LaneArgs a{};
  a.net = st.net;
  a.packed = st.packed;
  a.out = out;
  a.p = p;
  a.obs = obs;
  a.invalid = invalid;
  a.noise = noise;
  a.action_out = action_out;
  a.weights_out = weights_out;
  a.root_value_out = root_value_out;
  a.B = B;
  a.N = N;
  a.dump_tree = getenv("MZ_FUSED_NO_DUMP") ? 0 : 1;
  a.K = std::min(16, kGNoiseFloats / A);
  if (const char* k = getenv("MZ_GROUP_K")) a.K = std::max(0, std::min(a.K, atoi(k)));
  if (p.policy == MZ_POLICY_MUZERO && NS > 0 && a.K > 0) {
    const size_t pairs = (size_t)B * NS;
    if (pairs > st.noise_capacity) {
      if (st.noise_table) cudaFree(st.noise_table);
      if (st.cont_keys) cudaFree(st.cont_keys);
      st.noise_table = nullptr;
      st.cont_keys = nullptr;
      if (cudaMalloc((void**)&st.noise_table, pairs * kGNoiseFloats * 4) != cudaSuccess ||
          cudaMalloc((void**)&st.cont_keys, pairs * 8) != cudaSuccess) {
        *err = "lane engine: cudaMalloc(noise table) failed";
        return 1;
      }
      st.noise_capacity = pairs;
    }
    noise_table_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, stream>>>(p, B, A, a.K, st.noise_table, st.cont_keys);
    *launches += 1;
    a.noise_table = st.noise_table;
    a.cont_keys = st.cont_keys;
  }
  const size_t smem = lane_smem_bytes(st.net, N, NS);
  const int grid = (B + kLT - 1) / kLT;
  void* args[] = {&a};
  const cudaError_t e = cudaLaunchKernel((void*)lane_search_kernel, dim3(grid), dim3(32 * st.warps), args, smem, stream);
  *launches += 1;
  if (e != cudaSuccess) {
    *err = std::string("lane engine launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

}  // namespace mz

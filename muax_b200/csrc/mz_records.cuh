// mz_records.cuh — trees as packed 16-byte records in global memory (HBM / L2 / the owning SM's L1), and the
// per-tree device functions that walk them: policy prologue, `simulate`, `expand` + `backward`, policy epilogue
// (SURVEY.md Appendix A.2-A.6).  Shared by the CTA-resident engine (mz_resident.cu: CTA-phased dense layers, weights
// streamed when they do not fit on chip) and the tree-warp engine (mz_treewarp.cu: a warp owns its trees for the
// whole act, no CTA barrier in the simulation loop).
#pragma once
#include "mz_device.cuh"

namespace mz {

// ------------------------------------------------------------------------------------------ per-row warp functions

// muax/nn.py:37-44 on one row of width n in shared memory, by one warp.
__device__ __forceinline__ void min_max_row_warp(float* s, int n, int lane) {
  float lo = mz_inf(), hi = -mz_inf();
  for (int i = lane; i < n; i += 32) {
    lo = fminf(lo, s[i]);
    hi = fmaxf(hi, s[i]);
  }
  lo = gmin<32>(lo, 0xffffffffu);
  hi = gmax<32>(hi, 0xffffffffu);
  float scale = MZ_SUB(hi, lo);
  if (scale < 1e-5f) scale = MZ_ADD(scale, 1e-5f);
  for (int i = lane; i < n; i += 32) s[i] = MZ_DIV(MZ_SUB(s[i], lo), scale);
}

// support_to_scalar(softmax(logits)) (muax/model.py:260,273-274 + muax/utils.py:94-102) for one row by one warp:
// the exponentials, quotients and products are evaluated one per lane, the two float sums run left to right through
// shuffles — the same operations in the same order as support_to_scalar_row.  Every lane returns the result.
__device__ __forceinline__ float support_to_scalar_warp(const float* logits, int S, int lane) {
  const int F = 2 * S + 1;
  constexpr int kV = 4;  // values per lane: F <= 128
  if (F > 32 * kV) {
    float r = 0.0f;
    if (lane == 0) r = support_to_scalar_row(logits, S);
    return __shfl_sync(0xffffffffu, r, 0);
  }
  float l[kV], e[kV];
  float mx = -mz_inf();
#pragma unroll
  for (int v = 0; v < kV; ++v) {
    const int i = v * 32 + lane;
    l[v] = i < F ? logits[i] : -mz_inf();
    mx = fmaxf(mx, l[v]);
  }
  mx = gmax<32>(mx, 0xffffffffu);
#pragma unroll
  for (int v = 0; v < kV; ++v) e[v] = (v * 32 + lane) < F ? mz_expf(MZ_SUB(l[v], mx)) : 0.0f;
  float sum = 0.0f;
#pragma unroll
  for (int v = 0; v < kV; ++v) {
    const int cnt = min(32, F - v * 32);
    for (int i = 0; i < cnt; ++i) sum = MZ_ADD(sum, __shfl_sync(0xffffffffu, e[v], i));
  }
  float x = 0.0f;
#pragma unroll
  for (int v = 0; v < kV; ++v) {
    const int cnt = min(32, F - v * 32);
    const float term = (v * 32 + lane) < F ? MZ_MUL((float)(v * 32 + lane - S), MZ_DIV(e[v], sum)) : 0.0f;
    for (int i = 0; i < cnt; ++i) x = MZ_ADD(x, __shfl_sync(0xffffffffu, term, i));
  }
  return mz_inv_scaling(x);
}


// ------------------------------------------------------------------------------------------ tree records
// Working layout of the trees in HBM: 16-byte records, so that one level of `simulate` is one node load plus one
// 16-byte load per lane, one level of `backward` two loads and two stores, and every field of a node sits at a
// constant offset from one address.
//   node  n     : { visits (int), node_value, raw_value, parent << 8 | action  (0xFFFFFFFF: none) }
//   child (n, a): { child index << 16 | visits  (index 0xFFFF: unvisited), prior prob, value, reward }
//   logit (n, a): prior logit (own array: the MuZero selection never reads it, the Gumbel selectors do)
// children_discounts is not stored: on this path it is the constant gamma for every expanded edge (model.py:275) and
// an unexpanded edge has reward = value = 0, so reward + gamma * value is the same +0 as mctx's 0 + 0 * 0.
// The mctx SoA view of the C ABI (mz_get_tree) is produced on demand by resident_unpack_kernel.
//
// Cache policy: the records are the only data with reuse (every simulation re-walks the top of its tree); embeddings
// and the tie-break noise are touched once per simulation and stream (ld.cs / st.cs) so that they do not push the
// records out of the 126 MB L2 — measured on B200 at the C3 shapes: 24.5 -> 22.3 ms per act (an additional L2
// evict_last hint on the record accesses, MZ_RES_REC_HINT, changes nothing on top of that and stays off).
//
// Warp discipline: every lane of a warp runs the same loops (a group without a live tree, or whose walk has ended,
// is predicated off), so all shuffles use the full mask with width G — no per-group mask convergence checks.

constexpr uint32_t kRecNoChild = 0xFFFFu;
constexpr uint32_t kRecNoParent = 0xFFFFFFFFu;
constexpr unsigned kFull = 0xffffffffu;

// MZ_RES_WARP_UNIFORM = 1: all lanes of a warp run the walk loops together and shuffle with the full mask;
// 0: every lane group runs its own loops and shuffles with its group mask (measured faster on B200: a finished
// group leaves the loop instead of idling through the deepest walk of its warp).
#ifndef MZ_RES_WARP_UNIFORM
#define MZ_RES_WARP_UNIFORM 0
#endif
template <int G>
__device__ __forceinline__ unsigned walk_mask() {
#if MZ_RES_WARP_UNIFORM
  return kFull;
#else
  return group_mask<G>();
#endif
}
__device__ __forceinline__ bool walk_continues(bool active) {
#if MZ_RES_WARP_UNIFORM
  return __any_sync(kFull, active);
#else
  return active;
#endif
}

__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#ifndef MZ_RES_PREFETCH_CHILDREN
#define MZ_RES_PREFETCH_CHILDREN 0  // measured: C3 25.5 vs 22.3 ms with/without, C5 7.60 vs 7.75: off
#endif
#ifndef MZ_RES_REC_HINT
#define MZ_RES_REC_HINT 0
#endif
#ifndef MZ_RES_STREAM_NOISE
#define MZ_RES_STREAM_NOISE 1
#endif
#ifndef MZ_RES_STREAM_EMB
#define MZ_RES_STREAM_EMB 1
#endif
#if MZ_RES_STREAM_NOISE
#define MZ_LD_NOISE(p) __ldcs(p)
#else
#define MZ_LD_NOISE(p) (*(p))
#endif
#if MZ_RES_STREAM_EMB
#define MZ_LD_EMB(p) __ldcs(p)
#define MZ_ST_EMB(p, v) __stcs(p, v)
#else
#define MZ_LD_EMB(p) (*(p))
#define MZ_ST_EMB(p, v) (*(p) = (v))
#endif
#if !MZ_RES_REC_HINT
__device__ __forceinline__ float4 rec_ld(const float4* p, uint64_t) { return *p; }
__device__ __forceinline__ void rec_st(float4* p, const float4& v, uint64_t) { *p = v; }
#else
__device__ __forceinline__ float4 rec_ld(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol)
               : "memory");
  return v;
}
__device__ __forceinline__ void rec_st(float4* p, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
#endif

struct RecTrees {    // the records of the trees one CTA owns (local tree index 0..R-1)
  float4* nodes;     // [R][N]
  float4* childs;    // [R][N][A]
  float* logits;     // [R][N][A]
  float* emb;        // [R][embN][E]  (the handle's SoA embeddings)
  float* root_noise; // [R][A]
  uint8_t* root_invalid;
  int32_t* sim_depth; // [R][NS]
  uint64_t pol;       // L2 evict_last access policy for the records
  int32_t N, A, E;    // N = record stride (nodes of this search)
  int32_t embN;       // node stride of the embeddings (the handle's capacity)
};

__device__ __forceinline__ ChildRow rec_child_row(const float4& h0, float logit, float gamma, bool ok) {
  ChildRow c;
  c.visits = ok ? (int)(__float_as_uint(h0.x) & 0xFFFFu) : 0;
  c.logit = ok ? logit : 0.0f;
  c.prob = ok ? h0.y : 0.0f;
  c.value = ok ? h0.z : 0.0f;
  c.reward = ok ? h0.w : 0.0f;
  c.discount = ok ? gamma : 0.0f;
  return c;
}

// Policy prologue (A.2 / A.4) + instantiate_tree_from_root (A.3) for one tree.  `has`: this group owns tree b
// (groups without a tree run the arithmetic on a clamped row and store nothing).
template <int G>
__device__ __forceinline__ void rec_begin(const RecTrees& t, const SearchParams& p, int b, bool has, long gb,
                                          const float* root_logits, float root_value, const float* root_emb,
                                          const uint8_t* invalid, const float* noise, int a) {
  const int A = t.A;
  float logit, prob, nz;
  bool inv;
  group_begin_compute<G>(p, A, gb, root_logits, invalid, noise, a, walk_mask<G>(), logit, prob, nz, inv);
  if (!has) return;
  if (a < A) {
    rec_st(t.childs + (size_t)b * t.N * A + a, make_float4(__uint_as_float(kRecNoChild << 16), prob, 0.0f, 0.0f), t.pol);
    t.logits[(size_t)b * t.N * A + a] = logit;
    t.root_noise[b * A + a] = nz;
    t.root_invalid[b * A + a] = inv ? 1 : 0;
  }
  float* emb = t.emb + (size_t)b * t.embN * t.E;
  for (int e = a; e < t.E; e += G) MZ_ST_EMB(emb + e, root_emb[e]);
  if (a == 0)
    rec_st(t.nodes + (size_t)b * t.N, make_float4(__int_as_float(1), root_value, root_value, __uint_as_float(kRecNoParent)),
           t.pol);
}

// `simulate` (A.3) for one tree per lane group; all lanes of the warp stay in the level loop until every group of
// the warp has reached its leaf.  `fresh`: the selected edge was unvisited (the new node gets index sim + 1).
template <int G>
__device__ __forceinline__ void rec_simulate(const RecTrees& t, const SearchParams& p, int b, bool has, int sim, int a,
                                             int& parent, int& action_out, int& next, int& depth_out, bool& fresh,
                                             const SelectAux& aux, uint32_t* path) {
  const int A = t.A;
  const bool ok = a < A;
  const bool muzero = p.policy == MZ_POLICY_MUZERO;
  const bool table = aux.noise_row != nullptr;
  uint32_t k0 = 0, k1 = 0;
  if (muzero && !table)
    split_key(p.sim_keys[2 * sim], p.sim_keys[2 * sim + 1], (uint32_t)p.global_batch, (uint32_t)(p.batch_offset + b),
              p.prng_mode, k0, k1);
  const int max_depth = p.max_depth > 0 ? p.max_depth : p.num_simulations;
  const float4* nodes = t.nodes + (size_t)b * t.N;
  const float4* ch = t.childs + (size_t)b * t.N * A;
  const float* lg = t.logits + (size_t)b * t.N * A;
  const bool root_inv = has && ok && t.root_invalid[b * A + a] != 0;
  const float root_gumbel = (has && !muzero && ok) ? t.root_noise[b * A + a] : 0.0f;
  int node = 0;
  bool active = has;
  parent = 0; action_out = 0; next = 0; depth_out = 0; fresh = false;
  // `level` is warp-uniform: every group still walking is at the same depth, so the branches on it (root vs interior
  // selectors, table vs inline noise) never split a warp around a shuffle
  const unsigned wm = walk_mask<G>();
  for (int level = 0; walk_continues(active); ++level) {
    float4 nd = make_float4(0.0f, 0.0f, 0.0f, 0.0f), h0 = nd;
    float logit = 0.0f;
    if (active) {
      nd = rec_ld(nodes + node, t.pol);
      if (ok) {
        h0 = rec_ld(ch + node * A + a, t.pol);
        if (!muzero) logit = lg[node * A + a];
#if MZ_RES_PREFETCH_CHILDREN
        // the walk is a pointer chase with one HBM/L2 round trip per level: every lane pulls the records of ITS child
        // towards L1 while the scores are computed, so the level that follows the argmax finds them on the way
        const uint32_t cia = __float_as_uint(h0.x) >> 16;
        if (cia != kRecNoChild) {
          prefetch_l1(nodes + cia);
          prefetch_l1(ch + cia * A);
          if (A > 8) prefetch_l1(ch + cia * A + 8);
          if (A > 16) prefetch_l1(ch + cia * A + 16);
        }
#endif
      }
    }
    uint32_t s0 = 0, s1 = 0;
    bool have_noise = false;
    float nz = 0.0f;
    if (muzero) {
      if (table && level < aux.K) {
        have_noise = true;
        if (active) nz = MZ_LD_NOISE(aux.noise_row + level * A + (ok ? a : 0));
      } else {  // past the table (or no table): continue the jax key chain inline
        if (table && level == aux.K) {
          k0 = aux.cont0;
          k1 = aux.cont1;
        }
        group_split2<G>(k0, k1, p.prng_mode, a, wm, k0, k1, s0, s1);
      }
    }
    const ChildRow c = rec_child_row(h0, logit, p.discount, ok && active);
    const int action = group_select_score<G>(p, A, c, ok, nd.y, nd.z, __float_as_int(nd.x), level, root_inv, root_gumbel,
                                             s0, s1, a, wm, have_noise, nz, aux.pbc);
    const uint32_t ci = __shfl_sync(wm, __float_as_uint(h0.x) >> 16, action, G);
    if (active) {
      if (a == 0) path[level] = ((uint32_t)node << 8) | (uint32_t)action;
      if (ci == kRecNoChild || level + 1 >= max_depth) {
        active = false;
        parent = node;
        action_out = action;
        depth_out = level + 1;
        fresh = ci == kRecNoChild;
        next = fresh ? sim + 1 : (int)ci;
      } else {
        node = (int)ci;
      }
    }
  }
}

// `expand` scatter (A.3) + `backward` for one tree, walking the path recorded by the selection: the records of
// level d-1 are loaded while level d's mean update is computed.
template <int G>
__device__ __forceinline__ void rec_expand_backup(const RecTrees& t, int b, bool has, int parent, int action, int next,
                                                  bool fresh, float reward, float gamma, float value, float logit_a,
                                                  const float* next_emb, int a, const uint32_t* path, int depth) {
  const int A = t.A;
  const bool ok = a < A;
  float4* nodes = t.nodes + (size_t)b * t.N;
  float4* ch = t.childs + (size_t)b * t.N * A;
  const float prob = group_softmax<G>(logit_a, ok, A, walk_mask<G>());
  if (!has) return;
  if (ok) {
    float4 h0 = make_float4(__uint_as_float(kRecNoChild << 16), prob, 0.0f, 0.0f);
    if (!fresh) {  // max_depth re-expansion: priors are overwritten, the edge statistics stay (update_tree_node)
      h0 = rec_ld(ch + next * A + a, t.pol);
      h0.y = prob;
    }
    rec_st(ch + next * A + a, h0, t.pol);
    t.logits[((size_t)b * t.N + next) * A + a] = logit_a;
  }
  if (next_emb != nullptr) {  // null: the caller has already stored the new embedding in place
    float* emb = t.emb + ((size_t)b * t.embN + next) * t.E;
    for (int e = a; e < t.E; e += G) MZ_ST_EMB(emb + e, next_emb[e]);
  }
  if (a == 0) {
    const int old_visits = fresh ? 0 : __float_as_int(rec_ld(nodes + next, t.pol).x);
    rec_st(nodes + next,
           make_float4(__int_as_float(old_visits + 1), value, value, __uint_as_float(((uint32_t)parent << 8) | (uint32_t)action)),
           t.pol);
    // backward: path[d] = (node << 8 | action) of the edge selected at depth d; path[depth - 1] = (parent, action)
    float G_ = value, child_value = value;
    int d = depth - 1;
    int pn = parent, e2 = parent * A + action;
    float4 nd = rec_ld(nodes + pn, t.pol);
    float4 c = rec_ld(ch + e2, t.pol);
    c.x = __uint_as_float(((uint32_t)next << 16) | (__float_as_uint(c.x) & 0xFFFFu));  // children_index[parent, action]
    c.w = reward;                                                                      // children_rewards[parent, action]
    for (;;) {
      int n_pn = 0, n_e2 = 0;
      float4 n_nd = nd, n_c = c;
      if (d > 0) {
        const uint32_t pa = path[d - 1];
        n_pn = (int)(pa >> 8);
        n_e2 = n_pn * A + (int)(pa & 0xffu);
        n_nd = rec_ld(nodes + n_pn, t.pol);
        n_c = rec_ld(ch + n_e2, t.pol);
      }
      const int count_i = __float_as_int(nd.x);
      const float count = (float)count_i;
      G_ = MZ_ADD(c.w, MZ_MUL(gamma, G_));
      const float pv = MZ_DIV(MZ_ADD(MZ_MUL(nd.y, count), G_), MZ_ADD(count, 1.0f));
      rec_st(nodes + pn, make_float4(__int_as_float(count_i + 1), pv, nd.z, nd.w), t.pol);
      c.x = __uint_as_float(__float_as_uint(c.x) + 1u);  // children_visits += 1 (low 16 bits)
      c.z = child_value;
      rec_st(ch + e2, c, t.pol);
      child_value = pv;
      if (d == 0) break;
      --d;
      pn = n_pn; e2 = n_e2; nd = n_nd; c = n_c;
    }
  }
}

// Policy epilogue for one tree (A.2 / A.4).
template <int G>
__device__ __forceinline__ void rec_finish(const RecTrees& t, const SearchParams& p, int b, bool has, long gb,
                                           bool has_invalid, int a, int& action, float& weight) {
  const int A = t.A;
  const bool ok = a < A;
  const bool muzero = p.policy == MZ_POLICY_MUZERO;
  float4 nd = make_float4(0.0f, 0.0f, 0.0f, 0.0f), h0 = nd;
  float logit = 0.0f;
  bool root_inv = false;
  float root_gumbel = 0.0f;
  if (has) {
    nd = rec_ld(t.nodes + (size_t)b * t.N, t.pol);
    if (ok) {
      h0 = rec_ld(t.childs + (size_t)b * t.N * A + a, t.pol);
      if (!muzero) {
        logit = t.logits[(size_t)b * t.N * A + a];
        root_inv = t.root_invalid[b * A + a] != 0;
        root_gumbel = t.root_noise[b * A + a];
      }
    }
  }
  const ChildRow c = rec_child_row(h0, logit, p.discount, ok && has);
  group_finish_score<G>(p, A, c, ok, nd.y, nd.z, root_inv, root_gumbel, gb, has_invalid, a, walk_mask<G>(), action, weight);
}

}  // namespace mz

// mz_fused.cuh — fused persistent engine: ONE launch per act.
//
// A CTA owns T trees for the whole search.  Their SoA tree (5.1 KB per tree at the CartPole shapes), the
// complete weight blob (TMA bulk copy, 6.3 KB) and the MLP staging all live in shared memory, so the
// num_simulations x (select -> recurrent -> backup) loop never touches HBM and never leaves the SM:
// trees are independent, so the only synchronisation is the CTA barrier around the batched MLP.
// HBM traffic per act = observations in, (action, action_weights, root_value) out, plus the optional tree dump.
//
// Thread roles per simulation:
//   select / expand+backup : lane group per tree (G lanes, T*G <= blockDim) — mz_device.cuh group_* functions
//   recurrent_fn           : all threads, (row, output neuron) work items, heads evaluated pairwise
#pragma once
#include <algorithm>
#include <cstdlib>
#include <string>

#include "mz_device.cuh"

namespace mz {

struct FusedArgs {
  Net net;
  const float* weights;  // global fp32 blob
  int32_t weight_bytes;  // multiple of 16
  Tree out;              // global SoA tree (mctx layout) the smem trees are dumped to
  SearchParams p;
  const float* obs;
  const uint8_t* invalid;
  const float* noise;
  int32_t* action_out;
  float* weights_out;
  float* root_value_out;
  int32_t T;      // trees per CTA
  int32_t B;      // rows of this handle
  int32_t N;      // nodes per tree = num_simulations + 1 (<= out.N)
  int32_t ld;     // MLP staging row stride (floats)
  int32_t dump_tree;
};

struct FusedLayout {  // offsets in floats from the dynamic smem base
  int weights, node_i, child_i, node_f, child_f, emb, noise, invalid, mlp, sel, total_floats;
};


__host__ __device__ inline FusedLayout fused_layout(int weight_bytes, int T, int N, int A, int E, int ld) {
  FusedLayout L;
  int off = 0;
  L.weights = off; off += round_up(weight_bytes / 4, 4);
  L.node_i = off;  off += 3 * T * N;          // node_visits, parents, action_from_parent
  L.child_i = off; off += 2 * T * N * A;      // children_index, children_visits
  L.node_f = off;  off += 2 * T * N;          // raw_values, node_values
  L.child_f = off; off += 5 * T * N * A;      // logits, probs, values, rewards, discounts
  L.emb = off;     off += T * N * E;
  L.noise = off;   off += T * A;
  L.invalid = off; off += round_up(T * A, 4) / 4 + 1;
  L.mlp = off;     off += 8 * T * ld;         // x, ns, headA, headB, tmp0A, tmp1A, tmp0B, tmp1B
  L.sel = off;     off += 5 * T + 4;          // parent, action, next, reward, value
  L.total_floats = round_up(off, 4);
  return L;
}

// Two hk.Sequential heads that read the same input, evaluated in lockstep: one barrier per layer instead of two.
// Falls back to one-after-the-other when their depths differ.
template <bool kLdg>
__device__ __forceinline__ void dual_stack_forward_cta(const mz_stack& sa, const mz_stack& sb, const float* w, int act,
                                                       const float* x, int ldx, int in_x, const int* onehot,
                                                       float* outa, float* outb, int ldo, float* ta0, float* ta1,
                                                       float* tb0, float* tb1, int ldt, int R) {
  if (sa.n_layers != sb.n_layers) {
    stack_forward_cta<kLdg>(sa, w, act, x, ldx, in_x, onehot, outa, ldo, ta0, ta1, ldt, R);
    stack_forward_cta<kLdg>(sb, w, act, x, ldx, in_x, onehot, outb, ldo, tb0, tb1, ldt, R);
    return;
  }
  const float *srca = x, *srcb = x;
  int lds = ldx;
  for (int l = 0; l < sa.n_layers; ++l) {
    const bool last = l == sa.n_layers - 1;
    float* dsta = last ? outa : ((l & 1) ? ta1 : ta0);
    float* dstb = last ? outb : ((l & 1) ? tb1 : tb0);
    const int ldd = last ? ldo : ldt;
    const int nina = l == 0 ? in_x : sa.in_dim[l];
    const int ninb = l == 0 ? in_x : sb.in_dim[l];
    const int na = sa.out_dim[l], nb = sb.out_dim[l];
    const int ntot = na + nb;
    for (int idx = threadIdx.x; idx < R * ntot; idx += blockDim.x) {
      const int r = idx / ntot;
      int j = idx - r * ntot;
      const bool second = j >= na;
      if (second) j -= na;
      const mz_stack& s = second ? sb : sa;
      const int nin = second ? ninb : nina;
      const int nout = second ? nb : na;
      const float* W = w + s.w_off[l];
      const float* xr = (second ? srcb : srca) + r * lds;
      float acc = 0.0f;
      for (int k = 0; k < nin; ++k) acc = MZ_FMA(xr[k], ldw<kLdg>(W + (long)k * nout + j), acc);
      if (l == 0 && onehot != nullptr) acc = MZ_ADD(acc, ldw<kLdg>(W + (long)(nin + onehot[r]) * nout + j));
      float y = MZ_ADD(acc, ldw<kLdg>(w + s.b_off[l] + j));
      if (!last) y = activate(y, act);
      (second ? dstb : dsta)[r * ldd + j] = y;
    }
    __syncthreads();
    srca = dsta;
    srcb = dstb;
    lds = ldd;
  }
}

template <int G>
__global__ void __launch_bounds__(256) fused_search_kernel(FusedArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  const int T = a.T, N = a.N, A = a.net.num_actions, E = a.net.embed_dim, ld = a.ld;
  const int row0 = blockIdx.x * T;
  const int R = min(T, a.B - row0);
  const int tid = threadIdx.x;
  const FusedLayout L = fused_layout(a.weight_bytes, T, N, A, E, ld);

  // ---- stage the weights with one TMA bulk copy while the trees are initialised
  float* w = smem + L.weights;
  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_expect_tx(&wbar, (uint32_t)a.weight_bytes);
    tma_bulk_g2s(w, a.weights, (uint32_t)a.weight_bytes, &wbar);
  }

  Tree t;
  t.B = T; t.N = N; t.A = A; t.E = E;
  int32_t* ip = reinterpret_cast<int32_t*>(smem + L.node_i);
  t.node_visits = ip;
  t.parents = ip + T * N;
  t.action_from_parent = ip + 2 * T * N;
  int32_t* cp = reinterpret_cast<int32_t*>(smem + L.child_i);
  t.children_index = cp;
  t.children_visits = cp + T * N * A;
  t.raw_values = smem + L.node_f;
  t.node_values = smem + L.node_f + T * N;
  float* cf = smem + L.child_f;
  t.children_prior_logits = cf;
  t.children_prior_probs = cf + T * N * A;
  t.children_values = cf + 2 * T * N * A;
  t.children_rewards = cf + 3 * T * N * A;
  t.children_discounts = cf + 4 * T * N * A;
  t.embeddings = smem + L.emb;
  t.root_noise = smem + L.noise;
  t.root_invalid = reinterpret_cast<uint8_t*>(smem + L.invalid);
  t.sim_depth = a.out.sim_depth + (long)row0 * a.p.num_simulations;

  // mctx initial state (Appendix A.1): zeros, parents / action_from_parent / children_index = -1
  for (int i = tid; i < 3 * T * N; i += blockDim.x) ip[i] = i < T * N ? 0 : -1;
  for (int i = tid; i < 2 * T * N * A; i += blockDim.x) cp[i] = i < T * N * A ? -1 : 0;
  for (int i = tid; i < 2 * T * N; i += blockDim.x) smem[L.node_f + i] = 0.0f;
  for (int i = tid; i < 5 * T * N * A; i += blockDim.x) cf[i] = 0.0f;
  if (a.p.max_depth > 0)
    for (int i = tid; i < T * N * E; i += blockDim.x) t.embeddings[i] = 0.0f;

  float* x = smem + L.mlp;
  float* ns = x + T * ld;
  float* headA = ns + T * ld;
  float* headB = headA + T * ld;
  float* ta0 = headB + T * ld;
  float* ta1 = ta0 + T * ld;
  float* tb0 = ta1 + T * ld;
  float* tb1 = tb0 + T * ld;
  int32_t* sel_parent = reinterpret_cast<int32_t*>(smem + L.sel);
  int32_t* sel_action = sel_parent + T;
  int32_t* sel_next = sel_action + T;
  float* rec_reward = reinterpret_cast<float*>(sel_next + T);
  float* rec_value = rec_reward + T;

  SearchParams p = a.p;
  p.batch_offset += row0;  // PRNG draws are indexed by global row

  const int obs_dim = a.net.obs_dim;
  for (int i = tid; i < R * obs_dim; i += blockDim.x) {
    const int r = i / obs_dim, k = i - r * obs_dim;
    x[r * ld + k] = a.obs[(long)(row0 + r) * obs_dim + k];
  }
  __syncthreads();  // thread 0 initialised the mbarrier: it must exist before any other thread polls it
  mbar_wait(&wbar, 0);
  __syncthreads();

  // ---- root inference (muax/model.py:251-263); the root embedding lands in `ns`
  stack_forward_cta<false>(a.net.repr, w, a.net.activation, x, ld, obs_dim, nullptr, ns, ld, ta0, ta1, ld, R);
  if (a.net.repr_minmax) min_max_normalize_cta(ns, ld, E, R);
  dual_stack_forward_cta<false>(a.net.pred_v, a.net.pred_pi, w, a.net.activation, ns, ld, E, nullptr, headA, headB, ld,
                                ta0, ta1, tb0, tb1, ld, R);
  if (tid < R) {
    const float v = support_to_scalar_row(headA + tid * ld, a.net.support_size);
    rec_value[tid] = v;
    if (a.root_value_out != nullptr) a.root_value_out[row0 + tid] = v;  // raw network value (model.py:243)
  }
  __syncthreads();

  // ---- policy prologue + tree instantiation, then the simulations
  const int gi = tid / G;            // tree handled by this lane group
  const int ga = tid & (G - 1);      // action handled by this lane
  const bool walker = gi < R;
  const unsigned gm = group_mask<G>();
  if (walker) {
    const long ba = (long)(row0 + gi) * A;
    group_begin<G>(t, p, gi, (long)p.batch_offset + gi, headB + gi * ld, rec_value[gi], ns + gi * ld,
                   a.invalid != nullptr ? a.invalid + ba : nullptr, a.noise != nullptr ? a.noise + ba : nullptr, ga, gm);
  }
  __syncthreads();

  const int NS = p.num_simulations;
  for (int sim = 0; sim < NS; ++sim) {
    if (walker) {
      int parent, action, next, depth;
      group_simulate<G>(t, p, gi, sim, ga, gm, parent, action, next, depth);
      if (ga == 0) {
        sel_parent[gi] = parent;
        sel_action[gi] = action;
        sel_next[gi] = next;
        t.sim_depth[(long)gi * NS + sim] = depth;
      }
      for (int e = ga; e < E; e += G) x[gi * ld + e] = t.embeddings[((long)gi * N + parent) * E + e];
    }
    __syncthreads();
    // recurrent_fn (muax/model.py:265-282): Dynamic (both heads) -> min-max -> Prediction (both heads)
    dual_stack_forward_cta<false>(a.net.dyn_ns, a.net.dyn_r, w, a.net.activation, x, ld, E, sel_action, ns, headA, ld,
                                  ta0, ta1, tb0, tb1, ld, R);
    if (tid < R) rec_reward[tid] = support_to_scalar_row(headA + tid * ld, a.net.support_size);
    if (a.net.dyn_minmax)
      min_max_normalize_cta(ns, ld, E, R);
    else
      __syncthreads();  // headA (reward logits) is about to be reused by the value head
    dual_stack_forward_cta<false>(a.net.pred_v, a.net.pred_pi, w, a.net.activation, ns, ld, E, nullptr, headA, headB,
                                  ld, ta0, ta1, tb0, tb1, ld, R);
    if (tid < R) rec_value[tid] = support_to_scalar_row(headA + tid * ld, a.net.support_size);
    __syncthreads();
    if (walker) {
      const float logit = ga < A ? headB[gi * ld + ga] : 0.0f;
      group_expand_backup<G>(t, gi, sel_parent[gi], sel_action[gi], sel_next[gi], rec_reward[gi], p.discount,
                             rec_value[gi], logit, ns + gi * ld, ga, gm);
    }
    // no CTA barrier needed here: the next select of a tree runs on the same lanes that just backed it up,
    // and the MLP staging is only rewritten after the barrier that follows the next select
    __syncwarp();
  }

  // ---- policy epilogue
  if (walker) {
    int action;
    float weight;
    group_finish<G>(t, p, gi, (long)p.batch_offset + gi, a.invalid != nullptr, ga, gm, action, weight);
    if (ga < A) a.weights_out[(long)(row0 + gi) * A + ga] = weight;
    if (ga == 0) a.action_out[row0 + gi] = action;
  }
  __syncthreads();

  // ---- dump the trees to the global SoA arrays (mctx layout; the tree view of the C ABI)
  if (a.dump_tree) {
    const Tree& o = a.out;
    const int ON = o.N;
    for (int i = tid; i < R * N; i += blockDim.x) {
      const int r = i / N, n = i - r * N;
      const long g = (long)(row0 + r) * ON + n;
      o.node_visits[g] = t.node_visits[i];
      o.parents[g] = t.parents[i];
      o.action_from_parent[g] = t.action_from_parent[i];
      o.raw_values[g] = t.raw_values[i];
      o.node_values[g] = t.node_values[i];
    }
    for (int i = tid; i < R * N * A; i += blockDim.x) {
      const int r = i / (N * A), k = i - r * (N * A);
      const long g = (long)(row0 + r) * ON * A + k;
      o.children_index[g] = t.children_index[i];
      o.children_visits[g] = t.children_visits[i];
      o.children_prior_logits[g] = t.children_prior_logits[i];
      o.children_prior_probs[g] = t.children_prior_probs[i];
      o.children_values[g] = t.children_values[i];
      o.children_rewards[g] = t.children_rewards[i];
      o.children_discounts[g] = t.children_discounts[i];
    }
    for (int i = tid; i < R * N * E; i += blockDim.x) {
      const int r = i / (N * E), k = i - r * (N * E);
      o.embeddings[(long)(row0 + r) * ON * E + k] = t.embeddings[i];
    }
    for (int i = tid; i < R * A; i += blockDim.x) {
      o.root_noise[(long)row0 * A + i] = t.root_noise[i];
      o.root_invalid[(long)row0 * A + i] = t.root_invalid[i];
    }
  }
}

// ------------------------------------------------------------------------------------------ host side

struct FusedState {
  bool available = false;
  int max_smem = 0;
  int num_sms = 0;
  int G = 0;
  int threads = 128;
  int trees_per_cta = 0;  // 0 = choose per launch
};

inline void* fused_kernel_ptr(int G) {
  switch (G) {
    case 2: return (void*)fused_search_kernel<2>;
    case 4: return (void*)fused_search_kernel<4>;
    case 8: return (void*)fused_search_kernel<8>;
    case 16: return (void*)fused_search_kernel<16>;
    default: return (void*)fused_search_kernel<32>;
  }
}

inline int fused_ld(const Net& net) { return round_up(net.max_width, 4); }

inline size_t fused_smem_bytes(const Net& net, int weight_bytes, int T, int N) {
  return (size_t)fused_layout(weight_bytes, T, N, net.num_actions, net.embed_dim, fused_ld(net)).total_floats * 4;
}

inline int fused_init(FusedState& st, const Net& net, int /*batch*/, int /*max_sims*/, int device, std::string* err) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    *err = "cudaGetDeviceProperties failed";
    return 1;
  }
  st.max_smem = (int)prop.sharedMemPerBlockOptin;
  st.num_sms = prop.multiProcessorCount;
  int G = 2;
  while (G < net.num_actions) G <<= 1;
  st.G = G;
  // static shared memory (the weight mbarrier) counts against the opt-in limit
  const cudaError_t e = cudaFuncSetAttribute(fused_kernel_ptr(G), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             st.max_smem - 1024);
  if (e != cudaSuccess) {
    *err = std::string("fused engine: cudaFuncSetAttribute failed: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return 1;
  }
  if (const char* e = getenv("MZ_FUSED_THREADS")) st.threads = atoi(e);
  if (const char* e = getenv("MZ_FUSED_TREES")) st.trees_per_cta = atoi(e);
  st.available = net.obs_dim > 0;
  return 0;
}

inline void fused_destroy(FusedState&) {}

// Trees per CTA: as few as keeps every SM covered (smaller CTAs couple fewer trees at the MLP barrier and let
// several CTAs per SM overlap their phases), bounded by the lane budget and by shared memory.
inline int fused_pick_trees(const FusedState& st, const Net& net, int weight_bytes, int B, int N) {
  const int max_by_threads = st.threads / st.G;
  int best = 0;
  for (int T = 1; T <= max_by_threads; ++T) {
    const size_t bytes = fused_smem_bytes(net, weight_bytes, T, N) + 1024;
    if (bytes > (size_t)st.max_smem) break;
    best = T;
  }
  if (best == 0) return 0;
  if (st.trees_per_cta > 0) return st.trees_per_cta <= best ? st.trees_per_cta : 0;
  // target ~4 resident CTAs per SM when the batch is large enough, never below 4 trees per CTA
  int T = (B + st.num_sms * 4 - 1) / (st.num_sms * 4);
  if (T < 4) T = 4;
  if (T > best) T = best;
  return T;
}

inline bool fused_supported(const FusedState& st, const Net& net, const SearchParams& p) {
  if (!st.available) return false;
  const int F = 2 * net.support_size + 1;
  int wbytes = 0;
  const mz_stack* stacks[5] = {&net.repr, &net.pred_v, &net.pred_pi, &net.dyn_ns, &net.dyn_r};
  for (const mz_stack* s : stacks)
    for (int l = 0; l < s->n_layers; ++l) {
      wbytes = std::max(wbytes, (int)(s->w_off[l] + (int64_t)s->in_dim[l] * s->out_dim[l]) * 4);
      wbytes = std::max(wbytes, (int)(s->b_off[l] + s->out_dim[l]) * 4);
    }
  (void)F;
  return fused_pick_trees(st, net, round_up(wbytes, 16), 1 << 30, p.num_simulations + 1) > 0;
}

inline int fused_launch(FusedState& st, const Net& net, const float* weights, const Tree& out, const SearchParams& p,
                        const float* obs, const uint8_t* invalid, const float* noise, int32_t* action_out,
                        float* weights_out, float* root_value_out, cudaStream_t stream, std::string* err) {
  int wbytes = 0;
  const mz_stack* stacks[5] = {&net.repr, &net.pred_v, &net.pred_pi, &net.dyn_ns, &net.dyn_r};
  for (const mz_stack* s : stacks)
    for (int l = 0; l < s->n_layers; ++l) {
      wbytes = std::max(wbytes, (int)(s->w_off[l] + (int64_t)s->in_dim[l] * s->out_dim[l]) * 4);
      wbytes = std::max(wbytes, (int)(s->b_off[l] + s->out_dim[l]) * 4);
    }
  wbytes = round_up(wbytes, 16);
  const int N = p.num_simulations + 1;
  const int T = fused_pick_trees(st, net, wbytes, out.B, N);
  if (T <= 0) {
    *err = "fused engine: trees + weights do not fit in shared memory";
    return 1;
  }
  FusedArgs a;
  a.net = net;
  a.weights = weights;
  a.weight_bytes = wbytes;
  a.out = out;
  a.p = p;
  a.obs = obs;
  a.invalid = invalid;
  a.noise = noise;
  a.action_out = action_out;
  a.weights_out = weights_out;
  a.root_value_out = root_value_out;
  a.T = T;
  a.B = out.B;
  a.N = N;
  a.ld = fused_ld(net);
  a.dump_tree = getenv("MZ_FUSED_NO_DUMP") ? 0 : 1;
  const size_t smem = fused_smem_bytes(net, wbytes, T, N);
  const int grid = (out.B + T - 1) / T;
  void* args[] = {&a};
  cudaError_t e = cudaLaunchKernel(fused_kernel_ptr(st.G), dim3(grid), dim3(st.threads), args, smem, stream);
  if (e != cudaSuccess) {
    *err = std::string("fused engine launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

}  // namespace mz

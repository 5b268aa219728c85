// mz_resident.cuh — CTA-resident engine: ONE launch per act for shapes whose trees do not fit shared memory.
//
// The stepwise engine pays three launches per simulation and the shared-memory engines need the whole tree of
// every tree they own on chip (C3: 78 KB per tree, C5: 75 KB).  Trees are independent, so nothing forces a
// grid-wide step: here a CTA owns T trees for the whole act, the trees stay in the handle's SoA arrays in HBM
// (the rows a CTA owns are touched by that CTA only, so they live in its SM's L1 / in L2 between visits), the
// weights are staged once into shared memory by a TMA bulk copy when they fit (else read through the read-only
// path, L2 resident), and the loop  select -> recurrent_fn -> expand + backup  runs num_simulations times with
// CTA barriers only.  All B trees are in flight at once: grid = ceil(B / T) with T chosen so that every SM holds
// as many CTAs as shared memory allows — latency of one CTA's tree walk hides behind another CTA's MLP.
//
// Arithmetic and orders are the shared device functions of mz_device.cuh: bit-identical to the other engines.
#pragma once
#include "mz_fused.cuh"

namespace mz {

struct ResidentArgs {
  Net net;
  const float* weights;  // global fp32 blob
  int32_t weight_bytes;  // multiple of 16
  int32_t weights_in_smem;
  Tree t;                // the handle's SoA tree (whole batch)
  SearchParams p;
  const float* obs;          // [B,obs_dim] or null
  const float* root_emb;     // [B,E] when obs is null
  const float* root_logits;  // [B,A] or null (then Prediction runs here)
  const float* root_value;   // [B]   or null
  const uint8_t* invalid;
  const float* noise;
  int32_t* action_out;
  float* weights_out;
  float* root_value_out;
  int32_t T;   // trees per CTA
  int32_t ld;  // MLP staging row stride (floats)
  int32_t clear_embeddings;
};

struct ResidentLayout {  // offsets in floats from the dynamic smem base
  int weights, mlp, sel, total_floats;
};

__host__ __device__ inline ResidentLayout resident_layout(int weight_bytes_in_smem, int T, int ld) {
  ResidentLayout L;
  int off = 0;
  L.weights = off; off += round_up(weight_bytes_in_smem / 4, 4);
  L.mlp = off;     off += 8 * T * ld;  // x, ns, headA, headB, tmp0A, tmp1A, tmp0B, tmp1B
  L.sel = off;     off += 5 * T + 4;   // parent, action, next, reward, value
  L.total_floats = round_up(off, 4);
  return L;
}

// View of the T trees starting at global row `row0` (local tree index 0..R-1 inside the CTA).
__device__ __forceinline__ Tree tree_rows(const Tree& g, int row0, int num_sims) {
  Tree t = g;
  const long n0 = (long)row0 * g.N, c0 = n0 * g.A;
  t.node_visits += n0; t.parents += n0; t.action_from_parent += n0; t.raw_values += n0; t.node_values += n0;
  t.children_index += c0; t.children_visits += c0; t.children_prior_logits += c0; t.children_prior_probs += c0;
  t.children_values += c0; t.children_rewards += c0; t.children_discounts += c0;
  t.embeddings += n0 * g.E;
  t.root_noise += (long)row0 * g.A;
  t.root_invalid += (long)row0 * g.A;
  t.sim_depth += (long)row0 * num_sims;
  return t;
}

template <int G, bool kWSmem>
__global__ void __launch_bounds__(256) resident_search_kernel(ResidentArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  constexpr bool kLdg = !kWSmem;
  const int T = a.T, A = a.net.num_actions, E = a.net.embed_dim, ld = a.ld;
  const int row0 = blockIdx.x * T;
  const int R = min(T, a.t.B - row0);
  const int tid = threadIdx.x;
  const int NS = a.p.num_simulations;
  const ResidentLayout L = resident_layout(kWSmem ? a.weight_bytes : 0, T, ld);

  const float* w = a.weights;
  if constexpr (kWSmem) {
    float* ws = smem + L.weights;
    if (tid == 0) {
      mbar_init(&wbar, 1);
      mbar_expect_tx(&wbar, (uint32_t)a.weight_bytes);
      tma_bulk_g2s(ws, a.weights, (uint32_t)a.weight_bytes, &wbar);
    }
    w = ws;
  }

  const Tree t = tree_rows(a.t, row0, NS);
  const int N = t.N;

  // mctx initial state (Appendix A.1) for this CTA's rows: zeros, parents / action_from_parent / children_index = -1
  {
    const int RN = R * N, RNA = RN * A;
    for (int i = tid; i < RN; i += blockDim.x) {
      t.node_visits[i] = 0;
      t.parents[i] = -1;
      t.action_from_parent[i] = -1;
      t.raw_values[i] = 0.0f;
      t.node_values[i] = 0.0f;
    }
    for (int i = tid; i < RNA; i += blockDim.x) {
      t.children_index[i] = -1;
      t.children_visits[i] = 0;
      t.children_prior_logits[i] = 0.0f;
      t.children_prior_probs[i] = 0.0f;
      t.children_values[i] = 0.0f;
      t.children_rewards[i] = 0.0f;
      t.children_discounts[i] = 0.0f;
    }
    if (a.clear_embeddings) {
      float4* e4 = reinterpret_cast<float4*>(t.embeddings);  // row0 * N * E * 4 bytes: 16-byte aligned when E % 4 == 0
      const long n = (long)RN * E;
      if ((E & 3) == 0) {
        for (long i = tid; i < n / 4; i += blockDim.x) e4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      } else {
        for (long i = tid; i < n; i += blockDim.x) t.embeddings[i] = 0.0f;
      }
    }
  }

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

  // ---- root inference (muax/model.py:251-263); the root embedding lands in `ns`, the prior logits in `headB`
  const int obs_dim = a.net.obs_dim;
  if (a.obs != nullptr) {
    for (int i = tid; i < R * obs_dim; i += blockDim.x) {
      const int r = i / obs_dim, k = i - r * obs_dim;
      x[r * ld + k] = a.obs[(long)(row0 + r) * obs_dim + k];
    }
  } else {
    for (int i = tid; i < R * E; i += blockDim.x) {
      const int r = i / E, k = i - r * E;
      ns[r * ld + k] = a.root_emb[(long)(row0 + r) * E + k];
    }
  }
  if constexpr (kWSmem) mbar_wait(&wbar, 0);
  __syncthreads();
  if (a.obs != nullptr) {
    stack_forward_cta<kLdg>(a.net.repr, w, a.net.activation, x, ld, obs_dim, nullptr, ns, ld, ta0, ta1, ld, R);
    if (a.net.repr_minmax) min_max_normalize_cta(ns, ld, E, R);
  }
  if (a.obs != nullptr || a.root_logits == nullptr) {
    dual_stack_forward_cta<kLdg>(a.net.pred_v, a.net.pred_pi, w, a.net.activation, ns, ld, E, nullptr, headA, headB, ld,
                                 ta0, ta1, tb0, tb1, ld, R);
    if (tid < R) rec_value[tid] = support_to_scalar_row(headA + tid * ld, a.net.support_size);
  } else {
    for (int i = tid; i < R * A; i += blockDim.x) {
      const int r = i / A, k = i - r * A;
      headB[r * ld + k] = a.root_logits[(long)(row0 + r) * A + k];
    }
    if (tid < R) rec_value[tid] = a.root_value[row0 + tid];
  }
  __syncthreads();
  if (tid < R && a.root_value_out != nullptr) a.root_value_out[row0 + tid] = rec_value[tid];  // raw value (model.py:243)

  // ---- policy prologue + tree instantiation
  const int ngroups = blockDim.x / G;
  const int gi = tid / G;        // lane group of this thread
  const int ga = tid & (G - 1);  // action handled by this lane
  const unsigned gm = group_mask<G>();
  for (int b = gi; b < R; b += ngroups) {
    const long ba = (long)(row0 + b) * A;
    group_begin<G>(t, p, b, (long)p.batch_offset + b, headB + b * ld, rec_value[b], ns + b * ld,
                   a.invalid != nullptr ? a.invalid + ba : nullptr, a.noise != nullptr ? a.noise + ba : nullptr, ga, gm);
  }
  __syncthreads();

  // ---- simulations
  for (int sim = 0; sim < NS; ++sim) {
    for (int b = gi; b < R; b += ngroups) {
      int parent, action, next, depth;
      group_simulate<G>(t, p, b, sim, ga, gm, parent, action, next, depth);
      if (ga == 0) {
        sel_parent[b] = parent;
        sel_action[b] = action;
        sel_next[b] = next;
        t.sim_depth[(long)b * NS + sim] = depth;
      }
      for (int e = ga; e < E; e += G) x[b * ld + e] = t.embeddings[((long)b * N + parent) * E + e];
    }
    __syncthreads();
    // recurrent_fn (muax/model.py:265-282): Dynamic (both heads) -> min-max -> Prediction (both heads)
    dual_stack_forward_cta<kLdg>(a.net.dyn_ns, a.net.dyn_r, w, a.net.activation, x, ld, E, sel_action, ns, headA, ld, ta0,
                                 ta1, tb0, tb1, ld, R);
    if (tid < R) rec_reward[tid] = support_to_scalar_row(headA + tid * ld, a.net.support_size);
    if (a.net.dyn_minmax)
      min_max_normalize_cta(ns, ld, E, R);
    else
      __syncthreads();  // headA (reward logits) is about to be reused by the value head
    dual_stack_forward_cta<kLdg>(a.net.pred_v, a.net.pred_pi, w, a.net.activation, ns, ld, E, nullptr, headA, headB, ld,
                                 ta0, ta1, tb0, tb1, ld, R);
    if (tid < R) rec_value[tid] = support_to_scalar_row(headA + tid * ld, a.net.support_size);
    __syncthreads();
    for (int b = gi; b < R; b += ngroups) {
      const float logit = ga < A ? headB[b * ld + ga] : 0.0f;
      group_expand_backup<G>(t, b, sel_parent[b], sel_action[b], sel_next[b], rec_reward[b], p.discount, rec_value[b],
                             logit, ns + b * ld, ga, gm);
    }
    // the next select of a tree runs on the lanes of the same group (same warp): a warp-level fence orders the
    // backup's global writes before it; the staging buffers are only rewritten after the next CTA barrier
    __syncwarp();
  }

  // ---- policy epilogue
  for (int b = gi; b < R; b += ngroups) {
    int action;
    float weight;
    group_finish<G>(t, p, b, (long)p.batch_offset + b, a.invalid != nullptr, ga, gm, action, weight);
    if (ga < A) a.weights_out[(long)(row0 + b) * A + ga] = weight;
    if (ga == 0) a.action_out[row0 + b] = action;
  }
}

// ------------------------------------------------------------------------------------------ host side

struct ResidentState {
  bool available = false;
  int max_smem = 0, num_sms = 0, G = 0;
  int threads = 256;
  int trees_per_cta = 0;  // 0 = choose per launch
  int force_global_weights = 0;
};

inline void* resident_kernel_ptr(int G, bool wsmem) {
#define MZ_RES_CASE(g) case g: return wsmem ? (void*)resident_search_kernel<g, true> : (void*)resident_search_kernel<g, false>
  switch (G) {
    MZ_RES_CASE(2);
    MZ_RES_CASE(4);
    MZ_RES_CASE(8);
    MZ_RES_CASE(16);
    default: return wsmem ? (void*)resident_search_kernel<32, true> : (void*)resident_search_kernel<32, false>;
  }
#undef MZ_RES_CASE
}

inline int net_weight_bytes(const Net& net) {
  int64_t wfloats = 0;
  const mz_stack* stacks[5] = {&net.repr, &net.pred_v, &net.pred_pi, &net.dyn_ns, &net.dyn_r};
  for (const mz_stack* s : stacks)
    for (int l = 0; l < s->n_layers; ++l) {
      wfloats = std::max(wfloats, s->w_off[l] + (int64_t)s->in_dim[l] * s->out_dim[l]);
      wfloats = std::max(wfloats, s->b_off[l] + (int64_t)s->out_dim[l]);
    }
  return round_up((int)wfloats * 4, 16);
}

inline int resident_init(ResidentState& st, const Net& net, int device, std::string* err) {
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
  for (int ws = 0; ws < 2; ++ws) {
    const cudaError_t e = cudaFuncSetAttribute(resident_kernel_ptr(G, ws != 0),
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, st.max_smem - 1024);
    if (e != cudaSuccess) {
      *err = std::string("resident engine: cudaFuncSetAttribute failed: ") + cudaGetErrorString(e);
      cudaGetLastError();
      return 1;
    }
  }
  if (const char* e = getenv("MZ_RESIDENT_TREES")) st.trees_per_cta = atoi(e);
  if (const char* e = getenv("MZ_RESIDENT_GLOBAL_WEIGHTS")) st.force_global_weights = atoi(e);
  st.available = true;
  return 0;
}

struct ResidentPlan {
  int T = 0, grid = 0, wsmem = 0;
  size_t smem = 0;
};

// Trees per CTA: every SM should hold as many co-resident CTAs as shared memory (weights + MLP staging) and the
// thread budget allow, and all of B should be in flight at once.
inline ResidentPlan resident_plan(const ResidentState& st, const Net& net, int B) {
  ResidentPlan best;
  const int ld = fused_ld(net);
  const int wbytes = net_weight_bytes(net);
  const int budget = st.max_smem - 1024;      // per CTA (opt-in limit, minus the static mbarrier + slack)
  const int sm_budget = 227 * 1024;           // per SM
  for (int ws = st.force_global_weights ? 0 : 1; ws >= 0; --ws) {
    auto bytes = [&](int T) { return (size_t)resident_layout(ws ? wbytes : 0, T, ld).total_floats * 4; };
    if (bytes(1) > (size_t)budget) continue;
    int T;
    if (st.trees_per_cta > 0) {
      T = st.trees_per_cta;
    } else {
      // smallest T (>= 1) such that ceil(B / T) CTAs are co-resident: ctas_per_sm(T) * num_sms * T >= B
      T = 0;
      for (int cand = 1; cand <= 64; ++cand) {
        if (bytes(cand) > (size_t)budget) break;
        int per_sm = (int)(sm_budget / (bytes(cand) + 1024));
        per_sm = std::min(per_sm, 2048 / st.threads);
        per_sm = std::min(per_sm, 8);
        if ((long)per_sm * st.num_sms * cand >= B) {
          T = cand;
          break;
        }
        T = cand;  // largest that fits so far (several waves if nothing covers B)
      }
    }
    if (T <= 0 || bytes(T) > (size_t)budget) continue;
    // with the weights in shared memory a tiny T replicates them per CTA for nothing: keep T >= 4 when B allows
    best.T = T;
    best.wsmem = ws;
    best.smem = bytes(T);
    best.grid = (B + T - 1) / T;
    return best;
  }
  return best;
}

inline bool resident_supported(const ResidentState& st, const Net& net, int B) {
  return st.available && resident_plan(st, net, B).T > 0;
}

inline int resident_launch(ResidentState& st, const Net& net, const float* weights, const Tree& tree,
                           const SearchParams& p, const float* obs, const float* root_emb, const float* root_logits,
                           const float* root_value, const uint8_t* invalid, const float* noise, int32_t* action_out,
                           float* weights_out, float* root_value_out, cudaStream_t stream, std::string* err) {
  const ResidentPlan plan = resident_plan(st, net, tree.B);
  if (plan.T <= 0) {
    *err = "resident engine: MLP staging does not fit in shared memory";
    return 1;
  }
  ResidentArgs a{};
  a.net = net;
  a.weights = weights;
  a.weight_bytes = net_weight_bytes(net);
  a.weights_in_smem = plan.wsmem;
  a.t = tree;
  a.p = p;
  a.obs = obs;
  a.root_emb = root_emb;
  a.root_logits = root_logits;
  a.root_value = root_value;
  a.invalid = invalid;
  a.noise = noise;
  a.action_out = action_out;
  a.weights_out = weights_out;
  a.root_value_out = root_value_out;
  a.T = plan.T;
  a.ld = fused_ld(net);
  a.clear_embeddings = (p.max_depth > 0 || p.num_simulations + 1 < tree.N) ? 1 : 0;
  void* args[] = {&a};
  const cudaError_t e = cudaLaunchKernel(resident_kernel_ptr(st.G, plan.wsmem != 0), dim3(plan.grid), dim3(st.threads),
                                         args, plan.smem, stream);
  if (e != cudaSuccess) {
    *err = std::string("resident engine launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

}  // namespace mz

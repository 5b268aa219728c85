// mzsearch.cu — libmzsearch.so: kernels + C ABI (include/mzsearch.h) of the B200 batched MuZero search.
//
// Stepwise engine (this file): per simulation three kernels over trees resident in HBM —
//   select  (mctx simulate + action selection, Appendix A.3-A.6)         warp lane-group per tree
//   recurrent (muax/model.py:265-282: Dynamic + Prediction + support transform) CTA-cooperative fp32 MLP
//   expand_backup (mctx expand scatter + backward, Appendix A.3)          warp lane-group per tree
// bracketed by root (muax/model.py:251-263), begin (policy prologue A.2/A.4 + instantiate_tree_from_root)
// and finish (summary + temperature + categorical, or the Gumbel epilogue).
// The one-launch engines live in their own translation units: mz_warp.cu (headline shapes, trees in shared memory),
// mz_treewarp.cu (any net whose weights fit shared memory, trees as records in L1/L2), mz_resident.cu (wide nets,
// weights streamed), mz_recurrent_tc.cu (tcgen05 recurrent_fn of the bf16 throughput mode).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <algorithm>

#include "mz_device.cuh"
#include "mz_recurrent_tc.cuh"
#include "mz_resident.cuh"
#include "mz_treewarp.cuh"
#include "mz_warp.cuh"

namespace mz {

// ------------------------------------------------------------------------------------------ error plumbing

static thread_local std::string g_last_error;

static int fail(const std::string& msg) {  // runtime failure (CUDA error, unsupported configuration, misuse of state)
  g_last_error = msg;
  return MZ_ERR_RUNTIME;
}
static int fail_arg(const std::string& msg) {  // the caller passed an invalid argument (Python: ValueError)
  g_last_error = msg;
  return MZ_ERR_INVALID_ARGUMENT;
}

#define MZ_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t err__ = (expr);                                                                         \
    if (err__ != cudaSuccess)                                                                           \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(err__) + " (" + __FILE__ + ":" +       \
                  std::to_string(__LINE__) + ")");                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------ host threefry (key chain)

static inline uint32_t h_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void h_threefry(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = h_rotl(x1, R[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  o0 = x0;
  o1 = x1;
}

// jax.random.split(key, num)[j]
static void h_split_key(const uint32_t key[2], uint32_t num, uint32_t j, int mode, uint32_t out[2]) {
  if (mode == MZ_PRNG_THREEFRY_LEGACY) {
    const uint32_t n = 2 * num, half = num;
    for (int w = 0; w < 2; ++w) {
      const uint32_t m = 2 * j + w;
      const uint32_t i = m < half ? m : m - half;
      uint32_t y0, y1;
      h_threefry(key[0], key[1], i, half + i < n ? half + i : 0u, y0, y1);
      out[w] = m < half ? y0 : y1;
    }
  } else {
    h_threefry(key[0], key[1], 0u, j, out[0], out[1]);
  }
}

// seq_halving.get_sequence_of_considered_visits (Appendix A.4)
static void h_considered_sequence(int m, int n, int32_t* seq) {
  if (m <= 1) {
    for (int i = 0; i < n; ++i) seq[i] = i;
    return;
  }
  const int log2max = (int)std::ceil(std::log2((double)m));
  std::vector<int32_t> visits(m, 0);
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

// ------------------------------------------------------------------------------------------ kernels

constexpr int kTreeThreads = 128;  // threads per CTA of the lane-group kernels
constexpr int kMlpThreads = 128;
constexpr int kMlpRows = 8;        // rows (trees) per CTA in the MLP kernels

struct RootIO {
  const float* obs;        // [B,obs_dim], or null when emb_in is given
  const float* emb_in;     // [B,E] root embedding computed by the caller (e.g. a conv torso), or null
  float *logits, *value, *emb;  // scratch outputs [B,A], [B], [B,E]
};

// muax/model.py:251-263 for kMlpRows rows per CTA.
__global__ void __launch_bounds__(kMlpThreads) root_kernel(Net net, const float* __restrict__ w, RootIO io, int B) {
  extern __shared__ float smem[];
  const int row0 = blockIdx.x * kMlpRows;
  const int R = min(kMlpRows, B - row0);
  const int ld = net.max_width;
  float* x = smem;                  // [R][ld]
  float* t0 = x + kMlpRows * ld;    // [R][ld]
  float* t1 = t0 + kMlpRows * ld;   // [R][ld]
  float* emb = t1 + kMlpRows * ld;  // [R][ld]
  float* head = emb + kMlpRows * ld;  // [R][ld]
  if (io.emb_in == nullptr) {
    for (int i = threadIdx.x; i < R * net.obs_dim; i += blockDim.x) {
      const int r = i / net.obs_dim, k = i - r * net.obs_dim;
      x[r * ld + k] = io.obs[(long)(row0 + r) * net.obs_dim + k];
    }
    __syncthreads();
    stack_forward_cta(net.repr, w, net.activation, x, ld, net.obs_dim, nullptr, emb, ld, t0, t1, ld, R);
    if (net.repr_minmax) min_max_normalize_cta(emb, ld, net.embed_dim, R);
  } else {
    for (int i = threadIdx.x; i < R * net.embed_dim; i += blockDim.x) {
      const int r = i / net.embed_dim, k = i - r * net.embed_dim;
      emb[r * ld + k] = io.emb_in[(long)(row0 + r) * net.embed_dim + k];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < R * net.embed_dim; i += blockDim.x) {
    const int r = i / net.embed_dim, k = i - r * net.embed_dim;
    io.emb[(long)(row0 + r) * net.embed_dim + k] = emb[r * ld + k];
  }
  stack_forward_cta(net.pred_v, w, net.activation, emb, ld, net.embed_dim, nullptr, head, ld, t0, t1, ld, R);
  if (threadIdx.x < R) io.value[row0 + threadIdx.x] = support_to_scalar_row(head + threadIdx.x * ld, net.support_size);
  __syncthreads();
  stack_forward_cta(net.pred_pi, w, net.activation, emb, ld, net.embed_dim, nullptr, head, ld, t0, t1, ld, R);
  for (int i = threadIdx.x; i < R * net.num_actions; i += blockDim.x) {
    const int r = i / net.num_actions, k = i - r * net.num_actions;
    io.logits[(long)(row0 + r) * net.num_actions + k] = head[r * ld + k];
  }
}

// Policy prologue (Appendix A.2 / A.4) + instantiate_tree_from_root (A.3): one lane group per tree.
template <int G>
__global__ void __launch_bounds__(kTreeThreads) begin_kernel(Tree t, SearchParams p, const float* root_logits,
                                                             const float* root_value, const float* root_emb,
                                                             const uint8_t* invalid, const float* noise) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (b >= t.B) return;
  const int a = threadIdx.x & (G - 1);
  const unsigned m = group_mask<G>();
  const long ba = (long)b * t.A;
  group_begin<G>(t, p, b, (long)p.batch_offset + b, root_logits + ba, root_value[b], root_emb + (long)b * t.E,
                 invalid != nullptr ? invalid + ba : nullptr, noise != nullptr ? noise + ba : nullptr, a, m);
}

struct SelectIO {
  int32_t *parent, *action, *next;  // [B] scratch
  int32_t* action_out;              // callback mode: user-visible copy of `action` (may be null)
  float* parent_emb_out;            // callback mode: [B,E] gathered parent embeddings (may be null)
};

template <int G>
__global__ void __launch_bounds__(kTreeThreads) select_kernel(Tree t, SearchParams p, int sim, SelectIO io) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (b >= t.B) return;
  const int a = threadIdx.x & (G - 1);
  const unsigned m = group_mask<G>();
  int parent, action, next, depth;
  group_simulate<G>(t, p, b, sim, a, m, parent, action, next, depth);
  if (a == 0) {
    io.parent[b] = parent;
    io.action[b] = action;
    io.next[b] = next;
    t.sim_depth[(long)b * p.num_simulations + sim] = depth;
    if (io.action_out != nullptr) io.action_out[b] = action;
  }
  if (io.parent_emb_out != nullptr)
    for (int e = a; e < t.E; e += G)
      io.parent_emb_out[(long)b * t.E + e] = t.embeddings[((long)b * t.N + parent) * t.E + e];
}

struct RecurrentIO {
  const int32_t *parent, *action;          // [B]
  float *reward, *value, *logits, *next_emb;  // [B], [B], [B,A], [B,E]
};

// muax/model.py:265-282 for kMlpRows rows per CTA: gather parent embedding -> Dynamic -> (min-max) ->
// Prediction -> 2x support_to_scalar(softmax(.)).
__global__ void __launch_bounds__(kMlpThreads) recurrent_kernel(Net net, const float* __restrict__ w, Tree t,
                                                                RecurrentIO io) {
  extern __shared__ float smem[];
  __shared__ int s_action[kMlpRows];
  const int row0 = blockIdx.x * kMlpRows;
  const int R = min(kMlpRows, t.B - row0);
  const int ld = net.max_width;
  const int E = net.embed_dim;
  float* x = smem;
  float* t0 = x + kMlpRows * ld;
  float* t1 = t0 + kMlpRows * ld;
  float* ns = t1 + kMlpRows * ld;
  float* head = ns + kMlpRows * ld;
  if (threadIdx.x < R) s_action[threadIdx.x] = io.action[row0 + threadIdx.x];
  for (int i = threadIdx.x; i < R * E; i += blockDim.x) {
    const int r = i / E, k = i - r * E;
    const int b = row0 + r;
    x[r * ld + k] = t.embeddings[((long)b * t.N + io.parent[b]) * E + k];
  }
  __syncthreads();
  stack_forward_cta(net.dyn_r, w, net.activation, x, ld, E, s_action, head, ld, t0, t1, ld, R);
  if (threadIdx.x < R) io.reward[row0 + threadIdx.x] = support_to_scalar_row(head + threadIdx.x * ld, net.support_size);
  __syncthreads();
  stack_forward_cta(net.dyn_ns, w, net.activation, x, ld, E, s_action, ns, ld, t0, t1, ld, R);
  if (net.dyn_minmax) min_max_normalize_cta(ns, ld, E, R);
  for (int i = threadIdx.x; i < R * E; i += blockDim.x) {
    const int r = i / E, k = i - r * E;
    io.next_emb[(long)(row0 + r) * E + k] = ns[r * ld + k];
  }
  stack_forward_cta(net.pred_v, w, net.activation, ns, ld, E, nullptr, head, ld, t0, t1, ld, R);
  if (threadIdx.x < R) io.value[row0 + threadIdx.x] = support_to_scalar_row(head + threadIdx.x * ld, net.support_size);
  __syncthreads();
  stack_forward_cta(net.pred_pi, w, net.activation, ns, ld, E, nullptr, head, ld, t0, t1, ld, R);
  for (int i = threadIdx.x; i < R * net.num_actions; i += blockDim.x) {
    const int r = i / net.num_actions, k = i - r * net.num_actions;
    io.logits[(long)(row0 + r) * net.num_actions + k] = head[r * ld + k];
  }
}

struct ExpandIO {
  const int32_t *parent, *action, *next;                     // [B]
  const float *reward, *discount, *value, *logits, *next_emb;  // discount may be null -> constant
};

template <int G>
__global__ void __launch_bounds__(kTreeThreads) expand_backup_kernel(Tree t, SearchParams p, ExpandIO io) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (b >= t.B) return;
  const int a = threadIdx.x & (G - 1);
  const unsigned m = group_mask<G>();
  const float logit = a < t.A ? io.logits[(long)b * t.A + a] : 0.0f;
  const float discount = io.discount != nullptr ? io.discount[b] : p.discount;
  group_expand_backup<G>(t, b, io.parent[b], io.action[b], io.next[b], io.reward[b], discount, io.value[b], logit,
                         io.next_emb + (long)b * t.E, a, m);
}

// MuZero: summary().visit_probs -> _apply_temperature -> random.categorical (A.2).
// Gumbel: argmax of gumbel + logits + completed Q over the most-visited actions; softmax(logits + Q) (A.4).
template <int G>
__global__ void __launch_bounds__(kTreeThreads) finish_kernel(Tree t, SearchParams p, bool has_invalid,
                                                              int32_t* action_out, float* weights_out) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (b >= t.B) return;
  const int a = threadIdx.x & (G - 1);
  const unsigned m = group_mask<G>();
  int action;
  float weight;
  group_finish<G>(t, p, b, (long)p.batch_offset + b, has_invalid, a, m, action, weight);
  if (a < t.A) weights_out[(long)b * t.A + a] = weight;
  if (a == 0) action_out[b] = action;
}

__global__ void math_probe_kernel(int kind, const float* x, float* y, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = kind == 5 ? 0.0f : x[i];
  float r;
  switch (kind) {
    case 0: r = mz_expf(v); break;
    case 1: r = mz_logf(v); break;
    case 2: r = mz_expm1f(v); break;
    case 3: r = mz_inv_scaling(v); break;
    case 5: {  // x holds (a, b) pairs for 2n inputs; this thread handles pair i (n = number of pairs)
      const float a = x[2 * i], b = x[2 * i + 1];
      bool bad = false;
      r = div_try(a, b, bad);
      r = bad ? __uint_as_float(0x7fc00001u) : r;  // marker NaN: "the kernels would take the IEEE slow path here"
      break;
    }
    default: r = mz_bits_to_gumbel(__float_as_uint(v)); break;
  }
  y[i] = r;
}

}  // namespace mz

// ------------------------------------------------------------------------------------------ handle

struct mz_handle {
  mz_config cfg{};
  mz::Net net{};
  mz::Tree tree{};
  int N = 0, G = 0;
  float* weights = nullptr;
  size_t n_weights = 0;
  // scratch
  int32_t *sel_parent = nullptr, *sel_action = nullptr, *sel_next = nullptr;
  int32_t* sel5 = nullptr;  // 5 x B scratch of the batched tree-warp kernels (throughput mode)
  bool tree16 = false;      // the last search kept its embeddings in the tcgen05 kernel's bf16 rows (mz_get_tree exports them)
  float *rec_reward = nullptr, *rec_value = nullptr, *rec_logits = nullptr, *rec_emb = nullptr;
  float *root_logits = nullptr, *root_value = nullptr, *root_emb = nullptr;
  uint32_t* sim_keys_dev = nullptr;
  int32_t* table_dev = nullptr;
  int table_m = -1, table_n = -1;
  // pinned ring for the per-call simulate keys
  static constexpr int kSlots = 8;
  uint32_t* key_slots = nullptr;
  cudaEvent_t slot_done[kSlots]{};
  int next_slot = 0;
  // host staging for mz_search_host
  float *h_obs = nullptr, *h_noise = nullptr, *h_weights_out = nullptr, *h_value_out = nullptr;
  uint8_t* h_invalid = nullptr;
  int32_t* h_action_out = nullptr;
  float *d_obs = nullptr, *d_noise = nullptr, *d_weights_out = nullptr, *d_value_out = nullptr;
  uint8_t* d_invalid = nullptr;
  int32_t* d_action_out = nullptr;
  // bookkeeping
  mz::SearchParams params{};
  bool has_invalid = false;
  int64_t launches = 0;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  bool timed = false;
  std::vector<void*> allocs;
  mz::WarpState warpeng;
  mz::TreeWarpState treewarp;
  mz::RecurrentTcState rtc;
  mz::SimKeys host_keys{};      // simulate keys of the call in flight (host copy)
  bool keys_on_device = false;  // ... and whether they have been staged into sim_keys_dev
  int key_slot = 0;
  bool tree_valid = false;      // the SoA tree view describes the last search (want_tree, or an engine that always has it)
  std::vector<int64_t> peer_deltas;  // mz_set_peer_outputs
  mz::ResidentState resident;
};

namespace mz {

template <typename T>
static int dev_alloc(mz_handle* h, T** p, size_t count) {
  void* q = nullptr;
  MZ_CUDA(cudaMalloc(&q, count * sizeof(T) + 16));
  h->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}

static int stack_max_width(const mz_stack& s, int cur) {
  for (int l = 0; l < s.n_layers; ++l) {
    cur = std::max(cur, (int)s.in_dim[l]);
    cur = std::max(cur, (int)s.out_dim[l]);
  }
  return cur;
}

static int validate_stack(const mz_stack& s, int in, int out, const char* name, size_t* need) {
  if (s.n_layers < 1 || s.n_layers > MZ_MAX_LAYERS) return fail_arg(std::string(name) + ": n_layers out of range");
  if (s.in_dim[0] != in) return fail_arg(std::string(name) + ": first layer input width mismatch");
  if (s.out_dim[s.n_layers - 1] != out) return fail_arg(std::string(name) + ": last layer output width mismatch");
  for (int l = 0; l < s.n_layers; ++l) {
    if (s.in_dim[l] < 1 || s.out_dim[l] < 1) return fail_arg(std::string(name) + ": bad layer width");
    if (l > 0 && s.in_dim[l] != s.out_dim[l - 1]) return fail_arg(std::string(name) + ": layer widths do not chain");
    if (s.w_off[l] < 0 || s.b_off[l] < 0) return fail_arg(std::string(name) + ": negative offset");
    *need = std::max(*need, (size_t)s.w_off[l] + (size_t)s.in_dim[l] * s.out_dim[l]);
    *need = std::max(*need, (size_t)s.b_off[l] + (size_t)s.out_dim[l]);
  }
  return 0;
}

static size_t mlp_smem_bytes(const Net& net) { return (size_t)5 * kMlpRows * net.max_width * sizeof(float); }

static int tree_blocks(const mz_handle* h) { return (h->cfg.batch * h->G + kTreeThreads - 1) / kTreeThreads; }

#define MZ_DISPATCH_G(h, KERNEL, GRID, STREAM, ...)                                              \
  do {                                                                                           \
    switch ((h)->G) {                                                                            \
      case 2: KERNEL<2><<<GRID, kTreeThreads, 0, STREAM>>>(__VA_ARGS__); break;                  \
      case 4: KERNEL<4><<<GRID, kTreeThreads, 0, STREAM>>>(__VA_ARGS__); break;                  \
      case 8: KERNEL<8><<<GRID, kTreeThreads, 0, STREAM>>>(__VA_ARGS__); break;                  \
      case 16: KERNEL<16><<<GRID, kTreeThreads, 0, STREAM>>>(__VA_ARGS__); break;                \
      default: KERNEL<32><<<GRID, kTreeThreads, 0, STREAM>>>(__VA_ARGS__); break;                \
    }                                                                                            \
    (h)->launches += 1;                                                                          \
  } while (0)

static int check_args(const mz_handle* h, const mz_search_args* a) {
  if (a == nullptr) return fail_arg("args is NULL");
  if (a->policy != MZ_POLICY_MUZERO && a->policy != MZ_POLICY_GUMBEL) return fail_arg("unknown policy");
  if (a->qtransform != 0 && a->qtransform != 1) return fail_arg("unknown qtransform");
  if (a->precision != MZ_PRECISION_FP32 && a->precision != MZ_PRECISION_BF16) return fail_arg("unknown precision");
  if (a->num_decision_actions < 0 || a->num_decision_actions >= h->cfg.num_actions)
    return fail_arg("num_decision_actions must be in [0, num_actions)");
  if (a->num_decision_actions > 0 && a->policy != MZ_POLICY_MUZERO)
    return fail_arg("num_decision_actions (stochastic MuZero) needs policy = MZ_POLICY_MUZERO");
  if (a->num_simulations < 0 || a->num_simulations > h->cfg.max_num_simulations)
    return fail_arg("num_simulations exceeds the handle's max_num_simulations");
  if (a->policy == MZ_POLICY_GUMBEL && (a->max_considered < 0 || a->max_considered > 1024))
    return fail_arg("max_num_considered_actions out of range");
  const int gbatch = a->global_batch > 0 ? a->global_batch : h->cfg.batch;
  if (a->batch_offset < 0 || a->batch_offset + h->cfg.batch > gbatch)
    return fail_arg("batch_offset + batch exceeds global_batch");
  return 0;
}

// Derive all keys of one act on the host (scalar chain, identical for every tree).  The per-simulation simulate keys
// stay on the host (h->host_keys) until an engine needs them: engines that take them in their kernel parameters
// (SimKeys, searches of at most kInlineSims simulations) never pay an H2D copy; the others call ensure_device_keys.
static int stage_keys(mz_handle* h, const mz_search_args* a, cudaStream_t stream) {
  SearchParams& p = h->params;
  const int mode = h->cfg.prng_mode;
  const int NS = a->num_simulations;
  p.policy = a->policy;
  p.qtransform = a->qtransform;
  p.num_simulations = NS;
  p.max_depth = a->max_depth;
  p.max_considered = a->max_considered;
  p.global_batch = a->global_batch > 0 ? a->global_batch : h->cfg.batch;
  p.batch_offset = a->batch_offset;
  p.prng_mode = mode;
  p.temperature = a->temperature;
  p.dirichlet_fraction = a->dirichlet_fraction;
  p.dirichlet_alpha = a->dirichlet_alpha;
  p.pb_c_init = a->pb_c_init;
  p.pb_c_base = a->pb_c_base;
  p.gumbel_scale = a->gumbel_scale;
  p.value_scale = a->value_scale;
  p.maxvisit_init = a->maxvisit_init;
  p.discount = h->cfg.discount;
  p.stoch_A = a->num_decision_actions;
  const uint32_t rng[2] = {a->key0, a->key1};
  uint32_t search_key[2], aux[2], fin[2] = {0, 0};
  if (a->policy == MZ_POLICY_MUZERO) {  // rng_key, dirichlet_key, search_key = split(rng_key, 3)
    h_split_key(rng, 3, 0, mode, fin);
    h_split_key(rng, 3, 1, mode, aux);
    h_split_key(rng, 3, 2, mode, search_key);
  } else {  // rng_key, gumbel_key = split(rng_key)
    h_split_key(rng, 2, 0, mode, search_key);
    h_split_key(rng, 2, 1, mode, aux);
  }
  p.aux_key0 = aux[0];
  p.aux_key1 = aux[1];
  p.final_key0 = fin[0];
  p.final_key1 = fin[1];
  // simulate keys: rng_key, simulate_key, expand_key = split(rng_key, 3) per simulation
  const int slot = h->next_slot;
  h->next_slot = (slot + 1) % mz_handle::kSlots;
  h->key_slot = slot;
  const bool inline_ok = NS <= kInlineSims;
  uint32_t* keys = h->host_keys.w;
  if (!inline_ok) {  // long searches stage through the pinned ring: the slot must have been consumed by its last copy
    MZ_CUDA(cudaEventSynchronize(h->slot_done[slot]));
    keys = h->key_slots + (size_t)slot * 2 * (h->cfg.max_num_simulations + 1);
  }
  uint32_t cur[2] = {search_key[0], search_key[1]};
  for (int s = 0; s < NS; ++s) {
    uint32_t nxt[2];
    h_split_key(cur, 3, 0, mode, nxt);
    h_split_key(cur, 3, 1, mode, keys + 2 * s);
    cur[0] = nxt[0];
    cur[1] = nxt[1];
  }
  h->keys_on_device = false;
  p.sim_keys = nullptr;
  p.considered_table = nullptr;
  if (a->policy == MZ_POLICY_GUMBEL) {
    const int M = a->max_considered;
    if (h->table_m != M || h->table_n != NS) {
      std::vector<int32_t> table((size_t)(M + 1) * std::max(NS, 1), 0);
      for (int m = 0; m <= M; ++m) h_considered_sequence(m, NS, table.data() + (size_t)m * NS);
      if (h->table_dev != nullptr) {
        MZ_CUDA(cudaStreamSynchronize(stream));
        MZ_CUDA(cudaFree(h->table_dev));
        h->table_dev = nullptr;
      }
      MZ_CUDA(cudaMalloc((void**)&h->table_dev, table.size() * sizeof(int32_t)));
      MZ_CUDA(cudaMemcpyAsync(h->table_dev, table.data(), table.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                              stream));
      MZ_CUDA(cudaStreamSynchronize(stream));  // `table` is pageable and dies at scope exit
      h->table_m = M;
      h->table_n = NS;
    }
    p.considered_table = h->table_dev;
  }
  return 0;
}

// The simulate keys of the call in flight -> device memory (h->params.sim_keys), once per call.
static int ensure_device_keys(mz_handle* h, cudaStream_t stream) {
  if (h->keys_on_device) return 0;
  const int NS = h->params.num_simulations;
  const int slot = h->key_slot;
  uint32_t* pinned = h->key_slots + (size_t)slot * 2 * (h->cfg.max_num_simulations + 1);
  if (NS <= kInlineSims) {  // the keys were derived into host_keys: move them into this call's pinned slot
    MZ_CUDA(cudaEventSynchronize(h->slot_done[slot]));
    std::memcpy(pinned, h->host_keys.w, sizeof(uint32_t) * 2 * NS);
  }
  if (NS > 0)
    MZ_CUDA(cudaMemcpyAsync(h->sim_keys_dev, pinned, sizeof(uint32_t) * 2 * NS, cudaMemcpyHostToDevice, stream));
  MZ_CUDA(cudaEventRecord(h->slot_done[slot], stream));
  h->params.sim_keys = h->sim_keys_dev;
  h->keys_on_device = true;
  return 0;
}

// mctx zero/-1 initialises the whole tree (A.1); the embeddings only need it when max_depth can leave slots unused.
static int clear_tree(mz_handle* h, int NS, bool clear_embeddings, cudaStream_t stream) {
  const Tree& t = h->tree;
  const size_t BN = (size_t)t.B * t.N, BNA = BN * t.A;
  MZ_CUDA(cudaMemsetAsync(t.node_visits, 0, BN * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.raw_values, 0, BN * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.node_values, 0, BN * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.parents, 0xFF, BN * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.action_from_parent, 0xFF, BN * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.children_index, 0xFF, BNA * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.children_visits, 0, BNA * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.children_prior_logits, 0, BNA * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.children_prior_probs, 0, BNA * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.children_values, 0, BNA * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.children_rewards, 0, BNA * 4, stream));
  MZ_CUDA(cudaMemsetAsync(t.children_discounts, 0, BNA * 4, stream));
  if (clear_embeddings) MZ_CUDA(cudaMemsetAsync(t.embeddings, 0, BN * t.E * 4, stream));
  if (NS > 0) MZ_CUDA(cudaMemsetAsync(t.sim_depth, 0, (size_t)t.B * NS * 4, stream));
  return 0;
}

static int launch_begin(mz_handle* h, const float* root_logits, const float* root_value, const float* root_emb,
                        const uint8_t* invalid, const float* noise, cudaStream_t stream) {
  h->has_invalid = invalid != nullptr;
  if (ensure_device_keys(h, stream)) return 1;  // select_kernel derives the per-tree keys from them
  if (clear_tree(h, h->params.num_simulations, h->params.max_depth > 0, stream)) return 1;
  MZ_DISPATCH_G(h, begin_kernel, tree_blocks(h), stream, h->tree, h->params, root_logits, root_value, root_emb,
                invalid, noise);
  MZ_CUDA(cudaGetLastError());
  return 0;
}

static int launch_root(mz_handle* h, const float* obs, const float* emb_in, cudaStream_t stream) {
  if (emb_in == nullptr && h->cfg.obs_dim <= 0)
    return fail("handle was created with obs_dim = 0: supply the root embedding instead of obs");
  if (h->weights == nullptr) return fail("mz_set_weights has not been called");
  RootIO io{obs, emb_in, h->root_logits, h->root_value, h->root_emb};
  const int grid = (h->cfg.batch + kMlpRows - 1) / kMlpRows;
  root_kernel<<<grid, kMlpThreads, mlp_smem_bytes(h->net), stream>>>(h->net, h->weights, io, h->cfg.batch);
  h->launches += 1;
  MZ_CUDA(cudaGetLastError());
  return 0;
}

// Throughput mode (precision = bf16): per simulation  select -> tcgen05 recurrent kernel -> expand + backup  as three
// launches.  The tree phases are the tree-warp engine's walks on packed records (mz_treewarp.cu: warp-uniform level
// loop, tie-break noise from the pre-pass table, parallel backup); the root runs on the tensor-core kernel too.
static int search_tc(mz_handle* h, const float* obs, const float* root_logits, const float* root_value,
                     const float* root_emb, const uint8_t* invalid, const float* noise, int32_t* action_out,
                     float* weights_out, float* root_value_out, cudaStream_t stream) {
  if (h->weights == nullptr) return fail("mz_set_weights has not been called");
  if (!h->rtc.available)
    return fail("precision = bf16: the tcgen05 recurrent kernel does not cover this network (" + h->rtc.why + ")");
  if (!h->treewarp.available) return fail("precision = bf16: the tree-warp kernels are unavailable on this device");
  if (h->params.num_simulations + 1 >= 0xFFFF) return fail("precision = bf16: at most 65533 simulations");
  std::string err;
  const int NS = h->params.num_simulations;
  if (obs != nullptr || root_logits == nullptr) {  // Prediction (and Representation) run in the library
    const bool from_obs = obs != nullptr;
    if (recurrent_tc_has_root(h->rtc, from_obs)) {
      if (recurrent_tc_root(h->rtc, h->net, h->cfg.batch, obs, root_emb, h->root_value, h->root_logits, h->root_emb,
                            stream, &h->launches, &err))
        return fail(err);
      if (from_obs) root_emb = h->root_emb;  // the caller's embedding is used where it lies
    } else {
      if (launch_root(h, obs, from_obs ? nullptr : root_emb, stream)) return 1;
      root_emb = h->root_emb;
    }
    root_logits = h->root_logits;
    root_value = h->root_value;
  }
  if (root_value_out != nullptr)
    MZ_CUDA(cudaMemcpyAsync(root_value_out, root_value, sizeof(float) * h->cfg.batch, cudaMemcpyDefault, stream));
  if (ensure_device_keys(h, stream)) return 1;
  if (treewarp_batched_begin(h->treewarp, h->resident, h->tree, h->params, root_logits, root_value, root_emb, invalid,
                             noise, h->sel5, stream, &h->launches, &err))
    return fail(err);
  const int B = h->cfg.batch;
  // tree embeddings in bf16 (what the tensor core reads): the recurrent kernel gathers and stores them itself
  const bool clear16 = h->params.max_depth > 0 || NS + 1 < h->N;
  if (recurrent_tc_tree_begin(h->rtc, h->net, B, h->N, root_emb, clear16, stream, &h->launches, &err)) return fail(err);
  // per simulation two launches: [backup of the previous simulation + selection] -> tcgen05 recurrent kernel
  static const bool split = getenv("MZ_TC_SPLIT_BACKUP") != nullptr && atoi(getenv("MZ_TC_SPLIT_BACKUP")) != 0;  // A/B
  for (int sim = 0; sim < NS; ++sim) {
    if (sim == 0 || split) {
      if (treewarp_batched_select(h->treewarp, sim, stream, &h->launches, &err)) return fail(err);
    } else if (treewarp_batched_backup_select(h->treewarp, sim, h->rec_reward, h->rec_value, h->rec_logits, nullptr,
                                              stream, &h->launches, &err)) {
      return fail(err);
    }
    if (recurrent_tc_tree_launch(h->rtc, h->net, B, h->N, h->sel5, h->sel5 + B, h->sel5 + 2 * B, h->rec_reward,
                                 h->rec_value, h->rec_logits, stream, &h->launches, &err))
      return fail(err);
    if ((split || sim == NS - 1) &&
        treewarp_batched_backup(h->treewarp, h->rec_reward, h->rec_value, h->rec_logits, nullptr, stream, &h->launches,
                                &err))
      return fail(err);
  }
  h->tree16 = true;
  if (treewarp_batched_finish(h->treewarp, action_out, weights_out, stream, &h->launches, &err)) return fail(err);
  return 0;
}

// Callback-free stepwise loop with the tcgen05 recurrent kernel between the generic SoA tree kernels (kept as the
// cross-check of search_tc: MZ_ENGINE_STEPWISE + precision = bf16).
static int run_stepwise_tc(mz_handle* h, cudaStream_t stream) {
  if (!h->rtc.available)
    return fail("precision = bf16: the tcgen05 recurrent kernel does not cover this network (" + h->rtc.why + ")");
  const int NS = h->params.num_simulations;
  std::string err;
  for (int sim = 0; sim < NS; ++sim) {
    SelectIO sio{h->sel_parent, h->sel_action, h->sel_next, nullptr, nullptr};
    MZ_DISPATCH_G(h, select_kernel, tree_blocks(h), stream, h->tree, h->params, sim, sio);
    if (recurrent_tc_launch(h->rtc, h->net, h->tree, h->sel_parent, h->sel_action, h->rec_reward, h->rec_value,
                            h->rec_logits, h->rec_emb, stream, &h->launches, &err))
      return fail(err);
    ExpandIO eio{h->sel_parent, h->sel_action, h->sel_next, h->rec_reward, nullptr, h->rec_value, h->rec_logits,
                 h->rec_emb};
    MZ_DISPATCH_G(h, expand_backup_kernel, tree_blocks(h), stream, h->tree, h->params, eio);
  }
  MZ_CUDA(cudaGetLastError());
  return 0;
}

static int run_stepwise(mz_handle* h, cudaStream_t stream, int precision) {
  if (h->weights == nullptr) return fail("mz_set_weights has not been called");
  if (precision == MZ_PRECISION_BF16) return run_stepwise_tc(h, stream);
  const int NS = h->params.num_simulations;
  const int grid_mlp = (h->cfg.batch + kMlpRows - 1) / kMlpRows;
  for (int sim = 0; sim < NS; ++sim) {
    SelectIO sio{h->sel_parent, h->sel_action, h->sel_next, nullptr, nullptr};
    MZ_DISPATCH_G(h, select_kernel, tree_blocks(h), stream, h->tree, h->params, sim, sio);
    RecurrentIO rio{h->sel_parent, h->sel_action, h->rec_reward, h->rec_value, h->rec_logits, h->rec_emb};
    recurrent_kernel<<<grid_mlp, kMlpThreads, mlp_smem_bytes(h->net), stream>>>(h->net, h->weights, h->tree, rio);
    h->launches += 1;
    ExpandIO eio{h->sel_parent, h->sel_action, h->sel_next, h->rec_reward, nullptr, h->rec_value, h->rec_logits,
                 h->rec_emb};
    MZ_DISPATCH_G(h, expand_backup_kernel, tree_blocks(h), stream, h->tree, h->params, eio);
  }
  MZ_CUDA(cudaGetLastError());
  return 0;
}

static int launch_finish(mz_handle* h, int32_t* action_out, float* weights_out, cudaStream_t stream) {
  MZ_DISPATCH_G(h, finish_kernel, tree_blocks(h), stream, h->tree, h->params, h->has_invalid, action_out, weights_out);
  MZ_CUDA(cudaGetLastError());
  return 0;
}

static int search_device(mz_handle* h, const float* obs, const float* root_logits, const float* root_value,
                         const float* root_emb, const uint8_t* invalid, const float* noise, const mz_search_args* args,
                         int32_t* action_out, float* weights_out, float* root_value_out, cudaStream_t stream) {
  if (int rc = check_args(h, args)) return rc;
  if (action_out == nullptr || weights_out == nullptr) return fail_arg("action / action_weights outputs are required");
  if (obs == nullptr && root_emb == nullptr)
    return fail_arg("give obs, or root_emb (optionally with root_logits + root_value)");
  if (obs == nullptr && ((root_logits == nullptr) != (root_value == nullptr)))
    return fail_arg("root_logits and root_value must be given together");
  if (obs != nullptr && h->cfg.obs_dim <= 0)
    return fail_arg("handle was created with obs_dim = 0: supply the root embedding instead of obs");
  if (args->num_decision_actions > 0)
    return fail_arg("stochastic MuZero (num_decision_actions > 0) runs in the callback mode: mz_begin / mz_select / "
                    "mz_expand_backup / mz_finish");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  if (int rc = stage_keys(h, args, stream)) return rc;
  h->resident.dirty = false;  // whatever runs next owns the SoA tree view
  h->tree_valid = false;
  h->tree16 = false;
  const bool want_tree = (args->flags & MZ_FLAG_WANT_TREE) != 0;
  MZ_CUDA(cudaEventRecord(h->ev_start, stream));
  int engine = args->engine;
  const bool have_w = h->weights != nullptr;
  const int B = h->cfg.batch, NS = h->params.num_simulations;
  const bool warp_ok = have_w && obs != nullptr && warp_supported(h->warpeng, h->params, B);
  const bool treewarp_ok = have_w && treewarp_supported(h->treewarp, h->net, B, NS, h->params.max_depth);
  const bool resident_ok = have_w && resident_supported(h->resident, h->net, B, NS);
  if (args->precision == MZ_PRECISION_BF16) {
    // throughput mode: tree-warp select / backup kernels around the tcgen05 recurrent kernel (search_tc); STEPWISE
    // keeps the generic SoA tree kernels around the same recurrent kernel
    if (engine != MZ_ENGINE_AUTO && engine != MZ_ENGINE_STEPWISE)
      return fail_arg("precision = bf16: engine must be AUTO or STEPWISE");
    if (engine == MZ_ENGINE_AUTO) {
      h->has_invalid = invalid != nullptr;
      if (int rc = search_tc(h, obs, root_logits, root_value, root_emb, invalid, noise, action_out, weights_out,
                             root_value_out, stream))
        return rc;
      h->tree_valid = true;  // records, unpacked lazily by mz_get_tree
      MZ_CUDA(cudaEventRecord(h->ev_stop, stream));
      h->timed = true;
      return 0;
    }
  }
  // AUTO: the warp engine when its compile-time shapes match and the trees fit on chip; else the tree-warp engine
  // (weights in shared memory, trees as records in L1/L2); else the CTA-resident engine (weights streamed); the
  // stepwise engine remains for the callback mode, the bf16 throughput mode and as the reference implementation.
  if (engine == MZ_ENGINE_AUTO || engine == MZ_ENGINE_FUSED)
    engine = warp_ok ? MZ_ENGINE_FUSED_WARP
                     : (treewarp_ok ? MZ_ENGINE_TREEWARP : (resident_ok ? MZ_ENGINE_RESIDENT : MZ_ENGINE_STEPWISE));
  if ((!h->peer_deltas.empty() || h->warpeng.flag != nullptr) && engine != MZ_ENGINE_FUSED_WARP)
    return fail("peer outputs (mz_set_peer_outputs) are written by the warp engine only; this configuration runs on "
                "another engine — clear them and exchange the outputs with a collective");
  h->has_invalid = invalid != nullptr;
  std::string err;
  if (engine == MZ_ENGINE_RESIDENT || engine == MZ_ENGINE_TREEWARP) {
    if (engine == MZ_ENGINE_RESIDENT ? !resident_ok : !treewarp_ok)
      return fail("this one-launch engine does not support this configuration (see DESIGN.md)");
    if (ensure_device_keys(h, stream)) return 1;
    const int rc =
        engine == MZ_ENGINE_RESIDENT
            ? resident_launch(h->resident, h->net, h->weights, h->tree, h->params, obs, root_emb, root_logits,
                              root_value, invalid, noise, action_out, weights_out, root_value_out, stream, &h->launches,
                              &err)
            : treewarp_launch(h->treewarp, h->resident, h->net, h->weights, h->tree, h->params, obs, root_emb,
                              root_logits, root_value, invalid, noise, action_out, weights_out, root_value_out, stream,
                              &h->launches, &err);
    if (rc) return fail(err);
    h->tree_valid = true;  // the records are unpacked lazily by mz_get_tree
  } else if (engine == MZ_ENGINE_FUSED_WARP) {
    if (!warp_ok)
      return fail(std::string("this one-launch engine does not support this configuration (see DESIGN.md)") +
                  (!h->warpeng.available ? " [warp engine: no compiled shape variant]"
                                         : " [warp engine: policy / qtransform / root mode / size]"));
    if (want_tree && args->num_simulations + 1 < h->N && clear_tree(h, args->num_simulations, true, stream)) return 1;
    const bool inline_keys = NS <= kInlineSims && h->warpeng.producers > 0;
    if (!inline_keys && ensure_device_keys(h, stream)) return 1;
    if (warp_launch(h->warpeng, h->tree, h->params, inline_keys ? &h->host_keys : nullptr, obs, invalid, noise,
                    action_out, weights_out, root_value_out, h->peer_deltas, want_tree, stream, &h->launches, &err))
      return fail(err);
    h->tree_valid = want_tree;
  } else if (engine == MZ_ENGINE_STEPWISE) {
    if (obs != nullptr || root_logits == nullptr) {  // Prediction (and Representation) run in the library
      if (launch_root(h, obs, obs != nullptr ? nullptr : root_emb, stream)) return 1;
      root_logits = h->root_logits;
      root_value = h->root_value;
      root_emb = h->root_emb;
    }
    if (root_value_out != nullptr)
      MZ_CUDA(cudaMemcpyAsync(root_value_out, root_value, sizeof(float) * h->cfg.batch, cudaMemcpyDefault,
                              stream));
    if (launch_begin(h, root_logits, root_value, root_emb, invalid, noise, stream)) return 1;
    if (run_stepwise(h, stream, args->precision)) return 1;
    if (launch_finish(h, action_out, weights_out, stream)) return 1;
    h->tree_valid = true;
  } else {
    return fail_arg("unknown engine");
  }
  MZ_CUDA(cudaEventRecord(h->ev_stop, stream));
  h->timed = true;
  return 0;
}

}  // namespace mz

// ------------------------------------------------------------------------------------------ C ABI

extern "C" {

const char* mz_last_error(void) { return mz::g_last_error.c_str(); }

void mz_default_args(mz_search_args* a) {
  if (a == nullptr) return;
  std::memset(a, 0, sizeof(*a));
  a->policy = MZ_POLICY_MUZERO;
  a->qtransform = MZ_QTRANSFORM_BY_PARENT_AND_SIBLINGS;
  a->num_simulations = 5;  // muax/model.py:86
  a->max_depth = 0;
  a->max_considered = 16;
  a->temperature = 1.0f;
  a->dirichlet_fraction = 0.25f;
  a->dirichlet_alpha = 0.3f;
  a->pb_c_init = 1.25f;
  a->pb_c_base = 19652.0f;
  a->gumbel_scale = 1.0f;
  a->value_scale = 0.1f;
  a->maxvisit_init = 50.0f;
  a->engine = MZ_ENGINE_AUTO;
  a->flags = 0;  // no tree view unless asked for (MZ_FLAG_WANT_TREE): muax never reads PolicyOutput.search_tree
  a->precision = MZ_PRECISION_FP32;
  a->num_decision_actions = 0;
}

int mz_create(mz_handle** out, const mz_config* cfg) {
  using namespace mz;
  if (out == nullptr || cfg == nullptr) return fail_arg("mz_create: NULL argument");
  *out = nullptr;
  if (cfg->batch < 1) return fail_arg("batch must be >= 1");
  if (cfg->num_actions < 1 || cfg->num_actions > MZ_MAX_ACTIONS) return fail_arg("num_actions must be in 1..32");
  if (cfg->embed_dim < 1) return fail_arg("embed_dim must be >= 1");
  if (cfg->support_size < 0) return fail_arg("support_size must be >= 0");
  if (cfg->max_num_simulations < 0) return fail_arg("max_num_simulations must be >= 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: libmzsearch has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail_arg("device ordinal out of range");
  const int F = 2 * cfg->support_size + 1;
  size_t need = 0;
  if (cfg->obs_dim > 0)
    if (int rc = validate_stack(cfg->repr, cfg->obs_dim, cfg->embed_dim, "repr", &need)) return rc;
  if (int rc = validate_stack(cfg->pred_v, cfg->embed_dim, F, "pred_v", &need)) return rc;
  if (int rc = validate_stack(cfg->pred_pi, cfg->embed_dim, cfg->num_actions, "pred_pi", &need)) return rc;
  if (int rc = validate_stack(cfg->dyn_ns, cfg->embed_dim + cfg->num_actions, cfg->embed_dim, "dyn_ns", &need)) return rc;
  if (int rc = validate_stack(cfg->dyn_r, cfg->embed_dim + cfg->num_actions, F, "dyn_r", &need)) return rc;
  MZ_CUDA(cudaSetDevice(cfg->device));
  mz_handle* h = new mz_handle();
  h->cfg = *cfg;
  h->n_weights = need;
  h->N = cfg->max_num_simulations + 1;
  int G = 2;
  while (G < cfg->num_actions) G <<= 1;
  h->G = G;
  Net& net = h->net;
  net.repr = cfg->repr;
  net.pred_v = cfg->pred_v;
  net.pred_pi = cfg->pred_pi;
  net.dyn_ns = cfg->dyn_ns;
  net.dyn_r = cfg->dyn_r;
  net.activation = cfg->activation;
  net.repr_minmax = cfg->repr_minmax;
  net.dyn_minmax = cfg->dyn_minmax;
  net.support_size = cfg->support_size;
  net.obs_dim = cfg->obs_dim;
  net.embed_dim = cfg->embed_dim;
  net.num_actions = cfg->num_actions;
  int mw = std::max(std::max(cfg->embed_dim + cfg->num_actions, F), cfg->obs_dim);
  if (cfg->obs_dim > 0) mw = stack_max_width(cfg->repr, mw);
  mw = stack_max_width(cfg->pred_v, mw);
  mw = stack_max_width(cfg->pred_pi, mw);
  mw = stack_max_width(cfg->dyn_ns, mw);
  mw = stack_max_width(cfg->dyn_r, mw);
  net.max_width = mw;
  const size_t smem = mlp_smem_bytes(net);
  if (smem > 200 * 1024) {
    delete h;
    return fail_arg("layer width too large for the MLP kernels' shared-memory staging");
  }
  auto bail = [&](int rc) {
    if (rc) mz_destroy(h);
    return rc;
  };
#define MZ_TRY(x) \
  if (bail(x)) return 1
  MZ_TRY((cudaFuncSetAttribute(root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
             ? fail("cudaFuncSetAttribute(root_kernel) failed")
             : 0);
  MZ_TRY((cudaFuncSetAttribute(recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess)
             ? fail("cudaFuncSetAttribute(recurrent_kernel) failed")
             : 0);
  Tree& t = h->tree;
  t.B = cfg->batch;
  t.N = h->N;
  t.A = cfg->num_actions;
  t.E = cfg->embed_dim;
  const size_t B = cfg->batch, BN = B * h->N, BNA = BN * cfg->num_actions, BA = B * cfg->num_actions;
  MZ_TRY(dev_alloc(h, &t.node_visits, BN));
  MZ_TRY(dev_alloc(h, &t.parents, BN));
  MZ_TRY(dev_alloc(h, &t.action_from_parent, BN));
  MZ_TRY(dev_alloc(h, &t.raw_values, BN));
  MZ_TRY(dev_alloc(h, &t.node_values, BN));
  MZ_TRY(dev_alloc(h, &t.children_index, BNA));
  MZ_TRY(dev_alloc(h, &t.children_visits, BNA));
  MZ_TRY(dev_alloc(h, &t.children_prior_logits, BNA));
  MZ_TRY(dev_alloc(h, &t.children_prior_probs, BNA));
  MZ_TRY(dev_alloc(h, &t.children_values, BNA));
  MZ_TRY(dev_alloc(h, &t.children_rewards, BNA));
  MZ_TRY(dev_alloc(h, &t.children_discounts, BNA));
  MZ_TRY(dev_alloc(h, &t.embeddings, BN * cfg->embed_dim));
  MZ_TRY(dev_alloc(h, &t.root_noise, BA));
  MZ_TRY(dev_alloc(h, &t.root_invalid, BA));
  MZ_TRY(dev_alloc(h, &t.sim_depth, B * std::max(cfg->max_num_simulations, 1)));
  MZ_TRY(dev_alloc(h, &h->sel_parent, B));
  MZ_TRY(dev_alloc(h, &h->sel_action, B));
  MZ_TRY(dev_alloc(h, &h->sel_next, B));
  MZ_TRY(dev_alloc(h, &h->sel5, 5 * B));
  MZ_TRY(dev_alloc(h, &h->rec_reward, B));
  MZ_TRY(dev_alloc(h, &h->rec_value, B));
  MZ_TRY(dev_alloc(h, &h->rec_logits, BA));
  MZ_TRY(dev_alloc(h, &h->rec_emb, B * cfg->embed_dim));
  MZ_TRY(dev_alloc(h, &h->root_logits, BA));
  MZ_TRY(dev_alloc(h, &h->root_value, B));
  MZ_TRY(dev_alloc(h, &h->root_emb, B * cfg->embed_dim));
  MZ_TRY(dev_alloc(h, &h->sim_keys_dev, (size_t)2 * (cfg->max_num_simulations + 1)));
  MZ_TRY(dev_alloc(h, &h->d_obs, B * std::max(cfg->obs_dim, 1)));
  MZ_TRY(dev_alloc(h, &h->d_noise, BA));
  MZ_TRY(dev_alloc(h, &h->d_invalid, BA));
  MZ_TRY(dev_alloc(h, &h->d_action_out, B));
  MZ_TRY(dev_alloc(h, &h->d_weights_out, BA));
  MZ_TRY(dev_alloc(h, &h->d_value_out, B));
  auto host_alloc = [&](void** p, size_t bytes) -> int {
    MZ_CUDA(cudaMallocHost(p, bytes + 16));
    return 0;
  };
  MZ_TRY(host_alloc((void**)&h->key_slots,
                    sizeof(uint32_t) * 2 * (cfg->max_num_simulations + 1) * mz_handle::kSlots));
  MZ_TRY(host_alloc((void**)&h->h_obs, sizeof(float) * B * std::max(cfg->obs_dim, 1)));
  MZ_TRY(host_alloc((void**)&h->h_noise, sizeof(float) * BA));
  MZ_TRY(host_alloc((void**)&h->h_invalid, BA));
  MZ_TRY(host_alloc((void**)&h->h_action_out, sizeof(int32_t) * B));
  MZ_TRY(host_alloc((void**)&h->h_weights_out, sizeof(float) * BA));
  MZ_TRY(host_alloc((void**)&h->h_value_out, sizeof(float) * B));
  for (int i = 0; i < mz_handle::kSlots; ++i)
    MZ_TRY(cudaEventCreateWithFlags(&h->slot_done[i], cudaEventDisableTiming) != cudaSuccess
               ? fail("cudaEventCreate failed")
               : 0);
  MZ_TRY(cudaEventCreate(&h->ev_start) != cudaSuccess ? fail("cudaEventCreate failed") : 0);
  MZ_TRY(cudaEventCreate(&h->ev_stop) != cudaSuccess ? fail("cudaEventCreate failed") : 0);
  {
    std::string err;
    if (warp_init(h->warpeng, h->net, cfg->device, &err) || treewarp_init(h->treewarp, h->net, cfg->device, &err) ||
        resident_init(h->resident, h->net, cfg->device, &err) ||
        recurrent_tc_init(h->rtc, h->net, cfg->batch, cfg->device, &err)) {
      mz_destroy(h);
      return fail(err);
    }
  }
#undef MZ_TRY
  *out = h;
  return 0;
}

int mz_destroy(mz_handle* h) {
  if (h == nullptr) return 0;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  mz::warp_destroy(h->warpeng);
  mz::resident_destroy(h->resident);
  mz::treewarp_destroy(h->treewarp);
  mz::recurrent_tc_destroy(h->rtc);
  for (void* p : h->allocs) cudaFree(p);
  if (h->weights) cudaFree(h->weights);
  if (h->table_dev) cudaFree(h->table_dev);
  void* pinned[] = {h->key_slots, h->h_obs, h->h_noise, h->h_invalid, h->h_action_out, h->h_weights_out,
                    h->h_value_out};
  for (void* p : pinned)
    if (p) cudaFreeHost(p);
  for (auto& e : h->slot_done)
    if (e) cudaEventDestroy(e);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->ev_stop) cudaEventDestroy(h->ev_stop);
  delete h;
  return 0;
}

int mz_set_weights(mz_handle* h, const float* blob, size_t n_floats, int on_device, void* stream) {
  using namespace mz;
  if (h == nullptr || blob == nullptr) return fail_arg("mz_set_weights: NULL argument");
  if (n_floats < h->n_weights) return fail_arg("weight blob is smaller than the layer offsets require");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (h->weights == nullptr) MZ_CUDA(cudaMalloc((void**)&h->weights, std::max(h->n_weights, (size_t)4) * sizeof(float) + 16));
  MZ_CUDA(cudaMemcpyAsync(h->weights, blob, h->n_weights * sizeof(float),
                          on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (warp_pack(h->warpeng, h->weights, s, &h->launches) ||
      recurrent_tc_pack(h->rtc, h->net, h->weights, s, &h->launches))
    return fail("re-packing the weights for the warp engine / the tcgen05 recurrent kernel failed");
  if (!on_device) MZ_CUDA(cudaStreamSynchronize(s));  // the host blob may be pageable and short-lived
  return 0;
}

int mz_search(mz_handle* h, const float* obs_dev, const float* root_logits_dev, const float* root_value_dev,
              const float* root_emb_dev, const uint8_t* invalid_dev, const float* noise_dev,
              const mz_search_args* args, int32_t* action_out_dev, float* action_weights_out_dev,
              float* root_value_out_dev, void* stream) {
  if (h == nullptr) return mz::fail_arg("mz_search: NULL handle");
  return mz::search_device(h, obs_dev, root_logits_dev, root_value_dev, root_emb_dev, invalid_dev, noise_dev, args,
                           action_out_dev, action_weights_out_dev, root_value_out_dev, (cudaStream_t)stream);
}

int mz_set_peer_outputs(mz_handle* h, int32_t n, const int64_t* byte_deltas) {
  using namespace mz;
  if (h == nullptr || n < 0 || n > 7 || (n > 0 && byte_deltas == nullptr))
    return fail_arg("mz_set_peer_outputs: expected 0..7 byte offsets");
  h->peer_deltas.assign(byte_deltas, byte_deltas + n);
  return 0;
}

int mz_set_peer_flags(mz_handle* h, int32_t* flags_dev, int32_t rank, int32_t step) {
  using namespace mz;
  if (h == nullptr || (flags_dev != nullptr && (rank < 0 || rank > 7))) return fail_arg("mz_set_peer_flags: bad argument");
  h->warpeng.flag = flags_dev;
  h->warpeng.flag_rank = rank;
  h->warpeng.flag_step = step;
  return 0;
}

int mz_peer_wait(mz_handle* h, const int32_t* flags_dev, int32_t world, int32_t step, void* stream) {
  using namespace mz;
  if (h == nullptr || flags_dev == nullptr || world < 1 || world > 8) return fail_arg("mz_peer_wait: bad argument");
  if (!h->warpeng.available) return fail("mz_peer_wait: the warp engine is unavailable for this handle");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  if (warp_peer_wait(h->warpeng, flags_dev, world, step, (cudaStream_t)stream, &h->launches))
    return fail("mz_peer_wait: launch failed");
  return 0;
}

int mz_search_host(mz_handle* h, const float* obs_host, const uint8_t* invalid_host, const float* noise_host,
                   const mz_search_args* args, int32_t* action_out_host, float* action_weights_out_host,
                   float* root_value_out_host, void* stream) {
  using namespace mz;
  if (h == nullptr || obs_host == nullptr) return fail_arg("mz_search_host: NULL argument");
  if (action_out_host == nullptr) return fail_arg("mz_search_host: action output is required");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t B = h->cfg.batch, BA = B * h->cfg.num_actions, BO = B * h->cfg.obs_dim;
  // Zero-copy host I/O: the pinned staging buffers are mapped into the device's address space (unified addressing),
  // so the kernels write action / action_weights / root_value straight into them — no D2H copy commands — and small
  // observation batches are read in place by the search kernel's prologue instead of an H2D copy (four copy commands
  // of ~7 us each were 10 % of an act at the headline shapes).  Large observation batches still go through the copy
  // engine.  MZ_HOST_COPIES=1 restores the explicit copies.
  static const bool explicit_copies = getenv("MZ_HOST_COPIES") != nullptr && atoi(getenv("MZ_HOST_COPIES")) != 0;
  const bool obs_in_place = !explicit_copies && BO * sizeof(float) <= (size_t)256 * 1024;
  std::memcpy(h->h_obs, obs_host, BO * sizeof(float));
  if (!obs_in_place) MZ_CUDA(cudaMemcpyAsync(h->d_obs, h->h_obs, BO * sizeof(float), cudaMemcpyHostToDevice, s));
  if (invalid_host != nullptr) {
    std::memcpy(h->h_invalid, invalid_host, BA);
    MZ_CUDA(cudaMemcpyAsync(h->d_invalid, h->h_invalid, BA, cudaMemcpyHostToDevice, s));
  }
  if (noise_host != nullptr) {
    std::memcpy(h->h_noise, noise_host, BA * sizeof(float));
    MZ_CUDA(cudaMemcpyAsync(h->d_noise, h->h_noise, BA * sizeof(float), cudaMemcpyHostToDevice, s));
  }
  if (int rc = search_device(h, obs_in_place ? h->h_obs : h->d_obs, nullptr, nullptr, nullptr,
                             invalid_host ? h->d_invalid : nullptr, noise_host ? h->d_noise : nullptr, args,
                             explicit_copies ? h->d_action_out : h->h_action_out,
                             explicit_copies ? h->d_weights_out : h->h_weights_out,
                             explicit_copies ? h->d_value_out : h->h_value_out, s))
    return rc;
  if (explicit_copies) {
    MZ_CUDA(cudaMemcpyAsync(h->h_action_out, h->d_action_out, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    MZ_CUDA(cudaMemcpyAsync(h->h_weights_out, h->d_weights_out, BA * sizeof(float), cudaMemcpyDeviceToHost, s));
    MZ_CUDA(cudaMemcpyAsync(h->h_value_out, h->d_value_out, B * sizeof(float), cudaMemcpyDeviceToHost, s));
  }
  MZ_CUDA(cudaStreamSynchronize(s));
  std::memcpy(action_out_host, h->h_action_out, B * sizeof(int32_t));
  if (action_weights_out_host) std::memcpy(action_weights_out_host, h->h_weights_out, BA * sizeof(float));
  if (root_value_out_host) std::memcpy(root_value_out_host, h->h_value_out, B * sizeof(float));
  return 0;
}

int mz_recurrent(mz_handle* h, const int32_t* action_dev, const float* embedding_dev, int32_t precision,
                 float* reward_out_dev, float* value_out_dev, float* prior_logits_out_dev, float* next_embedding_out_dev,
                 void* stream) {
  using namespace mz;
  if (h == nullptr || action_dev == nullptr || embedding_dev == nullptr || reward_out_dev == nullptr ||
      value_out_dev == nullptr || prior_logits_out_dev == nullptr || next_embedding_out_dev == nullptr)
    return fail_arg("mz_recurrent: NULL argument");
  if (precision != MZ_PRECISION_FP32 && precision != MZ_PRECISION_BF16) return fail_arg("unknown precision");
  if (h->weights == nullptr) return fail("mz_set_weights has not been called");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  // the kernels gather embeddings[b, parent[b]] from a tree: present the batch as B one-node trees with parent = 0
  Tree view = h->tree;
  view.embeddings = const_cast<float*>(embedding_dev);
  view.N = 1;
  MZ_CUDA(cudaMemsetAsync(h->sel_parent, 0, sizeof(int32_t) * h->cfg.batch, s));
  if (precision == MZ_PRECISION_BF16) {
    if (!h->rtc.available)
      return fail("precision = bf16: the tcgen05 recurrent kernel does not cover this network (" + h->rtc.why + ")");
    std::string err;
    if (recurrent_tc_launch(h->rtc, h->net, view, h->sel_parent, action_dev, reward_out_dev, value_out_dev,
                            prior_logits_out_dev, next_embedding_out_dev, s, &h->launches, &err))
      return fail(err);
  } else {
    RecurrentIO rio{h->sel_parent, action_dev, reward_out_dev, value_out_dev, prior_logits_out_dev, next_embedding_out_dev};
    const int grid_mlp = (h->cfg.batch + kMlpRows - 1) / kMlpRows;
    recurrent_kernel<<<grid_mlp, kMlpThreads, mlp_smem_bytes(h->net), s>>>(h->net, h->weights, view, rio);
    h->launches += 1;
  }
  MZ_CUDA(cudaGetLastError());
  return 0;
}

int mz_begin(mz_handle* h, const float* root_logits_dev, const float* root_value_dev, const float* root_emb_dev,
             const uint8_t* invalid_dev, const float* noise_dev, const mz_search_args* args, void* stream) {
  using namespace mz;
  if (h == nullptr || root_logits_dev == nullptr || root_value_dev == nullptr || root_emb_dev == nullptr)
    return fail_arg("mz_begin: NULL argument");
  if (int rc = check_args(h, args)) return rc;
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  if (stage_keys(h, args, (cudaStream_t)stream)) return 1;
  h->resident.dirty = false;
  h->tree_valid = true;
  return launch_begin(h, root_logits_dev, root_value_dev, root_emb_dev, invalid_dev, noise_dev, (cudaStream_t)stream);
}

int mz_select(mz_handle* h, int32_t sim, int32_t* action_out_dev, float* parent_emb_out_dev, void* stream) {
  using namespace mz;
  if (h == nullptr || action_out_dev == nullptr || parent_emb_out_dev == nullptr)
    return fail_arg("mz_select: NULL argument");
  if (sim < 0 || sim >= h->params.num_simulations) return fail_arg("mz_select: sim out of range");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  SelectIO sio{h->sel_parent, h->sel_action, h->sel_next, action_out_dev, parent_emb_out_dev};
  MZ_DISPATCH_G(h, select_kernel, tree_blocks(h), (cudaStream_t)stream, h->tree, h->params, (int)sim, sio);
  MZ_CUDA(cudaGetLastError());
  return 0;
}

int mz_expand_backup(mz_handle* h, int32_t sim, const float* reward_dev, const float* discount_dev,
                     const float* prior_logits_dev, const float* value_dev, const float* next_emb_dev, void* stream) {
  using namespace mz;
  if (h == nullptr || reward_dev == nullptr || prior_logits_dev == nullptr || value_dev == nullptr ||
      next_emb_dev == nullptr)
    return fail_arg("mz_expand_backup: NULL argument");
  if (sim < 0 || sim >= h->params.num_simulations) return fail_arg("mz_expand_backup: sim out of range");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  ExpandIO eio{h->sel_parent, h->sel_action, h->sel_next, reward_dev, discount_dev, value_dev, prior_logits_dev,
               next_emb_dev};
  MZ_DISPATCH_G(h, expand_backup_kernel, tree_blocks(h), (cudaStream_t)stream, h->tree, h->params, eio);
  MZ_CUDA(cudaGetLastError());
  return 0;
}

int mz_finish(mz_handle* h, int32_t* action_out_dev, float* action_weights_out_dev, void* stream) {
  using namespace mz;
  if (h == nullptr || action_out_dev == nullptr || action_weights_out_dev == nullptr)
    return fail_arg("mz_finish: NULL argument");
  MZ_CUDA(cudaSetDevice(h->cfg.device));
  return launch_finish(h, action_out_dev, action_weights_out_dev, (cudaStream_t)stream);
}

int mz_get_tree(mz_handle* h, mz_tree_view* v) {
  if (h == nullptr || v == nullptr) return mz::fail_arg("mz_get_tree: NULL argument");
  if (!h->tree_valid)
    return mz::fail("mz_get_tree: the last search did not keep its tree (set MZ_FLAG_WANT_TREE in mz_search_args.flags)");
  const mz::Tree& t = h->tree;
  {  // the CTA-resident engine keeps packed records; the mctx SoA view is produced when somebody asks for it
    std::string err;
    if (mz::resident_unpack(h->resident, t, h->cfg.discount, &err)) return mz::fail(err);
    if (h->tree16) {  // throughput mode: the embeddings live in bf16 rows
      if (mz::recurrent_tc_tree_export(h->rtc, h->net, t.B, t.N, t.embeddings, h->resident.last_stream, &err))
        return mz::fail(err);
      h->tree16 = false;  // exported once; the fp32 view stays valid until the next search
    }
  }
  v->batch = t.B;
  v->num_nodes = t.N;
  v->num_actions = t.A;
  v->embed_dim = t.E;
  v->node_visits = t.node_visits;
  v->parents = t.parents;
  v->action_from_parent = t.action_from_parent;
  v->children_index = t.children_index;
  v->children_visits = t.children_visits;
  v->raw_values = t.raw_values;
  v->node_values = t.node_values;
  v->children_prior_logits = t.children_prior_logits;
  v->children_values = t.children_values;
  v->children_rewards = t.children_rewards;
  v->children_discounts = t.children_discounts;
  v->embeddings = t.embeddings;
  v->root_noise = t.root_noise;
  v->sim_depth = t.sim_depth;
  return 0;
}

int mz_launch_count(mz_handle* h, int64_t* count) {
  if (h == nullptr || count == nullptr) return mz::fail_arg("mz_launch_count: NULL argument");
  *count = h->launches;
  return 0;
}

int mz_last_kernel_ms(mz_handle* h, float* ms) {
  using namespace mz;
  if (h == nullptr || ms == nullptr) return fail_arg("mz_last_kernel_ms: NULL argument");
  if (!h->timed) return fail("no search has been timed yet");
  MZ_CUDA(cudaEventSynchronize(h->ev_stop));
  MZ_CUDA(cudaEventElapsedTime(ms, h->ev_start, h->ev_stop));
  return 0;
}

int mz_math_probe(int32_t kind, const float* x_dev, float* y_dev, int64_t n, void* stream) {
  using namespace mz;
  if (x_dev == nullptr || y_dev == nullptr || n < 0) return fail_arg("mz_math_probe: bad argument");
  if (n == 0) return 0;
  math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, x_dev, y_dev, (long)n);
  MZ_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

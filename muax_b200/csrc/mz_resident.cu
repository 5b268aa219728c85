// mz_resident.cu — CTA-resident engine (interface and rationale: mz_resident.cuh).
//
// Per simulation a CTA runs, for its T trees:
//   A  select        one lane group (G lanes) per tree, groups dealt round-robin to the warps so that trees do not
//                    serialise behind each other; one memory round trip per level (the child index travels with the
//                    child row), tie-break noise from a table produced ahead of the search (the jax key chain depends
//                    only on key / global row / simulation / depth), sqrt(n)*pb_c(n) from a shared-memory table; the
//                    selected path is recorded in shared memory
//   B  Dynamic       all threads: (row tile, output unit) work items, k-unrolled FMA chains in haiku's order
//   C  min-max + reward support transform   one warp per row
//   D  Prediction    as B
//   F  value support transform (warp of the tree) + expand + backup along the recorded path with the next level's
//                    operands prefetched while the current level's mean update is computed
// Arithmetic and orders are the shared device functions of mz_device.cuh / mz_math.h: bit-identical to the other
// engines and to the CPU checkers.
#include "mz_resident.cuh"

#include <algorithm>
#include <cstdlib>


namespace mz {

// ------------------------------------------------------------------------------------------ shared-memory accessors
// The dense layers live in __noinline__ functions (one copy per (weights location, row tile), shared by every lane
// group instantiation of the kernel); explicit shared-space loads keep them LDS/STS instead of generic accesses.

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------------ dense layers

struct DenseJob {    // one hk.Linear applied to R rows staged in shared memory
  const float* W;    // [nin (+ one-hot rows)][nout], shared memory (kLdg = false) or global (kLdg = true)
  const float* bias; // [nout], same space as W
  const float* src;  // shared, row-major, row stride lds (floats, multiple of 4)
  float* dst;        // shared, row stride ldd
  int32_t nin, nout, lds, ldd;
};

// Two layers that run in the same phase (the two heads of a module); B.nout == 0 for a single layer.
// Work item = (tile of RT rows, output unit j); y[r][j] = (sum_k fma(x[r][k], W[k][j])) (+ W[nin + onehot[r]][j]) + b[j]
// with k ascending — the accumulation order of the CPU checkers.  RT rows share every weight load.
template <bool kLdg, int RT>
__device__ __noinline__ void dense_pair(const DenseJob A, const DenseJob B, int R, const int32_t* onehot, int act_kind,
                                        int apply_act) {
  const int na = A.nout, ntot = na + B.nout;
  const int tiles = (R + RT - 1) / RT;
  for (int item = threadIdx.x; item < tiles * ntot; item += blockDim.x) {
    const int tile = item / ntot;
    int j = item - tile * ntot;
    const bool second = j >= na;
    if (second) j -= na;
    const DenseJob& J = second ? B : A;
    const int nin = J.nin, nout = J.nout;
    const int r0 = tile * RT;
    uint32_t xa[RT];
    float acc[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      xa[i] = smem_u32(J.src + min(r0 + i, R - 1) * J.lds);
      acc[i] = 0.0f;
    }
    const float* Wg = J.W + j;
    const uint32_t Ws = kLdg ? 0u : smem_u32(Wg);
    int k = 0;
#pragma unroll 1
    for (; k + 4 <= nin; k += 4) {
      float w0, w1, w2, w3;
      if constexpr (kLdg) {
        w0 = __ldg(Wg + (size_t)k * nout);
        w1 = __ldg(Wg + (size_t)(k + 1) * nout);
        w2 = __ldg(Wg + (size_t)(k + 2) * nout);
        w3 = __ldg(Wg + (size_t)(k + 3) * nout);
      } else {
        w0 = lds_f32(Ws + (uint32_t)(k * nout) * 4u);
        w1 = lds_f32(Ws + (uint32_t)((k + 1) * nout) * 4u);
        w2 = lds_f32(Ws + (uint32_t)((k + 2) * nout) * 4u);
        w3 = lds_f32(Ws + (uint32_t)((k + 3) * nout) * 4u);
      }
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float4 xv = lds_v4(xa[i] + (uint32_t)k * 4u);
        acc[i] = MZ_FMA(xv.x, w0, acc[i]);
        acc[i] = MZ_FMA(xv.y, w1, acc[i]);
        acc[i] = MZ_FMA(xv.z, w2, acc[i]);
        acc[i] = MZ_FMA(xv.w, w3, acc[i]);
      }
    }
    for (; k < nin; ++k) {
      const float wk = kLdg ? __ldg(Wg + (size_t)k * nout) : lds_f32(Ws + (uint32_t)(k * nout) * 4u);
#pragma unroll
      for (int i = 0; i < RT; ++i) acc[i] = MZ_FMA(lds_f32(xa[i] + (uint32_t)k * 4u), wk, acc[i]);
    }
    if (onehot != nullptr) {  // [x, one_hot(action)] @ W = x @ W[:nin] + W[nin + action]  (muax/nn.py:105-108)
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const int row = nin + onehot[min(r0 + i, R - 1)];
        const float wo = kLdg ? __ldg(Wg + (size_t)row * nout) : lds_f32(Ws + (uint32_t)(row * nout) * 4u);
        acc[i] = MZ_ADD(acc[i], wo);
      }
    }
    const float bj = kLdg ? __ldg(J.bias + j) : lds_f32(smem_u32(J.bias + j));
    const uint32_t da = smem_u32(J.dst + j);
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      if (r0 + i < R) {
        float y = MZ_ADD(acc[i], bj);
        if (apply_act) y = activate(y, act_kind);
        sts_f32(da + (uint32_t)((r0 + i) * J.ldd) * 4u, y);
      }
    }
  }
}

// Smallest row tile that covers the layer in one pass of the CTA (more rows per thread = fewer weight loads, fewer
// threads busy): parallelism first, register tiling only once every thread has work.
template <bool kLdg>
__device__ __forceinline__ void dense_pair_auto(const DenseJob& A, const DenseJob& B, int R, const int32_t* onehot,
                                                int act_kind, int apply_act) {
  const int ntot = A.nout + B.nout;
  int rt = 1;
  while (rt < 4 && ((R + rt - 1) / rt) * ntot > (int)blockDim.x) rt <<= 1;  // 4 rows is what 64 registers hold
  if (rt == 1)
    dense_pair<kLdg, 1>(A, B, R, onehot, act_kind, apply_act);
  else if (rt == 2)
    dense_pair<kLdg, 2>(A, B, R, onehot, act_kind, apply_act);
  else
    dense_pair<kLdg, 4>(A, B, R, onehot, act_kind, apply_act);
}

__device__ __forceinline__ DenseJob make_job(const mz_stack& s, int l, const float* w, const float* src, int lds,
                                             int in_x, float* dst, int ldd) {
  DenseJob J;
  J.W = w + s.w_off[l];
  J.bias = w + s.b_off[l];
  J.src = src;
  J.dst = dst;
  J.nin = l == 0 ? in_x : s.in_dim[l];
  J.nout = s.out_dim[l];
  J.lds = lds;
  J.ldd = ldd;
  return J;
}

// One hk.Sequential (sb == nullptr) or the two heads of a module evaluated in lockstep (one barrier per layer).
// Ends with a CTA barrier.
template <bool kLdg>
__device__ __noinline__ void run_lockstep(const mz_stack& sa, const mz_stack* sb, const float* w, int act_kind,
                                             const float* x, int ldx, int in_x, const int32_t* onehot, float* outa,
                                             float* outb, int ldo, float* ta0, float* ta1, float* tb0, float* tb1,
                                             int ldt, int R) {
  const float *srca = x, *srcb = x;
  int lds = ldx;
  for (int l = 0; l < sa.n_layers; ++l) {
    const bool last = l == sa.n_layers - 1;
    float* dsta = last ? outa : ((l & 1) ? ta1 : ta0);
    float* dstb = last ? outb : ((l & 1) ? tb1 : tb0);
    const int ldd = last ? ldo : ldt;
    const DenseJob A = make_job(sa, l, w, srca, lds, in_x, dsta, ldd);
    DenseJob B = A;
    B.nout = 0;
    if (sb != nullptr) B = make_job(*sb, l, w, srcb, lds, in_x, dstb, ldd);
    dense_pair_auto<kLdg>(A, B, R, l == 0 ? onehot : nullptr, act_kind, last ? 0 : 1);
    __syncthreads();
    srca = dsta;
    srcb = dstb;
    lds = ldd;
  }
}

template <bool kLdg>
__device__ __forceinline__ void run_stacks(const mz_stack& sa, const mz_stack* sb, const float* w, int act_kind,
                                           const float* x, int ldx, int in_x, const int32_t* onehot, float* outa,
                                           float* outb, int ldo, float* ta0, float* ta1, float* tb0, float* tb1, int ldt,
                                           int R) {
  if (sb != nullptr && sa.n_layers != sb->n_layers) {  // heads of different depth: one after the other
    run_lockstep<kLdg>(sa, nullptr, w, act_kind, x, ldx, in_x, onehot, outa, nullptr, ldo, ta0, ta1, nullptr, nullptr, ldt, R);
    run_lockstep<kLdg>(*sb, nullptr, w, act_kind, x, ldx, in_x, onehot, outb, nullptr, ldo, tb0, tb1, nullptr, nullptr, ldt, R);
    return;
  }
  run_lockstep<kLdg>(sa, sb, w, act_kind, x, ldx, in_x, onehot, outa, outb, ldo, ta0, ta1, tb0, tb1, ldt, R);
}

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

// ------------------------------------------------------------------------------------------ tie-break noise pre-pass

// One thread per (tree, simulation): per-tree key = split(sim_key, B_global)[global row]; then per level
// (key, sel) = split(key); noise[a] = 1e-7 * uniform(sel, (A,))[a]  (Appendix A.3, A.5, A.7).  Row = K levels x A.
__global__ void __launch_bounds__(128) resident_noise_kernel(SearchParams p, int B, int A, int K,
                                                             float* __restrict__ table, uint32_t* __restrict__ cont) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int NS = p.num_simulations;
  if (idx >= B * NS) return;
  const int b = idx / NS, sim = idx - b * NS;
  uint32_t k0, k1;
  split_key(p.sim_keys[2 * sim], p.sim_keys[2 * sim + 1], (uint32_t)p.global_batch, (uint32_t)(p.batch_offset + b),
            p.prng_mode, k0, k1);
  float* row = table + (size_t)idx * K * A;
  for (int d = 0; d < K; ++d) {
    uint32_t s0, s1;
    split_key(k0, k1, 2u, 1u, p.prng_mode, s0, s1);
    split_key(k0, k1, 2u, 0u, p.prng_mode, k0, k1);
    for (int x = 0; x < A; ++x) row[d * A + x] = tie_break_noise(bits_word(s0, s1, (uint32_t)A, (uint32_t)x, p.prng_mode));
  }
  cont[2 * (size_t)idx] = k0;
  cont[2 * (size_t)idx + 1] = k1;
}

// ------------------------------------------------------------------------------------------ tree records
// Working layout of the trees in HBM: 16-byte records, so that one level of `simulate` is one node load plus one
// (MuZero) or two (Gumbel: + prior logit) 16-byte loads per lane, one level of `backward` two loads and two stores,
// and every field of a node sits at a constant offset from one address.
//   node  n        : { visits (int), node_value, raw_value, parent << 8 | action  (0xFFFFFFFF: none) }
//   child (n, a) h0: { child index << 16 | visits  (index 0xFFFF: unvisited), prior prob, value, reward }
//   child (n, a) h1: { prior logit, -, -, - }
// children_discounts is not stored: on this path it is the constant gamma for every expanded edge (model.py:275) and
// an unexpanded edge has reward = value = 0, so reward + gamma * value is the same +0 as mctx's 0 + 0 * 0.
// The mctx SoA view of the C ABI (mz_get_tree) is produced on demand by resident_unpack_kernel.

constexpr uint32_t kRecNoChild = 0xFFFFu;
constexpr uint32_t kRecNoParent = 0xFFFFFFFFu;

struct RecTrees {    // the records of the trees one CTA owns (local tree index 0..R-1)
  float4* nodes;     // [R][N]
  float4* childs;    // [R][N][A][2]
  float* emb;        // [R][N][E]  (the handle's SoA embeddings)
  float* root_noise; // [R][A]
  uint8_t* root_invalid;
  int32_t* sim_depth; // [R][NS]
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

// Policy prologue (A.2 / A.4) + instantiate_tree_from_root (A.3) for one tree.
template <int G>
__device__ __forceinline__ void rec_begin(const RecTrees& t, const SearchParams& p, int b, long gb,
                                          const float* root_logits, float root_value, const float* root_emb,
                                          const uint8_t* invalid, const float* noise, int a, unsigned m) {
  const int A = t.A;
  float logit, prob, nz;
  bool inv;
  group_begin_compute<G>(p, A, gb, root_logits, invalid, noise, a, m, logit, prob, nz, inv);
  float4* ch = t.childs + (size_t)b * t.N * A * 2;
  if (a < A) {
    ch[a * 2] = make_float4(__uint_as_float(kRecNoChild << 16), prob, 0.0f, 0.0f);
    ch[a * 2 + 1] = make_float4(logit, 0.0f, 0.0f, 0.0f);
    t.root_noise[b * A + a] = nz;
    t.root_invalid[b * A + a] = inv ? 1 : 0;
  }
  float* emb = t.emb + (size_t)b * t.embN * t.E;
  for (int e = a; e < t.E; e += G) emb[e] = root_emb[e];
  if (a == 0)
    t.nodes[(size_t)b * t.N] = make_float4(__int_as_float(1), root_value, root_value, __uint_as_float(kRecNoParent));
}

// `simulate` (A.3) for one tree.  `fresh`: the selected edge was unvisited (the new node gets index sim + 1).
template <int G>
__device__ __forceinline__ void rec_simulate(const RecTrees& t, const SearchParams& p, int b, int sim, int a, unsigned m,
                                             int& parent, int& action, int& next, int& depth_out, bool& fresh,
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
  const float4* ch = t.childs + (size_t)b * t.N * A * 2;
  const bool root_inv = ok && t.root_invalid[b * A + a] != 0;
  const float root_gumbel = (!muzero && ok) ? t.root_noise[b * A + a] : 0.0f;
  int node = 0, depth = 0;
  uint32_t ci;
  for (;;) {
    const float4 nd = nodes[node];
    float4 h0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float logit = 0.0f;
    if (ok) {
      h0 = ch[(node * A + a) * 2];
      if (!muzero) logit = ch[(node * A + a) * 2 + 1].x;
    }
    uint32_t s0 = 0, s1 = 0;
    bool have_noise = false;
    float nz = 0.0f;
    if (muzero) {
      if (table && depth < aux.K) {
        have_noise = true;
        nz = aux.noise_row[depth * A + (ok ? a : 0)];
      } else {
        if (table && depth == aux.K) {
          k0 = aux.cont0;
          k1 = aux.cont1;
        }
        group_split2<G>(k0, k1, p.prng_mode, a, m, k0, k1, s0, s1);
      }
    }
    const ChildRow c = rec_child_row(h0, logit, p.discount, ok);
    action = group_select_score<G>(p, A, c, ok, nd.y, nd.z, __float_as_int(nd.x), depth, root_inv, root_gumbel, s0, s1, a,
                                   m, have_noise, nz, aux.pbc);
    ci = __shfl_sync(m, __float_as_uint(h0.x) >> 16, action, G);
    if (a == 0) path[depth] = ((uint32_t)node << 8) | (uint32_t)action;
    ++depth;
    if (ci == kRecNoChild || depth >= max_depth) break;
    node = (int)ci;
  }
  parent = node;
  depth_out = depth;
  fresh = ci == kRecNoChild;
  next = fresh ? sim + 1 : (int)ci;
}

// `expand` scatter (A.3) + `backward` for one tree, walking the path recorded by the selection: the records of
// level d-1 are loaded while level d's mean update is computed.
template <int G>
__device__ __forceinline__ void rec_expand_backup(const RecTrees& t, int b, int parent, int action, int next, bool fresh,
                                                  float reward, float gamma, float value, float logit_a,
                                                  const float* next_emb, int a, unsigned m, const uint32_t* path,
                                                  int depth) {
  const int A = t.A;
  const bool ok = a < A;
  float4* nodes = t.nodes + (size_t)b * t.N;
  float4* ch = t.childs + (size_t)b * t.N * A * 2;
  const float prob = group_softmax<G>(logit_a, ok, A, m);
  if (ok) {
    float4 h0 = make_float4(__uint_as_float(kRecNoChild << 16), prob, 0.0f, 0.0f);
    if (!fresh) {  // max_depth re-expansion: priors are overwritten, the edge statistics stay (update_tree_node)
      h0 = ch[(next * A + a) * 2];
      h0.y = prob;
    }
    ch[(next * A + a) * 2] = h0;
    ch[(next * A + a) * 2 + 1] = make_float4(logit_a, 0.0f, 0.0f, 0.0f);
  }
  float* emb = t.emb + ((size_t)b * t.embN + next) * t.E;
  for (int e = a; e < t.E; e += G) emb[e] = next_emb[e];
  if (a == 0) {
    const int old_visits = fresh ? 0 : __float_as_int(nodes[next].x);
    nodes[next] = make_float4(__int_as_float(old_visits + 1), value, value,
                              __uint_as_float(((uint32_t)parent << 8) | (uint32_t)action));
    // backward: path[d] = (node << 8 | action) of the edge selected at depth d; path[depth - 1] = (parent, action)
    float G_ = value, child_value = value;
    int d = depth - 1;
    int pn = parent, e2 = (parent * A + action) * 2;
    float4 nd = nodes[pn];
    float4 c = ch[e2];
    c.x = __uint_as_float(((uint32_t)next << 16) | (__float_as_uint(c.x) & 0xFFFFu));  // children_index[parent, action]
    c.w = reward;                                                                      // children_rewards[parent, action]
    for (;;) {
      int n_pn = 0, n_e2 = 0;
      float4 n_nd = nd, n_c = c;
      if (d > 0) {
        const uint32_t pa = path[d - 1];
        n_pn = (int)(pa >> 8);
        n_e2 = (n_pn * A + (int)(pa & 0xffu)) * 2;
        n_nd = nodes[n_pn];
        n_c = ch[n_e2];
      }
      const int count_i = __float_as_int(nd.x);
      const float count = (float)count_i;
      G_ = MZ_ADD(c.w, MZ_MUL(gamma, G_));
      const float pv = MZ_DIV(MZ_ADD(MZ_MUL(nd.y, count), G_), MZ_ADD(count, 1.0f));
      nodes[pn] = make_float4(__int_as_float(count_i + 1), pv, nd.z, nd.w);
      c.x = __uint_as_float(__float_as_uint(c.x) + 1u);  // children_visits += 1 (low 16 bits)
      c.z = child_value;
      ch[e2] = c;
      child_value = pv;
      if (d == 0) break;
      --d;
      pn = n_pn; e2 = n_e2; nd = n_nd; c = n_c;
    }
  }
}

// Policy epilogue for one tree (A.2 / A.4).
template <int G>
__device__ __forceinline__ void rec_finish(const RecTrees& t, const SearchParams& p, int b, long gb, bool has_invalid,
                                           int a, unsigned m, int& action, float& weight) {
  const int A = t.A;
  const bool ok = a < A;
  const bool muzero = p.policy == MZ_POLICY_MUZERO;
  const float4* ch = t.childs + (size_t)b * t.N * A * 2;
  const float4 nd = t.nodes[(size_t)b * t.N];
  float4 h0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  float logit = 0.0f;
  if (ok) {
    h0 = ch[a * 2];
    if (!muzero) logit = ch[a * 2 + 1].x;
  }
  const ChildRow c = rec_child_row(h0, logit, p.discount, ok);
  const bool root_inv = !muzero && ok && t.root_invalid[b * A + a] != 0;
  const float root_gumbel = (!muzero && ok) ? t.root_noise[b * A + a] : 0.0f;
  group_finish_score<G>(p, A, c, ok, nd.y, nd.z, root_inv, root_gumbel, gb, has_invalid, a, m, action, weight);
}

// Records -> the mctx SoA arrays of the handle (mz_get_tree view).  One thread per (tree, node); nodes that were
// never expanded (visits == 0) read as mctx's initial state (A.1).
__global__ void __launch_bounds__(256) resident_unpack_kernel(const float4* __restrict__ nodes,
                                                             const float4* __restrict__ childs, Tree o, int N_used,
                                                             float gamma) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)o.B * o.N) return;
  const int A = o.A;
  const int n = (int)(i % o.N);
  const long rec = (i / o.N) * N_used + n;  // records are packed with the stride of the search that wrote them
  float4 nd = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(kRecNoParent));
  if (n < N_used) nd = nodes[rec];
  const int visits = __float_as_int(nd.x);
  const uint32_t pa = __float_as_uint(nd.w);
  o.node_visits[i] = visits;
  o.parents[i] = pa == kRecNoParent ? -1 : (int)(pa >> 8);
  o.action_from_parent[i] = pa == kRecNoParent ? -1 : (int)(pa & 0xFFu);
  o.raw_values[i] = visits > 0 ? nd.z : 0.0f;
  o.node_values[i] = visits > 0 ? nd.y : 0.0f;
  for (int x = 0; x < A; ++x) {
    const long g = i * A + x;
    float4 h0 = make_float4(__uint_as_float(kRecNoChild << 16), 0.0f, 0.0f, 0.0f);
    float logit = 0.0f;
    if (visits > 0) {
      h0 = childs[(rec * A + x) * 2];
      logit = childs[(rec * A + x) * 2 + 1].x;
    }
    const uint32_t cx = __float_as_uint(h0.x);
    const bool has = (cx >> 16) != kRecNoChild;
    o.children_index[g] = has ? (int)(cx >> 16) : -1;
    o.children_visits[g] = (int)(cx & 0xFFFFu);
    o.children_prior_logits[g] = logit;
    o.children_prior_probs[g] = h0.y;
    o.children_values[g] = h0.z;
    o.children_rewards[g] = h0.w;
    o.children_discounts[g] = has ? gamma : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------ kernel

struct ResidentArgs {
  Net net;
  const float* weights;  // global fp32 blob
  int32_t weight_bytes;  // multiple of 16
  Tree t;                // the handle's SoA tree (embeddings, root_noise, root_invalid, sim_depth are used in place)
  float4* rec_nodes;     // [B][N]
  float4* rec_childs;    // [B][N][A][2]
  SearchParams p;
  const float* obs;          // [B,obs_dim] or null
  const float* root_emb;     // [B,E] when obs is null
  const float* root_logits;  // [B,A] or null (then Prediction runs here)
  const float* root_value;   // [B]   or null
  const uint8_t* invalid;
  const float* noise;
  const float* noise_table;  // [B][NS][K][A] or null
  const uint32_t* cont_keys; // [B][NS][2]
  int32_t K;
  int32_t* action_out;
  float* weights_out;
  float* root_value_out;
  int32_t T;   // trees per CTA
  int32_t ld;  // MLP staging row stride (floats, multiple of 4)
  int32_t PL;  // path slots per tree
  int32_t clear_embeddings;
};

struct ResidentLayout {  // offsets in floats from the dynamic smem base
  int weights, pbc, mlp, sel, path, total_floats;
};

constexpr int kResBufs = 9;  // x, ns, headV, headP, headR, tmp0A, tmp1A, tmp0B, tmp1B

__host__ __device__ inline ResidentLayout resident_layout(int weight_bytes_in_smem, int NS, int T, int ld, int PL) {
  ResidentLayout L;
  int off = 0;
  L.weights = off; off += round_up(weight_bytes_in_smem / 4, 4);
  L.pbc = off;     off += round_up(NS + 2, 4);
  L.mlp = off;     off += kResBufs * T * ld;
  L.sel = off;     off += round_up(7 * T, 4);  // parent, action, next, depth, fresh, reward, value
  L.path = off;    off += round_up(T * PL, 4);
  L.total_floats = round_up(off, 4);
  return L;
}

#ifndef MZ_RES_MIN_CTAS
#define MZ_RES_MIN_CTAS 4  // 64 registers per thread: four 256-thread CTAs per SM
#endif

template <int G, bool kWSmem>
__global__ void __launch_bounds__(256, MZ_RES_MIN_CTAS) resident_search_kernel(ResidentArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  constexpr bool kLdg = !kWSmem;
  const int T = a.T, A = a.net.num_actions, E = a.net.embed_dim, ld = a.ld, S = a.net.support_size;
  const int row0 = blockIdx.x * T;
  const int R = min(T, a.t.B - row0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int NS = a.p.num_simulations;
  const int N = NS + 1;  // record stride: nodes of this search
  const ResidentLayout L = resident_layout(kWSmem ? a.weight_bytes : 0, NS, T, ld, a.PL);
  const int act_kind = a.net.activation;

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

  RecTrees t;
  t.N = N; t.A = A; t.E = E; t.embN = a.t.N;
  t.nodes = a.rec_nodes + (size_t)row0 * N;
  t.childs = a.rec_childs + (size_t)row0 * N * A * 2;
  t.emb = a.t.embeddings + (size_t)row0 * a.t.N * E;  // NB: the SoA embeddings keep the handle's node stride
  t.root_noise = a.t.root_noise + (size_t)row0 * A;
  t.root_invalid = a.t.root_invalid + (size_t)row0 * A;
  t.sim_depth = a.t.sim_depth + (size_t)row0 * NS;

  // node records start as "never expanded" (visits = 0); child records are written when their node is expanded
  for (int i = tid; i < R * N; i += blockDim.x) t.nodes[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(kRecNoParent));
  if (a.clear_embeddings) {
    const long n = (long)R * a.t.N * E;
    if ((E & 3) == 0) {  // row0 * N * E * 4 bytes is a multiple of 16
      float4* e4 = reinterpret_cast<float4*>(t.emb);
      for (long i = tid; i < n / 4; i += blockDim.x) e4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    } else {
      for (long i = tid; i < n; i += blockDim.x) t.emb[i] = 0.0f;
    }
  }

  float* pbc = smem + L.pbc;
  for (int n = tid; n < NS + 2; n += blockDim.x) pbc[n] = pbc_explore((float)n, a.p.pb_c_init, a.p.pb_c_base);
  __syncthreads();  // the mbarrier initialised by thread 0 must exist before any other thread polls it

  float* x = smem + L.mlp;
  float* ns = x + T * ld;
  float* headV = ns + T * ld;
  float* headP = headV + T * ld;
  float* headR = headP + T * ld;
  float* ta0 = headR + T * ld;
  float* ta1 = ta0 + T * ld;
  float* tb0 = ta1 + T * ld;
  float* tb1 = tb0 + T * ld;
  int32_t* sel_parent = reinterpret_cast<int32_t*>(smem + L.sel);
  int32_t* sel_action = sel_parent + T;
  int32_t* sel_next = sel_action + T;
  int32_t* sel_depth = sel_next + T;
  int32_t* sel_fresh = sel_depth + T;
  float* rec_reward = reinterpret_cast<float*>(sel_fresh + T);
  float* rec_value = rec_reward + T;
  uint32_t* path = reinterpret_cast<uint32_t*>(smem + L.path);

  SearchParams p = a.p;
  p.batch_offset += row0;  // PRNG draws are indexed by global row

  // ---- root inference (muax/model.py:251-263); the root embedding lands in `ns`, the prior logits in `headP`
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
    run_stacks<kLdg>(a.net.repr, nullptr, w, act_kind, x, ld, obs_dim, nullptr, ns, nullptr, ld, ta0, ta1, nullptr, nullptr,
                     ld, R);
    if (a.net.repr_minmax) {
      for (int r = warp; r < R; r += nwarps) min_max_row_warp(ns + r * ld, E, lane);
      __syncthreads();
    }
  }
  if (a.obs != nullptr || a.root_logits == nullptr) {
    run_stacks<kLdg>(a.net.pred_v, &a.net.pred_pi, w, act_kind, ns, ld, E, nullptr, headV, headP, ld, ta0, ta1, tb0, tb1,
                     ld, R);
    for (int r = warp; r < R; r += nwarps) {
      const float v = support_to_scalar_warp(headV + r * ld, S, lane);
      if (lane == 0) rec_value[r] = v;
    }
  } else {
    for (int i = tid; i < R * A; i += blockDim.x) {
      const int r = i / A, k = i - r * A;
      headP[r * ld + k] = a.root_logits[(long)(row0 + r) * A + k];
    }
    if (tid < R) rec_value[tid] = a.root_value[row0 + tid];
  }
  __syncthreads();
  if (tid < R && a.root_value_out != nullptr) a.root_value_out[row0 + tid] = rec_value[tid];  // raw value (model.py:243)

  // ---- lane groups: group g of the CTA owns trees g, g + ngroups, ...  Trees that share a warp advance level by
  // level together (same loop body), so packing them costs no latency and divides the instruction count.
  const int ngroups = blockDim.x / G;
  const int gi = tid / G;
  const int ga = tid & (G - 1);  // action handled by this lane
  const unsigned gm = group_mask<G>();
  const bool use_table = a.noise_table != nullptr && a.K > 0 && p.policy == MZ_POLICY_MUZERO;
  const size_t nz_row = (size_t)a.K * A;

  for (int b = gi; b < R; b += ngroups) {
    const long ba = (long)(row0 + b) * A;
    rec_begin<G>(t, p, b, (long)p.batch_offset + b, headP + b * ld, rec_value[b], ns + b * ld,
                 a.invalid != nullptr ? a.invalid + ba : nullptr, a.noise != nullptr ? a.noise + ba : nullptr, ga, gm);
  }
  __syncthreads();

  // ---- simulations
  for (int sim = 0; sim < NS; ++sim) {
    // A: select
    for (int b = gi; b < R; b += ngroups) {
      SelectAux aux;
      aux.noise_row = nullptr;
      aux.K = 0;
      aux.cont0 = aux.cont1 = 0u;
      aux.pbc = pbc;
      if (use_table) {
        const size_t pair = (size_t)(row0 + b) * NS + sim;
        aux.noise_row = a.noise_table + pair * nz_row;
        aux.K = a.K;
        aux.cont0 = a.cont_keys[2 * pair];
        aux.cont1 = a.cont_keys[2 * pair + 1];
        if (sim + 1 < NS && ga == 0) prefetch_l1(aux.noise_row + nz_row);
      }
      int parent, action, next, depth;
      bool fresh;
      rec_simulate<G>(t, p, b, sim, ga, gm, parent, action, next, depth, fresh, aux, path + b * a.PL);
      if (ga == 0) {
        sel_parent[b] = parent;
        sel_action[b] = action;
        sel_next[b] = next;
        sel_depth[b] = depth;
        sel_fresh[b] = fresh ? 1 : 0;
        t.sim_depth[(size_t)b * NS + sim] = depth;
      }
      const float* pe = t.emb + ((size_t)b * t.embN + parent) * E;
      for (int e = ga; e < E; e += G) x[b * ld + e] = pe[e];
    }
    __syncthreads();
    // B: Dynamic (muax/model.py:269-271): next state -> ns, reward logits -> headR
    run_stacks<kLdg>(a.net.dyn_ns, &a.net.dyn_r, w, act_kind, x, ld, E, sel_action, ns, headR, ld, ta0, ta1, tb0, tb1, ld, R);
    // C: min-max of the next state + reward support transform, one warp per row
    for (int r = warp; r < R; r += nwarps) {
      if (a.net.dyn_minmax) min_max_row_warp(ns + r * ld, E, lane);
      const float rv = support_to_scalar_warp(headR + r * ld, S, lane);
      if (lane == 0) rec_reward[r] = rv;
    }
    __syncthreads();
    // D: Prediction (model.py:272): value logits -> headV, policy logits -> headP
    run_stacks<kLdg>(a.net.pred_v, &a.net.pred_pi, w, act_kind, ns, ld, E, nullptr, headV, headP, ld, ta0, ta1, tb0, tb1,
                     ld, R);
    // E: value support transform, one warp per row
    for (int r = warp; r < R; r += nwarps) {
      const float v = support_to_scalar_warp(headV + r * ld, S, lane);
      if (lane == 0) rec_value[r] = v;
    }
    __syncthreads();
    // F: expand + backup by the tree's lane group
    for (int b = gi; b < R; b += ngroups) {
      const float logit = ga < A ? headP[b * ld + ga] : 0.0f;
      rec_expand_backup<G>(t, b, sel_parent[b], sel_action[b], sel_next[b], sel_fresh[b] != 0, rec_reward[b], p.discount,
                           rec_value[b], logit, ns + b * ld, ga, gm, path + b * a.PL, sel_depth[b]);
    }
    // the next select of a tree runs on the lanes of the same group: a warp-level fence orders the backup's global
    // writes before it; the staging buffers are only rewritten after the next CTA barrier
    __syncwarp();
  }

  // ---- policy epilogue
  for (int b = gi; b < R; b += ngroups) {
    int action;
    float weight;
    rec_finish<G>(t, p, b, (long)p.batch_offset + b, a.invalid != nullptr, ga, gm, action, weight);
    if (ga < A) a.weights_out[(long)(row0 + b) * A + ga] = weight;
    if (ga == 0) a.action_out[row0 + b] = action;
  }
}

// ------------------------------------------------------------------------------------------ host side

static void* resident_kernel_ptr(int G, bool wsmem) {
#define MZ_RES_CASE(g) \
  case g: return wsmem ? (void*)resident_search_kernel<g, true> : (void*)resident_search_kernel<g, false>
  switch (G) {
    MZ_RES_CASE(2);
    MZ_RES_CASE(4);
    MZ_RES_CASE(8);
    MZ_RES_CASE(16);
    default: return wsmem ? (void*)resident_search_kernel<32, true> : (void*)resident_search_kernel<32, false>;
  }
#undef MZ_RES_CASE
}

static int net_weight_bytes(const Net& net) {
  int64_t wfloats = 0;
  const mz_stack* stacks[5] = {&net.repr, &net.pred_v, &net.pred_pi, &net.dyn_ns, &net.dyn_r};
  for (const mz_stack* s : stacks)
    for (int l = 0; l < s->n_layers; ++l) {
      wfloats = std::max(wfloats, s->w_off[l] + (int64_t)s->in_dim[l] * s->out_dim[l]);
      wfloats = std::max(wfloats, s->b_off[l] + (int64_t)s->out_dim[l]);
    }
  return round_up((int)wfloats * 4, 16);
}

int resident_init(ResidentState& st, const Net& net, int device, std::string* err) {
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
  if (const char* e = getenv("MZ_RESIDENT_K")) st.noise_levels = std::max(0, atoi(e));
  if (const char* e = getenv("MZ_RESIDENT_THREADS")) {
    const int n = atoi(e);
    if (n == 64 || n == 128 || n == 256) st.threads = n;
  }
  st.available = true;
  return 0;
}

void resident_destroy(ResidentState& st) {
  if (st.noise_table) cudaFree(st.noise_table);
  if (st.cont_keys) cudaFree(st.cont_keys);
  if (st.rec_nodes) cudaFree(st.rec_nodes);
  if (st.rec_childs) cudaFree(st.rec_childs);
  st.noise_table = nullptr;
  st.cont_keys = nullptr;
  st.rec_nodes = st.rec_childs = nullptr;
  st.noise_capacity = st.cont_capacity = st.rec_capacity = 0;
  st.dirty = false;
}

// Records of the last search -> the handle's SoA arrays (lazily, when the tree view is asked for).
int resident_unpack(ResidentState& st, const Tree& tree, float gamma, std::string* err) {
  if (!st.dirty) return 0;
  const long n = (long)tree.B * tree.N;
  resident_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st.last_stream>>>(
      reinterpret_cast<const float4*>(st.rec_nodes), reinterpret_cast<const float4*>(st.rec_childs), tree,
      st.last_num_sims + 1, gamma);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st.last_stream);
  if (e != cudaSuccess) {
    *err = std::string("resident engine: unpacking the tree records failed: ") + cudaGetErrorString(e);
    return 1;
  }
  st.dirty = false;
  return 0;
}

struct ResidentPlan {
  int T = 0, grid = 0, wsmem = 0, PL = 1;
  size_t smem = 0;
};

// Trees per CTA: the smallest T for which all ceil(B / T) CTAs are co-resident (occupancy from the runtime: shared
// memory for weights + staging, registers, threads), so that every tree of the batch is in flight at once and the
// SMs hold as many independent CTAs as possible to overlap one CTA's tree walk with another's dense layers.
static ResidentPlan resident_plan(const ResidentState& st, const Net& net, int B, int NS, int max_depth) {
  ResidentPlan best;
  const int ld = round_up(net.max_width, 4);
  const int wbytes = net_weight_bytes(net);
  const int budget = st.max_smem - 1024;  // per CTA (opt-in limit minus the static mbarrier + slack)
  const int PL = std::max(1, std::min(max_depth > 0 ? max_depth : NS, NS));
  for (int ws = st.force_global_weights ? 0 : 1; ws >= 0; --ws) {
    auto bytes = [&](int T) { return (size_t)resident_layout(ws ? wbytes : 0, NS, T, ld, PL).total_floats * 4; };
    if (bytes(1) > (size_t)budget) continue;
    int T = 0;
    if (st.trees_per_cta > 0) {
      T = st.trees_per_cta;
    } else {
      for (int cand = 1; cand <= 64; ++cand) {
        if (bytes(cand) > (size_t)budget) break;
        T = cand;  // the largest that fits, unless a smaller one already covers B in one wave
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resident_kernel_ptr(st.G, ws != 0), st.threads,
                                                          bytes(cand)) != cudaSuccess) {
          cudaGetLastError();
          per_sm = 1;
        }
        if ((long)per_sm * st.num_sms * cand >= B) break;
      }
    }
    if (T <= 0 || bytes(T) > (size_t)budget) continue;
    best.T = T;
    best.wsmem = ws;
    best.smem = bytes(T);
    best.grid = (B + T - 1) / T;
    best.PL = PL;
    return best;
  }
  return best;
}

bool resident_supported(const ResidentState& st, const Net& net, int B, int num_simulations) {
  // child records pack the child index and the visit count into 16 bits each
  return st.available && num_simulations + 1 < (int)kRecNoChild && resident_plan(st, net, B, num_simulations, 0).T > 0;
}

int resident_launch(ResidentState& st, const Net& net, const float* weights, const Tree& tree, const SearchParams& p,
                    const float* obs, const float* root_emb, const float* root_logits, const float* root_value,
                    const uint8_t* invalid, const float* noise, int32_t* action_out, float* weights_out,
                    float* root_value_out, cudaStream_t stream, int64_t* launches, std::string* err) {
  const int B = tree.B, NS = p.num_simulations, A = net.num_actions;
  const ResidentPlan plan = resident_plan(st, net, B, NS, p.max_depth);
  if (plan.T <= 0) {
    *err = "resident engine: MLP staging does not fit in shared memory";
    return 1;
  }
  ResidentArgs a{};
  a.net = net;
  a.weights = weights;
  a.weight_bytes = net_weight_bytes(net);
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
  a.ld = round_up(net.max_width, 4);
  a.PL = plan.PL;
  a.clear_embeddings = (p.max_depth > 0 || NS + 1 < tree.N) ? 1 : 0;
  {
    const size_t nodes = (size_t)B * (NS + 1);
    if (nodes > st.rec_capacity) {
      if (st.rec_nodes) cudaFree(st.rec_nodes);
      if (st.rec_childs) cudaFree(st.rec_childs);
      st.rec_nodes = st.rec_childs = nullptr;
      st.rec_capacity = 0;
      if (cudaMalloc(&st.rec_nodes, nodes * 16) != cudaSuccess || cudaMalloc(&st.rec_childs, nodes * A * 32) != cudaSuccess) {
        cudaGetLastError();
        *err = "resident engine: cudaMalloc(tree records) failed";
        return 1;
      }
      st.rec_capacity = nodes;
    }
    a.rec_nodes = reinterpret_cast<float4*>(st.rec_nodes);
    a.rec_childs = reinterpret_cast<float4*>(st.rec_childs);
  }
  // tie-break noise ahead of the search (MuZero policy only: the Gumbel selectors ignore their key)
  if (p.policy == MZ_POLICY_MUZERO && NS > 0 && st.noise_levels > 0) {
    const size_t pairs = (size_t)B * NS;
    const size_t cap_bytes = (size_t)1 << 30;  // at most 1 GiB of table
    int K = std::min(st.noise_levels, plan.PL);
    K = (int)std::min<size_t>((size_t)K, cap_bytes / (pairs * A * 4));
    if (K > 0) {
      const size_t need = pairs * (size_t)K * A;
      if (need > st.noise_capacity) {
        if (st.noise_table) cudaFree(st.noise_table);
        st.noise_table = nullptr;
        st.noise_capacity = 0;
        if (cudaMalloc((void**)&st.noise_table, need * 4) != cudaSuccess) {
          cudaGetLastError();
          *err = "resident engine: cudaMalloc(noise table) failed";
          return 1;
        }
        st.noise_capacity = need;
      }
      if (pairs > st.cont_capacity) {
        if (st.cont_keys) cudaFree(st.cont_keys);
        st.cont_keys = nullptr;
        st.cont_capacity = 0;
        if (cudaMalloc((void**)&st.cont_keys, pairs * 8) != cudaSuccess) {
          cudaGetLastError();
          *err = "resident engine: cudaMalloc(carry keys) failed";
          return 1;
        }
        st.cont_capacity = pairs;
      }
      resident_noise_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, stream>>>(p, B, A, K, st.noise_table, st.cont_keys);
      *launches += 1;
      a.noise_table = st.noise_table;
      a.cont_keys = st.cont_keys;
      a.K = K;
    }
  }
  void* args[] = {&a};
  const cudaError_t e = cudaLaunchKernel(resident_kernel_ptr(st.G, plan.wsmem != 0), dim3(plan.grid), dim3(st.threads),
                                         args, plan.smem, stream);
  *launches += 1;
  if (e != cudaSuccess) {
    *err = std::string("resident engine launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  st.dirty = true;
  st.last_stream = stream;
  st.last_num_sims = NS;
  return 0;
}

}  // namespace mz

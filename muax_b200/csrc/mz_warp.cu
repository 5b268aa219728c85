// mz_warp.cu — warp-autonomous, compile-time specialised search engine for the stock muax MLP family
// (the configuration BASELINE.json's metric is quoted on).
//
// The previous headline engine ("lane2", retired in round 2) ran 32 trees per CTA in lock step: nine CTA barriers per simulation, every phase as long as
// the slowest tree's, 12 of 16 warps idle during the tree walks (profiles/r01_lane2_*: 33 % of the stall samples on
// one barrier, 17k cycles per simulation).  A tree's simulations are a dependent chain, so the only way to finish an
// act sooner is to shorten that chain.  Here a warp owns 32 / G trees (G = 16 lanes per tree by default, 8 as a
// knob) for the whole act and never meets a CTA barrier after the prologue:
//   * dense layers: the two heads of a module form one 32-column matrix (re-laid out in shared memory once per CTA
//     from the TMA-staged blob), a lane owns U = 32 / G adjacent columns -> one LDS.64 / LDS.128 of weights per k (the
//     lanes of a tree read one 128-byte row), activations are broadcast LDS.128 from the tree's scratch; 4 layers =
//     4 `__syncwarp`s; ELU is evaluated branch-free on the U units at once;
//   * categorical heads: reward head on the even lanes, value head on the odd lanes — exps and quotients dealt out
//     G / 2 ways, the two left-to-right float sums recomputed by every lane (same order as every other engine);
//   * selection is lane-parallel over the actions (lane x scores child x; min / max and the index-ordered argmax go
//     through shuffles inside the tree's lanes); expand / backup are lane2's scalar code on 16-byte records, executed
//     redundantly by the lanes of a tree (identical values, identical stores), so no broadcast or vote is needed and
//     the trees of a warp only wait for each other's path length, not for 31 other trees;
//   * tie-break noise: `producers` extra warps per CTA run the jax key chain of (tree, simulation) pairs ahead of the
//     search into a ring of kWRing simulations guarded by full / empty mbarriers (lane = tree; every producer lane
//     arrives on `full`, every search lane on `empty`); with 0 producers the
//     rows come from noise_table_kernel's table in HBM;
//   * every warp unpacks its own trees into the mctx SoA view when it finishes (no barrier before the dump); in the
//     sharded multi-GPU act the outputs are also stored into the peers' gather buffers (LaneArgs::peer_delta).
//   * the per-simulation simulate keys travel in the kernel parameters (SimKeys) for searches of at most kInlineSims
//     simulations: no H2D copy and no event on the act's critical path.
// Same arithmetic in the same order as every other engine / the CPU checkers: bit-identical (tests/test_gpu_parity.py).
#include "mz_warp.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace mz {

constexpr int kLMaxLayers = 4;
constexpr int kGNoiseFloats = 32;  // tie-break noise row per (tree, simulation): K * A <= 32 floats

struct PackSrc {  // a raw hk.Linear in the fp32 blob
  int64_t w_off, b_off;
  int32_t in, out;
};


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
  int32_t walkers;  // search warps per CTA (the remaining warps produce tie-break noise); 0 = all of them
  // warp engine: action / action_weights / root_value are also stored at (pointer + peer_delta[i]) — the same slots of
  // the peer GPUs' gather buffers (NVLink peer stores; muax_b200/sharded.py)
  int32_t n_peers;
  int64_t peer_delta[7];
  // completion flag of the sharded act (mz_set_peer_flags): the grid's last CTA, once every output store (local and
  // peer) is fenced system-wide, stores flag_step at flag[flag_rank] here and at the same slot of every peer
  int32_t* flag;
  int32_t flag_rank, flag_step;
  uint32_t* done_counter;  // CTAs of this launch that have finished (reset by the last one)
  SimKeys ik;  // simulate keys by value (p.sim_keys == nullptr), else p.sim_keys points at them in device memory
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


// Tie-break noise of one (tree, simulation) for the first `levels` levels of its walk: per-tree key =
// split(sim_key, B_global)[global row]; then per level (key, sel) = split(key); noise[a] = 1e-7 *
// uniform(sel, (A,))[a]  (Appendix A.3, A.5, A.7).  `cont` receives the key the chain continues from.
__device__ __forceinline__ void noise_row(const SearchParams& p, int A, int levels, uint32_t sk0, uint32_t sk1,
                                          uint32_t global_row, float* __restrict__ row, uint32_t* __restrict__ cont) {
  uint32_t k0, k1;
  split_key(sk0, sk1, (uint32_t)p.global_batch, global_row, p.prng_mode, k0, k1);
  const int half = (A + 1) >> 1;
  for (int d = 0; d < levels; ++d) {
    uint32_t n0, n1, s0, s1;
    if (p.prng_mode == MZ_PRNG_THREEFRY_LEGACY) {
      uint32_t p0, p1, q0, q1;
      threefry2x32(k0, k1, 0u, 2u, p0, p1);
      threefry2x32(k0, k1, 1u, 3u, q0, q1);
      n0 = p0; n1 = q0; s0 = p1; s1 = q1;
      for (int i = 0; i < half; ++i) {
        uint32_t y0, y1;
        threefry2x32(s0, s1, (uint32_t)i, (uint32_t)(half + i < A ? half + i : 0), y0, y1);
        row[d * A + i] = tie_break_noise(y0);
        if (half + i < A) row[d * A + half + i] = tie_break_noise(y1);
      }
    } else {
      threefry2x32(k0, k1, 0u, 0u, n0, n1);
      threefry2x32(k0, k1, 0u, 1u, s0, s1);
      for (int i = 0; i < A; ++i) {
        uint32_t y0, y1;
        threefry2x32(s0, s1, 0u, (uint32_t)i, y0, y1);
        row[d * A + i] = tie_break_noise(y0 ^ y1);
      }
    }
    k0 = n0;
    k1 = n1;
  }
  cont[0] = k0;
  cont[1] = k1;
}

// One thread per (tree, simulation).
__global__ void __launch_bounds__(128) noise_table_kernel(SearchParams p, int B, int A, int K, float* __restrict__ table,
                                                          uint32_t* __restrict__ cont) {
  // programmatic dependent launch: the search kernel that follows may start its prologue (weight staging, tree
  // initialisation, root inference) right away; it executes griddepcontrol.wait before it first reads the table
  asm volatile("griddepcontrol.launch_dependents;");
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int NS = p.num_simulations;
  if (idx >= B * NS) return;
  // simulation-major: a warp works on one simulation of 32 trees, so the depth bound below is warp-uniform
  const int sim = idx / B, b = idx - sim * B;
  const size_t pair = (size_t)b * NS + sim;
  // simulation `sim` walks a tree of sim + 1 nodes: its path has at most sim + 1 levels (and then never reaches the
  // continuation key, which is only read at depth K)
  noise_row(p, A, min(K, sim + 1), p.sim_keys[2 * sim], p.sim_keys[2 * sim + 1], (uint32_t)(p.batch_offset + b),
            table + pair * kGNoiseFloats, cont + 2 * pair);
}


// MZ_WARP_CACHED (default 1): selection scores cached per (node, child) and refreshed by a lane-parallel backup (see
// the simulation loop of warp_search_kernel); 0 = the round-1 loop (scores computed on the walk, serial backup) for A/B.
#ifndef MZ_WARP_CACHED
#define MZ_WARP_CACHED 1
#endif

template <int A, int E>
struct L2Layout {  // float offsets inside one tree block (all 16-byte aligned)
  int nodes, childs, raw, logits, emb, root, scores, path, stride;
  __host__ __device__ explicit L2Layout(int N) {
    int o = 0;
    nodes = o; o += 4 * N;
    childs = o; o += 4 * N * A;
    raw = o; o += round_up(N, 4);
    logits = o; o += round_up(N * A, 4);
    emb = o; o += round_up(N * E, 4);
    root = o; o += round_up(2 * A, 4);  // root_noise[A], root_invalid[A] (as floats 0/1)
    scores = o; o += MZ_WARP_CACHED ? round_up(N * A, 4) : 0;  // selection scores without the tie-break noise
    path = o; o += MZ_WARP_CACHED ? round_up(N, 4) : 0;        // (node << 8 | action) per level of the last walk
    while (o % 32 != 4) o += 4;
    stride = o;
  }
};

constexpr uint32_t kNoChild = 0xFFFFu;


// Cold path of the selection: tie-break noise for a level the pre-computed table does not cover (depth >= K, or
// no table at all).  Kept out of line so the hot loop stays small.  k0/k1 carry the jax key chain between levels.
// All state goes through this thread's shared-memory slot {key0, key1, noise[A]} (no stack frame).
template <int A>
__device__ __noinline__ void l2_noise_cold(uint32_t sk0, uint32_t sk1, uint32_t global_batch, uint32_t global_row,
                                           int prng_mode, int depth, int K, const uint32_t* cont, float* slot) {
  uint32_t k0 = __float_as_uint(slot[0]), k1 = __float_as_uint(slot[1]);
  if (K < 0 && depth == 0) split_key(sk0, sk1, global_batch, global_row, prng_mode, k0, k1);
  if (K >= 0 && depth == K) {
    k0 = cont[0];  // generic loads: the warp engine keeps the continuation keys in shared memory
    k1 = cont[1];
  }
  uint32_t s0, s1;
  lt_split2(k0, k1, prng_mode, k0, k1, s0, s1);
  slot[0] = __uint_as_float(k0);
  slot[1] = __uint_as_float(k1);
#pragma unroll
  for (int x = 0; x < A; ++x) slot[2 + x] = tie_break_noise(bits_word(s0, s1, A, x, prng_mode));
}




constexpr int kWMaxWarps = 16;      // search warps per CTA (8 when a tree takes 8 lanes)
constexpr int kWMaxProducers = 4;   // noise-producer warps per CTA
constexpr int kWRing = 4;           // simulations of tie-break noise in flight per CTA
constexpr int kWRow = kGNoiseFloats + 1;  // ring row stride: lane-strided row writes land on distinct banks

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// G lanes per tree (8 or 16), U = 32 / G adjacent columns of every 32-column layer per lane, 32 / G trees per warp.
template <int U>
__device__ __forceinline__ void w_ldv(const float* p, float (&a)[U]) {
  if constexpr (U == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
  } else {
    const float2 v = *reinterpret_cast<const float2*>(p);
    a[0] = v.x; a[1] = v.y;
  }
}
template <int U>
__device__ __forceinline__ void w_stv(float* p, const float (&a)[U]) {
  if constexpr (U == 4) *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(a[0], a[1]);
}

template <int A, int E, int H, int S>
struct WShape {
  static constexpr int F = 2 * S + 1;
  static constexpr int F4 = (F + 3) / 4 * 4;
  static constexpr int A4 = (A + 3) / 4 * 4;
  static constexpr int E4 = (E + 3) / 4 * 4;
  static_assert(2 * H == 32, "the two hidden layers of a module (H units each) form one 32-column matrix");
  static_assert(E % 4 == 0 && E + F4 <= 32, "Dynamic's output columns (next state | reward logits) must fit 32");
  static_assert(F4 + A4 <= 32, "Prediction's output columns (value logits | policy logits) must fit 32");
  // re-laid-out weight matrices, [rows][32] floats each
  static constexpr int D1 = 0;                          // Dynamic layer 1: E x-rows, A one-hot rows, bias
  static constexpr int D2 = D1 + (E + A + 1) * 32;      // Dynamic layer 2: H rows, bias
  static constexpr int P1 = D2 + (H + 1) * 32;          // Prediction layer 1: E rows, bias
  static constexpr int P2 = P1 + (E + 1) * 32;          // Prediction layer 2: H rows, bias
  static constexpr int WQ = P2 + (H + 1) * 32;
  // per-tree scratch (float offsets)
  static constexpr int sH = 0;            // hidden units [32]
  static constexpr int sO1 = 32;          // Dynamic out: next state [E] | reward logits [F4]
  static constexpr int sO2 = 64;          // Prediction out: value logits [F4] | policy logits [A4]
  static constexpr int sE = 96;           // exps: reward [24] | value [24]
  static constexpr int sP = 144;          // (j - S) * prob: reward [24] | value [24]
  static constexpr int sNz = 192;         // tie-break noise row of this simulation [32]
  static constexpr int sCold = 224;       // cold noise slot {key0, key1, noise[A]}
  static constexpr int sStride = 264;     // == 8 (mod 32): the 4 trees of a warp sit on different banks
  static_assert(F4 <= 24 && A + 2 <= 8, "scratch sized for F <= 24, A <= 6");
};

template <int A, int E, int H, int S>
struct WSmem {  // float offsets of the CTA's shared memory
  int w, wq, pbc, ring, ringc, scratch, blocks, total;
  __host__ __device__ WSmem(int packed_floats, int NS, int N, int trees, bool producers) {
    using W = WShape<A, E, H, S>;
    int o = 0;
    w = o; o += round_up(packed_floats, 4);
    wq = o; o += W::WQ;
    pbc = o; o += round_up(NS + 2, 4);
    ring = o; o += producers ? round_up(kWRing * trees * kWRow, 4) : 0;    // noise rows [kWRing][trees][kWRow]
    ringc = o; o += producers ? round_up(kWRing * trees * 2, 4) : 0;      // continuation keys [kWRing][trees][2]
    scratch = o; o += trees * W::sStride;
    blocks = o;
    total = o + trees * L2Layout<A, E>(N).stride;
  }
};

// U adjacent columns [U*l, U*l+U) of a [rows][32] matrix: acc_j = (sum_k fma(x_k, W_kj)) [+ one-hot row] (+ bias later).
template <int K, int U>
__device__ __forceinline__ void w_dense(const float* __restrict__ Wm, int l, const float (&x)[K], int extra_row,
                                        float (&a)[U]) {
  const float* wj = Wm + U * l;
#pragma unroll
  for (int u = 0; u < U; ++u) a[u] = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float wv[U];
    w_ldv<U>(wj + k * 32, wv);
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = MZ_FMA(x[k], wv[u], a[u]);
  }
  if (extra_row >= 0) {
    float wv[U];
    w_ldv<U>(wj + extra_row * 32, wv);
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = MZ_ADD(a[u], wv[u]);
  }
}

template <int K>
__device__ __forceinline__ void w_load(const float* src, float (&x)[K]) {
  static_assert(K % 4 == 0, "vector loads");
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + k);
    x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
  }
}

// Activation of U independent units.  mz_elu's `x > 0 ? x : expm1(x)` compiles to a branch per unit (serialised,
// both sides taken in nearly every warp); here expm1 runs unconditionally on all U units (interleaved by the
// scheduler) and the result is selected — the same values bit for bit.
template <int U>
__device__ __forceinline__ void w_activate(float (&a)[U], int kind) {
  if (kind == MZ_ACT_ELU) {
    float e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      e[u] = mz_expm1f(a[u]);
      asm volatile("" : "+f"(e[u]));  // keep the evaluation out of a conditional block
    }
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = a[u] > 0.0f ? a[u] : e[u];
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = a[u] > 0.0f ? a[u] : 0.0f;
  }
}

// div_try for operands known to be non-negative (a negative or -0 numerator is merely sent to the exact path):
// a == +0 or a in [2^-30, 2^31) on the raw bits; `b_ok` = the caller knows b is in [2^-30, 2^31).
__device__ __forceinline__ float w_div_nn(float a, float b, bool b_ok, bool& bad) {
  const uint32_t ua = __float_as_uint(a);
  bool ok = ua == 0u || (ua - 0x30800000u) < 0x1E800000u;
  if (!b_ok) ok = ok && (__float_as_uint(b) - 0x30800000u) < 0x1E800000u;
  bad = bad || !ok;
  return div_core(a, b);
}

template <int U>
__device__ __forceinline__ void w_bias_act_store(const float* __restrict__ bias_row, int l, float (&a)[U], bool act,
                                                 int act_kind, float* dst) {
  float bv[U];
  w_ldv<U>(bias_row + U * l, bv);
#pragma unroll
  for (int u = 0; u < U; ++u) a[u] = MZ_ADD(a[u], bv[u]);
  if (act) w_activate<U>(a, act_kind);
  w_stv<U>(dst + U * l, a);
}

// muax/nn.py:37-44 on a register row (same operations as l2_minmax_col).
template <int N_>
__device__ __forceinline__ void w_minmax(float (&v)[N_], bool enabled) {
  if (!enabled) return;
  float lo = v[0], hi = v[0];
#pragma unroll
  for (int k = 1; k < N_; ++k) {
    lo = fminf(lo, v[k]);
    hi = fmaxf(hi, v[k]);
  }
  float scale = MZ_SUB(hi, lo);
  if (scale < 1e-5f) scale = MZ_ADD(scale, 1e-5f);
  float num[N_];
  bool bad = !((__float_as_uint(scale) - 0x30800000u) < 0x1E800000u);
#pragma unroll
  for (int k = 0; k < N_; ++k) {
    num[k] = MZ_SUB(v[k], lo);
    v[k] = w_div_nn(num[k], scale, true, bad);
  }
  if (bad) {
#pragma unroll
    for (int k = 0; k < N_; ++k) v[k] = MZ_DIV(num[k], scale);
  }
}

// The same normalisation dealt out over the lanes of a tree: every lane finds min / max of the raw row (shared memory
// at `raw`), lane k divides element k, the quotients meet in `tmp` and every lane reloads the row — one division per
// lane instead of N_ (identical operations per element).
template <int N_, int G_>
__device__ __forceinline__ void w_minmax_dist(const float* raw, float* tmp, int l, bool enabled, float (&v)[N_]) {
  static_assert((N_ & (N_ - 1)) == 0, "power-of-two row");
  w_load<N_>(raw, v);
  if (!enabled) return;
  float lo = v[0], hi = v[0];
#pragma unroll
  for (int k = 1; k < N_; ++k) {
    lo = fminf(lo, v[k]);
    hi = fmaxf(hi, v[k]);
  }
  float scale = MZ_SUB(hi, lo);
  if (scale < 1e-5f) scale = MZ_ADD(scale, 1e-5f);
#pragma unroll
  for (int k0 = 0; k0 < N_; k0 += G_) {  // one pass when the tree's lanes cover the row
    const int k = (k0 + l) & (N_ - 1);
    const float num = MZ_SUB(raw[k], lo);
    bool bad = !((__float_as_uint(scale) - 0x30800000u) < 0x1E800000u);
    float qv = w_div_nn(num, scale, true, bad);
    if (bad) qv = MZ_DIV(num, scale);
    tmp[k] = qv;
  }
  __syncwarp();
  w_load<N_>(tmp, v);
}

// Both categorical heads of one tree (muax/utils.py:94-102 on softmax(logits)): even lanes take the reward head
// (logits at sc + sO1 + E), odd lanes the value head (sc + sO2); returns this lane's head scalar.
template <int A, int E, int H, int S, int G>
__device__ __forceinline__ float w_heads(float* sc, int l) {
  using W = WShape<A, E, H, S>;
  constexpr int F = W::F, F4 = W::F4;
  constexpr int HL = G / 2;               // lanes per head
  constexpr int R = (F + HL - 1) / HL;    // quotient / exp rounds per lane
  const int hd = l & 1, q = l >> 1;
  const float* lg = hd ? sc + W::sO2 : sc + W::sO1 + E;
  float* eb = sc + W::sE + hd * 24;
  float* pb = sc + W::sP + hd * 24;
  {
    float v[F4];
    w_load<F4>(lg, v);
    float mx = v[0];
#pragma unroll
    for (int j = 1; j < F; ++j) mx = fmaxf(mx, v[j]);
#pragma unroll
    float ex[R];
#pragma unroll
    for (int i = 0; i < R; ++i) ex[i] = mz_expf(MZ_SUB(lg[min(q + HL * i, F - 1)], mx));
#pragma unroll
    for (int i = 0; i < R; ++i)
      if (q + HL * i < F) eb[q + HL * i] = ex[i];
  }
  __syncwarp();
  {
    float e[F4];
    w_load<F4>(eb, e);
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < F; ++j) s = MZ_ADD(s, e[j]);
#pragma unroll
    // s >= 1 (the largest logit contributes exp(0)) and s <= F: only the numerators need the range check
    float ej[R], pr[R];
    bool bad = !(s >= 1.0f && s <= (float)F);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      ej[i] = eb[min(q + HL * i, F - 1)];
      pr[i] = w_div_nn(ej[i], s, true, bad);
    }
    if (bad) {
#pragma unroll
      for (int i = 0; i < R; ++i) pr[i] = MZ_DIV(ej[i], s);
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
      if (q + HL * i < F) pb[q + HL * i] = MZ_MUL((float)(q + HL * i - S), pr[i]);
  }
  __syncwarp();
  float pv[F4];
  w_load<F4>(pb, pv);
  float x = 0.0f;
#pragma unroll
  for (int j = 0; j < F; ++j) x = MZ_ADD(x, pv[j]);
  return mz_inv_scaling(x);
}

// Prediction (muax/nn.py:73-90) on the register row v -> value / policy logits in sc + sO2.
template <int A, int E, int H, int S, int G>
__device__ __forceinline__ void w_prediction(const float* wq, float* sc, int l, const float (&v)[E], int act_kind) {
  using W = WShape<A, E, H, S>;
  constexpr int U = 32 / G;
  float acc[U];
  w_dense<E, U>(wq + W::P1, l, v, -1, acc);
  w_bias_act_store<U>(wq + W::P1 + E * 32, l, acc, true, act_kind, sc + W::sH);
  __syncwarp();
  float h[H];
  w_load<H>(sc + W::sH + (U * l < W::F4 ? 0 : H), h);
  w_dense<H, U>(wq + W::P2, l, h, -1, acc);
  w_bias_act_store<U>(wq + W::P2 + H * 32, l, acc, false, 0, sc + W::sO2);
  __syncwarp();
}

// muzero_action_selection (A.5) with qtransform_by_parent_and_siblings (A.6) for ALL children of node n, by one lane:
// score[a] = value_score + prior_score, i.e. what the walk adds the tie-break noise to.  The operations (and their
// order per child) are those of the on-the-walk selection; min / max over the visited children are order-independent.
// A node's scores change only when a backup passes through it or when it is (re)expanded.
template <int A>
__device__ __forceinline__ void w_node_scores(const float4* nodes, const float4* childs, float* tsc, int n, float gamma) {
  const float4 nd = nodes[n];
  float4 ch[A];
#pragma unroll
  for (int x = 0; x < A; ++x) ch[x] = childs[n * A + x];
  float q[A], lo = nd.y, hi = nd.y;
#pragma unroll
  for (int x = 0; x < A; ++x) {
    q[x] = MZ_ADD(ch[x].w, MZ_MUL(gamma, ch[x].z));
    const float qn = (__float_as_uint(ch[x].x) & 0xFFFFu) != 0u ? q[x] : mz_nan();  // fminf / fmaxf drop the NaN
    lo = fminf(lo, qn);
    hi = fmaxf(hi, qn);
  }
  const float denom = fmaxf(MZ_SUB(hi, lo), 1e-8f);
#pragma unroll
  for (int x = 0; x < A; ++x) {
    const int vis = (int)(__float_as_uint(ch[x].x) & 0xFFFFu);
    const float vnum = MZ_SUB(vis > 0 ? q[x] : lo, lo);
    const float pnum = MZ_MUL(nd.z, ch[x].y);
    const float pden = (float)(vis + 1);  // in [1, 65536]
    bool bad = false;
    float vsv = w_div_nn(vnum, denom, false, bad);
    float psv = w_div_nn(pnum, pden, true, bad);
    if (bad) {
      vsv = MZ_DIV(vnum, denom);
      psv = MZ_DIV(pnum, pden);
    }
    tsc[n * A + x] = MZ_ADD(vsv, psv);
  }
}

template <int A, int E, int H, int S, int G>
__global__ void __launch_bounds__(G == 8 ? 32 * (8 + kWMaxProducers) : 32 * (kWMaxWarps + kWMaxProducers))
    warp_search_kernel(const __grid_constant__ LaneArgs a) {
  using W = WShape<A, E, H, S>;
  constexpr int F = W::F, F4 = W::F4;
  constexpr int kWG = G, kWT = 32 / G, U = 32 / G;
  static_assert(G == 8 || G == 16, "8 or 16 lanes per tree");
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  __shared__ unsigned cta_done;  // search warps of this CTA whose outputs are stored
  const LaneNet& net = a.net;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int l = lane & (kWG - 1);
  // warp roles: [0, SW) search (a warp owns kWT trees), [SW, nwarps) produce tie-break noise into the ring
  const int SW = a.walkers > 0 ? min(a.walkers, nwarps) : nwarps;
  const int PW = nwarps - SW;
  const int trees = kWT * SW;
  __shared__ __align__(8) uint64_t nz_full[kWRing], nz_empty[kWRing];
  const int N = a.N, NS = a.p.num_simulations;
  const int row0 = blockIdx.x * trees;
  const float gamma = a.p.discount;
  const int act_kind = net.activation;
  const WSmem<A, E, H, S> M(net.packed_floats, NS, N, trees, PW > 0);
  const L2Layout<A, E> L(N);
  float* w = smem + M.w;
  float* wq = smem + M.wq;
  float* pbc = smem + M.pbc;
  float* blocks = smem + M.blocks;

  // ---- prologue: weights by one TMA bulk copy, pb_c table, tree init, re-layout of the 4 two-head matrices
  if (tid == 0) {
    cta_done = 0u;
    for (int i = 0; i < kWRing; ++i) {
      mbar_init(&nz_full[i], 32);          // every lane of the producing warp arrives
      mbar_init(&nz_empty[i], (uint32_t)(32 * SW));  // every lane of every search warp arrives (releases its own reads)
    }
    mbar_init(&wbar, 1);
    mbar_expect_tx(&wbar, (uint32_t)(round_up(net.packed_floats, 4) * 4));
    tma_bulk_g2s(w, a.packed, (uint32_t)(round_up(net.packed_floats, 4) * 4), &wbar);
  }
  for (int n = tid; n < NS + 2; n += blockDim.x) pbc[n] = pbc_explore((float)n, a.p.pb_c_init, a.p.pb_c_base);
  for (int tr = warp; tr < trees; tr += nwarps) {  // a warp per tree: no per-element division
    uint32_t* ub = reinterpret_cast<uint32_t*>(blocks + (size_t)tr * L.stride);
    for (int o = lane; o < L.stride; o += 32) {
      uint32_t v = 0u;
      if (o >= L.childs && o < L.raw && ((o - L.childs) & 3) == 0) v = kNoChild << 16;
      if (o < L.childs && (o & 3) == 3) v = 0xFFFFFFFFu;
      ub[o] = v;
    }
  }
  const bool searcher = warp < SW;
  const int ti = searcher ? warp * kWT + lane / kWG : 0;  // tree inside the CTA
  const bool live = row0 + ti < a.B;
  const int b = min(row0 + ti, a.B - 1);  // surplus groups shadow the last tree (no global writes)
  float* sc = smem + M.scratch + (size_t)ti * W::sStride;
  if (searcher) {
    for (int i = l; i < W::sStride; i += kWG) sc[i] = 0.0f;
    __syncwarp();
    for (int i = l; i < net.obs_dim; i += kWG) sc[W::sO1 + i] = a.obs[(size_t)b * net.obs_dim + i];  // obs_dim <= 160
  }
  __syncthreads();  // thread 0 initialised the mbarrier: it must exist before any other thread polls it
  mbar_wait(&wbar, 0);
  for (int i = tid; i < W::WQ; i += blockDim.x) {
    const int m = i < W::D2 ? 0 : (i < W::P1 ? 1 : (i < W::P2 ? 2 : 3));
    const int base = m == 0 ? W::D1 : (m == 1 ? W::D2 : (m == 2 ? W::P1 : W::P2));
    const int r = (i - base) >> 5, c = (i - base) & 31;
    float v = 0.0f;
    if (m == 0) {
      v = w[(c < H ? net.dyn_ns[0] : net.dyn_r[0]).off + r * H + (c & (H - 1))];
    } else if (m == 1) {
      if (c < E) v = w[net.dyn_ns[1].off + r * W::E4 + c];
      else if (c - E < F4) v = w[net.dyn_r[1].off + r * F4 + (c - E)];
    } else if (m == 2) {
      v = w[(c < H ? net.pred_v[0] : net.pred_pi[0]).off + r * H + (c & (H - 1))];
    } else {
      if (c < F4) v = w[net.pred_v[1].off + r * F4 + c];
      else if (c - F4 < W::A4) v = w[net.pred_pi[1].off + r * W::A4 + (c - F4)];
    }
    wq[i] = v;
  }
  __syncthreads();  // the last CTA-wide barrier before the dump

  const bool in_kernel_noise = PW > 0 && NS > 0 && a.K > 0;
  float* ring = smem + M.ring;
  uint32_t* ringc = reinterpret_cast<uint32_t*>(smem + M.ringc);
  if (!searcher) {
    // ---- noise producers: warp pw fills the ring slot of simulations pw, pw + PW, ... — lane t = tree t of the CTA.
    // The key chain depends only on (key, global row, simulation, depth), never on the tree, so it runs ahead of the
    // search in otherwise idle issue slots (before: a 37 us pre-pass kernel and a 26 MB table in HBM).
    if (in_kernel_noise) {
      const int pw = warp - SW;
      const bool tv = lane < trees;
      const uint32_t grow = (uint32_t)(a.p.batch_offset + min(row0 + lane, a.B - 1));
      for (int s = pw; s < NS; s += PW) {
        const int slot = s % kWRing, use = s / kWRing;
        if (use > 0) mbar_wait(&nz_empty[slot], (uint32_t)((use - 1) & 1));
        const uint32_t sk0 = a.p.sim_keys != nullptr ? a.p.sim_keys[2 * s] : a.ik.w[2 * s];
        const uint32_t sk1 = a.p.sim_keys != nullptr ? a.p.sim_keys[2 * s + 1] : a.ik.w[2 * s + 1];
        if (tv)
          noise_row(a.p, A, min(a.K, s + 1), sk0, sk1, grow, ring + (size_t)(slot * trees + lane) * kWRow,
                    ringc + (size_t)(slot * trees + lane) * 2);
        mbar_arrive(&nz_full[slot]);
      }
    }
  } else {
  float* blk = blocks + (size_t)ti * L.stride;
  float4* nodes = reinterpret_cast<float4*>(blk + L.nodes);
  float4* childs = reinterpret_cast<float4*>(blk + L.childs);
  float* traw = blk + L.raw;
  float* tlog = blk + L.logits;
  float* temb = blk + L.emb;
  float* troot = blk + L.root;
#if MZ_WARP_CACHED
  float* tsc = blk + L.scores;
  uint32_t* tpath = reinterpret_cast<uint32_t*>(blk + L.path);
#endif
  SearchParams p = a.p;
  p.batch_offset += b;

  // ---- root inference (muax/model.py:251-263): repr (runtime obs_dim, one layer) -> min-max -> pred -> value head
  {
    const LLayer& l0 = net.repr[0];
    if (U * l < E) {
      const float* wj = w + l0.off + U * l;
      float acc[U], wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = 0.0f;
      for (int k = 0; k < l0.K; ++k) {
        const float x = sc[W::sO1 + k];
        w_ldv<U>(wj + k * l0.out4, wv);
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = MZ_FMA(x, wv[u], acc[u]);
      }
      w_ldv<U>(wj + l0.K * l0.out4, wv);
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = MZ_ADD(acc[u], wv[u]);
      w_stv<U>(sc + W::sH + U * l, acc);
    }
  }
  __syncwarp();
  float v[E];
  w_load<E>(sc + W::sH, v);
  w_minmax<E>(v, net.repr_minmax != 0);
  __syncwarp();  // every lane has read the Representation output before Prediction overwrites the hidden buffer
  for (int i = l; i < 32; i += kWG) sc[W::sO1 + i] = 0.0f;  // the reward head of the root reads zeros, result unused
  w_prediction<A, E, H, S, G>(wq, sc, l, v, act_kind);
  {
    const float hs = w_heads<A, E, H, S, G>(sc, l);
    const float rv = __shfl_sync(0xffffffffu, hs, (lane & ~(kWG - 1)) + 1);
    const float* bufP = sc + W::sO2 + F4;
    // policy prologue (A.2) + node 0: every lane of the tree computes and stores the same values
    if (live && l == 0 && a.root_value_out != nullptr) {
      a.root_value_out[b] = rv;
      for (int q = 0; q < a.n_peers; ++q)
        *reinterpret_cast<float*>(reinterpret_cast<char*>(a.root_value_out + b) + a.peer_delta[q]) = rv;
    }
    float mx = -mz_inf();
#pragma unroll
    for (int x = 0; x < A; ++x) mx = fmaxf(mx, bufP[x]);
    float sum = 0.0f;
#pragma unroll
    for (int x = 0; x < A; ++x) sum = MZ_ADD(sum, mz_expf(MZ_SUB(bufP[x], mx)));
    const uint8_t* inv = a.invalid != nullptr ? a.invalid + (size_t)b * A : nullptr;
    const float* inj = a.noise != nullptr ? a.noise + (size_t)b * A : nullptr;
    float g[A], gsum = 0.0f;
    if (inj == nullptr) {
#pragma unroll
      for (int x = 0; x < A; ++x) {
        g[x] = gamma_draw(p.aux_key0, p.aux_key1, (uint32_t)((long)p.batch_offset * A + x), p.dirichlet_alpha);
        gsum = MZ_ADD(gsum, g[x]);
      }
    }
    float lgt[A], lmax = -mz_inf();
#pragma unroll
    for (int x = 0; x < A; ++x) {
      const float prob = MZ_DIV(mz_expf(MZ_SUB(bufP[x], mx)), sum);
      const float nzv = inj != nullptr ? inj[x] : (gsum > 0.0f ? MZ_DIV(g[x], gsum) : MZ_DIV(1.0f, (float)A));
      troot[x] = nzv;
      const float noisy = MZ_ADD(MZ_MUL(MZ_SUB(1.0f, p.dirichlet_fraction), prob), MZ_MUL(p.dirichlet_fraction, nzv));
      lgt[x] = mz_logf(fmaxf(noisy, MZ_F32_TINY));
      lmax = fmaxf(lmax, lgt[x]);
    }
    float m2 = -mz_inf();
#pragma unroll
    for (int x = 0; x < A; ++x) {
      const bool iv = inv != nullptr && inv[x] != 0;
      if (inv != nullptr) lgt[x] = iv ? -MZ_F32_MAX : MZ_SUB(lgt[x], lmax);
      troot[A + x] = iv ? 1.0f : 0.0f;
      tlog[x] = lgt[x];
      m2 = fmaxf(m2, lgt[x]);
    }
    float s2 = 0.0f;
#pragma unroll
    for (int x = 0; x < A; ++x) s2 = MZ_ADD(s2, mz_expf(MZ_SUB(lgt[x], m2)));
#pragma unroll
    for (int x = 0; x < A; ++x) {
      float4 c = childs[x];
      c.y = MZ_DIV(mz_expf(MZ_SUB(lgt[x], m2)), s2);
      childs[x] = c;
    }
    *reinterpret_cast<float4*>(temb) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
    for (int e = 4; e < E; e += 4) *reinterpret_cast<float4*>(temb + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
    traw[0] = rv;
    nodes[0] = make_float4(__int_as_float(1), rv, pbc[1], __uint_as_float(0xFFFFFFFFu));
#if MZ_WARP_CACHED
    w_node_scores<A>(nodes, childs, tsc, 0, gamma);  // every lane of the tree: same values, same stores
#endif
  }
  __syncwarp();

  const bool use_table = (a.noise_table != nullptr || in_kernel_noise) && NS > 0;
  static_assert(kGNoiseFloats == 32, "a noise row is dealt out U floats per lane");
  using NzVec = typename std::conditional<U == 4, float4, float2>::type;
  const NzVec* ntab = reinterpret_cast<const NzVec*>(a.noise_table) + (size_t)b * NS * G;
  NzVec nz_next{};
  if (use_table && !in_kernel_noise) {
    // launched with programmatic stream serialisation: everything above overlapped the noise pre-pass
    asm volatile("griddepcontrol.wait;" ::: "memory");
    nz_next = __ldcs(ntab + l);
  }
  const int max_depth = p.max_depth > 0 ? p.max_depth : NS;
  // selection is lane-parallel over the actions: lane x of every AP-lane subgroup scores child x
  constexpr int AP = A <= 2 ? 2 : (A <= 4 ? 4 : 8);
  static_assert(A <= 8 && AP <= G, "one lane per action inside a tree group");
  const int ax = l & (AP - 1);
  const bool axv = ax < A;
  const int axs = axv ? ax : A - 1;
  const int abase = lane & ~(AP - 1);
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(G - 1));
  const bool my_root_invalid = axv && troot[A + axs] != 0.0f;
  const float* nzrow = sc + W::sNz;
  const uint32_t* contp = nullptr;
  float* slot = sc + W::sCold;

  // ---- simulations: no CTA barrier from here to the dump
  for (int sim = 0; sim < NS; ++sim) {
    if (in_kernel_noise) {
      const int rs = sim % kWRing;
      mbar_wait(&nz_full[rs], (uint32_t)((sim / kWRing) & 1));
      nzrow = ring + (size_t)(rs * trees + ti) * kWRow;
      contp = ringc + (size_t)(rs * trees + ti) * 2;
    } else if (use_table) {
      reinterpret_cast<NzVec*>(sc + W::sNz)[l] = nz_next;
      __syncwarp();
      if (sim + 1 < NS) nz_next = __ldcs(ntab + (size_t)(sim + 1) * G + l);
      contp = a.cont_keys + ((size_t)b * NS + sim) * 2;
    }
    int parent, action = 0, next = 0, depth = 0;
    {
      // simulate (A.3) with muzero_action_selection (A.5) + qtransform_by_parent_and_siblings (A.6)
      int node = 0;
      for (;;) {
#if MZ_WARP_CACHED
        // the level's scores were left by the backup that last changed this node (w_node_scores): one load + the noise
        const float base = tsc[node * A + axs];
        const uint32_t chx = __float_as_uint(childs[node * A + axs].x);
        const float* nzp = nzrow + depth * A;
        if (!(use_table && depth < a.K)) {
          l2_noise_cold<A>(a.p.sim_keys != nullptr ? a.p.sim_keys[2 * sim] : a.ik.w[2 * sim],
                           a.p.sim_keys != nullptr ? a.p.sim_keys[2 * sim + 1] : a.ik.w[2 * sim + 1],
                           (uint32_t)p.global_batch, (uint32_t)p.batch_offset, p.prng_mode, depth,
                           use_table ? a.K : -1, contp, slot);
          nzp = slot + 2;
        }
        float s = MZ_ADD(base, nzp[axs]);
        if (depth == 0 && my_root_invalid) s = -mz_inf();
#else
        const float4 nd = nodes[node];
        const float4 ch = childs[node * A + axs];
        const uint32_t chx = __float_as_uint(ch.x);
        const float* nzp = nzrow + depth * A;
        if (!(use_table && depth < a.K)) {
          l2_noise_cold<A>(a.p.sim_keys != nullptr ? a.p.sim_keys[2 * sim] : a.ik.w[2 * sim],
                           a.p.sim_keys != nullptr ? a.p.sim_keys[2 * sim + 1] : a.ik.w[2 * sim + 1],
                           (uint32_t)p.global_batch, (uint32_t)p.batch_offset, p.prng_mode, depth,
                           use_table ? a.K : -1, contp, slot);
          nzp = slot + 2;
        }
        const float nz = nzp[axs];
        const int vis = (int)(__float_as_uint(ch.x) & 0xFFFFu);
        const bool seen = axv && vis > 0;
        const float q = MZ_ADD(ch.w, MZ_MUL(gamma, ch.z));
        // min / max over the parent value and the visited children's q: unvisited lanes contribute NaN, which
        // fminf / fmaxf drop
        const float qn = seen ? q : mz_nan();
        float lo = fminf(nd.y, qn), hi = fmaxf(nd.y, qn);
#pragma unroll
        for (int r = 1; r < AP; ++r) {
          const float o = __shfl_xor_sync(gmask, qn, r);
          lo = fminf(lo, o);
          hi = fmaxf(hi, o);
        }
        const float denom = fmaxf(MZ_SUB(hi, lo), 1e-8f);
        const float vnum = MZ_SUB(seen ? q : lo, lo);
        const float pnum = MZ_MUL(nd.z, ch.y);
        const float pden = (float)(vis + 1);  // in [1, 65536]
        bool bad = false;
        float vsv = w_div_nn(vnum, denom, false, bad);
        float psv = w_div_nn(pnum, pden, true, bad);
        if (bad) {
          vsv = MZ_DIV(vnum, denom);
          psv = MZ_DIV(pnum, pden);
        }
        float s = MZ_ADD(MZ_ADD(vsv, psv), nz);
        if (depth == 0 && my_root_invalid) s = -mz_inf();
#endif
        float bestv = __shfl_sync(gmask, s, abase);
        uint32_t cx = __shfl_sync(gmask, chx, abase);
        int best = 0;
#pragma unroll
        for (int j = 1; j < A; ++j) {
          const float sj = __shfl_sync(gmask, s, abase + j);
          const uint32_t cj = __shfl_sync(gmask, chx, abase + j);
          const bool better = sj > bestv;
          bestv = better ? sj : bestv;
          best = better ? j : best;
          cx = better ? cj : cx;
        }
        action = best;
        const uint32_t ci = cx >> 16;
#if MZ_WARP_CACHED
        tpath[depth] = ((uint32_t)node << 8) | (uint32_t)best;  // every lane of the tree stores the same word
#endif
        ++depth;
        if (ci == kNoChild || depth >= max_depth) {
          next = ci == kNoChild ? sim + 1 : (int)ci;
          break;
        }
        node = (int)ci;
      }
      parent = node;
      if (live && l == 0) a.out.sim_depth[(size_t)b * NS + sim] = depth;
    }
    __syncwarp();
    if (in_kernel_noise) mbar_arrive(&nz_empty[sim % kWRing]);  // this lane is done with the ring slot
    // recurrent_fn (muax/model.py:265-282): Dynamic -> min-max -> Prediction -> reward / value transforms
    float x[E];
    w_load<E>(temb + parent * E, x);
    float acc[U];
    w_dense<E, U>(wq + W::D1, l, x, E + action, acc);
    w_bias_act_store<U>(wq + W::D1 + (E + A) * 32, l, acc, true, act_kind, sc + W::sH);
    __syncwarp();
    {
      float h[H];
      w_load<H>(sc + W::sH + (U * l < E ? 0 : H), h);
      w_dense<H, U>(wq + W::D2, l, h, -1, acc);
      w_bias_act_store<U>(wq + W::D2 + H * 32, l, acc, false, 0, sc + W::sO1);
    }
    __syncwarp();
    w_minmax_dist<E, G>(sc + W::sO1, sc + W::sE, l, net.dyn_minmax != 0, x);
    w_prediction<A, E, H, S, G>(wq, sc, l, x, act_kind);
    const float hs = w_heads<A, E, H, S, G>(sc, l);
    const float reward = __shfl_sync(0xffffffffu, hs, lane & ~(kWG - 1));
    const float value = __shfl_sync(0xffffffffu, hs, (lane & ~(kWG - 1)) + 1);
    {
      // expand (A.3)
      const float* bufP = sc + W::sO2 + F4;
      float lg[A], mx = -mz_inf();
#pragma unroll
      for (int xx = 0; xx < A; ++xx) {
        lg[xx] = bufP[xx];
        mx = fmaxf(mx, lg[xx]);
      }
      float ex[A], sum = 0.0f;
#pragma unroll
      for (int xx = 0; xx < A; ++xx) {
        ex[xx] = mz_expf(MZ_SUB(lg[xx], mx));
        sum = MZ_ADD(sum, ex[xx]);
      }
      float pb[A];
      bool badp = !(sum >= 1.0f && sum <= (float)A);  // the largest logit contributes exp(0) = 1
#pragma unroll
      for (int xx = 0; xx < A; ++xx) pb[xx] = w_div_nn(ex[xx], sum, true, badp);
      if (badp) {
#pragma unroll
        for (int xx = 0; xx < A; ++xx) pb[xx] = MZ_DIV(ex[xx], sum);
      }
#pragma unroll
      for (int xx = 0; xx < A; ++xx) {
        tlog[next * A + xx] = lg[xx];
        float4 c = childs[next * A + xx];
        c.y = pb[xx];
        childs[next * A + xx] = c;
      }
#pragma unroll
      for (int e = 0; e < E; e += 4)
        *reinterpret_cast<float4*>(temb + next * E + e) = make_float4(x[e], x[e + 1], x[e + 2], x[e + 3]);
      traw[next] = value;
      const int nvis = __float_as_int(nodes[next].x) + 1;
      nodes[next] = make_float4(__int_as_float(nvis), value, pbc[min(nvis, NS + 1)],
                                __uint_as_float(((uint32_t)parent << 8) | (uint32_t)action));
      {
        float4 c = childs[parent * A + action];
        c.x = __uint_as_float(((uint32_t)next << 16) | (__float_as_uint(c.x) & 0xFFFFu));
        c.w = reward;
        childs[parent * A + action] = c;
      }
#if MZ_WARP_CACHED
      // backward (A.3), lane-parallel over the levels of the path the walk recorded: level d (edge tpath[d]: node pn,
      // action) belongs to lane d of the tree's lanes, kWG levels per round starting from the leaf.  The only chain is
      // the discounted return G_d = r_d + gamma * G_(d+1): every lane runs it over the round's rewards (shared memory
      // broadcasts) and keeps its own level's value.  The value an edge stores for its child is the node value the
      // lane one level deeper has just computed (one shuffle).  A lane then refreshes the cached selection scores of
      // its node — its record and one of its children changed, by this very lane — and the lane at pseudo-level
      // `depth` those of the (re)expanded node.  Same operations per level as the serial loop, ~190 instructions per
      // simulation instead of ~35 per level + the scoring on the walk's dependent chain.
      {
        float* sr = sc + W::sE;  // the round's edge rewards (the heads' scratch is free until the next simulation)
        const int topmax = __reduce_max_sync(0xffffffffu, (depth / kWG) * kWG);  // the trees of a warp differ in depth
        float G_in = value, cv_in = value;
        for (int d0 = topmax; d0 >= 0; d0 -= kWG) {
          const int d = d0 + l;
          const bool valid = d < depth;
          const uint32_t pa = tpath[min(d, depth - 1)];
          const int pn = (int)(pa >> 8), e2 = pn * A + (int)(pa & 0xFFu);
          float4 c = childs[e2];
          const float4 pd = nodes[pn];
          sr[l] = c.w;
          __syncwarp();
          const int cnt = min(kWG, depth - d0);  // levels of this tree in the round (<= 0: none)
          const int cntmax = __reduce_max_sync(0xffffffffu, cnt);
          float Gr = G_in, myG = G_in;
          for (int e = cntmax - 1; e >= 0; --e) {
            if (e < cnt) Gr = MZ_ADD(sr[e], MZ_MUL(gamma, Gr));
            if (e == l) myG = Gr;
          }
          G_in = Gr;
          const int ci = __float_as_int(pd.x);
          const float count = (float)ci;
          const float pnum = MZ_ADD(MZ_MUL(pd.y, count), myG), pden = MZ_ADD(count, 1.0f);
          // pden = count + 1 with count in [1, 65535]: only the numerator (any sign) needs the range check
          const uint32_t un = __float_as_uint(pnum) & 0x7fffffffu;
          float pv = div_core(pnum, pden);
          if (!(un == 0u || (un - 0x30800000u) < 0x1E800000u)) pv = MZ_DIV(pnum, pden);
          float cv = __shfl_down_sync(0xffffffffu, pv, 1, kWG);  // the node one level deeper, after its update
          if (l == kWG - 1) cv = cv_in;                           // ... computed in the round before
          if (d == depth - 1) cv = value;                         // ... or the new node
          if (valid) {
            nodes[pn] = make_float4(__int_as_float(ci + 1), pv, pbc[min(ci + 1, NS + 1)], pd.w);
            c.x = __uint_as_float(__float_as_uint(c.x) + 1u);
            c.z = cv;
            childs[e2] = c;
          }
          cv_in = __shfl_sync(0xffffffffu, pv, 0, kWG);
          if (d <= depth) w_node_scores<A>(nodes, childs, tsc, valid ? pn : next, gamma);
          __syncwarp();
        }
      }
    }
#else
      // backward (A.3)
      int index = next;
      float G_ = value, child_value = value;
      while (index != 0) {
        const uint32_t pa = __float_as_uint(nodes[index].w);
        const int pn = (int)(pa >> 8);
        const int e2 = pn * A + (int)(pa & 0xFFu);
        const float4 pd = nodes[pn];
        float4 c = childs[e2];
        const int ci = __float_as_int(pd.x);
        const float count = (float)ci;
        G_ = MZ_ADD(c.w, MZ_MUL(gamma, G_));
        const float pnum = MZ_ADD(MZ_MUL(pd.y, count), G_), pden = MZ_ADD(count, 1.0f);
        // pden = count + 1 with count in [1, 65535]: only the numerator (any sign) needs the range check
        const uint32_t un = __float_as_uint(pnum) & 0x7fffffffu;
        float pv = div_core(pnum, pden);
        if (!(un == 0u || (un - 0x30800000u) < 0x1E800000u)) pv = MZ_DIV(pnum, pden);
        nodes[pn] = make_float4(__int_as_float(ci + 1), pv, pbc[min(ci + 1, NS + 1)], pd.w);
        c.x = __uint_as_float(__float_as_uint(c.x) + 1u);
        c.z = child_value;
        childs[e2] = c;
        child_value = pv;
        index = pn;
      }
    }
#endif
    __syncwarp();
  }

  // ---- policy epilogue (A.2): visit_probs -> temperature -> categorical
  {
    float total = 0.0f;
    float vc[A];
#pragma unroll
    for (int x = 0; x < A; ++x) {
      vc[x] = (float)(__float_as_uint(childs[x].x) & 0xFFFFu);
      total = MZ_ADD(total, vc[x]);
    }
    float wgt[A], lw[A], lmax = -mz_inf();
#pragma unroll
    for (int x = 0; x < A; ++x) {
      wgt[x] = total > 0.0f ? MZ_DIV(vc[x], fmaxf(total, 1.0f)) : MZ_DIV(1.0f, (float)A);
      lw[x] = mz_logf(fmaxf(wgt[x], MZ_F32_TINY));
      lmax = fmaxf(lmax, lw[x]);
    }
    const float temp = fmaxf(MZ_F32_TINY, p.temperature);
    int best = 0;
    float bestv = 0.0f;
#pragma unroll
    for (int x = 0; x < A; ++x) {
      const uint32_t bits = bits_word(p.final_key0, p.final_key1, (uint32_t)p.global_batch * (uint32_t)A,
                                      (uint32_t)((long)p.batch_offset * A + x), p.prng_mode);
      const float s = MZ_ADD(mz_bits_to_gumbel(bits), MZ_DIV(MZ_SUB(lw[x], lmax), temp));
      if (x == 0 || s > bestv) {
        bestv = s;
        best = x;
      }
    }
    if (live && l == 0) {
#pragma unroll
      for (int x = 0; x < A; ++x) a.weights_out[(size_t)b * A + x] = wgt[x];
      a.action_out[b] = best;
      for (int q = 0; q < a.n_peers; ++q) {  // the all-gather of the sharded act, done by the kernel itself
        char* wp = reinterpret_cast<char*>(a.weights_out + (size_t)b * A) + a.peer_delta[q];
#pragma unroll
        for (int x = 0; x < A; ++x) reinterpret_cast<float*>(wp)[x] = wgt[x];
        *reinterpret_cast<int32_t*>(reinterpret_cast<char*>(a.action_out + b) + a.peer_delta[q]) = best;
      }
    }
  }

    // ---- completion flag of the sharded act: the exchange is this kernel's own peer stores, so "everybody's rows have
    // arrived" is a flag per source rank in every rank's gather buffer, set by the LAST CTA of the source's kernel.
    // Every lane fences its output stores system-wide, the warp's lane 0 counts the warp in, the CTA's last warp counts
    // the CTA in, the grid's last CTA publishes (the fence-then-count pattern of a last-block reduction, at system scope).
    if (a.flag != nullptr) {
      __threadfence_system();
      __syncwarp();
      if (lane == 0 && atomicAdd(&cta_done, 1u) == (unsigned)(SW - 1)) {
        __threadfence();
        if (atomicAdd(a.done_counter, 1u) == gridDim.x - 1) {
          atomicExch(a.done_counter, 0u);
          __threadfence_system();
          int32_t* mine = a.flag + a.flag_rank;
          asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(mine), "r"(a.flag_step) : "memory");
          for (int q = 0; q < a.n_peers; ++q)
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(reinterpret_cast<char*>(mine) + a.peer_delta[q]),
                         "r"(a.flag_step)
                         : "memory");
        }
      }
    }

    // ---- dump: every warp unpacks its own trees into the mctx SoA arrays as soon as it is done (no CTA barrier:
    // the act ends with its slowest warp, the others have written their trees by then)
    if (a.dump_tree) {
      __syncwarp();
      const Tree& o = a.out;
      for (int tt = 0; tt < kWT; ++tt) {
        const int tr = warp * kWT + tt;
        if (row0 + tr >= a.B) break;
        const float* tb = blocks + (size_t)tr * L.stride;
        for (int n = lane; n < N; n += 32) {
          const float4 nd = reinterpret_cast<const float4*>(tb + L.nodes)[n];
          const size_t g = (size_t)(row0 + tr) * o.N + n;
          const uint32_t pa = __float_as_uint(nd.w);
          o.node_visits[g] = __float_as_int(nd.x);
          o.parents[g] = pa == 0xFFFFFFFFu ? -1 : (int)(pa >> 8);
          o.action_from_parent[g] = pa == 0xFFFFFFFFu ? -1 : (int)(pa & 0xFFu);
          o.raw_values[g] = (tb + L.raw)[n];
          o.node_values[g] = nd.y;
        }
        for (int k = lane; k < N * A; k += 32) {
          const float4 c = reinterpret_cast<const float4*>(tb + L.childs)[k];
          const size_t g = (size_t)(row0 + tr) * o.N * A + k;
          const uint32_t cx = __float_as_uint(c.x);
          const bool has = (cx >> 16) != kNoChild;
          o.children_index[g] = has ? (int)(cx >> 16) : -1;
          o.children_visits[g] = (int)(cx & 0xFFFFu);
          o.children_prior_logits[g] = (tb + L.logits)[k];
          o.children_prior_probs[g] = c.y;
          o.children_values[g] = c.z;
          o.children_rewards[g] = c.w;
          o.children_discounts[g] = has ? gamma : 0.0f;
        }
        for (int k = lane; k < N * E; k += 32)
          o.embeddings[(size_t)(row0 + tr) * o.N * E + k] = (tb + L.emb)[k];
        if (lane < A) {
          const float* rt = tb + L.root;
          o.root_noise[(size_t)(row0 + tr) * A + lane] = rt[lane];
          o.root_invalid[(size_t)(row0 + tr) * A + lane] = rt[A + lane] != 0.0f ? 1 : 0;
        }
      }
    }
  }  // search warps
}

// ---------------------------------------------------------------------------------------- host side

struct WarpVariant {
  int A, E, H, S;
  void* fn;     // 8 lanes per tree
  void* fn16;   // 16 lanes per tree
  size_t (*smem)(int packed_floats, int NS, int N, int trees, bool producers);
};

template <int A, int E, int H, int S>
size_t warp_smem(int packed_floats, int NS, int N, int trees, bool producers) {
  return (size_t)WSmem<A, E, H, S>(packed_floats, NS, N, trees, producers).total * 4;
}

#define MZ_WARP_VARIANT(A, E, H, S) \
  WarpVariant { A, E, H, S, (void*)warp_search_kernel<A, E, H, S, 8>, (void*)warp_search_kernel<A, E, H, S, 16>, \
                &warp_smem<A, E, H, S> }

static const std::vector<WarpVariant>& warp_variants() {
  static const std::vector<WarpVariant> v = {
      MZ_WARP_VARIANT(2, 8, 16, 10),   // CartPole-v1 stock nets (README / BASELINE headline)
      MZ_WARP_VARIANT(4, 8, 16, 10),   // 4-action environments with the stock nets
      MZ_WARP_VARIANT(3, 8, 16, 10),
      MZ_WARP_VARIANT(5, 8, 16, 10),   // up to 6 actions (one lane per action inside an 8-lane subgroup)
      MZ_WARP_VARIANT(6, 8, 16, 10),
      MZ_WARP_VARIANT(2, 8, 16, 5),    // support_size 5: 11-logit heads leave room for 16-wide embeddings
      MZ_WARP_VARIANT(3, 8, 16, 5),
      MZ_WARP_VARIANT(4, 8, 16, 5),
      MZ_WARP_VARIANT(2, 16, 16, 5),
      MZ_WARP_VARIANT(4, 16, 16, 5),
  };
  return v;
}

struct WarpImpl {
  LaneNet net{};
  std::vector<LPackDesc> descs;
  const WarpVariant* variant = nullptr;
};

static bool plan_stack(const mz_stack& s, int in_x, int extra, LLayer* out, std::vector<LPackDesc>& descs, int& off,
                       int& hmax) {
  if (s.n_layers < 1 || s.n_layers > kLMaxLayers) return false;
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

int warp_init(WarpState& st, const Net& net, int device, std::string* err) {
  st.available = false;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    *err = "cudaGetDeviceProperties failed";
    return 1;
  }
  st.max_smem = (int)prop.sharedMemPerBlockOptin;
  st.num_sms = prop.multiProcessorCount;
  if (net.obs_dim <= 0 || net.obs_dim > 160) return 0;  // the Representation runs inside the kernel
  if (net.repr.n_layers != 1 || net.pred_v.n_layers != 2 || net.pred_pi.n_layers != 2 || net.dyn_ns.n_layers != 2 ||
      net.dyn_r.n_layers != 2)
    return 0;
  const int H = net.pred_v.out_dim[0];
  if (net.pred_pi.out_dim[0] != H || net.dyn_ns.out_dim[0] != H || net.dyn_r.out_dim[0] != H) return 0;
  const WarpVariant* variant = nullptr;
  for (const WarpVariant& v : warp_variants())
    if (v.A == net.num_actions && v.E == net.embed_dim && v.H == H && v.S == net.support_size) variant = &v;
  if (variant == nullptr) return 0;
  WarpImpl* impl = new WarpImpl();
  LaneNet& g = impl->net;
  g.obs_dim = net.obs_dim; g.E = net.embed_dim; g.A = net.num_actions; g.S = net.support_size;
  g.F = 2 * net.support_size + 1;
  g.activation = net.activation; g.repr_minmax = net.repr_minmax; g.dyn_minmax = net.dyn_minmax;
  int off = 0, hmax = 1;
  if (!plan_stack(net.repr, net.obs_dim, 0, g.repr, impl->descs, off, hmax) ||
      !plan_stack(net.pred_v, net.embed_dim, 0, g.pred_v, impl->descs, off, hmax) ||
      !plan_stack(net.pred_pi, net.embed_dim, 0, g.pred_pi, impl->descs, off, hmax) ||
      !plan_stack(net.dyn_ns, net.embed_dim, net.num_actions, g.dyn_ns, impl->descs, off, hmax) ||
      !plan_stack(net.dyn_r, net.embed_dim, net.num_actions, g.dyn_r, impl->descs, off, hmax)) {
    delete impl;
    return 0;
  }
  g.n_repr = 1; g.n_pred = 2; g.n_dyn = 2;
  g.Hmax = hmax;
  g.packed_floats = off;
  if (cudaFuncSetAttribute(variant->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, st.max_smem - 1024) != cudaSuccess ||
      cudaFuncSetAttribute(variant->fn16, cudaFuncAttributeMaxDynamicSharedMemorySize, st.max_smem - 1024) != cudaSuccess) {
    cudaGetLastError();
    delete impl;
    return 0;
  }
  if (cudaMalloc((void**)&st.packed, (size_t)round_up(off, 4) * 4 + 16) != cudaSuccess ||
      cudaMalloc((void**)&st.done_counter, 16) != cudaSuccess || cudaMemset(st.done_counter, 0, 16) != cudaSuccess) {
    cudaGetLastError();
    delete impl;
    *err = "warp engine: cudaMalloc(packed weights) failed";
    return 1;
  }
  impl->variant = variant;
  st.impl = impl;
  if (const char* wv = getenv("MZ_WARP_WARPS")) st.force_warps = std::max(0, std::min(kWMaxWarps, atoi(wv)));
  if (const char* wl = getenv("MZ_WARP_LANES")) st.lanes = atoi(wl) == 8 ? 8 : 16;
  if (const char* wp = getenv("MZ_WARP_PRODUCERS")) st.producers = std::max(0, std::min(kWMaxProducers, atoi(wp)));
  st.available = true;
  return 0;
}

void warp_destroy(WarpState& st) {
  if (st.packed) cudaFree(st.packed);
  if (st.done_counter) cudaFree(st.done_counter);
  st.done_counter = nullptr;
  if (st.noise_table) cudaFree(st.noise_table);
  if (st.cont_keys) cudaFree(st.cont_keys);
  st.packed = nullptr;
  st.noise_table = nullptr;
  st.cont_keys = nullptr;
  st.noise_capacity = 0;
  delete static_cast<WarpImpl*>(st.impl);
  st.impl = nullptr;
  st.available = false;
}

// Re-lays the raw fp32 blob out as row-padded [K + one-hot rows + bias][out4] matrices (one small kernel per layer).
int warp_pack(WarpState& st, const float* raw, cudaStream_t stream, int64_t* launches) {
  if (!st.available) return 0;
  const WarpImpl* impl = static_cast<const WarpImpl*>(st.impl);
  for (const LPackDesc& d : impl->descs) {
    const int total = (d.l.K + d.l.extra + 1) * d.l.out4;
    lane_pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(raw, st.packed, d);
    *launches += 1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Warps per CTA: enough that one wave of num_sms CTAs covers the batch, within shared memory.
static int warp_pick_warps(const WarpState& st, int B, int NS) {
  if (!st.available) return 0;
  const WarpImpl* impl = static_cast<const WarpImpl*>(st.impl);
  const int kWT = 32 / st.lanes;
  int fit = 0;
  for (int wv = 1; wv <= (st.lanes == 8 ? 8 : kWMaxWarps); ++wv)
    if (impl->variant->smem(impl->net.packed_floats, NS, NS + 1, kWT * wv, st.producers > 0) + 1024 <= (size_t)st.max_smem)
      fit = wv;
  if (fit == 0) return 0;
  if (st.force_warps > 0) return std::min(st.force_warps, fit);
  const int sms = std::max(1, st.num_sms);
  const int per_sm = (B + sms - 1) / sms;
  return std::max(1, std::min(fit, (per_sm + kWT - 1) / kWT));
}

bool warp_supported(const WarpState& st, const SearchParams& p, int B) {
  if (!st.available) return false;
  if (p.policy != MZ_POLICY_MUZERO || p.qtransform != MZ_QTRANSFORM_BY_PARENT_AND_SIBLINGS) return false;
  if (p.num_simulations + 1 >= (int)kNoChild) return false;
  return warp_pick_warps(st, B, p.num_simulations) > 0;
}

int warp_launch(WarpState& st, const Tree& out, const SearchParams& p, const SimKeys* inline_keys, const float* obs,
                const uint8_t* invalid, const float* noise, int32_t* action_out, float* weights_out,
                float* root_value_out, const std::vector<int64_t>& peers, bool dump_tree, cudaStream_t stream,
                int64_t* launches, std::string* err) {
  const WarpImpl* impl = static_cast<const WarpImpl*>(st.impl);
  const int B = out.B, NS = p.num_simulations, N = NS + 1, A = impl->net.A;
  LaneArgs a{};
  a.net = impl->net;
  a.packed = st.packed;
  a.out = out;
  a.p = p;
  if (inline_keys != nullptr) {  // the keys travel in the kernel parameters
    a.ik = *inline_keys;
    a.p.sim_keys = nullptr;
  }
  a.obs = obs;
  a.invalid = invalid;
  a.noise = noise;
  a.action_out = action_out;
  a.weights_out = weights_out;
  a.root_value_out = root_value_out;
  a.B = B;
  a.N = N;
  a.n_peers = (int)std::min<size_t>(peers.size(), 7);
  for (int i = 0; i < a.n_peers; ++i) a.peer_delta[i] = peers[i];
  a.flag = st.flag;
  a.flag_rank = st.flag_rank;
  a.flag_step = st.flag_step;
  a.done_counter = st.done_counter;
  a.dump_tree = dump_tree ? 1 : 0;
  a.K = std::min(16, kGNoiseFloats / A);
  if (const char* k = getenv("MZ_GROUP_K")) a.K = std::max(0, std::min(a.K, atoi(k)));
  const int warps = warp_pick_warps(st, B, NS);
  a.walkers = warps;
  if (NS > 0 && a.K > 0 && st.producers == 0) {
    if (p.sim_keys == nullptr) {
      *err = "warp engine: the noise pre-pass needs the simulate keys in device memory";
      return 1;
    }
    const size_t pairs = (size_t)B * NS;
    if (pairs > st.noise_capacity) {
      if (st.noise_table) cudaFree(st.noise_table);
      if (st.cont_keys) cudaFree(st.cont_keys);
      st.noise_table = nullptr;
      st.cont_keys = nullptr;
      st.noise_capacity = 0;
      if (cudaMalloc((void**)&st.noise_table, pairs * kGNoiseFloats * 4) != cudaSuccess ||
          cudaMalloc((void**)&st.cont_keys, pairs * 8) != cudaSuccess) {
        cudaGetLastError();
        *err = "warp engine: cudaMalloc(noise table) failed";
        return 1;
      }
      st.noise_capacity = pairs;
    }
    noise_table_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, stream>>>(p, B, A, a.K, st.noise_table, st.cont_keys);
    *launches += 1;
    a.noise_table = st.noise_table;
    a.cont_keys = st.cont_keys;
  }
  const int trees = (32 / st.lanes) * warps;
  const size_t smem = impl->variant->smem(impl->net.packed_floats, NS, N, trees, st.producers > 0);
  const int grid = (B + trees - 1) / trees;
  void* args[] = {&a};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * (warps + st.producers));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (a.noise_table != nullptr && !getenv("MZ_NO_PDL")) ? 1 : 0;  // overlap the prologue with the noise pre-pass
  const cudaError_t e = cudaLaunchKernelExC(&cfg, st.lanes == 8 ? impl->variant->fn : impl->variant->fn16, args);
  *launches += 1;
  if (e != cudaSuccess) {
    *err = std::string("warp engine launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

// One thread per source rank spins (system-scope acquire loads) until the rank's flag reaches `step`.  Bounded: a flag
// that never comes is a protocol bug or a dead peer, and must not hang the GPU.
__global__ void peer_wait_kernel(const int32_t* flags, int world, int step, int32_t* timed_out) {
  const int q = threadIdx.x;
  if (q >= world) return;
  for (unsigned spin = 0;; ++spin) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
    if (v >= step) return;
    if (spin > (1u << 24)) {  // ~ seconds
      if (timed_out != nullptr) atomicExch(timed_out, 1);
      return;
    }
    __nanosleep(200);
  }
}

int warp_peer_wait(WarpState& st, const int32_t* flags, int world, int step, cudaStream_t stream, int64_t* launches) {
  // same shared-memory carve-out as the search kernel it follows: a kernel with another preference makes the SMs
  // reconfigure their L1 / shared memory split before it may start
  static const bool once = [] {
    cudaFuncSetAttribute(peer_wait_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  (void)once;
  peer_wait_kernel<<<1, 32, 0, stream>>>(flags, world, step, reinterpret_cast<int32_t*>(st.done_counter) + 1);
  *launches += 1;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mz

// mz_lane2.cuh — compile-time specialisation of the lane engine for the stock muax MLP family
// (Representation: 1 linear; Prediction / Dynamic: two heads of Linear(H)-act-Linear(out); MuZero policy with
// qtransform_by_parent_and_siblings): the configuration BASELINE.json's metric is quoted on.
//
// What changes against mz_lane.cuh (profiles/r01_lane_v1_*: 240 instructions per selected level, 17k static
// instructions, `no_instruction` stalls, IPC 0.15 per warp):
//   * A, E, H, S are template parameters: every loop over actions / units / k is unrolled, no policy or
//     qtransform branches survive, the kernel is ~10x smaller;
//   * the tree is stored as 16-byte records — node {visits, value, sqrt(n)*pb_c(n), parent<<8|action},
//     child {index<<16|visits, prior prob, value, reward} — so one level of `simulate` is A+1 LDS.128 plus the
//     noise pair, and one level of `backward` is two LDS.128 + two STS.128;
//   * the exploration factor sqrt(n)*pb_c(n) is refreshed when the visit count changes (backup), taking the
//     dependent table look-up off the selection path;
//   * dense layers read all K activations up front (independent LDS) and run 4 FMA chains per warp.
// Same arithmetic, same order: bit-identical to every other engine (tests/test_gpu_parity.py).
#pragma once
#include "mz_lane.cuh"

namespace mz {

template <int A, int E>
struct L2Layout {  // float offsets inside one tree block (all 16-byte aligned)
  int nodes, childs, raw, logits, emb, root, stride;
  __host__ __device__ explicit L2Layout(int N) {
    int o = 0;
    nodes = o; o += 4 * N;
    childs = o; o += 4 * N * A;
    raw = o; o += round_up(N, 4);
    logits = o; o += round_up(N * A, 4);
    emb = o; o += round_up(N * E, 4);
    root = o; o += round_up(2 * A, 4);  // root_noise[A], root_invalid[A] (as floats 0/1)
    while (o % 32 != 4) o += 4;
    stride = o;
  }
};

constexpr uint32_t kNoChild = 0xFFFFu;

template <int A, int E, int H, int S>
struct L2Smem {  // float offsets of the CTA's shared memory
  static constexpr int F = 2 * S + 1;
  int w, pbc, in, hA, hB, ns, rlog, vlog, plog, er, ev, nz, sc, cold, blocks, total;
  __host__ __device__ L2Smem(int packed_floats, int NS, int obs_dim, int N) {
    int o = 0;
    w = o; o += round_up(packed_floats, 4);
    pbc = o; o += round_up(NS + 2, 4);
    in = o; o += (obs_dim > E ? obs_dim : E) * kLT;
    hA = o; o += H * kLT;
    hB = o; o += H * kLT;
    ns = o; o += E * kLT;
    rlog = o; o += F * kLT;
    vlog = o; o += F * kLT;
    plog = o; o += A * kLT;
    er = o; o += F * kLT;
    ev = o; o += F * kLT;
    nz = o; o += 2 * kLT * kGNoiseFloats;
    sc = o; o += 3 * kLT;
    cold = o; o += round_up(kLT * (A + 2), 4);
    blocks = o;
    total = o + kLT * L2Layout<A, E>(N).stride;
  }
};

// 4 output units [j0, j0+4) of a dense layer with compile-time K for this lane's tree.
template <int K, bool ACT>
__device__ __forceinline__ void l2_block(const float* __restrict__ wj, int out4, const float* in_col, int extra_row,
                                         int bias_row, int act_kind, float (&a)[4]) {
  float x[K];
#pragma unroll
  for (int k = 0; k < K; ++k) x[k] = in_col[k * kLT];
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float4 wv = *reinterpret_cast<const float4*>(wj + k * out4);
    a0 = MZ_FMA(x[k], wv.x, a0);
    a1 = MZ_FMA(x[k], wv.y, a1);
    a2 = MZ_FMA(x[k], wv.z, a2);
    a3 = MZ_FMA(x[k], wv.w, a3);
  }
  if (extra_row >= 0) {
    const float4 wv = *reinterpret_cast<const float4*>(wj + extra_row * out4);
    a0 = MZ_ADD(a0, wv.x); a1 = MZ_ADD(a1, wv.y); a2 = MZ_ADD(a2, wv.z); a3 = MZ_ADD(a3, wv.w);
  }
  const float4 bv = *reinterpret_cast<const float4*>(wj + bias_row * out4);
  a0 = MZ_ADD(a0, bv.x); a1 = MZ_ADD(a1, bv.y); a2 = MZ_ADD(a2, bv.z); a3 = MZ_ADD(a3, bv.w);
  if (ACT) {
    a0 = activate(a0, act_kind); a1 = activate(a1, act_kind); a2 = activate(a2, act_kind); a3 = activate(a3, act_kind);
  }
  a[0] = a0; a[1] = a1; a[2] = a2; a[3] = a3;
}

template <int OUT>
__device__ __forceinline__ void l2_store(float* out_col, int j0, const float (&a)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (j0 + i < OUT) out_col[(j0 + i) * kLT] = a[i];
}

// Layer 1 of a two-head module (both heads read `in`, K inputs [+ one-hot rows]) -> hA, hB ([H][32]).
template <int K, int H>
__device__ __forceinline__ void l2_layer1(const float* w, const LLayer& l0, const LLayer& l1, const float* in, int onehot,
                                          int act_kind, float* hA, float* hB, int lane, int warp, int nwarps) {
  constexpr int NB = H / 4;
  for (int g = warp; g < 2 * NB; g += nwarps) {
    const bool second = g >= NB;
    const LLayer& L = second ? l1 : l0;
    const int j0 = (second ? g - NB : g) * 4;
    float a[4];
    l2_block<K, true>(w + L.off + j0, H, in + lane, onehot >= 0 ? K + onehot : -1, K + L.extra, act_kind, a);
    l2_store<H>((second ? hB : hA) + lane, j0, a);
  }
}

// Layer 2: head 0 (OUT0 units from hA) and head 1 (OUT1 units from hB).
template <int H, int OUT0, int OUT1>
__device__ __forceinline__ void l2_layer2(const float* w, const LLayer& l0, const LLayer& l1, const float* hA,
                                          const float* hB, float* out0, float* out1, int lane, int warp, int nwarps) {
  constexpr int O40 = (OUT0 + 3) / 4 * 4, O41 = (OUT1 + 3) / 4 * 4;
  constexpr int NB0 = O40 / 4, NB1 = O41 / 4;
  for (int g = warp; g < NB0 + NB1; g += nwarps) {
    float a[4];
    if (g < NB0) {
      l2_block<H, false>(w + l0.off + g * 4, O40, hA + lane, -1, H, 0, a);
      l2_store<OUT0>(out0 + lane, g * 4, a);
    } else {
      const int j0 = (g - NB0) * 4;
      l2_block<H, false>(w + l1.off + j0, O41, hB + lane, -1, H, 0, a);
      l2_store<OUT1>(out1 + lane, j0, a);
    }
  }
}

template <int N_>
__device__ __forceinline__ void l2_minmax_col(const float* raw_col, bool enabled, float (&v)[N_]) {
#pragma unroll
  for (int k = 0; k < N_; ++k) v[k] = raw_col[k * kLT];
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
  bool bad = false;
#pragma unroll
  for (int k = 0; k < N_; ++k) {
    num[k] = MZ_SUB(v[k], lo);
    v[k] = div_try(num[k], scale, bad);
  }
  if (bad) {
#pragma unroll
    for (int k = 0; k < N_; ++k) v[k] = MZ_DIV(num[k], scale);
  }
}

template <int F>
__device__ __forceinline__ void l2_head_exps(const float* logits_col, float* e_col, int part, int parts) {
  float mx = logits_col[0];
#pragma unroll
  for (int j = 1; j < F; ++j) mx = fmaxf(mx, logits_col[j * kLT]);
  for (int j = part; j < F; j += parts) e_col[j * kLT] = mz_expf(MZ_SUB(logits_col[j * kLT], mx));
}

template <int F, int S>
__device__ __forceinline__ float l2_head_scalar(const float* e_col) {
  float e[F];
#pragma unroll
  for (int j = 0; j < F; ++j) e[j] = e_col[j * kLT];
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < F; ++j) s = MZ_ADD(s, e[j]);
  float pr[F];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < F; ++j) pr[j] = div_try(e[j], s, bad);
  if (bad) {
#pragma unroll
    for (int j = 0; j < F; ++j) pr[j] = MZ_DIV(e[j], s);
  }
  float x = 0.0f;
#pragma unroll
  for (int j = 0; j < F; ++j) x = MZ_ADD(x, MZ_MUL((float)(j - S), pr[j]));
  return mz_inv_scaling(x);
}

// The same head split over `parts` warps (the serial version above kept one warp busy for ~2.5k cycles per simulation
// while 7 idled): every warp recomputes the left-to-right softmax denominator (21 adds), then takes its share of the
// quotients and products (j - S) * (e[j] / s) -> prod_col; l2_head_final sums the products left to right.  Same
// operations in the same order as l2_head_scalar.
template <int F, int S>
__device__ __forceinline__ void l2_head_products(const float* e_col, float* prod_col, int part, int parts) {
  float e[F];
#pragma unroll
  for (int j = 0; j < F; ++j) e[j] = e_col[j * kLT];
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < F; ++j) s = MZ_ADD(s, e[j]);
  for (int j = part; j < F; j += parts) {
    bool bad = false;
    float pr = div_try(e_col[j * kLT], s, bad);
    if (bad) pr = MZ_DIV(e_col[j * kLT], s);
    prod_col[j * kLT] = MZ_MUL((float)(j - S), pr);
  }
}

template <int F, int S>
__device__ __forceinline__ float l2_head_final(const float* prod_col) {
  float pv[F];
#pragma unroll
  for (int j = 0; j < F; ++j) pv[j] = prod_col[j * kLT];
  float x = 0.0f;
#pragma unroll
  for (int j = 0; j < F; ++j) x = MZ_ADD(x, pv[j]);
  return mz_inv_scaling(x);
}

// CTA barrier, or a named barrier over the first `nthreads` threads of a warp role when the roles are split.
__device__ __forceinline__ void l2_bar(bool named, int id, int nthreads) {
  if (named)
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
  else
    __syncthreads();
}

// Cold path of the selection: tie-break noise for a level the pre-computed table does not cover (depth >= K, or
// no table at all).  Kept out of line so the hot loop stays small.  k0/k1 carry the jax key chain between levels.
// All state goes through this thread's shared-memory slot {key0, key1, noise[A]} (no stack frame).
template <int A>
__device__ __noinline__ void l2_noise_cold(const uint32_t* sim_key, uint32_t global_batch, uint32_t global_row,
                                           int prng_mode, int depth, int K, const uint32_t* cont, float* slot) {
  uint32_t k0 = __float_as_uint(slot[0]), k1 = __float_as_uint(slot[1]);
  if (K < 0 && depth == 0) split_key(sim_key[0], sim_key[1], global_batch, global_row, prng_mode, k0, k1);
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

#ifdef MZ_PHASE_CLOCKS
#define MZ_CLK(i) do { const long long t__ = clock64(); phase_acc[i] += t__ - phase_t; phase_t = t__; } while (0)
#else
#define MZ_CLK(i) do { } while (0)
#endif

template <int A, int E, int H, int S>
__global__ void __launch_bounds__(512) lane2_search_kernel(LaneArgs a) {
  constexpr int F = 2 * S + 1;
#ifdef MZ_PHASE_CLOCKS
  long long phase_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long phase_t = clock64();
  const long long kernel_t0 = phase_t;
#endif
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  const LaneNet& net = a.net;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int N = a.N, NS = a.p.num_simulations;
  const int row0 = blockIdx.x * kLT;
  const float gamma = a.p.discount;
  const int act_kind = net.activation;
  const L2Smem<A, E, H, S> M(net.packed_floats, NS, net.obs_dim, N);
  const L2Layout<A, E> L(N);
  float* w = smem + M.w;
  float* pbc = smem + M.pbc;
  float* bufIn = smem + M.in;
  float* hA = smem + M.hA;
  float* hB = smem + M.hB;
  float* bufNs = smem + M.ns;
  float* bufR = smem + M.rlog;
  float* bufV = smem + M.vlog;
  float* bufP = smem + M.plog;
  float* bufEr = smem + M.er;
  float* bufEv = smem + M.ev;
  float* nzbuf = smem + M.nz;
  float* sc_reward = smem + M.sc;
  float* sc_value = sc_reward + kLT;
  int32_t* sc_action = reinterpret_cast<int32_t*>(sc_value + kLT);
  float* blocks = smem + M.blocks;

  // ---- prologue
  const bool use_tma = a.dump_tree < 2;  // debug knob (MZ_NO_TMA): stage the weights with plain loads instead
  if (use_tma && tid == 0) {
    mbar_init(&wbar, 1);
    mbar_expect_tx(&wbar, (uint32_t)(round_up(net.packed_floats, 4) * 4));
    tma_bulk_g2s(w, a.packed, (uint32_t)(round_up(net.packed_floats, 4) * 4), &wbar);
  }
  if (!use_tma)
    for (int i = tid; i < round_up(net.packed_floats, 4); i += blockDim.x) w[i] = a.packed[i];
  for (int n = tid; n < NS + 2; n += blockDim.x) pbc[n] = pbc_explore((float)n, a.p.pb_c_init, a.p.pb_c_base);
  {  // tree init: zero everything; child records start as {kNoChild<<16, 0, 0, 0}; parents as 0xFFFFFFFF
    uint32_t* ub = reinterpret_cast<uint32_t*>(blocks);
    for (int i = tid; i < kLT * L.stride; i += blockDim.x) {
      const int o = i % L.stride;
      uint32_t v = 0u;
      if (o >= L.childs && o < L.raw && ((o - L.childs) & 3) == 0) v = kNoChild << 16;
      if (o < L.childs && (o & 3) == 3) v = 0xFFFFFFFFu;
      ub[i] = v;
    }
  }
  for (int i = tid; i < kLT * net.obs_dim; i += blockDim.x) {
    const int tr = i / net.obs_dim, k = i - tr * net.obs_dim;
    bufIn[k * kLT + tr] = a.obs[(size_t)min(row0 + tr, a.B - 1) * net.obs_dim + k];
  }
  __syncthreads();  // thread 0 initialised the mbarrier: it must exist before any other thread polls it
  if (use_tma) mbar_wait(&wbar, 0);
  __syncthreads();

  // Tree walks (select / expand / backup) are scalar per-lane code: the trees are packed onto `walkers` warps (one
  // per scheduler by default) instead of two lanes of every warp — the same walk costs 16 / walkers times fewer
  // warp-instructions, and with 16 walker warps the phase was issue-bound (profiles/r01_lane2_phase_clocks.txt).
  const int WW = a.walkers > 0 ? min(a.walkers, nwarps) : nwarps;
  const int tpw = kLT / WW > 0 ? kLT / WW : 1;
  const bool owner = warp < WW && lane < tpw && warp * tpw + lane < kLT;
  const int ti = owner ? warp * tpw + lane : 0;
  const bool live = owner && row0 + ti < a.B;
  const int b = min(row0 + ti, a.B - 1);
  float* blk = blocks + (size_t)ti * L.stride;
  float4* nodes = reinterpret_cast<float4*>(blk + L.nodes);
  float4* childs = reinterpret_cast<float4*>(blk + L.childs);
  float* traw = blk + L.raw;
  float* tlog = blk + L.logits;
  float* temb = blk + L.emb;
  float* troot = blk + L.root;
  SearchParams p = a.p;
  p.batch_offset += b;

  // ---- root inference (muax/model.py:251-263): repr (runtime obs_dim, one layer) -> min-max -> pred
  {
    const LLayer& l0 = net.repr[0];
    for (int g = warp; g < (E + 3) / 4; g += nwarps) {
      const float* wj = w + l0.off + g * 4;
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int k = 0; k < l0.K; ++k) {
        const float x = bufIn[k * kLT + lane];
        const float4 wv = *reinterpret_cast<const float4*>(wj + k * l0.out4);
        acc[0] = MZ_FMA(x, wv.x, acc[0]); acc[1] = MZ_FMA(x, wv.y, acc[1]);
        acc[2] = MZ_FMA(x, wv.z, acc[2]); acc[3] = MZ_FMA(x, wv.w, acc[3]);
      }
      const float4 bv = *reinterpret_cast<const float4*>(wj + l0.K * l0.out4);
      acc[0] = MZ_ADD(acc[0], bv.x); acc[1] = MZ_ADD(acc[1], bv.y);
      acc[2] = MZ_ADD(acc[2], bv.z); acc[3] = MZ_ADD(acc[3], bv.w);
      l2_store<E>(bufNs + lane, g * 4, acc);
    }
  }
  __syncthreads();
  {
    float v[E];
    l2_minmax_col<E>(bufNs + lane, net.repr_minmax != 0, v);
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < E; ++k) bufIn[k * kLT + lane] = v[k];
    }
  }
  __syncthreads();
  l2_layer1<E, H>(w, net.pred_v[0], net.pred_pi[0], bufIn, -1, act_kind, hA, hB, lane, warp, nwarps);
  __syncthreads();
  l2_layer2<H, F, A>(w, net.pred_v[1], net.pred_pi[1], hA, hB, bufV, bufP, lane, warp, nwarps);
  __syncthreads();
  if (warp < 2) l2_head_exps<F>(bufV + lane, bufEv + lane, warp, nwarps >= 2 ? 2 : 1);
  __syncthreads();
  if (warp == 0) sc_value[lane] = l2_head_scalar<F, S>(bufEv + lane);
  __syncthreads();
  if (owner) {
    // policy prologue (A.2) + node 0, through the generic scalar code on a temporary SoA view
    const float rv = sc_value[ti];
    if (live && a.root_value_out != nullptr) a.root_value_out[b] = rv;
    float mx = -mz_inf();
#pragma unroll
    for (int x = 0; x < A; ++x) mx = fmaxf(mx, bufP[x * kLT + ti]);
    float sum = 0.0f;
#pragma unroll
    for (int x = 0; x < A; ++x) sum = MZ_ADD(sum, mz_expf(MZ_SUB(bufP[x * kLT + ti], mx)));
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
    float lg[A], lmax = -mz_inf();
#pragma unroll
    for (int x = 0; x < A; ++x) {
      const float prob = MZ_DIV(mz_expf(MZ_SUB(bufP[x * kLT + ti], mx)), sum);
      const float nzv = inj != nullptr ? inj[x] : (gsum > 0.0f ? MZ_DIV(g[x], gsum) : MZ_DIV(1.0f, (float)A));
      troot[x] = nzv;
      const float noisy = MZ_ADD(MZ_MUL(MZ_SUB(1.0f, p.dirichlet_fraction), prob), MZ_MUL(p.dirichlet_fraction, nzv));
      lg[x] = mz_logf(fmaxf(noisy, MZ_F32_TINY));
      lmax = fmaxf(lmax, lg[x]);
    }
    float m2 = -mz_inf();
#pragma unroll
    for (int x = 0; x < A; ++x) {
      const bool iv = inv != nullptr && inv[x] != 0;
      if (inv != nullptr) lg[x] = iv ? -MZ_F32_MAX : MZ_SUB(lg[x], lmax);
      troot[A + x] = iv ? 1.0f : 0.0f;
      tlog[x] = lg[x];
      m2 = fmaxf(m2, lg[x]);
    }
    float s2 = 0.0f;
#pragma unroll
    for (int x = 0; x < A; ++x) s2 = MZ_ADD(s2, mz_expf(MZ_SUB(lg[x], m2)));
#pragma unroll
    for (int x = 0; x < A; ++x) {
      float4 c = childs[x];
      c.y = MZ_DIV(mz_expf(MZ_SUB(lg[x], m2)), s2);
      childs[x] = c;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) temb[e] = bufIn[e * kLT + ti];
    traw[0] = rv;
    nodes[0] = make_float4(__int_as_float(1), rv, pbc[1], __uint_as_float(0xFFFFFFFFu));
  }
  const bool use_table = a.noise_table != nullptr && NS > 0;
  auto prefetch_noise = [&](int sim, int which) {
    const int chunks = kLT * (kGNoiseFloats / 4);
    for (int c = tid; c < chunks; c += blockDim.x) {
      const int tr = c / (kGNoiseFloats / 4), q = c - tr * (kGNoiseFloats / 4);
      cp_async16(nzbuf + ((size_t)which * kLT + tr) * kGNoiseFloats + q * 4,
                 a.noise_table + ((size_t)min(row0 + tr, a.B - 1) * NS + sim) * kGNoiseFloats + q * 4);
    }
  };
  if (use_table) {
    // launched with programmatic stream serialisation: everything above overlapped the noise pre-pass; its table
    // is complete and visible after this wait (a no-op without the launch attribute)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    prefetch_noise(0, 0);
    cp_async_wait_all();
  }
  __syncthreads();

  const int max_depth = p.max_depth > 0 ? p.max_depth : NS;
  bool root_masked = false;
  if (owner) {
#pragma unroll
    for (int x = 0; x < A; ++x) root_masked = root_masked || troot[A + x] != 0.0f;
  }
  // Warp roles inside recurrent_fn: "main" warps [0, MW) run the dense layers and the value head, "aux" warps
  // [MW, nwarps) run the reward head concurrently (it only needs Dynamic's output), so it leaves the critical path.
  const int MW = nwarps >= 16 ? 8 : nwarps;
  const bool has_aux = nwarps > MW;
  const int AW = nwarps - MW;
  MZ_CLK(0);  // prologue + root
  // ---- simulations
  for (int sim = 0; sim < NS; ++sim) {
    int parent = 0, action = 0, next = 0;
    if (owner) {
      // simulate (A.3) with muzero_action_selection (A.5) + qtransform_by_parent_and_siblings (A.6)
      const float* row = nzbuf + ((size_t)(sim & 1) * kLT + ti) * kGNoiseFloats;
      float* slot = smem + M.cold + ti * (A + 2);
      int node = 0, depth = 0;
      for (;;) {
        const float4 nd = nodes[node];
        float4 ch[A];
#pragma unroll
        for (int x = 0; x < A; ++x) ch[x] = childs[node * A + x];
        const float* nzp = row + depth * A;
        if (!(use_table && depth < a.K)) {
          l2_noise_cold<A>(p.sim_keys + 2 * sim, (uint32_t)p.global_batch, (uint32_t)p.batch_offset, p.prng_mode, depth,
                           use_table ? a.K : -1, use_table ? a.cont_keys + ((size_t)b * NS + sim) * 2 : nullptr, slot);
          nzp = slot + 2;
        }
        float nz[A];
#pragma unroll
        for (int x = 0; x < A; ++x) nz[x] = nzp[x];
        int vis[A];
        float q[A];
        float lo = nd.y, hi = nd.y;
#pragma unroll
        for (int x = 0; x < A; ++x) {
          vis[x] = (int)(__float_as_uint(ch[x].x) & 0xFFFFu);
          q[x] = MZ_ADD(ch[x].w, MZ_MUL(gamma, ch[x].z));
          lo = vis[x] > 0 ? fminf(lo, q[x]) : lo;
          hi = vis[x] > 0 ? fmaxf(hi, q[x]) : hi;
        }
        const float denom = fmaxf(MZ_SUB(hi, lo), 1e-8f);
        float vnum[A], pnum[A], pden[A], vsv[A], psv[A];
        bool bad = false;
#pragma unroll
        for (int x = 0; x < A; ++x) {
          vnum[x] = MZ_SUB(vis[x] > 0 ? q[x] : lo, lo);
          pnum[x] = MZ_MUL(nd.z, ch[x].y);
          pden[x] = (float)(vis[x] + 1);
          vsv[x] = div_try(vnum[x], denom, bad);
          psv[x] = div_try(pnum[x], pden[x], bad);
        }
        if (bad) {
#pragma unroll
          for (int x = 0; x < A; ++x) {
            vsv[x] = MZ_DIV(vnum[x], denom);
            psv[x] = MZ_DIV(pnum[x], pden[x]);
          }
        }
        int best = 0;
        float bestv = 0.0f;
        uint32_t cx = 0u;
#pragma unroll
        for (int x = 0; x < A; ++x) {
          float s = MZ_ADD(MZ_ADD(vsv[x], psv[x]), nz[x]);
          if (root_masked && depth == 0 && troot[A + x] != 0.0f) s = -mz_inf();
          const bool better = x == 0 || s > bestv;
          bestv = better ? s : bestv;
          best = better ? x : best;
          cx = better ? __float_as_uint(ch[x].x) : cx;
        }
        action = best;
        const uint32_t ci = cx >> 16;
        ++depth;
        if (ci == kNoChild || depth >= max_depth) {
          next = ci == kNoChild ? sim + 1 : (int)ci;
          break;
        }
        node = (int)ci;
      }
      parent = node;
      sc_action[ti] = action;
      if (live) a.out.sim_depth[(size_t)b * NS + sim] = depth;
#pragma unroll
      for (int e = 0; e < E; ++e) bufIn[e * kLT + ti] = temb[parent * E + e];
    }
    MZ_CLK(1);  // select
    __syncthreads();
    MZ_CLK(2);  // barrier after select
    // recurrent_fn (muax/model.py:265-282)
    if (warp < MW) l2_layer1<E, H>(w, net.dyn_ns[0], net.dyn_r[0], bufIn, sc_action[lane], act_kind, hA, hB, lane, warp, MW);
    __syncthreads();
    MZ_CLK(3);  // dyn layer 1 + barrier
    l2_layer2<H, E, F>(w, net.dyn_ns[1], net.dyn_r[1], hA, hB, bufNs, bufR, lane, warp, nwarps);
    if (use_table && sim + 1 < NS) prefetch_noise(sim + 1, (sim + 1) & 1);
    __syncthreads();
    MZ_CLK(4);  // dyn layer 2 + barrier
    if (warp < MW) {
      // main warps: min-max (every warp for its own lane, in registers) -> Prediction -> value head
      float v[E];
      l2_minmax_col<E>(bufNs + lane, net.dyn_minmax != 0, v);
      constexpr int NB = H / 4;
      for (int g = warp; g < 2 * NB; g += MW) {
        const bool second = g >= NB;
        const LLayer& Ld = second ? net.pred_pi[0] : net.pred_v[0];
        const int j0 = (second ? g - NB : g) * 4;
        const float* wj = w + Ld.off + j0;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for (int k = 0; k < E; ++k) {
          const float4 wv = *reinterpret_cast<const float4*>(wj + k * H);
          a0 = MZ_FMA(v[k], wv.x, a0); a1 = MZ_FMA(v[k], wv.y, a1);
          a2 = MZ_FMA(v[k], wv.z, a2); a3 = MZ_FMA(v[k], wv.w, a3);
        }
        const float4 bv = *reinterpret_cast<const float4*>(wj + E * H);
        float r4[4] = {activate(MZ_ADD(a0, bv.x), act_kind), activate(MZ_ADD(a1, bv.y), act_kind),
                       activate(MZ_ADD(a2, bv.z), act_kind), activate(MZ_ADD(a3, bv.w), act_kind)};
        l2_store<H>((second ? hB : hA) + lane, j0, r4);
      }
      if (warp == MW - 1) {
#pragma unroll
        for (int k = 0; k < E; ++k) bufIn[k * kLT + lane] = v[k];
      }
      l2_bar(has_aux, 2, MW * 32);
      MZ_CLK(5);  // min-max + pred layer 1 + barrier
      l2_layer2<H, F, A>(w, net.pred_v[1], net.pred_pi[1], hA, hB, bufV, bufP, lane, warp, MW);
      l2_bar(has_aux, 2, MW * 32);
      MZ_CLK(6);  // pred layer 2 + barrier
      // categorical heads: exps -> barrier -> quotients and products (logit buffers are free again) -> barrier -> sum
      const int half = MW / 2;
      if (has_aux) {
        l2_head_exps<F>(bufV + lane, bufEv + lane, warp, MW);
      } else if (warp < half) {
        l2_head_exps<F>(bufR + lane, bufEr + lane, warp, half);
      } else {
        l2_head_exps<F>(bufV + lane, bufEv + lane, warp - half, MW - half);
      }
      l2_bar(has_aux, 2, MW * 32);
      MZ_CLK(7);  // exps + barrier
      if (has_aux) {
        l2_head_products<F, S>(bufEv + lane, bufV + lane, warp, MW);
      } else if (warp < half) {
        l2_head_products<F, S>(bufEr + lane, bufR + lane, warp, half);
      } else {
        l2_head_products<F, S>(bufEv + lane, bufV + lane, warp - half, MW - half);
      }
      l2_bar(has_aux, 2, MW * 32);
      if (warp == 0) sc_value[lane] = l2_head_final<F, S>(bufV + lane);
      if (!has_aux && warp == half) sc_reward[lane] = l2_head_final<F, S>(bufR + lane);
      MZ_CLK(8);  // head scalar
    } else {
      // aux warps: reward head, off the critical path
      l2_head_exps<F>(bufR + lane, bufEr + lane, warp - MW, AW);
      l2_bar(true, 1, AW * 32);
      l2_head_products<F, S>(bufEr + lane, bufR + lane, warp - MW, AW);
      l2_bar(true, 1, AW * 32);
      if (warp == MW) sc_reward[lane] = l2_head_final<F, S>(bufR + lane);
    }
    cp_async_wait_all();
    __syncthreads();
    MZ_CLK(9);  // final barrier of recurrent_fn
    if (owner) {
      // expand (A.3)
      const float reward = sc_reward[ti], value = sc_value[ti];
      float lg[A], mx = -mz_inf();
#pragma unroll
      for (int x = 0; x < A; ++x) {
        lg[x] = bufP[x * kLT + ti];
        mx = fmaxf(mx, lg[x]);
      }
      float ex[A], sum = 0.0f;
#pragma unroll
      for (int x = 0; x < A; ++x) {
        ex[x] = mz_expf(MZ_SUB(lg[x], mx));
        sum = MZ_ADD(sum, ex[x]);
      }
      float pb[A];
      bool badp = false;
#pragma unroll
      for (int x = 0; x < A; ++x) pb[x] = div_try(ex[x], sum, badp);
      if (badp) {
#pragma unroll
        for (int x = 0; x < A; ++x) pb[x] = MZ_DIV(ex[x], sum);
      }
#pragma unroll
      for (int x = 0; x < A; ++x) {
        tlog[next * A + x] = lg[x];
        float4 c = childs[next * A + x];
        c.y = pb[x];
        childs[next * A + x] = c;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) temb[next * E + e] = bufIn[e * kLT + ti];
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
        bool badb = false;
        float pv = div_try(pnum, pden, badb);
        if (badb) pv = MZ_DIV(pnum, pden);
        nodes[pn] = make_float4(__int_as_float(ci + 1), pv, pbc[min(ci + 1, NS + 1)], pd.w);
        c.x = __uint_as_float(__float_as_uint(c.x) + 1u);
        c.z = child_value;
        childs[e2] = c;
        child_value = pv;
        index = pn;
      }
    }
    MZ_CLK(10);  // expand + backup
  }
#ifdef MZ_PHASE_CLOCKS
  if (blockIdx.x == 1 && lane == 0 && (warp == 0 || warp == 3 || warp == nwarps - 1)) {
    printf("warp %d total %lld | root %lld select %lld bar %lld dynL1 %lld dynL2 %lld mm+predL1 %lld predL2 %lld exps %lld "
           "scalar %lld endbar %lld backup %lld\n", warp, clock64() - kernel_t0, phase_acc[0], phase_acc[1], phase_acc[2],
           phase_acc[3], phase_acc[4], phase_acc[5], phase_acc[6], phase_acc[7], phase_acc[8], phase_acc[9], phase_acc[10]);
  }
#endif

  // ---- policy epilogue (A.2): visit_probs -> temperature -> categorical
  if (owner) {
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
    if (live) {
#pragma unroll
      for (int x = 0; x < A; ++x) a.weights_out[(size_t)b * A + x] = wgt[x];
      a.action_out[b] = best;
    }
  }
  __syncthreads();

  // ---- dump: unpack the records into the mctx SoA arrays
  if (a.dump_tree) {
    const Tree& o = a.out;
    const int live_trees = min(kLT, a.B - row0);
    for (int i = tid; i < live_trees * N; i += blockDim.x) {
      const int tr = i / N, n = i - tr * N;
      const float* tb = blocks + (size_t)tr * L.stride;
      const float4 nd = reinterpret_cast<const float4*>(tb + L.nodes)[n];
      const size_t g = (size_t)(row0 + tr) * o.N + n;
      const uint32_t pa = __float_as_uint(nd.w);
      o.node_visits[g] = __float_as_int(nd.x);
      o.parents[g] = pa == 0xFFFFFFFFu ? -1 : (int)(pa >> 8);
      o.action_from_parent[g] = pa == 0xFFFFFFFFu ? -1 : (int)(pa & 0xFFu);
      o.raw_values[g] = (tb + L.raw)[n];
      o.node_values[g] = nd.y;
    }
    for (int i = tid; i < live_trees * N * A; i += blockDim.x) {
      const int tr = i / (N * A), k = i - tr * (N * A);
      const float* tb = blocks + (size_t)tr * L.stride;
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
    for (int i = tid; i < live_trees * N * E; i += blockDim.x) {
      const int tr = i / (N * E), k = i - tr * (N * E);
      o.embeddings[(size_t)(row0 + tr) * o.N * E + k] = (blocks + (size_t)tr * L.stride + L.emb)[k];
    }
    for (int i = tid; i < live_trees * A; i += blockDim.x) {
      const int tr = i / A, x = i - tr * A;
      const float* rt = blocks + (size_t)tr * L.stride + L.root;
      o.root_noise[(size_t)(row0 + tr) * A + x] = rt[x];
      o.root_invalid[(size_t)(row0 + tr) * A + x] = rt[A + x] != 0.0f ? 1 : 0;
    }
  }
}

// ---------------------------------------------------------------------------------------- host side

struct Lane2Variant {
  int A, E, H, S;
  void* fn;
  size_t (*smem)(int packed_floats, int NS, int obs_dim, int N);
};

template <int A, int E, int H, int S>
size_t lane2_smem(int packed_floats, int NS, int obs_dim, int N) {
  return (size_t)L2Smem<A, E, H, S>(packed_floats, NS, obs_dim, N).total * 4;
}

#define MZ_LANE2_VARIANT(A, E, H, S) \
  Lane2Variant { A, E, H, S, (void*)lane2_search_kernel<A, E, H, S>, &lane2_smem<A, E, H, S> }

inline const std::vector<Lane2Variant>& lane2_variants() {
  static const std::vector<Lane2Variant> v = {
      MZ_LANE2_VARIANT(2, 8, 16, 10),   // CartPole-v1 stock nets (README / BASELINE headline)
      MZ_LANE2_VARIANT(4, 8, 16, 10),   // 4-action environments with the stock nets
      MZ_LANE2_VARIANT(3, 8, 16, 10),
      MZ_LANE2_VARIANT(2, 8, 16, 5),
  };
  return v;
}

struct Lane2State {
  const Lane2Variant* variant = nullptr;
  int warps = 16;
  int walkers = 4;  // MZ_LANE2_WALKERS
};

inline void lane2_init(Lane2State& st, const LaneState& ls, const Net& net, int max_smem) {
  st.variant = nullptr;
  if (!ls.available) return;
  const LaneNet& g = ls.net;
  if (g.n_repr != 1 || g.n_pred != 2 || g.n_dyn != 2) return;
  const int H = net.pred_v.out_dim[0];
  if (net.pred_pi.out_dim[0] != H || net.dyn_ns.out_dim[0] != H || net.dyn_r.out_dim[0] != H) return;
  if (g.A > 255) return;
  for (const Lane2Variant& v : lane2_variants())
    if (v.A == g.A && v.E == g.E && v.H == H && v.S == g.S) {
      if (cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 1024) != cudaSuccess) {
        cudaGetLastError();
        return;
      }
      st.variant = &v;
    }
  if (const char* wv = getenv("MZ_LANE2_WARPS")) st.warps = atoi(wv);
  if (st.warps != 4 && st.warps != 8 && st.warps != 16) st.warps = 16;
  if (const char* wk = getenv("MZ_LANE2_WALKERS")) st.walkers = atoi(wk);
  if (st.walkers != 1 && st.walkers != 2 && st.walkers != 4 && st.walkers != 8 && st.walkers != 16) st.walkers = 4;
}

inline bool lane2_supported(const Lane2State& st, const LaneState& ls, const SearchParams& p) {
  if (st.variant == nullptr) return false;
  if (p.policy != MZ_POLICY_MUZERO || p.qtransform != MZ_QTRANSFORM_BY_PARENT_AND_SIBLINGS) return false;
  if (p.num_simulations + 1 >= (int)kNoChild) return false;
  return st.variant->smem(ls.net.packed_floats, p.num_simulations, ls.net.obs_dim, p.num_simulations + 1) + 1024 <=
         (size_t)ls.max_smem;
}

inline int lane2_launch(Lane2State& st, LaneState& ls, const Tree& out, const SearchParams& p, const float* obs,
                        const uint8_t* invalid, const float* noise, int32_t* action_out, float* weights_out,
                        float* root_value_out, cudaStream_t stream, int64_t* launches, std::string* err) {
  const int B = out.B, NS = p.num_simulations, N = NS + 1, A = ls.net.A;
  LaneArgs a{};
  a.net = ls.net;
  a.packed = ls.packed;
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
  if (getenv("MZ_NO_TMA")) a.dump_tree = 2;
  a.K = std::min(16, kGNoiseFloats / A);
  if (const char* k = getenv("MZ_GROUP_K")) a.K = std::max(0, std::min(a.K, atoi(k)));
  a.walkers = st.walkers;
  if (NS > 0 && a.K > 0) {
    const size_t pairs = (size_t)B * NS;
    if (pairs > ls.noise_capacity) {
      if (ls.noise_table) cudaFree(ls.noise_table);
      if (ls.cont_keys) cudaFree(ls.cont_keys);
      ls.noise_table = nullptr;
      ls.cont_keys = nullptr;
      if (cudaMalloc((void**)&ls.noise_table, pairs * kGNoiseFloats * 4) != cudaSuccess ||
          cudaMalloc((void**)&ls.cont_keys, pairs * 8) != cudaSuccess) {
        *err = "lane2 engine: cudaMalloc(noise table) failed";
        return 1;
      }
      ls.noise_capacity = pairs;
    }
    noise_table_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, stream>>>(p, B, A, a.K, ls.noise_table, ls.cont_keys);
    *launches += 1;
    a.noise_table = ls.noise_table;
    a.cont_keys = ls.cont_keys;
  }
  const size_t smem = st.variant->smem(ls.net.packed_floats, NS, ls.net.obs_dim, N);
  const int grid = (B + kLT - 1) / kLT;
  void* args[] = {&a};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * st.warps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (a.noise_table != nullptr && !getenv("MZ_NO_PDL")) ? 1 : 0;  // overlap the prologue with the noise pre-pass
  const cudaError_t e = cudaLaunchKernelExC(&cfg, st.variant->fn, args);
  *launches += 1;
  if (e != cudaSuccess) {
    *err = std::string("lane2 engine launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

}  // namespace mz

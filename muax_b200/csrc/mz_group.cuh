// mz_group.cuh — warp-autonomous fused engine ("group" engine): ONE search launch per act, no CTA barrier
// inside the simulation loop.
//
// Work decomposition (sized for the stock muax MLPs: embed <= 8, hidden 16, <= 8 actions):
//   * 8 lanes own one tree, a warp owns 4 trees, for the whole act.  The tree (SoA, ~5.7 KB at the CartPole
//     shapes), its MLP staging and its tie-break noise row live in shared memory; warps never wait for each other
//     after the prologue, so every tree runs at its own depth and the SM interleaves ~7 independent warps.
//   * recurrent_fn: each lane computes 4 output units per layer ("slots").  Weights are pre-packed on the device
//     (pack_weights_kernel) as [row][lane][slot] so a lane fetches its 4 weights with one LDS.128, the two heads
//     of a module (value/policy, next-state/reward) are one block-diagonal packed layer, and the packed blob is
//     staged into shared memory with a single TMA bulk copy per CTA.
//   * selection: the 1e-7 * uniform tie-break noise of every (tree, simulation, depth < K) is produced ahead of
//     the search by noise_table_kernel — the jax key chain split(sim_key, B)[b] -> split -> split ... depends only
//     on (key, global row, simulation, depth), never on the tree — so the dependent 2 x threefry per level leave
//     the critical path; deeper paths continue the chain inline from the stored carry key.  sqrt(n) * pb_c(n) is a
//     table over the visit count.
// Arithmetic is the same MZ_* sequence as everywhere else: results are bit-identical to the other engines.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "mz_device.cuh"
#include "mz_fused.cuh"

namespace mz {

constexpr int kGL = 8;          // lanes per tree
constexpr int kGU = 4;          // unit slots per lane (one float4 of weights per input row)
constexpr int kGW = kGL * kGU;  // units per packed layer
constexpr int kGMaxLayers = 4;  // layers per stack supported by this engine
constexpr int kGNoiseFloats = 32;  // tie-break noise row per (tree, simulation): K * A <= 32 floats

struct PLayer {   // one packed dense layer
  int32_t K;      // rows taken from the input vector (multiple of 4)
  int32_t extra;  // one-hot rows appended after K (Dynamic's first layer), else 0
  int32_t off;    // float offset: [(K + extra)][kGL][kGU] weights, then [kGL][kGU] bias
  int32_t act;    // activation after this layer
};

struct PackSrc {  // a raw hk.Linear in the fp32 blob
  int64_t w_off, b_off;
  int32_t in, out;
};

struct PackDesc {  // everything pack_weights_kernel needs for one packed layer
  PLayer pl;
  PackSrc h0, h1;   // h1.out == 0: single head
  int32_t U0, U1;   // output slots of head 0 / head 1
  int32_t first;    // 1: input is a plain vector shared by both heads; 0: previous packed layer's hidden [slot][lane]
  int32_t pU0, p_out0, p_out1;  // previous layer's slot split and true widths (first == 0)
  int32_t in_x;     // first == 1: number of real input rows (E or obs_dim)
};

struct GroupNet {
  PLayer repr[kGMaxLayers], pred[kGMaxLayers], dyn[kGMaxLayers];
  int32_t n_repr, n_pred, n_dyn;
  int32_t obs_dim, E, A, S, F, activation, repr_minmax, dyn_minmax;
  int32_t U_ns, U_f;  // slots holding the next-state units / one categorical head
  int32_t packed_floats;
};

struct GroupArgs {
  GroupNet net;
  const float* packed;  // device, packed weights
  Tree out;             // global SoA tree (dump target)
  SearchParams p;
  const float* obs;
  const uint8_t* invalid;
  const float* noise;          // injected root noise or null
  const float* noise_table;    // [B][NS][kGNoiseFloats] or null (Gumbel)
  const uint32_t* cont_keys;   // [B][NS][2]
  int32_t K;                   // levels covered by the noise table
  int32_t* action_out;
  float* weights_out;
  float* root_value_out;
  int32_t B, N, dump_tree, tree_stride;  // tree_stride: floats per tree block
};

// ---------------------------------------------------------------------------------------- weight packing

__global__ void pack_weights_kernel(const float* __restrict__ raw, float* __restrict__ packed, PackDesc d) {
  const int rows = d.pl.K + d.pl.extra;
  const int total = (rows + 1) * kGW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / kGW, l = (i % kGW) / kGU, u = i % kGU;
    const int h = u < d.U0 ? 0 : 1;
    const PackSrc& src = h == 0 ? d.h0 : d.h1;
    const int j = l + (u - (h == 0 ? 0 : d.U0)) * kGL;  // output unit of head h
    float v = 0.0f;
    if (u < d.U0 + d.U1 && j < src.out) {
      if (r == rows) {
        v = raw[src.b_off + j];
      } else {
        int k_src = -1;
        if (d.first) {
          if (r < d.pl.K) {
            if (r < d.in_x) k_src = r;
          } else {
            k_src = d.in_x + (r - d.pl.K);  // one-hot rows follow the embedding rows (muax/nn.py:105-108)
          }
        } else {  // r = slot' * kGL + lane' of the previous layer; block-diagonal: only the same head feeds
          const int up = r / kGL, lp = r % kGL;
          const int hp = up < d.pU0 ? 0 : 1;
          const int jp = lp + (up - (hp == 0 ? 0 : d.pU0)) * kGL;
          if (hp == h && jp < (hp == 0 ? d.p_out0 : d.p_out1)) k_src = jp;
        }
        if (k_src >= 0 && k_src < src.in) v = raw[src.w_off + (int64_t)k_src * src.out + j];
      }
    }
    packed[d.pl.off + i] = v;
  }
}

// ---------------------------------------------------------------------------------------- tie-break noise table

// Tie-break noise of one (tree, simulation) for the first `levels` levels of its walk: per-tree key =
// split(sim_key, B_global)[global row]; then per level (key, sel) = split(key); noise[a] = 1e-7 *
// uniform(sel, (A,))[a]  (Appendix A.3, A.5, A.7).  `cont` receives the key the chain continues from.
__device__ __forceinline__ void noise_row(const SearchParams& p, int A, int levels, int sim, uint32_t global_row,
                                          float* __restrict__ row, uint32_t* __restrict__ cont) {
  uint32_t k0, k1;
  split_key(p.sim_keys[2 * sim], p.sim_keys[2 * sim + 1], (uint32_t)p.global_batch, global_row, p.prng_mode, k0, k1);
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
  noise_row(p, A, min(K, sim + 1), sim, (uint32_t)(p.batch_offset + b), table + pair * kGNoiseFloats, cont + 2 * pair);
}

// ---------------------------------------------------------------------------------------- per-tree shared memory block

struct GroupLayout {  // float offsets inside one tree block
  int node_visits, parents, afp, children_index, children_visits, raw, values, logits, probs, cvalues, rewards,
      discounts, emb, root_noise, root_invalid, h0, h1, eb_r, eb_v, tb_r, tb_v, nzrow, xbuf, stride;
};

__host__ __device__ inline GroupLayout group_layout(int N, int A, int E, int F, int xk) {
  GroupLayout L;
  int o = 0;
  auto seg = [&](int n) { const int at = o; o += round_up(n, 4); return at; };
  L.node_visits = seg(N); L.parents = seg(N); L.afp = seg(N);
  L.children_index = seg(N * A); L.children_visits = seg(N * A);
  L.raw = seg(N); L.values = seg(N);
  L.logits = seg(N * A); L.probs = seg(N * A); L.cvalues = seg(N * A); L.rewards = seg(N * A);
  L.discounts = seg(N * A);
  L.emb = seg(N * E);
  L.root_noise = seg(A); L.root_invalid = seg((A + 3) / 4);
  L.h0 = seg(kGW); L.h1 = seg(kGW);
  L.eb_r = seg(F); L.eb_v = seg(F); L.tb_r = seg(F); L.tb_v = seg(F);
  L.nzrow = seg(kGNoiseFloats);
  L.xbuf = seg(xk);
  while (o % 32 != 8) o += 4;  // trees of one warp land on different banks
  L.stride = o;
  return L;
}

// ---------------------------------------------------------------------------------------- packed MLP (8 lanes per tree)

// register-array helpers with compile-time indices only (a runtime index would push the array to local memory)
__device__ __forceinline__ float slot_get(const float (&v)[kGU], int i) {
  return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3]));
}
__device__ __forceinline__ void slot_shift(const float (&v)[kGU], int by, float (&out)[kGU]) {
  out[0] = slot_get(v, by);
  out[1] = by + 1 < kGU ? slot_get(v, by + 1) : 0.0f;
  out[2] = by + 2 < kGU ? slot_get(v, by + 2) : 0.0f;
  out[3] = 0.0f;
}

// acc[u] = (sum_k fma(x_k, W[k][lane][u])) (+ one-hot row) + bias, k ascending; x: K floats in shared memory.
__device__ __forceinline__ void packed_layer(const float* wp, const PLayer& pl, const float* x, int onehot, int l,
                                             int act_kind, float (&acc)[kGU]) {
  const float4* w4 = reinterpret_cast<const float4*>(wp + pl.off) + l;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  const int K4 = pl.K >> 2;
#pragma unroll 2
  for (int k4 = 0; k4 < K4; ++k4) {
    const float4 xv = x4[k4];
    const float4 w0 = w4[(4 * k4 + 0) * kGL], w1 = w4[(4 * k4 + 1) * kGL], w2 = w4[(4 * k4 + 2) * kGL],
                 w3 = w4[(4 * k4 + 3) * kGL];
    a0 = MZ_FMA(xv.x, w0.x, a0); a1 = MZ_FMA(xv.x, w0.y, a1); a2 = MZ_FMA(xv.x, w0.z, a2); a3 = MZ_FMA(xv.x, w0.w, a3);
    a0 = MZ_FMA(xv.y, w1.x, a0); a1 = MZ_FMA(xv.y, w1.y, a1); a2 = MZ_FMA(xv.y, w1.z, a2); a3 = MZ_FMA(xv.y, w1.w, a3);
    a0 = MZ_FMA(xv.z, w2.x, a0); a1 = MZ_FMA(xv.z, w2.y, a1); a2 = MZ_FMA(xv.z, w2.z, a2); a3 = MZ_FMA(xv.z, w2.w, a3);
    a0 = MZ_FMA(xv.w, w3.x, a0); a1 = MZ_FMA(xv.w, w3.y, a1); a2 = MZ_FMA(xv.w, w3.z, a2); a3 = MZ_FMA(xv.w, w3.w, a3);
  }
  if (onehot >= 0) {
    const float4 w = w4[(pl.K + onehot) * kGL];
    a0 = MZ_ADD(a0, w.x); a1 = MZ_ADD(a1, w.y); a2 = MZ_ADD(a2, w.z); a3 = MZ_ADD(a3, w.w);
  }
  const float4 bias = w4[(pl.K + pl.extra) * kGL];
  a0 = MZ_ADD(a0, bias.x); a1 = MZ_ADD(a1, bias.y); a2 = MZ_ADD(a2, bias.z); a3 = MZ_ADD(a3, bias.w);
  if (pl.act) {
    a0 = activate(a0, act_kind); a1 = activate(a1, act_kind); a2 = activate(a2, act_kind); a3 = activate(a3, act_kind);
  }
  acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = a3;
}

// One module (both heads) for this lane's tree: hidden activations ping-pong through h0/h1 ([slot][lane] order).
__device__ __forceinline__ void packed_stack(const float* wp, const PLayer* layers, int n, const float* x, int onehot,
                                             float* h0, float* h1, int l, int act_kind, float (&acc)[kGU]) {
  const float* in = x;
  for (int i = 0; i < n; ++i) {
    packed_layer(wp, layers[i], in, i == 0 ? onehot : -1, l, act_kind, acc);
    if (i + 1 < n) {
      float* hb = (i & 1) ? h1 : h0;
#pragma unroll
      for (int u = 0; u < kGU; ++u) hb[u * kGL + l] = acc[u];
      __syncwarp();
      in = hb;
    }
  }
}

__device__ __forceinline__ float lmax8(float v) {
#pragma unroll
  for (int o = kGL / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, kGL));
  return v;
}
__device__ __forceinline__ float lmin8(float v) {
#pragma unroll
  for (int o = kGL / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o, kGL));
  return v;
}

// support_to_scalar(softmax(.)) of two categorical heads at once (reward and value logits held in lane slots
// r[0..UF), v[0..UF): unit j = lane + slot * 8).  Even lanes run the left-to-right sums of the first head, odd
// lanes those of the second, so both heads cost one pass.  muax/model.py:273-274, muax/utils.py:94-102.
__device__ __forceinline__ void joint_support(const float (&r)[kGU], const float (&v)[kGU], int UF, int F, int S, float* eb_r,
                                              float* eb_v, float* tb_r, float* tb_v, int l, float& out_r,
                                              float& out_v) {
  float mr = -mz_inf(), mv = -mz_inf();
#pragma unroll
  for (int u = 0; u < kGU; ++u)
    if (u < UF && l + u * kGL < F) {
      mr = fmaxf(mr, r[u]);
      mv = fmaxf(mv, v[u]);
    }
  mr = lmax8(mr);
  mv = lmax8(mv);
  float er[kGU], ev[kGU];
#pragma unroll
  for (int u = 0; u < kGU; ++u) {
    const int j = l + u * kGL;
    er[u] = ev[u] = 0.0f;
    if (u < UF && j < F) {
      er[u] = mz_expf(MZ_SUB(r[u], mr));
      ev[u] = mz_expf(MZ_SUB(v[u], mv));
      eb_r[j] = er[u];
      eb_v[j] = ev[u];
    }
  }
  __syncwarp();
  const bool odd = l & 1;
  const float* eb = odd ? eb_v : eb_r;
  float s = 0.0f;
  for (int j = 0; j < F; ++j) s = MZ_ADD(s, eb[j]);
  const float so = __shfl_xor_sync(0xffffffffu, s, 1);
  const float sr = odd ? so : s, sv = odd ? s : so;
#pragma unroll
  for (int u = 0; u < kGU; ++u) {
    const int j = l + u * kGL;
    if (u < UF && j < F) {
      tb_r[j] = MZ_MUL((float)(j - S), MZ_DIV(er[u], sr));
      tb_v[j] = MZ_MUL((float)(j - S), MZ_DIV(ev[u], sv));
    }
  }
  __syncwarp();
  const float* tb = odd ? tb_v : tb_r;
  float x = 0.0f;
  for (int j = 0; j < F; ++j) x = MZ_ADD(x, tb[j]);
  const float val = mz_inv_scaling(x);
  const float other = __shfl_xor_sync(0xffffffffu, val, 1);
  out_r = odd ? other : val;
  out_v = odd ? val : other;
}

// min_max_normalize (muax/nn.py:37-44) of a vector held in slots [0, U) of the tree's 8 lanes.
__device__ __forceinline__ void group_min_max(float (&acc)[kGU], int U, int n, int l) {
  float lo = mz_inf(), hi = -mz_inf();
#pragma unroll
  for (int u = 0; u < kGU; ++u)
    if (u < U && l + u * kGL < n) {
      lo = fminf(lo, acc[u]);
      hi = fmaxf(hi, acc[u]);
    }
  lo = lmin8(lo);
  hi = lmax8(hi);
  float scale = MZ_SUB(hi, lo);
  if (scale < 1e-5f) scale = MZ_ADD(scale, 1e-5f);
#pragma unroll
  for (int u = 0; u < kGU; ++u)
    if (u < U) acc[u] = MZ_DIV(MZ_SUB(acc[u], lo), scale);
}

// ---------------------------------------------------------------------------------------- the search kernel

template <int G>
__global__ void __launch_bounds__(256) group_search_kernel(GroupArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  const GroupNet& net = a.net;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int l = lane & (kGL - 1);       // lane inside the tree group
  const int ga = l & (G - 1);           // action owned inside a G-lane subgroup
  const bool writer = l < G;            // first subgroup performs the tree side effects
  const unsigned gm = group_mask<G>();
  const int N = a.N, A = net.A, E = net.E, F = net.F, NS = a.p.num_simulations;
  constexpr int TW = 32 / kGL;          // trees per warp
  const int trees_per_cta = TW * nwarps;

  // ---- prologue: TMA bulk copy of the packed weights; pb_c table
  float* wp = smem;
  const int wfloats = round_up(net.packed_floats, 4);
  float* pbc = smem + wfloats;                       // [NS + 2]
  float* blocks = pbc + round_up(NS + 2, 4);
  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_expect_tx(&wbar, (uint32_t)(wfloats * 4));
    tma_bulk_g2s(wp, a.packed, (uint32_t)(wfloats * 4), &wbar);
  }
  for (int n = tid; n < NS + 2; n += blockDim.x) pbc[n] = pbc_explore((float)n, a.p.pb_c_init, a.p.pb_c_base);

  const GroupLayout L = group_layout(N, A, E, F, round_up(max(net.obs_dim, 4), 4));
  const int tree_local = warp * TW + lane / kGL;
  const int brow = blockIdx.x * trees_per_cta + tree_local;
  const bool live = brow < a.B;
  const int b = live ? brow : a.B - 1;  // surplus groups shadow the last tree (no global writes)
  float* blk = blocks + (size_t)tree_local * L.stride;

  Tree t;
  t.B = 1; t.N = N; t.A = A; t.E = E;
  t.node_visits = reinterpret_cast<int32_t*>(blk + L.node_visits);
  t.parents = reinterpret_cast<int32_t*>(blk + L.parents);
  t.action_from_parent = reinterpret_cast<int32_t*>(blk + L.afp);
  t.children_index = reinterpret_cast<int32_t*>(blk + L.children_index);
  t.children_visits = reinterpret_cast<int32_t*>(blk + L.children_visits);
  t.raw_values = blk + L.raw;
  t.node_values = blk + L.values;
  t.children_prior_logits = blk + L.logits;
  t.children_prior_probs = blk + L.probs;
  t.children_values = blk + L.cvalues;
  t.children_rewards = blk + L.rewards;
  t.children_discounts = blk + L.discounts;
  t.embeddings = blk + L.emb;
  t.root_noise = blk + L.root_noise;
  t.root_invalid = reinterpret_cast<uint8_t*>(blk + L.root_invalid);
  t.sim_depth = a.out.sim_depth + (size_t)b * NS;
  float* h0 = blk + L.h0;
  float* h1 = blk + L.h1;
  float* nzrow = blk + L.nzrow;
  float* xbuf = blk + L.xbuf;

  // mctx initial state (Appendix A.1): zeros; parents / action_from_parent / children_index = -1
  for (int i = l; i < L.xbuf + round_up(max(net.obs_dim, 4), 4); i += kGL) blk[i] = 0.0f;
  __syncwarp();
  {
    int32_t* iblk = reinterpret_cast<int32_t*>(blk);
    for (int i = l; i < N; i += kGL) iblk[L.parents + i] = iblk[L.afp + i] = -1;
    for (int i = l; i < N * A; i += kGL) iblk[L.children_index + i] = -1;
  }
  for (int i = l; i < net.obs_dim; i += kGL) xbuf[i] = a.obs[(size_t)b * net.obs_dim + i];
  __syncthreads();  // thread 0 initialised the mbarrier: it must exist before any other thread polls it
  mbar_wait(&wbar, 0);
  __syncthreads();  // weights + pb_c table visible; the last CTA-wide barrier

  SearchParams p = a.p;
  p.batch_offset += b;  // the per-tree Tree uses local row 0; PRNG draws are indexed by the global row

  // ---- root inference (muax/model.py:251-263)
  float acc[kGU];
  packed_stack(wp, net.repr, net.n_repr, xbuf, -1, h0, h1, l, net.activation, acc);
  if (net.repr_minmax) group_min_max(acc, net.U_ns, E, l);
#pragma unroll
  for (int u = 0; u < kGU; ++u)
    if (u < net.U_ns && l + u * kGL < E) t.embeddings[l + u * kGL] = acc[u];
  __syncwarp();
  packed_stack(wp, net.pred, net.n_pred, t.embeddings, -1, h0, h1, l, net.activation, acc);
  float root_value, dummy;
  joint_support(acc, acc, net.U_f, F, net.S, blk + L.eb_r, blk + L.eb_v, blk + L.tb_r, blk + L.tb_v, l, dummy,
                root_value);
  if (live && l == 0 && a.root_value_out != nullptr) a.root_value_out[b] = root_value;  // raw value, model.py:243
  // policy logits sit in slot U_f of lanes 0..A-1: hand them to the begin step through shared memory
  if (l < A) h0[l] = slot_get(acc, net.U_f);
  __syncwarp();
  {
    const size_t ba = (size_t)b * A;
    group_begin<G>(t, p, 0, (long)p.batch_offset, h0, root_value, t.embeddings,
                   a.invalid != nullptr ? a.invalid + ba : nullptr, a.noise != nullptr ? a.noise + ba : nullptr, ga, gm);
  }
  __syncwarp();

  // ---- simulations
  SelectAux aux;
  aux.pbc = pbc;
  aux.K = a.K;
  aux.noise_row = a.noise_table != nullptr ? nzrow : nullptr;
  const float4* ntab = reinterpret_cast<const float4*>(a.noise_table) + (size_t)b * NS * (kGNoiseFloats / 4);
  const uint2* ctab = reinterpret_cast<const uint2*>(a.cont_keys) + (size_t)b * NS;
  float4 nz_next = make_float4(0.f, 0.f, 0.f, 0.f);
  uint2 ck_next = make_uint2(0u, 0u);
  if (a.noise_table != nullptr && NS > 0) {
    nz_next = __ldg(ntab + l);
    ck_next = __ldg(ctab);
  }
  for (int sim = 0; sim < NS; ++sim) {
    if (a.noise_table != nullptr) {
      reinterpret_cast<float4*>(nzrow)[l] = nz_next;
      aux.cont0 = ck_next.x;
      aux.cont1 = ck_next.y;
      __syncwarp();
      if (sim + 1 < NS) {  // prefetch the next simulation's row behind this simulation's work
        nz_next = __ldg(ntab + (size_t)(sim + 1) * (kGNoiseFloats / 4) + l);
        ck_next = __ldg(ctab + sim + 1);
      }
    }
    int parent, action, next, depth;
    group_simulate<G>(t, p, 0, sim, ga, gm, parent, action, next, depth, &aux);
    if (live && l == 0) t.sim_depth[sim] = depth;

    // recurrent_fn (muax/model.py:265-282)
    float dacc[kGU];
    packed_stack(wp, net.dyn, net.n_dyn, t.embeddings + parent * E, action, h0, h1, l, net.activation, dacc);
    if (net.dyn_minmax) group_min_max(dacc, net.U_ns, E, l);
    float* nemb = t.embeddings + next * E;
#pragma unroll
    for (int u = 0; u < kGU; ++u)
      if (u < net.U_ns && l + u * kGL < E) nemb[l + u * kGL] = dacc[u];
    __syncwarp();
    packed_stack(wp, net.pred, net.n_pred, nemb, -1, h0, h1, l, net.activation, acc);
    float reward, value, racc[kGU];
    slot_shift(dacc, net.U_ns, racc);
    joint_support(racc, acc, net.U_f, F, net.S, blk + L.eb_r, blk + L.eb_v, blk + L.tb_r, blk + L.tb_v, l, reward,
                  value);
    const float logit = __shfl_sync(0xffffffffu, slot_get(acc, net.U_f), (lane & ~(kGL - 1)) + ga);
    group_expand_backup<G>(t, 0, parent, action, next, reward, p.discount, value, logit, nullptr, ga, gm, writer);
    __syncwarp();
  }

  // ---- policy epilogue
  {
    int action;
    float weight;
    group_finish<G>(t, p, 0, (long)p.batch_offset, a.invalid != nullptr, ga, gm, action, weight);
    if (live && writer) {
      if (ga < A) a.weights_out[(size_t)b * A + ga] = weight;
      if (ga == 0) a.action_out[b] = action;
    }
  }
  __syncwarp();

  // ---- dump this tree to the global SoA arrays (mctx layout)
  if (a.dump_tree && live) {
    const Tree& o = a.out;
    const size_t gn = (size_t)b * o.N;
    for (int n = l; n < N; n += kGL) {
      o.node_visits[gn + n] = t.node_visits[n];
      o.parents[gn + n] = t.parents[n];
      o.action_from_parent[gn + n] = t.action_from_parent[n];
      o.raw_values[gn + n] = t.raw_values[n];
      o.node_values[gn + n] = t.node_values[n];
    }
    for (int i = l; i < N * A; i += kGL) {
      const size_t g = gn * A + i;
      o.children_index[g] = t.children_index[i];
      o.children_visits[g] = t.children_visits[i];
      o.children_prior_logits[g] = t.children_prior_logits[i];
      o.children_prior_probs[g] = t.children_prior_probs[i];
      o.children_values[g] = t.children_values[i];
      o.children_rewards[g] = t.children_rewards[i];
      o.children_discounts[g] = t.children_discounts[i];
    }
    for (int i = l; i < N * E; i += kGL) o.embeddings[gn * E + i] = t.embeddings[i];
    for (int i = l; i < A; i += kGL) {
      o.root_noise[(size_t)b * A + i] = t.root_noise[i];
      o.root_invalid[(size_t)b * A + i] = t.root_invalid[i];
    }
  }
}

// ---------------------------------------------------------------------------------------- host side

struct GroupState {
  bool available = false;   // network fits the packed 8-lane layout
  GroupNet net{};
  std::vector<PackDesc> descs;
  float* packed = nullptr;  // device
  float* noise_table = nullptr;
  uint32_t* cont_keys = nullptr;
  size_t noise_capacity = 0;  // (tree, simulation) pairs
  int max_smem = 0, num_sms = 0, G = 0, warps = 0;
  std::string why;
};

inline void* group_kernel_ptr(int G) {
  switch (G) {
    case 2: return (void*)group_search_kernel<2>;
    case 4: return (void*)group_search_kernel<4>;
    default: return (void*)group_search_kernel<8>;
  }
}

// Builds the packed-layer plan for one module: `s0`/`s1` are its two heads (s1 == nullptr: single head).
inline bool group_plan_module(const mz_stack* s0, const mz_stack* s1, int in_x, int extra, bool act_last,
                              PLayer* out_layers, int32_t* n_out, std::vector<PackDesc>& descs, int& off,
                              int* U0_last, std::string* why) {
  const int n = s0->n_layers;
  if (n < 1 || n > kGMaxLayers || (s1 != nullptr && s1->n_layers != n)) {
    *why = "module depth unsupported by the group engine";
    return false;
  }
  int pU0 = 0, p_out0 = 0, p_out1 = 0;
  for (int i = 0; i < n; ++i) {
    PackDesc d{};
    d.h0 = PackSrc{s0->w_off[i], s0->b_off[i], s0->in_dim[i], s0->out_dim[i]};
    if (s1 != nullptr) d.h1 = PackSrc{s1->w_off[i], s1->b_off[i], s1->in_dim[i], s1->out_dim[i]};
    d.U0 = (d.h0.out + kGL - 1) / kGL;
    d.U1 = s1 != nullptr ? (d.h1.out + kGL - 1) / kGL : 0;
    if (d.U0 + d.U1 > kGU) {
      *why = "layer too wide for 8 lanes x 4 slots";
      return false;
    }
    d.first = i == 0;
    d.in_x = in_x;
    d.pU0 = pU0;
    d.p_out0 = p_out0;
    d.p_out1 = p_out1;
    d.pl.K = i == 0 ? round_up(std::max(in_x, 4), 4) : kGW;
    d.pl.extra = i == 0 ? extra : 0;
    d.pl.act = (i + 1 < n) || act_last;
    d.pl.off = off;
    off += (d.pl.K + d.pl.extra + 1) * kGW;
    out_layers[i] = d.pl;
    descs.push_back(d);
    pU0 = d.U0;
    p_out0 = d.h0.out;
    p_out1 = d.h1.out;
  }
  *n_out = n;
  *U0_last = pU0;
  return true;
}

inline int group_init(GroupState& st, const Net& net, int device, std::string* err) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    *err = "cudaGetDeviceProperties failed";
    return 1;
  }
  st.max_smem = (int)prop.sharedMemPerBlockOptin;
  st.num_sms = prop.multiProcessorCount;
  st.available = false;
  GroupNet& g = st.net;
  g = GroupNet{};
  if (net.obs_dim <= 0) { st.why = "no Representation in the library (obs_dim = 0)"; return 0; }
  if (net.num_actions > kGL) { st.why = "more than 8 actions"; return 0; }
  if (net.embed_dim % 4 != 0) { st.why = "embed_dim not a multiple of 4"; return 0; }
  int G = 2;
  while (G < net.num_actions) G <<= 1;
  st.G = G;
  g.obs_dim = net.obs_dim; g.E = net.embed_dim; g.A = net.num_actions; g.S = net.support_size;
  g.F = 2 * net.support_size + 1;
  g.activation = net.activation; g.repr_minmax = net.repr_minmax; g.dyn_minmax = net.dyn_minmax;
  int off = 0, u_last = 0;
  st.descs.clear();
  if (!group_plan_module(&net.repr, nullptr, net.obs_dim, 0, false, g.repr, &g.n_repr, st.descs, off, &u_last, &st.why))
    return 0;
  g.U_ns = u_last;
  if (!group_plan_module(&net.pred_v, &net.pred_pi, net.embed_dim, 0, false, g.pred, &g.n_pred, st.descs, off, &u_last,
                         &st.why))
    return 0;
  g.U_f = u_last;
  if (!group_plan_module(&net.dyn_ns, &net.dyn_r, net.embed_dim, net.num_actions, false, g.dyn, &g.n_dyn, st.descs, off,
                         &u_last, &st.why))
    return 0;
  if (u_last != g.U_ns) { st.why = "internal: next-state slot mismatch"; return 0; }
  if (kGNoiseFloats / net.num_actions < 1) { st.why = "too many actions for the noise row"; return 0; }
  g.packed_floats = off;
  if (cudaMalloc((void**)&st.packed, (size_t)round_up(off, 4) * 4 + 16) != cudaSuccess) {
    *err = "cudaMalloc(packed weights) failed";
    return 1;
  }
  const cudaError_t e = cudaFuncSetAttribute(group_kernel_ptr(G), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             st.max_smem - 1024);
  if (e != cudaSuccess) {
    *err = std::string("group engine: cudaFuncSetAttribute failed: ") + cudaGetErrorString(e);
    return 1;
  }
  st.warps = 0;
  if (const char* w = getenv("MZ_GROUP_WARPS")) st.warps = atoi(w);
  st.available = true;
  return 0;
}

inline void group_destroy(GroupState& st) {
  if (st.packed) cudaFree(st.packed);
  if (st.noise_table) cudaFree(st.noise_table);
  if (st.cont_keys) cudaFree(st.cont_keys);
  st.packed = nullptr;
  st.noise_table = nullptr;
  st.cont_keys = nullptr;
}

// Re-pack after mz_set_weights (stream-ordered).
inline int group_pack(GroupState& st, const float* raw, cudaStream_t stream, int64_t* launches) {
  if (!st.available) return 0;
  for (const PackDesc& d : st.descs) {
    const int total = (d.pl.K + d.pl.extra + 1) * kGW;
    pack_weights_kernel<<<(total + 255) / 256, 256, 0, stream>>>(raw, st.packed, d);
    *launches += 1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

inline size_t group_smem_bytes(const GroupState& st, int N, int NS, int warps) {
  const GroupNet& g = st.net;
  const GroupLayout L = group_layout(N, g.A, g.E, g.F, round_up(std::max(g.obs_dim, 4), 4));
  const size_t floats = (size_t)round_up(g.packed_floats, 4) + round_up(NS + 2, 4) +
                        (size_t)(32 / kGL) * warps * L.stride;
  return floats * 4;
}

// Warps per CTA: one CTA per SM when the batch allows it (every SM gets the same number of trees), else fewer.
inline int group_pick_warps(const GroupState& st, int B, int N, int NS) {
  int best = 0;
  for (int w = 1; w <= 8; ++w)
    if (group_smem_bytes(st, N, NS, w) + 1024 <= (size_t)st.max_smem) best = w;
  if (best == 0) return 0;
  if (st.warps > 0) return st.warps <= best ? st.warps : 0;
  const int trees_per_sm = (B + st.num_sms - 1) / st.num_sms;
  int w = (trees_per_sm + (32 / kGL) - 1) / (32 / kGL);
  if (w < 1) w = 1;
  if (w > best) w = best;
  return w;
}

inline bool group_supported(const GroupState& st, const SearchParams& p, int B) {
  return st.available && group_pick_warps(st, B, p.num_simulations + 1, p.num_simulations) > 0;
}

inline int group_launch(GroupState& st, const Tree& out, const SearchParams& p, const float* obs,
                        const uint8_t* invalid, const float* noise, int32_t* action_out, float* weights_out,
                        float* root_value_out, cudaStream_t stream, int64_t* launches, std::string* err) {
  const int B = out.B, NS = p.num_simulations, N = NS + 1, A = st.net.A;
  const int warps = group_pick_warps(st, B, N, NS);
  if (warps <= 0) {
    *err = "group engine: tree does not fit in shared memory";
    return 1;
  }
  GroupArgs a{};
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
        *err = "group engine: cudaMalloc(noise table) failed";
        return 1;
      }
      st.noise_capacity = pairs;
    }
    noise_table_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, stream>>>(p, B, A, a.K, st.noise_table, st.cont_keys);
    *launches += 1;
    a.noise_table = st.noise_table;
    a.cont_keys = st.cont_keys;
  }
  const size_t smem = group_smem_bytes(st, N, NS, warps);
  const int trees_per_cta = (32 / kGL) * warps;
  const int grid = (B + trees_per_cta - 1) / trees_per_cta;
  void* args[] = {&a};
  const cudaError_t e = cudaLaunchKernel(group_kernel_ptr(st.G), dim3(grid), dim3(32 * warps), args, smem, stream);
  *launches += 1;
  if (e != cudaSuccess) {
    *err = std::string("group engine launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

}  // namespace mz

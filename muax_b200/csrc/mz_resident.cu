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
#include <cstdio>
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
// Work item = (tile of RT rows, quad of 4 output units): y[r][j] = (sum_k fma(x[r][k], W[k][j])) (+ W[nin + onehot[r]][j])
// + b[j] with k ascending — the accumulation order of the CPU checkers; RT x 4 independent chains per thread.
// kVec: W rows are 16-byte aligned and nout % 4 == 0 for both layers, so a quad of weights is one 128-bit load;
// otherwise four scalar loads at consecutive addresses (a partial last quad re-reads the last unit and drops it).
__device__ __forceinline__ float4 ld_quad_vec(const float* p, bool ldg) {
  if (ldg) return __ldg(reinterpret_cast<const float4*>(p));
  return lds_v4(smem_u32(p));
}

template <bool kLdg, int RT, bool kVec>
__device__ __noinline__ void dense_quads(const DenseJob A, const DenseJob B, int R, const int32_t* onehot, int act_kind,
                                         int apply_act) {
  const int qa = (A.nout + 3) >> 2, qtot = qa + ((B.nout + 3) >> 2);
  const int tiles = (R + RT - 1) / RT;
  for (int item = threadIdx.x; item < tiles * qtot; item += blockDim.x) {
    const int tile = item / qtot;
    int q = item - tile * qtot;
    const bool second = q >= qa;
    if (second) q -= qa;
    const DenseJob& J = second ? B : A;
    const int nin = J.nin, nout = J.nout;
    const int j0 = q * 4;
    const int r0 = tile * RT;
    // column offsets of the quad (clamped for a partial last quad; only used by the scalar path)
    const int c1 = min(j0 + 1, nout - 1) - j0, c2 = min(j0 + 2, nout - 1) - j0, c3 = min(j0 + 3, nout - 1) - j0;
    uint32_t xa[RT];
    float acc[RT][4];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      xa[i] = smem_u32(J.src + min(r0 + i, R - 1) * J.lds);
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
    }
    const float* wp = J.W + j0;  // row k of the quad: wp + k * nout
    auto load_quad = [&](const float* p) -> float4 {
      if constexpr (kVec) {
        return ld_quad_vec(p, kLdg);
      } else if constexpr (kLdg) {
        return make_float4(__ldg(p), __ldg(p + c1), __ldg(p + c2), __ldg(p + c3));
      } else {
        const uint32_t s = smem_u32(p);
        return make_float4(lds_f32(s), lds_f32(s + 4u * c1), lds_f32(s + 4u * c2), lds_f32(s + 4u * c3));
      }
    };
    int k = 0;
#pragma unroll 1
    for (; k + 4 <= nin; k += 4) {
      const float4 wa = load_quad(wp);
      const float4 wb = load_quad(wp + nout);
      const float4 wc = load_quad(wp + 2 * nout);
      const float4 wd = load_quad(wp + 3 * nout);
      wp += 4 * nout;
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float4 xv = lds_v4(xa[i] + (uint32_t)k * 4u);
        acc[i][0] = MZ_FMA(xv.x, wa.x, acc[i][0]); acc[i][1] = MZ_FMA(xv.x, wa.y, acc[i][1]);
        acc[i][2] = MZ_FMA(xv.x, wa.z, acc[i][2]); acc[i][3] = MZ_FMA(xv.x, wa.w, acc[i][3]);
        acc[i][0] = MZ_FMA(xv.y, wb.x, acc[i][0]); acc[i][1] = MZ_FMA(xv.y, wb.y, acc[i][1]);
        acc[i][2] = MZ_FMA(xv.y, wb.z, acc[i][2]); acc[i][3] = MZ_FMA(xv.y, wb.w, acc[i][3]);
        acc[i][0] = MZ_FMA(xv.z, wc.x, acc[i][0]); acc[i][1] = MZ_FMA(xv.z, wc.y, acc[i][1]);
        acc[i][2] = MZ_FMA(xv.z, wc.z, acc[i][2]); acc[i][3] = MZ_FMA(xv.z, wc.w, acc[i][3]);
        acc[i][0] = MZ_FMA(xv.w, wd.x, acc[i][0]); acc[i][1] = MZ_FMA(xv.w, wd.y, acc[i][1]);
        acc[i][2] = MZ_FMA(xv.w, wd.z, acc[i][2]); acc[i][3] = MZ_FMA(xv.w, wd.w, acc[i][3]);
      }
    }
    for (; k < nin; ++k) {
      const float4 wk = load_quad(wp);
      wp += nout;
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float xk = lds_f32(xa[i] + (uint32_t)k * 4u);
        acc[i][0] = MZ_FMA(xk, wk.x, acc[i][0]); acc[i][1] = MZ_FMA(xk, wk.y, acc[i][1]);
        acc[i][2] = MZ_FMA(xk, wk.z, acc[i][2]); acc[i][3] = MZ_FMA(xk, wk.w, acc[i][3]);
      }
    }
    if (onehot != nullptr) {  // [x, one_hot(action)] @ W = x @ W[:nin] + W[nin + action]  (muax/nn.py:105-108)
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float4 wo = load_quad(wp + (size_t)onehot[min(r0 + i, R - 1)] * nout);  // wp == row nin here
        acc[i][0] = MZ_ADD(acc[i][0], wo.x); acc[i][1] = MZ_ADD(acc[i][1], wo.y);
        acc[i][2] = MZ_ADD(acc[i][2], wo.z); acc[i][3] = MZ_ADD(acc[i][3], wo.w);
      }
    }
    float4 bq;
    {
      const float* bp = J.bias + j0;
      if constexpr (kLdg) {
        bq = make_float4(__ldg(bp), __ldg(bp + c1), __ldg(bp + c2), __ldg(bp + c3));
      } else {
        const uint32_t s = smem_u32(bp);
        bq = make_float4(lds_f32(s), lds_f32(s + 4u * c1), lds_f32(s + 4u * c2), lds_f32(s + 4u * c3));
      }
    }
    const int valid = min(4, nout - j0);
    const uint32_t da = smem_u32(J.dst + j0);
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      if (r0 + i < R) {
        float y[4] = {MZ_ADD(acc[i][0], bq.x), MZ_ADD(acc[i][1], bq.y), MZ_ADD(acc[i][2], bq.z), MZ_ADD(acc[i][3], bq.w)};
        const uint32_t d = da + (uint32_t)((r0 + i) * J.ldd) * 4u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < valid) {
            if (apply_act) y[c] = activate(y[c], act_kind);
            sts_f32(d + 4u * c, y[c]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ TMA-staged weights
// Wide nets (C5: 1.4 MB of fp32 weights) do not fit shared memory, and a plain load per k-step leaves every FMA
// waiting for an L2 round trip (measured: 170k cycles per Dynamic pass).  Here the weight rows of the two heads are
// streamed through a ring of kTmaStages shared-memory stages by TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx) while all threads run the FMA chains of the current chunk out of shared memory with 128-bit loads.
// Accumulation order is unchanged (k ascending).
//
// Every CTA needs the same rows at about the same time, so the kernel is launched in thread-block clusters when the
// batch divides evenly: the cluster's rank-0 CTA issues ONE multicast bulk copy per chunk that lands in all CTAs'
// rings (same CTA-relative offset) and completes the transaction on each CTA's own `full` mbarrier — L2 traffic for
// weights drops by the cluster size.  A stage is refilled only after every CTA of the cluster has finished reading
// it: each CTA's thread 0 arrives (remote mbarrier arrive over DSMEM) on the leader's `empty` mbarrier of that stage.
#ifndef MZ_TMA_KC
#define MZ_TMA_KC 16
#endif
#ifndef MZ_TMA_STAGES
#define MZ_TMA_STAGES 3
#endif
constexpr int kTmaKC = MZ_TMA_KC;          // weight rows per chunk (multiple of 4)
constexpr int kTmaStages = MZ_TMA_STAGES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                       uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// mbarrier wait that cannot hang the GPU: a phase that has not completed after ~1 s of polling is a bug -> trap.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 24)) __trap();
  }
}

struct TmaRing {
  float* stage;          // kTmaStages x stage_floats (shared), or null when the ring is disabled
  uint64_t* full;        // kTmaStages mbarriers: "chunk has landed" (one arrival + the chunk's bytes)
  uint64_t* empty;       // kTmaStages mbarriers, used in the cluster's rank-0 CTA: "all CTAs are done with the stage"
  int32_t stage_floats;
  uint32_t g;            // chunks consumed so far (stage = g % kTmaStages, parity = (g / kTmaStages) & 1); uniform
  uint32_t rank, ncta;   // position in the cluster (ncta == 1: no cluster)
  int32_t T;             // trees per CTA (uniform over the cluster; this CTA may own fewer live rows)
};

// One chunk, thread 0 only (gi = global chunk number):
//   1. arm this CTA's `full` barrier with the chunk's byte count;
//   2. (consumed_gi >= 0) tell the cluster leader that this CTA is done with the chunk it has just consumed;
//   3. the leader waits until the chunk's stage is free cluster-wide, then issues the (multicast) copies.
// 2 comes before 3 because the leader's own arrival is one of the arrivals it waits for.
__device__ __forceinline__ void tma_ring_step(TmaRing* ring, bool do_issue, uint32_t gi, const float* srcA,
                                              uint32_t bytesA, const float* srcB, uint32_t bytesB, uint32_t offB_floats,
                                              bool consumed, uint32_t consumed_gi) {
  constexpr uint32_t S = kTmaStages;
  const uint32_t st = gi % S;
  float* dst = ring->stage + (size_t)st * ring->stage_floats;
  if (do_issue) mbar_expect_tx(&ring->full[st], bytesA + bytesB);
  if (consumed && ring->ncta > 1) mbar_arrive_remote(&ring->empty[consumed_gi % S], 0u);
  if (!do_issue) return;
  if (ring->ncta == 1) {
    tma_bulk_g2s(dst, srcA, bytesA, &ring->full[st]);
    if (bytesB) tma_bulk_g2s(dst + offB_floats, srcB, bytesB, &ring->full[st]);
  } else if (ring->rank == 0) {
    // chunk gi reuses the stage of chunk gi - S: its (gi / S - 1)-th consumption must be complete in every CTA
    if (gi >= S) mbar_wait_bounded(&ring->empty[st], (gi / S - 1u) & 1u);
    const uint16_t mask = (uint16_t)((1u << ring->ncta) - 1u);
    tma_bulk_g2s_multicast(dst, srcA, bytesA, &ring->full[st], mask);
    if (bytesB) tma_bulk_g2s_multicast(dst + offB_floats, srcB, bytesB, &ring->full[st], mask);
  }
}

template <int RT, bool kVec>
__device__ __noinline__ void dense_tma(const DenseJob A, const DenseJob B, int R, int T, const int32_t* onehot,
                                       int act_kind, int apply_act, TmaRing* ring) {
  constexpr int KC = kTmaKC, S = kTmaStages;
  const int na = A.nout, nb = B.nout, nin = A.nin;
  const int qa = (na + 3) >> 2, qtot = qa + ((nb + 3) >> 2);
  // the pass structure depends on T (uniform over the cluster), not on this CTA's live rows
  const int tiles = (T + RT - 1) / RT, total = tiles * qtot;
  const int nchunks = (nin + KC - 1) / KC;
  const uint32_t stage0 = smem_u32(ring->stage);
  for (int pass0 = 0; pass0 < total; pass0 += blockDim.x) {
    const int item = min(pass0 + (int)threadIdx.x, total - 1);  // a thread without an item recomputes the last one
    const int tile = item / qtot;
    int q = item - tile * qtot;
    const bool second = q >= qa;
    if (second) q -= qa;
    const DenseJob& J = second ? B : A;
    const int nout = J.nout;
    const int j0 = q * 4;
    const int r0 = tile * RT;
    const bool valid = pass0 + (int)threadIdx.x < total && r0 < R;
    const int c1 = min(j0 + 1, nout - 1) - j0, c2 = min(j0 + 2, nout - 1) - j0, c3 = min(j0 + 3, nout - 1) - j0;
    const uint32_t col_off = (uint32_t)((second ? KC * na : 0) + j0) * 4u;  // byte offset of the quad inside a stage
    uint32_t xa[RT];
    float acc[RT][4];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      xa[i] = smem_u32(J.src + min(r0 + i, R - 1) * J.lds);
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
    }
    const uint32_t g0 = ring->g;
    // step(c, done): issue chunk c of this pass (if it exists) and/or report chunk `done` as consumed
    auto step = [&](int c, int done) {
      const bool do_issue = c < nchunks;
      const int cc = do_issue ? c : 0;
      const int rows = min(KC, nin - cc * KC);
      tma_ring_step(ring, do_issue, g0 + (uint32_t)cc, A.W + (size_t)cc * KC * na, (uint32_t)(rows * na) * 4u,
                    B.W + (size_t)cc * KC * nb, (uint32_t)(rows * nb) * 4u, (uint32_t)(KC * na), done >= 0,
                    g0 + (uint32_t)max(done, 0));
    };
    if (threadIdx.x == 0)
      for (int c = 0; c < min(S, nchunks); ++c) step(c, -1);
    for (int c = 0; c < nchunks; ++c) {
      const uint32_t gi = g0 + (uint32_t)c;
      mbar_wait_bounded(&ring->full[gi % S], (gi / S) & 1u);
      const uint32_t wbase = stage0 + (uint32_t)((gi % S) * ring->stage_floats) * 4u + col_off;
      const int rows = min(KC, nin - c * KC);
      const uint32_t row_bytes = (uint32_t)nout * 4u;
#pragma unroll 1
      for (int kk = 0; kk < rows; kk += 4) {
        float4 wq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t wa = wbase + (uint32_t)(kk + u) * row_bytes;
          if constexpr (kVec) {
            wq[u] = lds_v4(wa);
          } else {
            wq[u] = make_float4(lds_f32(wa), lds_f32(wa + 4u * c1), lds_f32(wa + 4u * c2), lds_f32(wa + 4u * c3));
          }
        }
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const float4 xv = lds_v4(xa[i] + (uint32_t)(c * KC + kk) * 4u);
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[i][0] = MZ_FMA(xs[u], wq[u].x, acc[i][0]);
            acc[i][1] = MZ_FMA(xs[u], wq[u].y, acc[i][1]);
            acc[i][2] = MZ_FMA(xs[u], wq[u].z, acc[i][2]);
            acc[i][3] = MZ_FMA(xs[u], wq[u].w, acc[i][3]);
          }
        }
      }
      __syncthreads();  // this CTA is done with chunk c
      if (threadIdx.x == 0) step(c + S, c);  // report chunk c consumed; refill its stage with chunk c + S
    }
    ring->g = g0 + (uint32_t)nchunks;
    const float* wrow = J.W + (size_t)nin * nout + j0;  // one-hot rows follow the nin input rows
    if (onehot != nullptr) {  // [x, one_hot(action)] @ W = x @ W[:nin] + W[nin + action]  (muax/nn.py:105-108)
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float* wo = wrow + (size_t)onehot[min(r0 + i, R - 1)] * nout;
        acc[i][0] = MZ_ADD(acc[i][0], __ldg(wo)); acc[i][1] = MZ_ADD(acc[i][1], __ldg(wo + c1));
        acc[i][2] = MZ_ADD(acc[i][2], __ldg(wo + c2)); acc[i][3] = MZ_ADD(acc[i][3], __ldg(wo + c3));
      }
    }
    const float* bp = J.bias + j0;
    const float4 bq = make_float4(__ldg(bp), __ldg(bp + c1), __ldg(bp + c2), __ldg(bp + c3));
    const int nvalid = valid ? min(4, nout - j0) : 0;
    const uint32_t da = smem_u32(J.dst + j0);
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      if (r0 + i < R) {
        float y[4] = {MZ_ADD(acc[i][0], bq.x), MZ_ADD(acc[i][1], bq.y), MZ_ADD(acc[i][2], bq.z), MZ_ADD(acc[i][3], bq.w)};
        const uint32_t d = da + (uint32_t)((r0 + i) * J.ldd) * 4u;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc < nvalid) {
            if (apply_act) y[cc] = activate(y[cc], act_kind);
            sts_f32(d + 4u * cc, y[cc]);
          }
        }
      }
    }
  }
}

// Row tile: as few rows per thread as keeps the layer to one pass of the CTA (parallelism first, register tiling —
// fewer weight loads — only once every thread has work); weights streamed from L2 always amortise over >= 2 rows.
template <bool kLdg>
__device__ __forceinline__ void dense_pair_auto(const DenseJob& A, const DenseJob& B, int R, const int32_t* onehot,
                                                int act_kind, int apply_act, TmaRing* ring) {
  const int qtot = ((A.nout + 3) >> 2) + ((B.nout + 3) >> 2);
  int rt = (kLdg && R > 1) ? 2 : 1;
  while (rt < 4 && ((R + rt - 1) / rt) * qtot > (int)blockDim.x) rt <<= 1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(A.W) | reinterpret_cast<uintptr_t>(B.W)) & 15) == 0;
  const bool vec = ((A.nout | B.nout) & 3) == 0 && aligned;
  if constexpr (kLdg) {
    // stream the weight rows through the TMA ring when the layer is long enough to pay for the pipeline
    if (ring != nullptr && ring->stage != nullptr && aligned && (A.nin & 3) == 0 && A.nin >= 2 * kTmaKC &&
        (B.nout == 0 || B.nin == A.nin) && kTmaKC * (A.nout + B.nout) <= ring->stage_floats) {
      if (ring->T > 4) {
        if (vec) dense_tma<4, true>(A, B, R, ring->T, onehot, act_kind, apply_act, ring);
        else dense_tma<4, false>(A, B, R, ring->T, onehot, act_kind, apply_act, ring);
      } else {
        if (vec) dense_tma<2, true>(A, B, R, ring->T, onehot, act_kind, apply_act, ring);
        else dense_tma<2, false>(A, B, R, ring->T, onehot, act_kind, apply_act, ring);
      }
      return;
    }
  }
#define MZ_DENSE(RT_)                                                                   \
  do {                                                                                  \
    if (vec)                                                                            \
      dense_quads<kLdg, RT_, true>(A, B, R, onehot, act_kind, apply_act);               \
    else                                                                                \
      dense_quads<kLdg, RT_, false>(A, B, R, onehot, act_kind, apply_act);              \
  } while (0)
  if (rt == 1)
    MZ_DENSE(1);
  else if (rt == 2)
    MZ_DENSE(2);
  else
    MZ_DENSE(4);
#undef MZ_DENSE
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
                                             float* outb, int ldoa, int ldob, float* ta0, float* ta1, float* tb0,
                                             float* tb1, int ldt, int R, TmaRing* ring) {
  const float *srca = x, *srcb = x;
  int lds = ldx;
  for (int l = 0; l < sa.n_layers; ++l) {
    const bool last = l == sa.n_layers - 1;
    float* dsta = last ? outa : ((l & 1) ? ta1 : ta0);
    float* dstb = last ? outb : ((l & 1) ? tb1 : tb0);
    const DenseJob A = make_job(sa, l, w, srca, lds, in_x, dsta, last ? ldoa : ldt);
    DenseJob B = A;
    B.nout = 0;
    if (sb != nullptr) B = make_job(*sb, l, w, srcb, lds, in_x, dstb, last ? ldob : ldt);
    dense_pair_auto<kLdg>(A, B, R, l == 0 ? onehot : nullptr, act_kind, last ? 0 : 1, ring);
    __syncthreads();
    srca = dsta;
    srcb = dstb;
    lds = ldt;
  }
}

template <bool kLdg>
__device__ __forceinline__ void run_stacks(const mz_stack& sa, const mz_stack* sb, const float* w, int act_kind,
                                           const float* x, int ldx, int in_x, const int32_t* onehot, float* outa,
                                           float* outb, int ldoa, int ldob, float* ta0, float* ta1, float* tb0,
                                           float* tb1, int ldt, int R, TmaRing* ring) {
  if (sb != nullptr && sa.n_layers != sb->n_layers) {  // heads of different depth: one after the other
    run_lockstep<kLdg>(sa, nullptr, w, act_kind, x, ldx, in_x, onehot, outa, nullptr, ldoa, ldoa, ta0, ta1, nullptr,
                       nullptr, ldt, R, ring);
    run_lockstep<kLdg>(*sb, nullptr, w, act_kind, x, ldx, in_x, onehot, outb, nullptr, ldob, ldob, tb0, tb1, nullptr,
                       nullptr, ldt, R, ring);
    return;
  }
  run_lockstep<kLdg>(sa, sb, w, act_kind, x, ldx, in_x, onehot, outa, outb, ldoa, ldob, ta0, ta1, tb0, tb1, ldt, R, ring);
}

}  // namespace mz
#include "mz_records.cuh"  // per-row warp functions + tree records (shared with the tree-warp engine)
namespace mz {

// ------------------------------------------------------------------------------------------ tie-break noise pre-pass

// One thread per (tree, simulation) of the simulations [sim0, sim1): per-tree key = split(sim_key, B_global)[global
// row]; then per level (key, sel) = split(key); noise[a] = 1e-7 * uniform(sel, (A,))[a]  (Appendix A.3, A.5, A.7).
// Row = K levels x A.  Tree-major (a warp works on 32 consecutive simulations of one tree: their rows are adjacent in
// the table).  Simulation s walks a tree of s + 1 nodes: its path has at most s + 1 levels (and then never reaches the
// continuation key, which is only read at depth K), so only the first min(K, s + 1) levels are produced.
__global__ void __launch_bounds__(128) resident_noise_kernel(SearchParams p, int B, int A, int K, int sim0, int sim1,
                                                             float* __restrict__ table, uint32_t* __restrict__ cont) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int NS = p.num_simulations, nsim = sim1 - sim0;
  if (idx >= B * nsim) return;
  const int b = idx / nsim, sim = sim0 + idx % nsim;
  const size_t pair = (size_t)b * NS + sim;
  uint32_t k0, k1;
  split_key(p.sim_keys[2 * sim], p.sim_keys[2 * sim + 1], (uint32_t)p.global_batch, (uint32_t)(p.batch_offset + b),
            p.prng_mode, k0, k1);
  float* row = table + pair * K * A;
  const int levels = min(K, sim + 1);
  for (int d = 0; d < levels; ++d) {
    uint32_t s0, s1;
    split_key(k0, k1, 2u, 1u, p.prng_mode, s0, s1);
    split_key(k0, k1, 2u, 0u, p.prng_mode, k0, k1);
    for (int x = 0; x < A; ++x) row[d * A + x] = tie_break_noise(bits_word(s0, s1, (uint32_t)A, (uint32_t)x, p.prng_mode));
  }
  cont[2 * pair] = k0;
  cont[2 * pair + 1] = k1;
}

// Records -> the mctx SoA arrays of the handle (mz_get_tree view).  One thread per (tree, node); nodes that were
// never expanded (visits == 0) read as mctx's initial state (A.1).
__global__ void __launch_bounds__(256) resident_unpack_kernel(const float4* __restrict__ nodes,
                                                             const float4* __restrict__ childs,
                                                             const float* __restrict__ logits, Tree o, int N_used,
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
      h0 = childs[rec * A + x];
      logit = logits[rec * A + x];
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
  float4* rec_childs;    // [B][N][A]
  float* rec_logits;     // [B][N][A]
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
  int32_t ldh; // row stride of the head outputs (value / policy / reward logits)
  int32_t PL;  // path slots per tree
  uint32_t* path;  // [B][PL] selected edges of the current simulation (global scratch, L1 resident)
  int32_t ring_stage_floats;  // floats per stage of the TMA weight ring (0: no ring)
  int32_t clear_embeddings;
};

struct ResidentLayout {  // offsets in floats from the dynamic smem base
  int weights, pbc, mlp, sel, ring, total_floats;
};

// staging per tree: x, ns, tmp0A, tmp1A, tmp0B, tmp1B (row stride ld) + headV, headP, headR (row stride ldh)
__host__ __device__ inline ResidentLayout resident_layout(int weight_bytes_in_smem, int NS, int T, int ld, int ldh,
                                                         int ring_stage_floats) {
  ResidentLayout L;
  int off = 0;
  L.weights = off; off += round_up(weight_bytes_in_smem / 4, 4);
  L.pbc = off;     off += round_up(NS + 2, 4);
  L.mlp = off;     off += T * (6 * ld + 3 * ldh);
  L.sel = off;     off += round_up(7 * T, 4);  // parent, action, next, depth, fresh, reward, value
  L.ring = off;    off += kTmaStages * ring_stage_floats;  // TMA weight ring (weights not resident in smem)
  L.total_floats = round_up(off, 4);
  return L;
}

#ifndef MZ_RES_MIN_CTAS
#define MZ_RES_MIN_CTAS 3  // 80 registers per thread, three 256-thread CTAs per SM (measured: 1-4% faster than 4 x 64)
#endif

// The dense-layer drivers are real (noinline) functions that take the layer stacks by reference.  Referencing the
// stacks inside the kernel parameter would make the compiler copy the whole 1.4 KB ResidentArgs into every thread's
// local memory (3 CTAs x 256 threads x 1.6 KB = 1.2 MB per SM: it thrashed L1 — ncu: 64 % local-load hit rate — and
// put an L2 round trip behind every parameter read of the tree walk).  The stacks are copied once into shared memory
// with constant indices instead, and nothing ever takes the address of the parameter.
struct NetStacks {
  mz_stack repr, pred_v, pred_pi, dyn_ns, dyn_r;
};
__device__ __forceinline__ void copy_stack(mz_stack& dst, const mz_stack& src) {
  dst.n_layers = src.n_layers;
#pragma unroll
  for (int l = 0; l < MZ_MAX_LAYERS; ++l) {
    dst.in_dim[l] = src.in_dim[l];
    dst.out_dim[l] = src.out_dim[l];
    dst.w_off[l] = src.w_off[l];
    dst.b_off[l] = src.b_off[l];
  }
}

template <int G, bool kWSmem>
__global__ void __launch_bounds__(256, MZ_RES_MIN_CTAS) resident_search_kernel(const __grid_constant__ ResidentArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t wbar;
  constexpr bool kLdg = !kWSmem;
  const int T = a.T, A = a.net.num_actions, E = a.net.embed_dim, ld = a.ld, S = a.net.support_size;
  const int row0 = blockIdx.x * T;
  const int R = min(T, a.t.B - row0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int NS = a.p.num_simulations;
  const int N = NS + 1;  // record stride: nodes of this search
  const int ldh = a.ldh;
  const ResidentLayout L = resident_layout(kWSmem ? a.weight_bytes : 0, NS, T, ld, ldh, kWSmem ? 0 : a.ring_stage_floats);
  __shared__ __align__(8) uint64_t ring_full[kTmaStages];
  __shared__ __align__(8) uint64_t ring_empty[kTmaStages];
  TmaRing ring;
  ring.stage = (!kWSmem && a.ring_stage_floats > 0) ? smem + L.ring : nullptr;
  ring.full = ring_full;
  ring.empty = ring_empty;
  ring.stage_floats = a.ring_stage_floats;
  ring.g = 0;
  ring.rank = kWSmem ? 0u : cluster_ctarank();
  ring.ncta = kWSmem ? 1u : cluster_nctarank();
  ring.T = T;
  if (!kWSmem && threadIdx.x == 0)
    for (int i = 0; i < kTmaStages; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], ring.ncta);
    }
  // barriers must exist cluster-wide before the first remote arrive / multicast (and nobody may leave early, below)
  if (!kWSmem && ring.ncta > 1) cluster_sync_all();
  const int act_kind = a.net.activation;
  __shared__ NetStacks net;
  if (threadIdx.x == 0) {
    copy_stack(net.repr, a.net.repr);
    copy_stack(net.pred_v, a.net.pred_v);
    copy_stack(net.pred_pi, a.net.pred_pi);
    copy_stack(net.dyn_ns, a.net.dyn_ns);
    copy_stack(net.dyn_r, a.net.dyn_r);
  }

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
  t.childs = a.rec_childs + (size_t)row0 * N * A;
  t.logits = a.rec_logits + (size_t)row0 * N * A;
  t.pol = l2_evict_last_policy();
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
  float* ta0 = ns + T * ld;
  float* ta1 = ta0 + T * ld;
  float* tb0 = ta1 + T * ld;
  float* tb1 = tb0 + T * ld;
  float* headV = tb1 + T * ld;
  float* headP = headV + T * ldh;
  float* headR = headP + T * ldh;
  int32_t* sel_parent = reinterpret_cast<int32_t*>(smem + L.sel);
  int32_t* sel_action = sel_parent + T;
  int32_t* sel_next = sel_action + T;
  int32_t* sel_depth = sel_next + T;
  int32_t* sel_fresh = sel_depth + T;
  float* rec_reward = reinterpret_cast<float*>(sel_fresh + T);
  float* rec_value = rec_reward + T;
  uint32_t* path = a.path + (size_t)row0 * a.PL;

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
    run_stacks<kLdg>(net.repr, nullptr, w, act_kind, x, ld, obs_dim, nullptr, ns, nullptr, ld, ld, ta0, ta1, nullptr,
                     nullptr, ld, R, &ring);
    if (a.net.repr_minmax) {
      for (int r = warp; r < R; r += nwarps) min_max_row_warp(ns + r * ld, E, lane);
      __syncthreads();
    }
  }
  if (a.obs != nullptr || a.root_logits == nullptr) {
    run_stacks<kLdg>(net.pred_v, &net.pred_pi, w, act_kind, ns, ld, E, nullptr, headV, headP, ldh, ldh, ta0, ta1, tb0,
                     tb1, ld, R, &ring);
    for (int r = warp; r < R; r += nwarps) {
      const float v = support_to_scalar_warp(headV + r * ldh, S, lane);
      if (lane == 0) rec_value[r] = v;
    }
  } else {
    for (int i = tid; i < R * A; i += blockDim.x) {
      const int r = i / A, k = i - r * A;
      headP[r * ldh + k] = a.root_logits[(long)(row0 + r) * A + k];
    }
    if (tid < R) rec_value[tid] = a.root_value[row0 + tid];
  }
  __syncthreads();
  if (tid < R && a.root_value_out != nullptr) a.root_value_out[row0 + tid] = rec_value[tid];  // raw value (model.py:243)

  // ---- lane groups: group g of the CTA owns trees g, g + ngroups, ...  Trees that share a warp advance level by
  // level together (same loop body), so packing them costs no latency and divides the instruction count.  The tree
  // loops are warp-uniform (`base`), a group without a live tree is predicated off (`has`).
  constexpr int gpw = 32 / G;
  const int ngroups = nwarps * gpw;
  const int gq = lane / G;       // group inside the warp
  const int ga = lane & (G - 1); // action handled by this lane
  const bool use_table = a.noise_table != nullptr && a.K > 0 && p.policy == MZ_POLICY_MUZERO;
  const size_t nz_row = (size_t)a.K * A;

  for (int base = warp * gpw; base < R; base += ngroups) {
    const bool has = base + gq < R;
    const int b = min(base + gq, R - 1);
    const long ba = (long)(row0 + b) * A;
    if (MZ_RES_WARP_UNIFORM || has)
      rec_begin<G>(t, p, b, has, (long)p.batch_offset + b, headP + b * ldh, rec_value[b], ns + b * ld,
                   a.invalid != nullptr ? a.invalid + ba : nullptr, a.noise != nullptr ? a.noise + ba : nullptr, ga);
  }
  __syncthreads();

#ifdef MZ_PHASE_CLOCKS
  long long phase_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long phase_t = clock64();
  long long levels = 0;
#define MZ_RCLK(i) do { const long long t__ = clock64(); phase_acc[i] += t__ - phase_t; phase_t = t__; } while (0)
#else
#define MZ_RCLK(i) do { } while (0)
#endif
  // ---- simulations
  for (int sim = 0; sim < NS; ++sim) {
    // A: select
    for (int base = warp * gpw; base < R; base += ngroups) {
      const bool has = base + gq < R;
      const int b = min(base + gq, R - 1);
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
      }
      int parent = 0, action = 0, next = 0, depth = 0;
      bool fresh = false;
      if (MZ_RES_WARP_UNIFORM || has)
        rec_simulate<G>(t, p, b, has, sim, ga, parent, action, next, depth, fresh, aux, path + b * a.PL);
      if (has && ga == 0) {
        sel_parent[b] = parent;
        sel_action[b] = action;
        sel_next[b] = next;
        sel_depth[b] = depth;
        sel_fresh[b] = fresh ? 1 : 0;
        t.sim_depth[(size_t)b * NS + sim] = depth;
      }
      if (has) {
        const float* pe = t.emb + ((size_t)b * t.embN + parent) * E;
        for (int e = ga; e < E; e += G) x[b * ld + e] = MZ_LD_EMB(pe + e);
      }
#ifdef MZ_PHASE_CLOCKS
      levels += depth;
#endif
    }
    MZ_RCLK(0);
    __syncthreads();
    MZ_RCLK(1);
    // B: Dynamic (muax/model.py:269-271): next state -> ns, reward logits -> headR
    run_stacks<kLdg>(net.dyn_ns, &net.dyn_r, w, act_kind, x, ld, E, sel_action, ns, headR, ld, ldh, ta0, ta1, tb0, tb1,
                     ld, R, &ring);
    MZ_RCLK(2);
    // C: min-max of the next state + reward support transform, one warp per row
    for (int r = warp; r < R; r += nwarps) {
      if (a.net.dyn_minmax) min_max_row_warp(ns + r * ld, E, lane);
      const float rv = support_to_scalar_warp(headR + r * ldh, S, lane);
      if (lane == 0) rec_reward[r] = rv;
    }
    __syncthreads();
    MZ_RCLK(3);
    // D: Prediction (model.py:272): value logits -> headV, policy logits -> headP
    run_stacks<kLdg>(net.pred_v, &net.pred_pi, w, act_kind, ns, ld, E, nullptr, headV, headP, ldh, ldh, ta0, ta1, tb0,
                     tb1, ld, R, &ring);
    MZ_RCLK(4);
    // E: value support transform, one warp per row
    for (int r = warp; r < R; r += nwarps) {
      const float v = support_to_scalar_warp(headV + r * ldh, S, lane);
      if (lane == 0) rec_value[r] = v;
    }
    __syncthreads();
    MZ_RCLK(5);
    // F: expand + backup by the tree's lane group
    for (int base = warp * gpw; base < R; base += ngroups) {
      const bool has = base + gq < R;
      const int b = min(base + gq, R - 1);
      const float logit = ga < A ? headP[b * ldh + ga] : 0.0f;
      if (MZ_RES_WARP_UNIFORM || has)
        rec_expand_backup<G>(t, b, has, sel_parent[b], sel_action[b], sel_next[b], sel_fresh[b] != 0, rec_reward[b],
                             p.discount, rec_value[b], logit, ns + b * ld, ga, path + b * a.PL, sel_depth[b]);
    }
    // the next select of a tree runs on the lanes of the same group: a warp-level fence orders the backup's global
    // writes before it; the staging buffers are only rewritten after the next CTA barrier
    __syncwarp();
    MZ_RCLK(6);
  }
#ifdef MZ_PHASE_CLOCKS
  if ((blockIdx.x == 1 || blockIdx.x == 100) && lane == 0 && (warp == 0 || warp == 1 || warp == nwarps - 1))
    printf("cta %d warp %d R %d | select %lld bar %lld dyn %lld C %lld pred %lld E %lld backup %lld | levels(lane0 trees) %lld\n",
           blockIdx.x, warp, R, phase_acc[0], phase_acc[1], phase_acc[2], phase_acc[3], phase_acc[4], phase_acc[5],
           phase_acc[6], levels);
#endif

  // ---- policy epilogue
  for (int base = warp * gpw; base < R; base += ngroups) {
    const bool has = base + gq < R;
    const int b = min(base + gq, R - 1);
    int action = 0;
    float weight = 0.0f;
    if (MZ_RES_WARP_UNIFORM || has)
      rec_finish<G>(t, p, b, has, (long)p.batch_offset + b, a.invalid != nullptr, ga, action, weight);
    if (has && ga < A) a.weights_out[(long)(row0 + b) * A + ga] = weight;
    if (has && ga == 0) a.action_out[row0 + b] = action;
  }
  // a CTA's shared memory (the leader's `empty` barriers, everybody's ring) must outlive its peers' last accesses
  if (!kWSmem && ring.ncta > 1) cluster_sync_all();
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

// Widest pair of layers evaluated in one phase (two heads of a module): columns of one TMA ring stage row.
static int net_max_pair_cols(const Net& net) {
  int cols = 0;
  auto pair = [&](const mz_stack& a, const mz_stack* b) {
    for (int l = 0; l < a.n_layers; ++l) {
      int c = a.out_dim[l];
      if (b != nullptr && b->n_layers == a.n_layers) c += b->out_dim[l];
      cols = std::max(cols, c);
    }
    if (b != nullptr && b->n_layers != a.n_layers)
      for (int l = 0; l < b->n_layers; ++l) cols = std::max(cols, (int)b->out_dim[l]);
  };
  pair(net.repr, nullptr);
  pair(net.pred_v, &net.pred_pi);
  pair(net.dyn_ns, &net.dyn_r);
  return round_up(cols, 4);
}

int net_weight_bytes(const Net& net) {
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
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, st.max_smem - 2048);
    if (e != cudaSuccess) {
      *err = std::string("resident engine: cudaFuncSetAttribute failed: ") + cudaGetErrorString(e);
      cudaGetLastError();
      return 1;
    }
  }
  if (const char* e = getenv("MZ_RESIDENT_TREES")) st.trees_per_cta = atoi(e);
  if (const char* e = getenv("MZ_RESIDENT_GLOBAL_WEIGHTS")) st.force_global_weights = atoi(e);
  if (const char* e = getenv("MZ_RESIDENT_K")) st.noise_levels = std::max(0, atoi(e));
  if (const char* e = getenv("MZ_RESIDENT_NO_TMA")) st.no_tma_ring = atoi(e);
  if (const char* e = getenv("MZ_RESIDENT_CLUSTER")) {
    const int n = atoi(e);
    if (n == 1 || n == 2 || n == 4 || n == 8) st.cluster_size = n;
  }
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
  if (st.rec_logits) cudaFree(st.rec_logits);
  st.rec_logits = nullptr;
  if (st.path) cudaFree(st.path);
  st.path = nullptr;
  st.path_capacity = 0;
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
      reinterpret_cast<const float4*>(st.rec_nodes), reinterpret_cast<const float4*>(st.rec_childs), st.rec_logits, tree,
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
  int T = 0, grid = 0, wsmem = 0, PL = 1, ring_stage_floats = 0, cluster = 1;
  size_t smem = 0;
};

// Trees per CTA: the smallest T for which all ceil(B / T) CTAs are co-resident (occupancy from the runtime: shared
// memory for weights + staging, registers, threads), so that every tree of the batch is in flight at once and the
// SMs hold as many independent CTAs as possible to overlap one CTA's tree walk with another's dense layers.
static ResidentPlan resident_plan(const ResidentState& st, const Net& net, int B, int NS, int max_depth) {
  ResidentPlan best;
  const int ld = round_up(net.max_width, 4);
  const int ldh = round_up(std::max(2 * net.support_size + 1, net.num_actions), 4);
  const int wbytes = net_weight_bytes(net);
  const int budget = st.max_smem - 2048;  // per CTA (opt-in limit minus static shared memory: layer stacks, mbarriers)
  const int PL = std::max(1, std::min(max_depth > 0 ? max_depth : NS, NS));
  for (int ws = st.force_global_weights ? 0 : 1; ws >= 0; --ws) {
    int ring_floats = ws ? 0 : kTmaKC * net_max_pair_cols(net);
    if (st.no_tma_ring) ring_floats = 0;
    auto bytes = [&](int T) { return (size_t)resident_layout(ws ? wbytes : 0, NS, T, ld, ldh, ring_floats).total_floats * 4; };
    if (ring_floats > 0 && bytes(1) > (size_t)budget) ring_floats = 0;  // no room for the ring: plain loads
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
      if (!ws) {
        // weights streamed from L2: every CTA re-reads the whole blob per simulation, so a CTA takes as many rows
        // as keeps one CTA on every SM
        int want = std::min((B + st.num_sms - 1) / st.num_sms, 16);
        while (want > T && bytes(want) > (size_t)budget) --want;
        T = std::max(T, want);
        // thread-block clusters share each weight chunk through one multicast copy: needs a uniform grid (every CTA
        // full, CTA count a multiple of the cluster size); take the nearest T that gives one, if any
        if (ring_floats > 0 && st.cluster_size > 1) {
          for (int cand = T; cand <= std::min(2 * T, 16); ++cand) {
            if (bytes(cand) > (size_t)budget) break;
            if (B % cand == 0 && (B / cand) % st.cluster_size == 0) {
              T = cand;
              best.cluster = st.cluster_size;
              break;
            }
          }
        }
      }
    }
    if (T <= 0 || bytes(T) > (size_t)budget) continue;
    best.T = T;
    best.ring_stage_floats = ring_floats;
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

// Record arrays + path scratch for B trees of NS + 1 nodes (grown on demand, shared with the tree-warp engine).
int records_reserve(ResidentState& st, int B, int NS, int A, int PL, std::string* err) {
  const size_t need_path = (size_t)B * std::max(PL, 1);
  if (need_path > st.path_capacity) {
    if (st.path) cudaFree(st.path);
    st.path = nullptr;
    st.path_capacity = 0;
    if (cudaMalloc((void**)&st.path, need_path * 4) != cudaSuccess) {
      cudaGetLastError();
      *err = "record engines: cudaMalloc(path scratch) failed";
      return 1;
    }
    st.path_capacity = need_path;
  }
  const size_t nodes = (size_t)B * (NS + 1);
  if (nodes > st.rec_capacity) {
    if (st.rec_nodes) cudaFree(st.rec_nodes);
    if (st.rec_childs) cudaFree(st.rec_childs);
    if (st.rec_logits) cudaFree(st.rec_logits);
    st.rec_nodes = st.rec_childs = nullptr;
    st.rec_logits = nullptr;
    st.rec_capacity = 0;
    if (cudaMalloc(&st.rec_nodes, nodes * 16) != cudaSuccess || cudaMalloc(&st.rec_childs, nodes * A * 16) != cudaSuccess ||
        cudaMalloc((void**)&st.rec_logits, nodes * A * 4) != cudaSuccess) {
      cudaGetLastError();
      *err = "record engines: cudaMalloc(tree records) failed";
      return 1;
    }
    st.rec_capacity = nodes;
  }
  return 0;
}

// Tie-break noise ahead of the search (MuZero policy only: the Gumbel selectors ignore their key): fills
// st.noise_table [B][NS][K][A] and st.cont_keys for the first K = min(levels, PL, 1 GiB cap) levels.  *K_out = 0 when
// no table is produced.  reserve + range: the same in two steps, for callers that produce the table in simulation
// ranges on a side stream (the throughput mode, mz_treewarp.cu).
int records_noise_reserve(ResidentState& st, const SearchParams& p, int B, int A, int levels, int PL, int* K_out,
                          std::string* err) {
  *K_out = 0;
  const int NS = p.num_simulations;
  if (p.policy != MZ_POLICY_MUZERO || NS <= 0 || levels <= 0) return 0;
  const size_t pairs = (size_t)B * NS;
  const size_t cap_bytes = (size_t)1 << 30;  // at most 1 GiB of table
  int K = std::min(levels, PL);
  K = (int)std::min<size_t>((size_t)K, cap_bytes / (pairs * A * 4));
  if (K <= 0) return 0;
  const size_t need = pairs * (size_t)K * A;
  if (need > st.noise_capacity) {
    if (st.noise_table) cudaFree(st.noise_table);
    st.noise_table = nullptr;
    st.noise_capacity = 0;
    if (cudaMalloc((void**)&st.noise_table, need * 4) != cudaSuccess) {
      cudaGetLastError();
      *err = "record engines: cudaMalloc(noise table) failed";
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
      *err = "record engines: cudaMalloc(carry keys) failed";
      return 1;
    }
    st.cont_capacity = pairs;
  }
  *K_out = K;
  return 0;
}

void records_noise_range(ResidentState& st, const SearchParams& p, int B, int A, int K, int sim0, int sim1,
                         cudaStream_t stream, int64_t* launches) {
  if (sim1 <= sim0 || K <= 0) return;
  const size_t pairs = (size_t)B * (sim1 - sim0);
  resident_noise_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, stream>>>(p, B, A, K, sim0, sim1, st.noise_table,
                                                                            st.cont_keys);
  *launches += 1;
}

int records_noise_prepass(ResidentState& st, const SearchParams& p, int B, int A, int levels, int PL,
                          cudaStream_t stream, int64_t* launches, int* K_out, std::string* err) {
  if (records_noise_reserve(st, p, B, A, levels, PL, K_out, err)) return 1;
  records_noise_range(st, p, B, A, *K_out, 0, p.num_simulations, stream, launches);
  return 0;
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
  a.ldh = round_up(std::max(2 * net.support_size + 1, net.num_actions), 4);
  a.PL = plan.PL;
  a.ring_stage_floats = plan.ring_stage_floats;
  if (records_reserve(st, B, NS, A, plan.PL, err)) return 1;
  a.path = st.path;
  a.clear_embeddings = (p.max_depth > 0 || NS + 1 < tree.N) ? 1 : 0;
  a.rec_nodes = reinterpret_cast<float4*>(st.rec_nodes);
  a.rec_childs = reinterpret_cast<float4*>(st.rec_childs);
  a.rec_logits = st.rec_logits;
  {
    int K = 0;
    if (records_noise_prepass(st, p, B, A, st.noise_levels, plan.PL, stream, launches, &K, err)) return 1;
    if (K > 0) {
      a.noise_table = st.noise_table;
      a.cont_keys = st.cont_keys;
      a.K = K;
    }
  }
  void* args[] = {&a};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(st.threads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = plan.cluster > 1 ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelExC(&cfg, resident_kernel_ptr(st.G, plan.wsmem != 0), args);
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

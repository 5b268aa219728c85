// mz_recurrent_tc.cu — muax's `_recurrent_inference` (muax/model.py:265-282) on the 5th-generation tensor cores
// (interface and rationale: mz_recurrent_tc.cuh).
//
// One CTA = one tile of 128 trees awaiting expansion (128 TMEM lanes = 128 rows).  The four heads of the two modules
// are a fixed program of GEMM steps   reward head -> next-state head -> value head -> policy head   (the reward head
// first, so that the next state may overwrite the [embedding, one-hot(action)] operand it no longer needs):
//   * A operand (activations): bf16 in shared memory in the canonical K-major, no-swizzle UMMA layout — 8 x 16-byte
//     core matrices, offset(row, k) = (k / 8) * 2048 + row * 16 + (k % 8) * 2 — which is exactly what a "thread = row"
//     epilogue writes without bank conflicts (a warp stores 32 consecutive 16-byte rows);
//   * B operand (weights, W^T, K-major): packed ONCE per mz_set_weights into the same layout, bf16, zero padded to
//     multiples of 16, cut into chunks of kTcKC k-values that one `cp.async.bulk` (TMA) each brings into a ring of
//     kTcStages shared-memory stages (full / empty mbarriers; the ring runs ahead across layer boundaries);
//   * D: fp32 accumulators in tensor memory, `tcgen05.mma.cta_group::1.kind::f16` M = 128, N = layer width (<= 256),
//     K = 16 per instruction, issued by ONE thread; completion reaches the epilogue through `tcgen05.commit` on an
//     mbarrier;
//   * epilogue (4 warps, thread = row, `tcgen05.ld.32x32b`): + bias (fp32), ELU / ReLU, round to bf16 into the next
//     step's A operand; the last layer of a head ends in registers: min-max normalisation + the new embedding (fp32 to
//     global, bf16 to the A operand of Prediction), softmax -> support_to_scalar -> reward / value, or the prior logits.
// Warp roles: 0-3 epilogue, 4 MMA issuer, 5 TMA producer.  No CTA barrier between the prologue and the TMEM
// deallocation: the roles meet on mbarriers only (every wait is bounded: a protocol bug traps instead of hanging).
//
// Numerics: operands are rounded to bf16 (8 bits of mantissa), products are exact, accumulation is fp32 in the order
// the tensor core chooses; the one-hot action column is exact.  This is the throughput mode — the trees it builds are
// NOT bit-identical to the fp32 engines (tests/test_gpu_tc.py states the tolerances).
#include "mz_recurrent_tc.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <vector>

namespace mz {

constexpr int kTcM = 128;        // rows per CTA
constexpr int kTcKC = 32;        // k values per weight chunk (two MMAs)
constexpr int kTcStages = 4;     // weight ring depth
constexpr int kTcMaxSteps = 32;  // 4 heads x MZ_MAX_LAYERS
constexpr int kTcThreads = 192;  // warps 0-3: epilogue; warp 4: MMA issuer; warp 5: TMA producer
constexpr uint32_t kTcChunkPitch = kTcM * 16;  // bytes between the 16-byte k-chunks of the A operand

enum { kEpiHidden = 0, kEpiNextState = 1, kEpiReward = 2, kEpiValue = 3, kEpiPolicy = 4 };
enum { kBufA = 0, kBufH0 = 1, kBufH1 = 2 };

struct TcStep {
  int32_t a_buf, out_buf;  // A operand of the GEMM / buffer the epilogue writes (hidden and next-state steps)
  int32_t k16;             // K / 16 after padding
  int32_t n, npad;         // true and padded (multiple of 16) output width
  int32_t epi;
  int64_t b_off;           // bias offset (floats) in the raw fp32 blob
  int64_t img_off;         // bf16-element offset of the layer's operand image
};

struct TcArgs {
  TcStep steps[kTcMaxSteps];
  int32_t n_steps;
  const __nv_bfloat16* images;
  const float* raw;         // raw fp32 blob (biases)
  const float* embeddings;  // [B][N][E]
  const int32_t* parent;    // [B]
  const int32_t* action;    // [B]
  float *reward, *value, *logits, *next_emb;
  int32_t B, N, E, A, S, act_kind, dyn_minmax;
  int32_t kx16;             // k16 of the [embedding, one-hot] operand
  int32_t bufA_bytes, bufH_bytes, bufH1_bytes, stage_bytes, tmem_cols;  // bufH1_bytes = 0 unless a head has >= 3 layers
};

// ------------------------------------------------------------------------------------------ PTX wrappers

__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
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
    if (spin > (1u << 26)) __trap();  // seconds of polling: a protocol bug, not a long kernel
  }
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, no swizzle.
//   bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (between the two 16-byte k-chunks of one MMA)
//   [32,46) stride byte offset >> 4 (between 8-row groups) | [46,48) version = 1 (Blackwell) | [61,64) layout = 0
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major, M = 128, N = npad.
__device__ __forceinline__ uint32_t tc_idesc(int npad) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t tc_pack2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ void tc_sts16(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float tc_act(float y, int kind) {
  if (kind == MZ_ACT_ELU) return y > 0.0f ? y : __expf(y) - 1.0f;
  return y > 0.0f ? y : 0.0f;
}

// ------------------------------------------------------------------------------------------ kernel

__global__ void __launch_bounds__(kTcThreads, 1) recurrent_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  __shared__ __align__(8) uint64_t bar_full[kTcStages], bar_empty[kTcStages], bar_acc, bar_aready;
  __shared__ uint32_t tmem_base_sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // dynamic shared memory: A operand (embedding | one-hot, later the next state) | hidden 0 | hidden 1 | weight ring
  uint8_t* stages = tc_smem + a.bufA_bytes + a.bufH_bytes + a.bufH1_bytes;
  const uint32_t smem_sh = smem_u32(tc_smem);
  auto buf_sh = [&](int which) -> uint32_t {
    return smem_sh + (which == kBufA ? 0u : (uint32_t)a.bufA_bytes + (which == kBufH0 ? 0u : (uint32_t)a.bufH_bytes));
  };
  const uint32_t stages_sh = smem_u32(stages);

  if (tid == 0) {
    for (int i = 0; i < kTcStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(&bar_acc, 1);
    mbar_init(&bar_aready, kTcM);
  }
  if (warp == 0) {  // one warp allocates the accumulator columns of tensor memory (and frees them at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;
  const int row0 = blockIdx.x * kTcM;

  if (warp == 5) {
    // ---- TMA producer: streams every layer's operand image, chunk by chunk, through the ring
    if (lane == 0) {
      uint32_t g = 0;
      for (int s = 0; s < a.n_steps; ++s) {
        const int kpad = a.steps[s].k16 * 16, npad = a.steps[s].npad;
        const __nv_bfloat16* img = a.images + a.steps[s].img_off;
        for (int k0 = 0; k0 < kpad; k0 += kTcKC, ++g) {
          const uint32_t st = g % kTcStages, round = g / kTcStages;
          tc_mbar_wait(&bar_empty[st], (round & 1u) ^ 1u);  // the MMAs that read the stage's last chunk are complete
          const uint32_t bytes = (uint32_t)(npad * min(kTcKC, kpad - k0)) * 2u;
          mbar_expect_tx(&bar_full[st], bytes);
          tma_bulk_g2s(stages + (size_t)st * a.stage_bytes, img + (size_t)k0 * npad, bytes, &bar_full[st]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ---- MMA issuer: one thread drives the tensor core
    if (lane == 0) {
      uint32_t g = 0;
      for (int s = 0; s < a.n_steps; ++s) {
        const int kpad = a.steps[s].k16 * 16, npad = a.steps[s].npad;
        const uint32_t idesc = tc_idesc(npad);
        const uint32_t a_sh = buf_sh(a.steps[s].a_buf);
        tc_mbar_wait(&bar_aready, (uint32_t)(s & 1));  // the A operand is in shared memory, the accumulators are drained
        tc_fence_after();
        for (int k0 = 0; k0 < kpad; k0 += kTcKC, ++g) {
          const uint32_t st = g % kTcStages, round = g / kTcStages;
          tc_mbar_wait(&bar_full[st], round & 1u);
          tc_fence_after();
          const uint32_t b_sh = stages_sh + st * (uint32_t)a.stage_bytes;
          const int kc = min(kTcKC, kpad - k0);
          for (int j = 0; j < kc; j += 16) {
            // one MMA consumes two 16-byte k-chunks of every row of A and of every row of W^T
            const uint64_t adesc = tc_smem_desc(a_sh + (uint32_t)((k0 + j) / 8) * kTcChunkPitch, kTcChunkPitch, 128u);
            const uint64_t bdesc = tc_smem_desc(b_sh + (uint32_t)(j / 8) * (uint32_t)npad * 16u, (uint32_t)npad * 16u, 128u);
            tc_mma(tmem_base, adesc, bdesc, idesc, (k0 + j) > 0 ? 1u : 0u);
          }
          tc_commit(&bar_empty[st]);  // frees the stage when these MMAs have read it
        }
        tc_commit(&bar_acc);  // the layer's accumulators are complete
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue warps: thread = row
    const int r = tid;  // 0..127 = TMEM lane
    const int row = row0 + r;
    const bool live = row < a.B;
    const int rb = min(row, a.B - 1);
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int E = a.E;
    {  // [embedding of the parent, one-hot(action)] -> A operand (muax/nn.py:105-108)
      const int parent = a.parent[rb], action = a.action[rb];
      const float* src = a.embeddings + ((size_t)rb * a.N + parent) * E;
      const int kpad = a.kx16 * 16;
      for (int k0 = 0; k0 < kpad; k0 += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = k0 + i;
          v[i] = (live && k < E) ? __ldcs(src + k) : ((live && k == E + action) ? 1.0f : 0.0f);
        }
        tc_sts16(buf_sh(kBufA) + (uint32_t)(k0 / 8) * kTcChunkPitch + (uint32_t)r * 16u, tc_pack2(v[0], v[1]),
                 tc_pack2(v[2], v[3]), tc_pack2(v[4], v[5]), tc_pack2(v[6], v[7]));
      }
      tc_fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's (async proxy) reads
      tc_mbar_arrive(&bar_aready);
    }
    for (int s = 0; s < a.n_steps; ++s) {
      const TcStep& st = a.steps[s];
      const int n = st.n, npad = st.npad;
      const float* bias = a.raw + st.b_off;
      tc_mbar_wait(&bar_acc, (uint32_t)(s & 1));
      tc_fence_after();
      if (st.epi == kEpiHidden) {
        const uint32_t out_sh = buf_sh(st.out_buf) + (uint32_t)r * 16u;
        for (int c0 = 0; c0 < npad; c0 += 16) {
          float v[16];
          tc_ld16(trow + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            v[i] = c < n ? tc_act(v[i] + __ldg(bias + c), a.act_kind) : 0.0f;
          }
          tc_sts16(out_sh + (uint32_t)(c0 / 8) * kTcChunkPitch, tc_pack2(v[0], v[1]), tc_pack2(v[2], v[3]),
                   tc_pack2(v[4], v[5]), tc_pack2(v[6], v[7]));
          tc_sts16(out_sh + (uint32_t)(c0 / 8 + 1) * kTcChunkPitch, tc_pack2(v[8], v[9]), tc_pack2(v[10], v[11]),
                   tc_pack2(v[12], v[13]), tc_pack2(v[14], v[15]));
        }
      } else if (st.epi == kEpiNextState) {
        // min_max_normalize (muax/nn.py:37-44) over the row, then the new embedding: fp32 to global (expand stores it
        // into the tree), bf16 into Prediction's A operand
        float lo = mz_inf(), hi = -mz_inf();
        if (a.dyn_minmax) {
          for (int c0 = 0; c0 < npad; c0 += 16) {
            float v[16];
            tc_ld16(trow + (uint32_t)c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = c0 + i;
              if (c < n) {
                const float y = v[i] + __ldg(bias + c);
                lo = fminf(lo, y);
                hi = fmaxf(hi, y);
              }
            }
          }
        }
        float scale = hi - lo;
        if (scale < 1e-5f) scale += 1e-5f;
        const uint32_t out_sh = buf_sh(st.out_buf) + (uint32_t)r * 16u;
        float* dst = a.next_emb + (size_t)rb * E;
        for (int c0 = 0; c0 < npad; c0 += 16) {
          float v[16];
          tc_ld16(trow + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            float y = 0.0f;
            if (c < n) {
              y = v[i] + __ldg(bias + c);
              if (a.dyn_minmax) y = (y - lo) / scale;
              if (live) dst[c] = y;
            }
            v[i] = y;
          }
          tc_sts16(out_sh + (uint32_t)(c0 / 8) * kTcChunkPitch, tc_pack2(v[0], v[1]), tc_pack2(v[2], v[3]),
                   tc_pack2(v[4], v[5]), tc_pack2(v[6], v[7]));
          tc_sts16(out_sh + (uint32_t)(c0 / 8 + 1) * kTcChunkPitch, tc_pack2(v[8], v[9]), tc_pack2(v[10], v[11]),
                   tc_pack2(v[12], v[13]), tc_pack2(v[14], v[15]));
        }
      } else if (st.epi == kEpiPolicy) {
        for (int c0 = 0; c0 < npad; c0 += 16) {
          float v[16];
          tc_ld16(trow + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            if (live && c < n) a.logits[(size_t)row * a.A + c] = v[i] + __ldg(bias + c);
          }
        }
      } else {
        // support_to_scalar(softmax(logits)) (muax/model.py:273-274 + muax/utils.py:94-102): max, sum of exps, expectation
        float mx = -mz_inf();
        for (int c0 = 0; c0 < npad; c0 += 16) {
          float v[16];
          tc_ld16(trow + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < n) mx = fmaxf(mx, v[i] + __ldg(bias + c0 + i));
        }
        float sum = 0.0f;
        for (int c0 = 0; c0 < npad; c0 += 16) {
          float v[16];
          tc_ld16(trow + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < n) sum = MZ_ADD(sum, mz_expf(MZ_SUB(v[i] + __ldg(bias + c0 + i), mx)));
        }
        float x = 0.0f;
        for (int c0 = 0; c0 < npad; c0 += 16) {
          float v[16];
          tc_ld16(trow + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < n) {
              const float pr = MZ_DIV(mz_expf(MZ_SUB(v[i] + __ldg(bias + c0 + i), mx)), sum);
              x = MZ_ADD(x, MZ_MUL((float)(c0 + i - a.S), pr));
            }
        }
        const float y = mz_inv_scaling(x);
        if (live) (st.epi == kEpiReward ? a.reward : a.value)[row] = y;
      }
      tc_fence_before();       // this thread's TMEM loads are complete (tcgen05.wait::ld) and ordered before ...
      tc_fence_proxy_async();  // ... and its shared-memory stores visible to ... the next step's MMAs
      tc_mbar_arrive(&bar_aready);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

// W [K][N] fp32 row-major -> the bf16 operand image of the layer: chunks of kTcKC k-values, inside a chunk
// [k / 8][npad rows][8 k-values], zero padded to kpad x npad.
__global__ void recurrent_tc_pack_kernel(const float* __restrict__ raw, __nv_bfloat16* __restrict__ img, int64_t w_off,
                                         int K, int N, int kpad, int npad) {
  const int total = kpad * npad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx / npad, n = idx - k * npad;
    const float v = (k < K && n < N) ? raw[w_off + (int64_t)k * N + n] : 0.0f;
    const int c = k / kTcKC, kk = k - c * kTcKC;
    const size_t off = (size_t)c * npad * kTcKC + ((size_t)(kk / 8) * npad + n) * 8 + (kk & 7);
    img[off] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------ host side

struct TcLayer {
  int64_t w_off, img_off;
  int K, N, kpad, npad;
};
struct TcImpl {
  TcArgs args{};
  std::vector<TcLayer> layers;
  __nv_bfloat16* images = nullptr;
  size_t image_elems = 0;
  size_t smem = 0;
};

int recurrent_tc_init(RecurrentTcState& st, const Net& net, int batch, int device, std::string* err) {
  (void)batch;
  st.available = false;
  st.why.clear();
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    *err = "cudaGetDeviceProperties failed";
    return 1;
  }
  if (prop.major != 10) {
    st.why = "tcgen05 needs an sm_100-class device";
    return 0;
  }
  TcImpl* impl = new TcImpl();
  TcArgs& a = impl->args;
  const int E = net.embed_dim, A = net.num_actions, F = 2 * net.support_size + 1;
  int max_npad = 16, max_hidden_pad = 16, deepest = 1;
  size_t img = 0;
  bool ok = true;
  // heads in execution order: reward, next state, value, policy
  struct Head { const mz_stack* s; int in, final_epi; };
  const Head heads[4] = {{&net.dyn_r, E + A, kEpiReward}, {&net.dyn_ns, E + A, kEpiNextState},
                         {&net.pred_v, E, kEpiValue}, {&net.pred_pi, E, kEpiPolicy}};
  int n_steps = 0;
  for (const Head& h : heads) {
    const mz_stack& s = *h.s;
    deepest = std::max(deepest, (int)s.n_layers);
    for (int l = 0; l < s.n_layers; ++l) {
      const int K = l == 0 ? h.in : s.in_dim[l], N = s.out_dim[l];
      const int kpad = round_up(K, 16), npad = round_up(N, 16);
      if (npad > 256) { ok = false; st.why = "a layer is wider than 256 units"; }
      if (n_steps >= kTcMaxSteps) { ok = false; st.why = "too many layers"; }
      if (!ok) break;
      const bool last = l == s.n_layers - 1;
      TcStep& t = a.steps[n_steps++];
      t.a_buf = l == 0 ? kBufA : ((l - 1) & 1 ? kBufH1 : kBufH0);
      t.out_buf = last ? kBufA : (l & 1 ? kBufH1 : kBufH0);
      t.k16 = kpad / 16;
      t.n = N;
      t.npad = npad;
      t.epi = last ? h.final_epi : kEpiHidden;
      t.b_off = s.b_off[l];
      t.img_off = (int64_t)img;
      impl->layers.push_back(TcLayer{s.w_off[l], (int64_t)img, K, N, kpad, npad});
      img += (size_t)kpad * npad;
      max_npad = std::max(max_npad, npad);
      if (!last) max_hidden_pad = std::max(max_hidden_pad, npad);
    }
    if (!ok) break;
  }
  if (ok && F != net.dyn_r.out_dim[net.dyn_r.n_layers - 1]) { ok = false; st.why = "support size mismatch"; }
  if (ok) {
    a.n_steps = n_steps;
    a.kx16 = round_up(E + A, 16) / 16;
    a.bufA_bytes = std::max(a.kx16 * 16, round_up(E, 16)) / 8 * (int)kTcChunkPitch;
    a.bufH_bytes = max_hidden_pad / 8 * (int)kTcChunkPitch;
    a.stage_bytes = max_npad * kTcKC * 2;
    int cols = 32;
    while (cols < max_npad) cols <<= 1;
    a.tmem_cols = cols;
    a.bufH1_bytes = deepest >= 3 ? a.bufH_bytes : 0;
    impl->smem = (size_t)a.bufA_bytes + a.bufH_bytes + a.bufH1_bytes + (size_t)kTcStages * a.stage_bytes + 128;
    if (impl->smem > (size_t)prop.sharedMemPerBlockOptin - 1024) {
      ok = false;
      st.why = "operands do not fit shared memory";
    }
  }
  if (ok && cudaFuncSetAttribute(recurrent_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)impl->smem) != cudaSuccess) {
    cudaGetLastError();
    ok = false;
    st.why = "cudaFuncSetAttribute failed";
  }
  if (ok && cudaMalloc((void**)&impl->images, img * sizeof(__nv_bfloat16) + 16) != cudaSuccess) {
    cudaGetLastError();
    delete impl;
    *err = "recurrent_tc: cudaMalloc(operand images) failed";
    return 1;
  }
  if (!ok) {
    delete impl;
    return 0;
  }
  impl->image_elems = img;
  a.images = impl->images;
  a.E = E;
  a.A = A;
  a.S = net.support_size;
  a.act_kind = net.activation;
  a.dyn_minmax = net.dyn_minmax;
  st.impl = impl;
  st.available = true;
  return 0;
}

void recurrent_tc_destroy(RecurrentTcState& st) {
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  if (impl != nullptr) {
    if (impl->images) cudaFree(impl->images);
    delete impl;
  }
  st.impl = nullptr;
  st.available = false;
}

int recurrent_tc_pack(RecurrentTcState& st, const Net& net, const float* raw_weights_dev, cudaStream_t stream,
                      int64_t* launches) {
  (void)net;
  if (!st.available) return 0;
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  for (const TcLayer& L : impl->layers) {
    const int total = L.kpad * L.npad;
    recurrent_tc_pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(raw_weights_dev, impl->images + L.img_off, L.w_off,
                                                                      L.K, L.N, L.kpad, L.npad);
    *launches += 1;
  }
  impl->args.raw = raw_weights_dev;
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int recurrent_tc_launch(RecurrentTcState& st, const Net& net, const Tree& t, const int32_t* parent,
                        const int32_t* action, float* reward, float* value, float* logits, float* next_emb,
                        cudaStream_t stream, int64_t* launches, std::string* err) {
  (void)net;
  if (!st.available) {
    *err = "recurrent_tc: unavailable (" + st.why + ")";
    return 1;
  }
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  TcArgs a = impl->args;
  a.embeddings = t.embeddings;
  a.parent = parent;
  a.action = action;
  a.reward = reward;
  a.value = value;
  a.logits = logits;
  a.next_emb = next_emb;
  a.B = t.B;
  a.N = t.N;
  void* args[] = {&a};
  const int grid = (t.B + kTcM - 1) / kTcM;
  const cudaError_t e = cudaLaunchKernel((void*)recurrent_tc_kernel, dim3(grid), dim3(kTcThreads), args, impl->smem, stream);
  *launches += 1;
  if (e != cudaSuccess) {
    *err = std::string("recurrent_tc launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

}  // namespace mz

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
//     multiples of 16, cut into chunks of as many k-values as fit a 16 KB stage (32 at N = 256, a whole narrow layer)
//     that one `cp.async.bulk` (TMA) each brings into a ring of up to 8 shared-memory stages (full / empty mbarriers;
//     the ring runs ahead across layer boundaries);
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
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace mz {

constexpr int kTcM = 128;        // rows per CTA
constexpr int kTcStageBytes = 16384;  // unit of the weight ring's stage size (16 or 32 KB): a layer is cut into chunks of as
                                      // many k values (a multiple of 16) as fit a stage — the whole layer for narrow ones
constexpr int kTcMaxStages = 8;
static int tc_stage_env() {  // MZ_TC_STAGE_KB: A/B knob; 0 = choose per program (tc_build)
  static const int v = getenv("MZ_TC_STAGE_KB") != nullptr ? std::max(8, std::min(96, atoi(getenv("MZ_TC_STAGE_KB")))) * 1024 : 0;
  return v;
}  // weight ring depth: as many stages (2 .. 8) as shared memory has room for
constexpr int kTcMaxSteps = 48;  // 4 heads x MZ_MAX_LAYERS, plus the second halves of the wide hidden layers (pipelined programs)
constexpr int kTcEpiWarps = 8;   // epilogue warps: warp w reads TMEM lanes 32 * (w % 4) .. + 31, column half w / 4
constexpr int kTcEpiThreads = 32 * kTcEpiWarps;
constexpr int kTcStoreWarps = 2;  // copy the new embedding rows out of the operand buffer into the bf16 tree
constexpr int kTcThreads = kTcEpiThreads + 64 + 32 * kTcStoreWarps;  // + warp 8: MMA issuer, warp 9: TMA producer, then the store warps
constexpr uint32_t kTcChunkPitch = kTcM * 16;   // bytes between the 16-byte k-chunks of the A operand

enum { kEpiHidden = 0, kEpiNextState = 1, kEpiReward = 2, kEpiValue = 3, kEpiPolicy = 4 };
enum { kBufA = 0, kBufH0 = 1, kBufH1 = 2 };

// One step = one GEMM (a whole layer, or one half of the output columns of a wide hidden layer) + its epilogue.  The
// steps of a program form a dataflow graph, not a chain: the MMA issuer, the weight producer and the epilogue warps
// all walk the table in order, every step has its own pair of single-use mbarriers (accumulators complete / epilogue
// done), and the issuer waits only for the ONE earlier epilogue (`dep`) that covers everything the step needs — the
// buffer it reads was written, the accumulator columns it overwrites were drained.  Independent heads are interleaved
// and wide layers cut in two column halves on alternate accumulator regions, so the tensor core works on one step
// while the epilogue warps finish another.
struct TcStep {
  int32_t a_buf, out_buf;  // A operand of the GEMM / buffer the epilogue writes (hidden and next-state steps)
  int32_t k16;             // K / 16 after padding
  int32_t kc;              // k values per weight chunk (multiple of 16): kc * npad * 2 bytes <= stage_bytes
  int32_t n, npad;         // true and padded (multiple of 16) width of the step's output columns
  int32_t epi;
  int32_t bias_sh;         // float offset of the step's (zero-padded) bias row in the shared-memory bias table
  int32_t minmax;          // kEpiNextState: min_max_normalize the row
  int32_t acc_col;         // first accumulator column of the step in tensor memory
  int32_t n_off;           // first output column of the layer this step computes (0 unless the layer is cut in two)
  int32_t dep;             // the epilogue the step's MMAs wait for (-1: only the gathered input operand)
  int64_t b_off;           // bias offset (floats) in the raw fp32 blob
  int64_t img_off;         // bf16-element offset of the step's operand image
};

struct TcArgs {
  TcStep steps[kTcMaxSteps];
  int32_t n_steps;
  const __nv_bfloat16* images;
  const float* bias_table;  // every step's bias, zero padded to npad, in step order (built by recurrent_tc_pack)
  // input rows: in + row * in_row_stride (+ parent[row] * in_dim when `parent` is given: embeddings[b, parent[b]])
  const float* in;
  int64_t in_row_stride;
  int32_t in_dim;           // E, or obs_dim in root mode
  const int32_t* parent;    // [B] or null
  const int32_t* action;    // [B] or null: one-hot(action) follows the in_dim input columns (muax/nn.py:105-108)
  // the throughput mode's tree embeddings, bf16 (the operand precision): node rows of `es` elements (E rounded up to
  // 8, zero padded), `tree16_stride` elements per tree.  in16: rows are gathered from [b][parent[b]] instead of `in`;
  // out16: the next state is stored at [b][next[b]] (16-byte pieces, coalesced through the A operand in shared memory)
  const __nv_bfloat16* in16;
  __nv_bfloat16* out16;
  const int32_t* next;
  int64_t tree16_stride;
  int32_t es;
  float *reward, *value, *logits, *next_emb;
  int32_t B, A, S, act_kind, out_dim;  // out_dim: width of next_emb rows (E)
  int32_t kx16;             // k16 of the input operand
  int32_t ns_step;          // the step that ends in kEpiNextState (-1: none)
  int32_t bufA_bytes, bufH_bytes, bufH1_bytes, stage_bytes, n_stages, bias_floats, tmem_cols;  // bufH1_bytes = 0 unless a head has >= 3 layers
  // hidden activations in tensor memory instead of shared memory (the A operand of the next layer is then read from
  // TMEM: `tcgen05.mma [d], [a], b-desc`): bf16 pairs packed into 32-bit columns h_col[0] / h_col[1] .. of the allocation
  int32_t h_tmem, h_col[2];
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
// One lane of a converged warp (cute::elect_one_sync).  A region guarded by this predicate is known to the compiler to
// run on a single thread: tcgen05.mma / commit / bulk copies inside it become straight uniform-datapath instructions.
// Guarded by `lane == 0` instead, every one of them is wrapped in an ELECT / BRA.U.ANY loop — 224 against 68 cycles per
// N = 32 MMA issued (tools/probes/mma_probe.cu, profiles/r02_mma_probe.txt).
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
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
// The same MMA with the A operand in tensor memory (lane = row, one 32-bit column = two consecutive k values).
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers.
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// 32 consecutive fp32 columns of this thread's TMEM lane: one wait per 32 columns.
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// W columns [c0, c0 + W) of this thread's accumulator row, W = 16 or 32.
template <int W>
__device__ __forceinline__ void tc_ldw(uint32_t taddr, float (&v)[W]) {
  if constexpr (W == 32) tc_ld32(taddr, v);
  else tc_ld16(taddr, v);
}
__device__ __forceinline__ uint32_t tc_pack2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ void tc_sts16(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float tc_ex2(float x) {  // 2^x on the SFU (flush-to-zero: inputs here are <= 0 or bounded)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kTcLog2e = 1.4426950408889634f;
// ELU / ReLU of W independent units, branch-free: a `y > 0 ? y : expf(y) - 1` per unit compiles to one divergent
// branch per element (BSSY / BSYNC around a dependent LDS -> FADD -> MUFU chain: 140 cycles per element, 19 k cycles
// per 256-wide layer — profiles/r02_recurrent_tc_timeline.txt); here the exponential runs on every unit and a select
// picks the result.
template <int W>
__device__ __forceinline__ void tc_activate(float (&v)[W], int kind) {
  if (kind == MZ_ACT_ELU) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const float e = tc_ex2(fminf(v[i], 0.0f) * kTcLog2e) - 1.0f;
      v[i] = v[i] > 0.0f ? v[i] : e;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
}
// v[i] += bias[c0 + i] with 128-bit shared-memory loads (the bias rows are zero padded to npad and 64-byte aligned).
template <int W>
__device__ __forceinline__ void tc_add_bias(float (&v)[W], const float* bias, int c0) {
  const float4* b4 = reinterpret_cast<const float4*>(bias + c0);
#pragma unroll
  for (int q = 0; q < W / 4; ++q) {
    const float4 b = b4[q];
    v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
  }
}

// ------------------------------------------------------------------------------------------ kernel

// Epilogue helper: columns [c0, c0 + W) of the accumulator row + bias -> activation -> bf16 into the A-format buffer.
// Padded columns need no mask: their weights and bias are zero, and both activations map 0 to 0.
template <int W>
__device__ __forceinline__ void tc_epi_hidden(uint32_t trow, int c0, const float* bias, int act_kind, uint32_t out_sh) {
  float v[W];
  tc_ldw<W>(trow + (uint32_t)c0, v);
  tc_add_bias<W>(v, bias, c0);
  tc_activate<W>(v, act_kind);
#pragma unroll
  for (int q = 0; q < W / 8; ++q)
    tc_sts16(out_sh + (uint32_t)(c0 / 8 + q) * kTcChunkPitch, tc_pack2(v[8 * q], v[8 * q + 1]),
             tc_pack2(v[8 * q + 2], v[8 * q + 3]), tc_pack2(v[8 * q + 4], v[8 * q + 5]), tc_pack2(v[8 * q + 6], v[8 * q + 7]));
}

// The same, with the bf16 pairs stored into tensor memory (columns h_taddr + c0 / 2 ..).
template <int W>
__device__ __forceinline__ void tc_epi_hidden_tmem(uint32_t trow, int c0, const float* bias, int act_kind, uint32_t h_taddr) {
  float v[W];
  tc_ldw<W>(trow + (uint32_t)c0, v);
  tc_add_bias<W>(v, bias, c0);
  tc_activate<W>(v, act_kind);
  uint32_t pk[W / 2];
#pragma unroll
  for (int i = 0; i < W / 2; ++i) pk[i] = tc_pack2(v[2 * i], v[2 * i + 1]);
  if constexpr (W == 32) tc_st16(h_taddr + (uint32_t)(c0 / 2), pk);
  else tc_st8(h_taddr + (uint32_t)(c0 / 2), pk);
}

// support_to_scalar(softmax(logits)) (muax/model.py:273-274 + muax/utils.py:94-102) of one accumulator row of NP
// (padded) columns held in registers: max, exponentials, sum and expectation without a branch or a second TMEM read.
template <int NP>
__device__ __forceinline__ float tc_head_scalar(uint32_t trow, const float* bias, int n, int S) {
  float v[NP];
#pragma unroll
  for (int c0 = 0; c0 < NP; c0 += 16) {
    float w[16];
    tc_ld16(trow + (uint32_t)c0, w);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[c0 + i] = w[i];
  }
  tc_add_bias<NP>(v, bias, 0);
  float mx = -mz_inf();
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    v[i] = i < n ? v[i] : -mz_inf();
    mx = fmaxf(mx, v[i]);
  }
  float sum = 0.0f, x = 0.0f;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const float e = tc_ex2((v[i] - mx) * kTcLog2e);  // padded columns: 2^-inf = 0
    sum += e;
    x += (float)(i - S) * e;
  }
  return mz_inv_scaling(x / sum);
}

// Next-state epilogue, pass 1: min / max of columns [c0, c0 + W) of the row (+ bias); pass 2: normalise, store fp32 to
// global (when asked) and bf16 into the A operand buffer.
template <int W>
__device__ __forceinline__ void tc_ns_minmax(uint32_t trow, int c0, const float* bias, int n, float& lo, float& hi) {
  float v[W];
  tc_ldw<W>(trow + (uint32_t)c0, v);
  tc_add_bias<W>(v, bias, c0);
  if (c0 + W <= n) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      lo = fminf(lo, v[i]);
      hi = fmaxf(hi, v[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const bool in = c0 + i < n;
      lo = fminf(lo, in ? v[i] : mz_inf());
      hi = fmaxf(hi, in ? v[i] : -mz_inf());
    }
  }
}
template <int W>
__device__ __forceinline__ void tc_ns_store(uint32_t trow, int c0, const float* bias, int n, float sub, float inv, float* dst,
                                            bool dst_vec, uint32_t out_sh) {
  float v[W];
  tc_ldw<W>(trow + (uint32_t)c0, v);
  tc_add_bias<W>(v, bias, c0);
#pragma unroll
  for (int i = 0; i < W; ++i) v[i] = c0 + i < n ? (v[i] - sub) * inv : 0.0f;
  if (dst != nullptr) {
    if (dst_vec && c0 + W <= n) {
#pragma unroll
      for (int q = 0; q < W / 4; ++q)
        __stcs(reinterpret_cast<float4*>(dst + c0) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (c0 + i < n) dst[c0 + i] = v[i];
    }
  }
#pragma unroll
  for (int q = 0; q < W / 8; ++q)
    tc_sts16(out_sh + (uint32_t)(c0 / 8 + q) * kTcChunkPitch, tc_pack2(v[8 * q], v[8 * q + 1]),
             tc_pack2(v[8 * q + 2], v[8 * q + 3]), tc_pack2(v[8 * q + 4], v[8 * q + 5]), tc_pack2(v[8 * q + 6], v[8 * q + 7]));
}

#ifdef MZ_TC_CLOCKS
// timeline of CTA 0 (cycles since kernel start): [0] prologue done, [1] A operand written; per step s:
// [4s+2] MMA issuer past its dependency, [4s+3] last MMA of the step issued, [4s+4] epilogue past bar_acc, [4s+5] epilogue done
#define MZ_TCCLK(i) do { if (blockIdx.x == 0) tc_clk[(i)] = clock64() - tc_t0; } while (0)
#else
#define MZ_TCCLK(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kTcThreads, 1) recurrent_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
#ifdef MZ_TC_CLOCKS
  __shared__ long long tc_clk[4 * kTcMaxSteps + 8];
  const long long tc_t0 = clock64();
#endif
  // bar_acc[s]: the accumulators of step s are complete (tcgen05.commit); bar_done[s]: every epilogue thread is through
  // step s; bar_in: the gathered input operand is in shared memory.  Each is used once per kernel (parity 0).
  __shared__ __align__(8) uint64_t bar_full[kTcMaxStages], bar_empty[kTcMaxStages], bar_acc[kTcMaxSteps], bar_done[kTcMaxSteps],
      bar_in;
  __shared__ uint32_t tmem_base_sh;
  // the step table, copied out of the kernel parameters: a dynamically indexed parameter read is a constant-cache
  // access of several hundred cycles (per field, per step and per weight chunk it cost ~700 cycles per layer on the MMA
  // issuer's critical path — profiles/r02_recurrent_tc_timeline_v2.txt)
  __shared__ TcStep steps_sh[kTcMaxSteps];
  __shared__ float row_lo[2][kTcM], row_hi[2][kTcM];  // partial row min / max of the two column halves
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // dynamic shared memory: A operand (input | one-hot, later the next state) | hidden 0 | hidden 1 | weight ring | biases
  uint8_t* stages = tc_smem + a.bufA_bytes + a.bufH_bytes + a.bufH1_bytes;
  float* bias_all = reinterpret_cast<float*>(stages + (size_t)a.n_stages * a.stage_bytes);
  const uint32_t n_stages = (uint32_t)a.n_stages;
  const uint32_t smem_sh = smem_u32(tc_smem);
  auto buf_sh = [&](int which) -> uint32_t {
    return smem_sh + (which == kBufA ? 0u : (uint32_t)a.bufA_bytes + (which == kBufH0 ? 0u : (uint32_t)a.bufH_bytes));
  };
  const uint32_t stages_sh = smem_u32(stages);

  if (tid == 0) {
    for (int i = 0; i < a.n_stages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < a.n_steps; ++i) {
      mbar_init(&bar_acc[i], 1);
      mbar_init(&bar_done[i], kTcEpiThreads);
    }
    mbar_init(&bar_in, kTcEpiThreads);
  }
  if (warp == 0) {  // one warp allocates the accumulator columns of tensor memory (and frees them at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // every layer's bias, zero padded to npad, once per CTA: the epilogues read it from shared memory.  One flat copy of
  // the packed table: all of a thread's loads are in flight together (a loop per layer paid one L2 round trip per layer)
  for (int i = tid; i < a.bias_floats; i += kTcThreads) bias_all[i] = __ldg(a.bias_table + i);
  {
    constexpr int kWords = (int)(sizeof(TcStep) / 4);
    const uint32_t* srcw = reinterpret_cast<const uint32_t*>(a.steps);
    uint32_t* dstw = reinterpret_cast<uint32_t*>(steps_sh);
    for (int i = tid; i < a.n_steps * kWords; i += kTcThreads) dstw[i] = srcw[i];
  }
  const int n_steps = a.n_steps, h_tmem = a.h_tmem, h_col0 = a.h_col[0], h_col1 = a.h_col[1], stage_bytes = a.stage_bytes;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;
  const int row0 = blockIdx.x * kTcM;
  if (tid == 0) MZ_TCCLK(0);

  if (warp == kTcEpiWarps + 1) {
    // ---- TMA producer: streams every layer's operand image, chunk by chunk, through the ring
    if (tc_elect_one()) {
      uint32_t g = 0;
      const __nv_bfloat16* images = a.images;
      uint32_t st = 0, round = 0;
      for (int s = 0; s < n_steps; ++s) {
        const int kpad = steps_sh[s].k16 * 16, npad = steps_sh[s].npad, kc = steps_sh[s].kc;
        const __nv_bfloat16* img = images + steps_sh[s].img_off;
        for (int k0 = 0; k0 < kpad; k0 += kc) {
          tc_mbar_wait(&bar_empty[st], (round & 1u) ^ 1u);  // the MMAs that read the stage's last chunk are complete
          const uint32_t bytes = (uint32_t)(npad * min(kc, kpad - k0)) * 2u;
          mbar_expect_tx(&bar_full[st], bytes);
          tma_bulk_g2s(stages + (size_t)st * stage_bytes, img + (size_t)k0 * npad, bytes, &bar_full[st]);
          if (++st == n_stages) { st = 0; ++round; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kTcEpiWarps) {
    // ---- MMA issuer: one thread drives the tensor core
    if (tc_elect_one()) {
      uint32_t st = 0, round = 0;
      for (int s = 0; s < n_steps; ++s) {
        const int kpad = steps_sh[s].k16 * 16, npad = steps_sh[s].npad, kc0 = steps_sh[s].kc, a_buf = steps_sh[s].a_buf;
        const int dep = steps_sh[s].dep;
        const uint32_t d_tmem = tmem_base + (uint32_t)steps_sh[s].acc_col;
        const uint32_t idesc = tc_idesc(npad);
        const bool a_tmem = h_tmem && a_buf != kBufA;
        // operands advance by a constant per MMA: two 16-byte k-chunks of A (shared memory: descriptor address field in
        // 16-byte units; tensor memory: 8 columns) and of W^T
        uint64_t adesc = tc_smem_desc(buf_sh(a_buf), kTcChunkPitch, 128u);
        uint32_t a_col = tmem_base + (uint32_t)(a_buf == kBufH1 ? h_col1 : h_col0);
        const uint64_t b_step = (uint64_t)((2u * (uint32_t)npad * 16u) >> 4);
        // the A operand is written, the accumulator columns are drained (epilogues finish in table order: one wait)
        tc_mbar_wait(dep < 0 ? &bar_in : &bar_done[dep], 0u);
        tc_fence_after();
        MZ_TCCLK(4 * s + 2);
        uint32_t acc = 0;
        for (int k0 = 0; k0 < kpad; k0 += kc0) {
          tc_mbar_wait(&bar_full[st], round & 1u);
          tc_fence_after();
          uint64_t bdesc = tc_smem_desc(stages_sh + st * (uint32_t)stage_bytes, (uint32_t)npad * 16u, 128u);
          const int kc = min(kc0, kpad - k0);
          for (int j = 0; j < kc; j += 16) {
            if (a_tmem) tc_mma_ts(d_tmem, a_col, bdesc, idesc, acc);
            else tc_mma(d_tmem, adesc, bdesc, idesc, acc);
            acc = 1;
            adesc += (uint64_t)((2u * kTcChunkPitch) >> 4);
            a_col += 8;
            bdesc += b_step;
          }
          tc_commit(&bar_empty[st]);  // frees the stage when these MMAs have read it
          if (++st == n_stages) { st = 0; ++round; }
        }
        tc_commit(&bar_acc[s]);  // the step's accumulators are complete
        MZ_TCCLK(4 * s + 3);
      }
    }
    __syncwarp();
  } else if (warp >= kTcEpiWarps + 2) {
    // ---- store warps: the new embedding rows, bf16, out of the A operand buffer (where the next-state epilogue has
    // put them for Prediction) into the tree at [b][next[b]].  Lane = (row of an 8-row group, piece mod 4): one store
    // instruction covers 8 rows x 64 contiguous bytes.  Off the epilogue warps' path: the copy used to hold them for
    // ~6 k cycles per simulation (profiles/r02_recurrent_tc_timeline_*).  Nothing writes the buffer again.
    if (a.out16 != nullptr && a.ns_step >= 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");  // `next` comes from the kernel before this one
      const int sw = warp - (kTcEpiWarps + 2), rsub = lane >> 2, q = lane & 3, pieces = a.es / 8;
      constexpr int kGroups = kTcM / 8 / kTcStoreWarps;  // 8-row groups per store warp
      int nxt[kGroups];
#pragma unroll
      for (int g = 0; g < kGroups; ++g) nxt[g] = a.next[min(row0 + (sw * kGroups + g) * 8 + rsub, a.B - 1)];
      tc_mbar_wait(&bar_done[a.ns_step], 0u);
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        const int rr = (sw * kGroups + g) * 8 + rsub;
        if (row0 + rr >= a.B) continue;
        __nv_bfloat16* dst16 = a.out16 + (size_t)(row0 + rr) * a.tree16_stride + (size_t)nxt[g] * a.es;
        const uint32_t src_sh = buf_sh(kBufA) + (uint32_t)rr * 16u;
        for (int c = q; c < pieces; c += 32) {  // eight pieces in flight
          uint4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + 4 * j < pieces)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                           : "r"(src_sh + (uint32_t)(c + 4 * j) * kTcChunkPitch));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + 4 * j < pieces) __stcs(reinterpret_cast<uint4*>(dst16) + c + 4 * j, v[j]);
        }
      }
    }
  } else {
    // ---- epilogue warps: thread = (row, column half)
    const int r = (warp & 3) * 32 + lane;  // 0..127 = TMEM lane
    const int half = warp >> 2;            // which half of a layer's columns this thread finishes
    const int row = row0 + r;
    const bool live = row < a.B;
    const int rb = min(row, a.B - 1);
    const uint32_t trow0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    {  // [input row, one-hot(action)] -> A operand (muax/nn.py:105-108); the two halves take alternate 16-byte chunks
      asm volatile("griddepcontrol.wait;" ::: "memory");  // programmatic dependent launch: everything above overlapped
                                                           // the kernel that selected (parent, action)
      // Only now may the next kernel of the stream begin its prologue: the backup + select kernel that follows copies
      // the tree records into shared memory before its own griddepcontrol.wait, so the kernel that wrote them (the
      // one this kernel has just waited for) must be complete when it starts.
      asm volatile("griddepcontrol.launch_dependents;");
      if (tid == 0) MZ_TCCLK(4 * kTcMaxSteps + 2);  // released by griddepcontrol.wait
      if (a.in16 != nullptr) {
        // bf16 rows: a 16-byte piece of a row IS a k-chunk of the operand.  Lane = (row of an 8-row group, piece mod 4):
        // one load instruction covers 8 rows x 64 contiguous bytes (8 L1 wavefronts, against 32 for a lane-per-row
        // walk), one store 8 rows x 16 bytes of 4 adjacent chunks.
        const int rsub = lane >> 2, q = lane & 3;
        const int D = a.in_dim, kpad = a.kx16 * 16, full = D / 8, chunks = kpad / 8;
        // both row groups of the thread together: their (parent, action) loads form one L2 round trip and the first
        // round of row pieces of both (sixteen 16-byte loads) the second — one group after the other was four
        const __nv_bfloat16* src2[2];
        uint32_t dst2[2];
        int action2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rr = warp * 16 + 8 * h + rsub;  // row of the tile
          const int gb = min(row0 + rr, a.B - 1);
          const int parent = a.parent != nullptr ? a.parent[gb] : 0;
          action2[h] = a.action != nullptr ? a.action[gb] : -1;
          src2[h] = a.in16 + (size_t)gb * a.tree16_stride + (size_t)parent * a.es;
          dst2[h] = buf_sh(kBufA) + (uint32_t)rr * 16u;
        }
        const bool round0 = q + 28 < full;
        if (round0) {  // eight pieces per group in flight
          uint4 v[2][8];
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[h][j] = __ldcs(reinterpret_cast<const uint4*>(src2[h]) + q + 4 * j);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              tc_sts16(dst2[h] + (uint32_t)(q + 4 * j) * kTcChunkPitch, v[h][j].x, v[h][j].y, v[h][j].z, v[h][j].w);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const __nv_bfloat16* src = src2[h];
          const uint32_t dst = dst2[h];
          const int action = action2[h];
          int c = q + (round0 ? 32 : 0);
          for (; c + 28 < full; c += 32) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldcs(reinterpret_cast<const uint4*>(src) + c + 4 * j);
#pragma unroll
            for (int j = 0; j < 8; ++j) tc_sts16(dst + (uint32_t)(c + 4 * j) * kTcChunkPitch, v[j].x, v[j].y, v[j].z, v[j].w);
          }
          for (; c < full; c += 4) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + c);
            tc_sts16(dst + (uint32_t)c * kTcChunkPitch, v.x, v.y, v.z, v.w);
          }
          for (; c < chunks; c += 4) {  // the pieces that hold the end of the row, the one-hot action and the padding
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int k = 8 * c + i;
              v[i] = k < D ? __bfloat162float(src[k]) : ((action >= 0 && k == D + action) ? 1.0f : 0.0f);
            }
            tc_sts16(dst + (uint32_t)c * kTcChunkPitch, tc_pack2(v[0], v[1]), tc_pack2(v[2], v[3]), tc_pack2(v[4], v[5]),
                     tc_pack2(v[6], v[7]));
          }
        }
      } else {
      const int parent = a.parent != nullptr ? a.parent[rb] : 0;
      const int action = a.action != nullptr ? a.action[rb] : -1;
      const int D = a.in_dim;
      const float* src = a.in + (size_t)rb * a.in_row_stride + (size_t)parent * D;  // surplus rows shadow the last row
      const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
      const int kpad = a.kx16 * 16;
      const uint32_t dst_sh = buf_sh(kBufA) + (uint32_t)r * 16u;
      int k0 = 8 * half;
      // eight chunks per round: sixteen independent 128-bit loads in flight before the first store (one chunk at a
      // time paid one L2 round trip per chunk: 7 us of a 256-wide row)
      for (; vec && k0 + 16 * 7 + 8 <= D; k0 += 16 * 8) {
        float4 lo4[8], hi4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          lo4[j] = __ldcs(reinterpret_cast<const float4*>(src + k0 + 16 * j));
          hi4[j] = __ldcs(reinterpret_cast<const float4*>(src + k0 + 16 * j + 4));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          tc_sts16(dst_sh + (uint32_t)((k0 + 16 * j) / 8) * kTcChunkPitch, tc_pack2(lo4[j].x, lo4[j].y),
                   tc_pack2(lo4[j].z, lo4[j].w), tc_pack2(hi4[j].x, hi4[j].y), tc_pack2(hi4[j].z, hi4[j].w));
      }
      for (; k0 < kpad; k0 += 16) {
        float v[8];
        if (vec && k0 + 8 <= D) {
          const float4 lo4 = __ldcs(reinterpret_cast<const float4*>(src + k0));
          const float4 hi4 = __ldcs(reinterpret_cast<const float4*>(src + k0 + 4));
          v[0] = lo4.x; v[1] = lo4.y; v[2] = lo4.z; v[3] = lo4.w; v[4] = hi4.x; v[5] = hi4.y; v[6] = hi4.z; v[7] = hi4.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            v[i] = k < D ? __ldcs(src + k) : ((action >= 0 && k == D + action) ? 1.0f : 0.0f);
          }
        }
        tc_sts16(dst_sh + (uint32_t)(k0 / 8) * kTcChunkPitch, tc_pack2(v[0], v[1]), tc_pack2(v[2], v[3]),
                 tc_pack2(v[4], v[5]), tc_pack2(v[6], v[7]));
      }
      }
      tc_fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's (async proxy) reads
      tc_mbar_arrive(&bar_in);
      if (tid == 0) MZ_TCCLK(1);
    }
    for (int s = 0; s < n_steps; ++s) {
      const TcStep& st = steps_sh[s];
      const int n = st.n, npad = st.npad;
      const float* bias = bias_all + st.bias_sh;
      const uint32_t trow = trow0 + (uint32_t)st.acc_col;  // this thread's row of the step's accumulators
      // this thread's columns [cb, ce): the step's columns are split in two halves of whole 32-column groups
      const int split = min(npad, ((npad / 2 + 31) / 32) * 32);
      const int cb = half == 0 ? 0 : split, ce = half == 0 ? split : npad;
      tc_mbar_wait(&bar_acc[s], 0u);
      tc_fence_after();
      if (tid == 0) MZ_TCCLK(4 * s + 4);
      if (st.epi == kEpiHidden && h_tmem) {
        const uint32_t h_taddr = trow0 + (uint32_t)((st.out_buf == kBufH1 ? h_col1 : h_col0) + st.n_off / 2);
        int c0 = cb;
        for (; c0 + 32 <= ce; c0 += 32) tc_epi_hidden_tmem<32>(trow, c0, bias, a.act_kind, h_taddr);
        for (; c0 < ce; c0 += 16) tc_epi_hidden_tmem<16>(trow, c0, bias, a.act_kind, h_taddr);
        tc_wait_st();
      } else if (st.epi == kEpiHidden) {
        const uint32_t out_sh = buf_sh(st.out_buf) + (uint32_t)r * 16u + (uint32_t)(st.n_off / 8) * kTcChunkPitch;
        int c0 = cb;
        for (; c0 + 32 <= ce; c0 += 32) tc_epi_hidden<32>(trow, c0, bias, a.act_kind, out_sh);
        for (; c0 < ce; c0 += 16) tc_epi_hidden<16>(trow, c0, bias, a.act_kind, out_sh);
      } else if (st.epi == kEpiNextState) {
        // min_max_normalize (muax/nn.py:37-44) over the row, then the new embedding: fp32 to global (expand stores it
        // into the tree), bf16 into Prediction's A operand.  The two column halves of a row meet through shared memory.
        float lo = mz_inf(), hi = -mz_inf();
        if (st.minmax) {
          int c0 = cb;
          for (; c0 + 32 <= ce; c0 += 32) tc_ns_minmax<32>(trow, c0, bias, n, lo, hi);
          for (; c0 < ce; c0 += 16) tc_ns_minmax<16>(trow, c0, bias, n, lo, hi);
          row_lo[half][r] = lo;
          row_hi[half][r] = hi;
          asm volatile("bar.sync 1, %0;" ::"r"(kTcEpiThreads) : "memory");  // the epilogue warps only
          lo = fminf(row_lo[0][r], row_lo[1][r]);
          hi = fmaxf(row_hi[0][r], row_hi[1][r]);
        }
        float scale = hi - lo;
        if (scale < 1e-5f) scale += 1e-5f;
        const float inv = st.minmax ? 1.0f / scale : 1.0f;
        const float sub = st.minmax ? lo : 0.0f;
        const uint32_t out_sh = buf_sh(st.out_buf) + (uint32_t)r * 16u;
        float* dst = (live && a.next_emb != nullptr) ? a.next_emb + (size_t)rb * a.out_dim : nullptr;
        const bool dst_vec = (a.out_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(a.next_emb) & 15) == 0;
        {
          int c0 = cb;
          for (; c0 + 32 <= ce; c0 += 32) tc_ns_store<32>(trow, c0, bias, n, sub, inv, dst, dst_vec, out_sh);
          for (; c0 < ce; c0 += 16) tc_ns_store<16>(trow, c0, bias, n, sub, inv, dst, dst_vec, out_sh);
        }
        // (the bf16 tree row is copied out of the operand buffer by the store warps, which wait for this step's barrier)
      } else if (st.epi == kEpiPolicy) {
        for (int c0 = cb; c0 < ce; c0 += 16) {
          float v[16];
          tc_ld16(trow + (uint32_t)c0, v);
          tc_add_bias<16>(v, bias, c0);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (live && c0 + i < n) a.logits[(size_t)row * a.A + c0 + i] = v[i];
        }
      } else if (half == 0) {
        // support_to_scalar(softmax(logits)): a narrow row (2S + 1 logits), finished by one thread in registers
        float y;
        if (npad <= 16) y = tc_head_scalar<16>(trow, bias, n, a.S);
        else if (npad <= 32) y = tc_head_scalar<32>(trow, bias, n, a.S);
        else if (npad <= 48) y = tc_head_scalar<48>(trow, bias, n, a.S);
        else if (npad <= 64) y = tc_head_scalar<64>(trow, bias, n, a.S);
        else {  // wide supports: three passes over tensor memory
          float mx = -mz_inf();
          for (int c0 = 0; c0 < npad; c0 += 16) {
            float v[16];
            tc_ld16(trow + (uint32_t)c0, v);
            tc_add_bias<16>(v, bias, c0);
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, c0 + i < n ? v[i] : -mz_inf());
          }
          float sum = 0.0f, x = 0.0f;
          for (int c0 = 0; c0 < npad; c0 += 16) {
            float v[16];
            tc_ld16(trow + (uint32_t)c0, v);
            tc_add_bias<16>(v, bias, c0);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float e = c0 + i < n ? tc_ex2((v[i] - mx) * kTcLog2e) : 0.0f;
              sum += e;
              x += (float)(c0 + i - a.S) * e;
            }
          }
          y = mz_inv_scaling(x / sum);
        }
        if (live) (st.epi == kEpiReward ? a.reward : a.value)[row] = y;
      }
      tc_fence_before();       // this thread's TMEM loads are complete (tcgen05.wait::ld) and ordered before ...
      tc_fence_proxy_async();  // ... and its shared-memory stores visible to ... the MMAs that wait for this step
      tc_mbar_arrive(&bar_done[s]);
      if (tid == 0) MZ_TCCLK(4 * s + 5);
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef MZ_TC_CLOCKS
  if (blockIdx.x == 0 && tid == 0) {
    printf("tc clk | prologue %lld released %lld A %lld |", tc_clk[0], tc_clk[4 * kTcMaxSteps + 2], tc_clk[1]);
    for (int s = 0; s < a.n_steps; ++s)
      printf(" s%d[k16=%d n=%d] mma_go %lld mma_issued %lld acc %lld epi_done %lld |", s, a.steps[s].k16, a.steps[s].npad,
             tc_clk[4 * s + 2], tc_clk[4 * s + 3], tc_clk[4 * s + 4], tc_clk[4 * s + 5]);
    printf(" end %lld\n", clock64() - tc_t0);
  }
#endif
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

// W [K][N] fp32 row-major -> the bf16 operand image of the layer: chunks of kc k-values, inside a chunk
// [k / 8][npad rows][8 k-values], zero padded to kpad x npad.
__global__ void recurrent_tc_pack_kernel(const float* __restrict__ raw, __nv_bfloat16* __restrict__ img, int64_t w_off,
                                         int K, int N, int ldn, int kpad, int npad, int kc) {
  const int total = kpad * npad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx / npad, n = idx - k * npad;
    const float v = (k < K && n < N) ? raw[w_off + (int64_t)k * ldn + n] : 0.0f;
    const int c = k / kc, kk = k - c * kc;
    const size_t off = (size_t)c * npad * kc + ((size_t)(kk / 8) * npad + n) * 8 + (kk & 7);
    img[off] = __float2bfloat16_rn(v);
  }
}

// The bias rows of every step of every program, zero padded to npad, in one table.
__global__ void recurrent_tc_bias_kernel(const float* __restrict__ raw, float* __restrict__ table, int64_t b_off, int N,
                                         int npad, int dst) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < npad; c += gridDim.x * blockDim.x)
    table[dst + c] = c < N ? raw[b_off + c] : 0.0f;
}

// Root embeddings fp32 [B][E] -> node 0 of the bf16 tree rows (zero padded to es).
__global__ void recurrent_tc_root16_kernel(const float* __restrict__ root_emb, __nv_bfloat16* __restrict__ emb16, int B,
                                           int E, int es, int64_t tree_stride) {
  const long total = (long)B * es;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / es), k = (int)(i - (long)b * es);
    emb16[(size_t)b * tree_stride + k] = __float2bfloat16_rn(k < E ? root_emb[(size_t)b * E + k] : 0.0f);
  }
}

// bf16 tree rows -> the fp32 [B][N][E] embeddings of the mctx tree view.
__global__ void recurrent_tc_export16_kernel(const __nv_bfloat16* __restrict__ emb16, float* __restrict__ out, long rows,
                                             int E, int es) {
  const long total = rows * E;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / E;
    const int k = (int)(i - r * E);
    out[i] = __bfloat162float(emb16[r * es + k]);
  }
}

// ------------------------------------------------------------------------------------------ host side

struct TcLayer {  // one step's operand image: columns [n_off, n_off + N) of a layer of ldn output units (the offsets
  int64_t w_off, img_off, b_off;  // already include n_off)
  int K, N, ldn, kpad, npad, kc, bias_dst;  // bias_dst: float offset of the step's padded bias row in the bias table
};
struct TcProgram {
  TcArgs args{};
  size_t smem = 0, bias_base = 0;  // bias_base: float offset of the program's rows in the bias table
  bool ok = false;
};
struct TcImpl {
  TcProgram rec;       // recurrent_fn: reward head, next-state head, value head, policy head
  TcProgram root_obs;  // root from observations: Representation, value head, policy head
  TcProgram root_emb;  // root from a caller-made embedding: value head, policy head
  std::vector<TcLayer> layers;
  __nv_bfloat16* images = nullptr;
  float* bias_table = nullptr;
  size_t image_elems = 0, bias_elems = 0;
  __nv_bfloat16* emb16 = nullptr;  // [B][N][es] tree embeddings of the throughput mode
  size_t emb16_elems = 0;
  int es = 0;
};

struct TcHead {
  const mz_stack* s;
  int in, final_epi, minmax;
};

// One program = a chain of heads over a shared input operand.  Every head's first layer reads buffer A; the head
// that ends in kEpiNextState overwrites buffer A with its (normalised) output, which the heads after it read.
static bool tc_build(TcImpl* impl, TcProgram& prog, const std::vector<TcHead>& heads, int in_dim, bool onehot_actions,
                     int num_actions, size_t* img, size_t* bias_base, int max_smem, std::string* why) {
  TcArgs& a = prog.args;
  prog.bias_base = *bias_base;
  const size_t layers0 = impl->layers.size(), img0 = *img;
  auto fail = [&](const char* msg) {  // a program that cannot be built leaves no layers behind
    impl->layers.resize(layers0);
    *img = img0;
    *why = msg;
    return false;
  };
  int max_npad = 16, max_hidden_pad = 16, deepest = 1, n_steps = 0, bias_floats = 0;
  int bufA_k = round_up(in_dim + (onehot_actions ? num_actions : 0), 16);
  for (const TcHead& h : heads) {
    const mz_stack& s = *h.s;
    deepest = std::max(deepest, (int)s.n_layers);
    for (int l = 0; l < s.n_layers; ++l) {
      const int npad = round_up(s.out_dim[l], 16);
      if (npad > 256) return fail("a layer is wider than 256 units");
      const bool last = l == s.n_layers - 1;
      max_npad = std::max(max_npad, npad);
      if (!last) max_hidden_pad = std::max(max_hidden_pad, npad);
      if (last && h.final_epi == kEpiNextState) bufA_k = std::max(bufA_k, npad);
    }
  }
  // groups of heads that read the same contents of buffer A: the head that ends in kEpiNextState rewrites it, the
  // heads after it read the new row
  std::vector<std::vector<int>> groups(1);
  for (size_t h = 0; h < heads.size(); ++h) {
    groups.back().push_back((int)h);
    if (heads[h].final_epi == kEpiNextState && h + 1 < heads.size()) groups.emplace_back();
  }
  size_t widest_group = 0;
  for (const auto& g : groups) widest_group = std::max(widest_group, g.size());
  // default on (bit-identical results; 133 against 160 cycles per N = 256 MMA and 64 KB of shared memory back for the
  // weight ring); MZ_TC_TMEM_H=0 keeps the hidden activations in shared memory
  static const bool want_h_tmem = getenv("MZ_TC_TMEM_H") == nullptr || atoi(getenv("MZ_TC_TMEM_H")) != 0;
  static const bool want_pipe = getenv("MZ_TC_PIPE") == nullptr || atoi(getenv("MZ_TC_PIPE")) != 0;
  const int hcols = round_up(max_hidden_pad / 2, 16);  // 16 accumulator-aligned columns per 32 hidden units
  // Pipelined program (heads of <= 2 layers, <= 2 heads per group): the heads of a group are interleaved layer by layer
  // (head q keeps its hidden activations in tensor-memory buffer q), hidden layers wider than 128 are cut into two
  // column halves, and the steps alternate between two accumulator regions of `rw` columns.
  const int rw = max_npad > 128 ? 128 : round_up(max_npad, 32);
  const bool pipe = want_pipe && want_h_tmem && deepest <= 2 && widest_group <= 2 && 2 * rw + 2 * hcols <= 512;
  a.h_tmem = 0;
  int need_cols = max_npad;
  if (pipe) {
    a.h_tmem = 1;
    a.h_col[0] = 2 * rw;
    a.h_col[1] = 2 * rw + hcols;
    need_cols = 2 * rw + 2 * hcols;
  } else if (want_h_tmem && deepest >= 2) {
    // hidden activations in tensor memory, after the accumulators
    const int total = round_up(max_npad, 16) + hcols * (deepest >= 3 ? 2 : 1);
    if (total <= 512) {
      a.h_tmem = 1;
      a.h_col[0] = round_up(max_npad, 16);
      a.h_col[1] = a.h_col[0] + hcols;
      need_cols = total;
    }
  }
  // the step table, in issue order
  struct Emit { int head, layer; };
  std::vector<Emit> order;
  for (const auto& g : groups) {
    if (pipe) {
      for (int l = 0; l < deepest; ++l)
        for (int h : g)
          if (l < heads[h].s->n_layers) order.push_back(Emit{h, l});
    } else {
      for (int h : g)
        for (int l = 0; l < heads[h].s->n_layers; ++l) order.push_back(Emit{h, l});
    }
  }
  int last_writer[3] = {-1, -1, -1};  // latest step whose epilogue writes buffer A / hidden 0 / hidden 1
  int region_user[2] = {-1, -1};      // latest step whose accumulators live in the region (pipelined programs)
  int next_region = 0;
  for (const Emit& e : order) {
    const TcHead& h = heads[e.head];
    const mz_stack& s = *h.s;
    const int l = e.layer;
    int q = 0;  // position of the head inside its group
    for (const auto& g : groups)
      for (size_t i = 0; i < g.size(); ++i)
        if (g[i] == e.head) q = (int)i;
    const int K = l == 0 ? h.in : s.in_dim[l], N = s.out_dim[l];
    const int kpad = round_up(K, 16), npad = round_up(N, 16);
    const bool last = l == s.n_layers - 1;
    const bool cut = pipe && !last && npad > 128;
    const int first = cut ? round_up(npad / 2, 32) : npad;
    for (int part = 0; part < (cut ? 2 : 1); ++part) {
      if (n_steps >= kTcMaxSteps) return fail("too many layers");
      const int n_off = part == 0 ? 0 : first, np = part == 0 ? first : npad - first;
      TcStep& t = a.steps[n_steps];
      if (pipe) {
        t.a_buf = l == 0 ? kBufA : (q ? kBufH1 : kBufH0);
        t.out_buf = last ? kBufA : (q ? kBufH1 : kBufH0);
      } else {
        t.a_buf = l == 0 ? kBufA : ((l - 1) & 1 ? kBufH1 : kBufH0);
        t.out_buf = last ? kBufA : (l & 1 ? kBufH1 : kBufH0);
      }
      t.k16 = kpad / 16;
      t.kc = 0;  // set below, once the stage size is known
      t.n = std::max(0, std::min(N - n_off, np));
      t.npad = np;
      t.n_off = n_off;
      t.epi = last ? h.final_epi : kEpiHidden;
      t.minmax = last ? h.minmax : 0;
      t.bias_sh = bias_floats;
      bias_floats += np;
      t.b_off = s.b_off[l] + n_off;
      t.img_off = (int64_t)*img;
      // what the step's MMAs must wait for: the epilogues that write its A operand and the epilogues that read the
      // accumulator columns it overwrites.  Epilogues complete in table order, so the latest of them covers all.
      int dep = last_writer[t.a_buf];
      if (pipe) {
        if (np > rw) {  // both regions
          t.acc_col = 0;
          dep = std::max(dep, std::max(region_user[0], region_user[1]));
          region_user[0] = region_user[1] = n_steps;
          next_region = 0;
        } else {
          const int reg = cut ? part : next_region;
          t.acc_col = reg * rw;
          dep = std::max(dep, region_user[reg]);
          region_user[reg] = n_steps;
          next_region = reg ^ 1;
        }
      } else {
        t.acc_col = 0;
        dep = n_steps - 1;
      }
      t.dep = dep;
      if (t.epi == kEpiHidden || t.epi == kEpiNextState) last_writer[t.out_buf] = n_steps;
      impl->layers.push_back(TcLayer{s.w_off[l] + n_off, (int64_t)*img, s.b_off[l] + n_off, K, t.n, N, kpad, np, t.kc,
                                     (int)(*bias_base + t.bias_sh)});
      *img += (size_t)kpad * np;
      ++n_steps;
    }
  }
  a.n_steps = n_steps;
  a.ns_step = -1;
  for (int i = 0; i < n_steps; ++i)
    if (a.steps[i].epi == kEpiNextState) a.ns_step = i;
  a.in_dim = in_dim;
  a.kx16 = round_up(in_dim + (onehot_actions ? num_actions : 0), 16) / 16;
  a.bufA_bytes = bufA_k / 8 * (int)kTcChunkPitch;
  a.bufH_bytes = a.h_tmem ? 0 : max_hidden_pad / 8 * (int)kTcChunkPitch;
  a.bufH1_bytes = (!a.h_tmem && deepest >= 3) ? a.bufH_bytes : 0;
  a.bias_floats = bias_floats;
  int cols = 32;
  while (cols < need_cols) cols <<= 1;
  a.tmem_cols = cols;
  const size_t fixed = (size_t)a.bufA_bytes + a.bufH_bytes + a.bufH1_bytes + (size_t)bias_floats * 4 + 128;
  const size_t budget = (size_t)max_smem - 2048;
  // stage size: the largest of 64 / 32 / 16 KB of which two fit.  Every chunk costs the issuer a wait / fence / commit
  // round (~330 cycles measured), so few large chunks beat many small ones even with a shallower ring — C5, one
  // 128-column step of K = 288: 3.65 k cycles with 16 KB stages, 3.0 k with 32 KB, 2.65 k with 64 KB (2.45 / 2.29 /
  // 2.23 ms per act)
  a.stage_bytes = tc_stage_env();
  if (a.stage_bytes == 0) {
    a.stage_bytes = kTcStageBytes;
    for (int cand = 4 * kTcStageBytes; cand > kTcStageBytes; cand /= 2)
      if (fixed + 2 * (size_t)cand <= budget) {
        a.stage_bytes = cand;
        break;
      }
  }
  if (fixed + 2 * (size_t)a.stage_bytes > budget) return fail("operands do not fit shared memory");
  a.n_stages = (int)std::min<size_t>(kTcMaxStages, (budget - fixed) / a.stage_bytes);
  prog.smem = fixed + (size_t)a.n_stages * a.stage_bytes;
  for (int i = 0; i < n_steps; ++i) {
    TcStep& t = a.steps[i];
    t.kc = std::min(t.k16 * 16, std::max(16, a.stage_bytes / (t.npad * 2) / 16 * 16));
    impl->layers[layers0 + i].kc = t.kc;
  }
  *bias_base += bias_floats;
  prog.ok = true;
  if (getenv("MZ_TC_DUMP") != nullptr) {  // the step table, for reading next to an MZ_TC_CLOCKS timeline
    fprintf(stderr, "tc program: %d steps, %s, tmem %d cols (hidden at %d / %d), %d stages of %d KB\n", n_steps,
            pipe ? "pipelined" : "serial", a.tmem_cols, a.h_col[0], a.h_col[1], a.n_stages, a.stage_bytes / 1024);
    for (int i = 0; i < n_steps; ++i) {
      const TcStep& t = a.steps[i];
      fprintf(stderr, "  s%d: A=buf%d k16=%d kc=%d cols [%d, +%d/%d) acc@%d epi=%d out=buf%d dep=%d\n", i, t.a_buf, t.k16, t.kc,
              t.n_off, t.n, t.npad, t.acc_col, t.epi, t.out_buf, t.dep);
    }
  }
  return true;
}

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
  const int E = net.embed_dim, A = net.num_actions, F = 2 * net.support_size + 1;
  const int max_smem = (int)prop.sharedMemPerBlockOptin;
  size_t img = 0, bias_base = 0;
  bool ok = F == net.dyn_r.out_dim[net.dyn_r.n_layers - 1];
  if (!ok) st.why = "support size mismatch";
  // heads in execution order; the reward head runs before the next-state head overwrites their shared input
  if (ok)
    ok = tc_build(impl, impl->rec, {{&net.dyn_r, E + A, kEpiReward, 0}, {&net.dyn_ns, E + A, kEpiNextState, net.dyn_minmax},
                                    {&net.pred_v, E, kEpiValue, 0}, {&net.pred_pi, E, kEpiPolicy, 0}},
                  E, true, A, &img, &bias_base, max_smem, &st.why);
  if (ok) {
    std::string why;  // the root programs are optional: without them the root runs on the fp32 kernels
    tc_build(impl, impl->root_emb, {{&net.pred_v, E, kEpiValue, 0}, {&net.pred_pi, E, kEpiPolicy, 0}}, E, false, A, &img,
             &bias_base, max_smem, &why);
    if (net.obs_dim > 0 && net.repr.n_layers > 0)
      tc_build(impl, impl->root_obs, {{&net.repr, net.obs_dim, kEpiNextState, net.repr_minmax}, {&net.pred_v, E, kEpiValue, 0},
                                      {&net.pred_pi, E, kEpiPolicy, 0}},
               net.obs_dim, false, A, &img, &bias_base, max_smem, &why);
  }
  size_t smem_max = 0;
  for (TcProgram* p : {&impl->rec, &impl->root_obs, &impl->root_emb})
    if (p->ok) smem_max = std::max(smem_max, p->smem);
  if (ok && cudaFuncSetAttribute(recurrent_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess) {
    cudaGetLastError();
    ok = false;
    st.why = "cudaFuncSetAttribute failed";
  }
  if (ok && (cudaMalloc((void**)&impl->images, img * sizeof(__nv_bfloat16) + 16) != cudaSuccess ||
             cudaMalloc((void**)&impl->bias_table, (bias_base + 4) * sizeof(float)) != cudaSuccess)) {
    if (impl->images) cudaFree(impl->images);
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
  impl->bias_elems = bias_base;
  for (TcProgram* p : {&impl->rec, &impl->root_obs, &impl->root_emb}) {
    TcArgs& a = p->args;
    a.images = impl->images;
    a.bias_table = impl->bias_table + p->bias_base;
    a.A = A;
    a.S = net.support_size;
    a.act_kind = net.activation;
    a.out_dim = E;
  }
  st.impl = impl;
  st.available = true;
  return 0;
}

void recurrent_tc_destroy(RecurrentTcState& st) {
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  if (impl != nullptr) {
    if (impl->images) cudaFree(impl->images);
    if (impl->bias_table) cudaFree(impl->bias_table);
    if (impl->emb16) cudaFree(impl->emb16);
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
                                                                      L.K, L.N, L.ldn, L.kpad, L.npad, L.kc);
    recurrent_tc_bias_kernel<<<1, 256, 0, stream>>>(raw_weights_dev, impl->bias_table, L.b_off, L.N, L.npad, L.bias_dst);
    *launches += 2;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

static int tc_run(const TcProgram& prog, TcArgs a, int B, cudaStream_t stream, int64_t* launches, std::string* err) {
  a.B = B;
  void* args[] = {&a};
  // programmatic dependent launch: the kernel may start while the kernel before it in the stream is still running —
  // barrier setup, TMEM allocation, the bias table and the first weight stages do not depend on it — and executes
  // griddepcontrol.wait before it reads (parent, action) and the embeddings
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((B + kTcM - 1) / kTcM);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = prog.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)recurrent_tc_kernel, args);
  *launches += 1;
  if (e != cudaSuccess) {
    *err = std::string("recurrent_tc launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

int recurrent_tc_launch(RecurrentTcState& st, const Net& net, const Tree& t, const int32_t* parent,
                        const int32_t* action, float* reward, float* value, float* logits, float* next_emb,
                        cudaStream_t stream, int64_t* launches, std::string* err) {
  if (!st.available) {
    *err = "recurrent_tc: unavailable (" + st.why + ")";
    return 1;
  }
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  TcArgs a = impl->rec.args;
  a.in = t.embeddings;
  a.in_row_stride = (int64_t)t.N * net.embed_dim;
  a.parent = parent;
  a.action = action;
  a.reward = reward;
  a.value = value;
  a.logits = logits;
  a.next_emb = next_emb;
  return tc_run(impl->rec, a, t.B, stream, launches, err);
}

int recurrent_tc_tree_begin(RecurrentTcState& st, const Net& net, int B, int N, const float* root_emb, bool clear,
                            cudaStream_t stream, int64_t* launches, std::string* err) {
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  const int E = net.embed_dim, es = round_up(E, 8);
  const size_t need = (size_t)B * N * es;
  if (need > impl->emb16_elems) {
    if (impl->emb16) cudaFree(impl->emb16);
    impl->emb16 = nullptr;
    impl->emb16_elems = 0;
    if (cudaMalloc((void**)&impl->emb16, need * sizeof(__nv_bfloat16) + 16) != cudaSuccess) {
      cudaGetLastError();
      *err = "recurrent_tc: cudaMalloc(bf16 tree embeddings) failed";
      return 1;
    }
    impl->emb16_elems = need;
  }
  impl->es = es;
  if (clear && cudaMemsetAsync(impl->emb16, 0, need * sizeof(__nv_bfloat16), stream) != cudaSuccess) {
    *err = "recurrent_tc: cudaMemsetAsync(bf16 tree embeddings) failed";
    return 1;
  }
  const long total = (long)B * es;
  recurrent_tc_root16_kernel<<<(unsigned)std::min<long>((total + 255) / 256, 4096), 256, 0, stream>>>(
      root_emb, impl->emb16, B, E, es, (int64_t)N * es);
  *launches += 1;
  return 0;
}

int recurrent_tc_tree_launch(RecurrentTcState& st, const Net& net, int B, int N, const int32_t* parent,
                             const int32_t* action, const int32_t* next, float* reward, float* value, float* logits,
                             cudaStream_t stream, int64_t* launches, std::string* err) {
  (void)net;
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  TcArgs a = impl->rec.args;
  a.in = nullptr;
  a.in_row_stride = 0;
  a.in16 = impl->emb16;
  a.out16 = impl->emb16;
  a.next = next;
  a.tree16_stride = (int64_t)N * impl->es;
  a.es = impl->es;
  a.parent = parent;
  a.action = action;
  a.reward = reward;
  a.value = value;
  a.logits = logits;
  a.next_emb = nullptr;
  return tc_run(impl->rec, a, B, stream, launches, err);
}

int recurrent_tc_tree_export(RecurrentTcState& st, const Net& net, int B, int N, float* embeddings, cudaStream_t stream,
                             std::string* err) {
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  if (impl == nullptr || impl->emb16 == nullptr) {
    *err = "recurrent_tc: no bf16 tree to export";
    return 1;
  }
  const long rows = (long)B * N;
  recurrent_tc_export16_kernel<<<(unsigned)std::min<long>((rows * net.embed_dim + 255) / 256, 8192), 256, 0, stream>>>(
      impl->emb16, embeddings, rows, net.embed_dim, impl->es);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
    *err = "recurrent_tc: exporting the bf16 tree embeddings failed";
    return 1;
  }
  return 0;
}

bool recurrent_tc_has_root(const RecurrentTcState& st, bool from_obs) {
  if (!st.available) return false;
  const TcImpl* impl = static_cast<const TcImpl*>(st.impl);
  return from_obs ? impl->root_obs.ok : impl->root_emb.ok;
}

int recurrent_tc_root(RecurrentTcState& st, const Net& net, int B, const float* obs, const float* emb_in, float* value,
                      float* logits, float* emb_out, cudaStream_t stream, int64_t* launches, std::string* err) {
  TcImpl* impl = static_cast<TcImpl*>(st.impl);
  const bool from_obs = obs != nullptr;
  const TcProgram& prog = from_obs ? impl->root_obs : impl->root_emb;
  TcArgs a = prog.args;
  a.in = from_obs ? obs : emb_in;
  a.in_row_stride = from_obs ? net.obs_dim : net.embed_dim;
  a.parent = nullptr;
  a.action = nullptr;
  a.reward = nullptr;
  a.value = value;
  a.logits = logits;
  a.next_emb = emb_out;  // written by the Representation's epilogue (root from observations only)
  return tc_run(prog, a, B, stream, launches, err);
}

}  // namespace mz

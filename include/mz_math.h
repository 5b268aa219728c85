/*
 * mz_math.h — elementary float32 math shared by the sm_100a kernels and the CPU checkers.
 *
 * Why this exists: the search is a long chain of argmax decisions.  If expf/logf differ by one ulp
 * between two implementations, a near-tie flips and the two trees diverge.  Everything the search
 * path needs beyond IEEE-754 {+,-,*,/,sqrt,fma} is therefore defined HERE, once, using only those
 * correctly-rounded primitives, so that gcc on the host and nvcc on the device produce the same bits.
 *
 * Rules for users of this header:
 *   - host: compile with -ffp-contract=off (gcc's default would fuse a*b+c);
 *   - device: the MZ_* primitives map to __fmul_rn/__fadd_rn/... which ptxas never contracts;
 *   - never write a bare `a*b+c` on a value that feeds the tree; use MZ_MUL/MZ_ADD or MZ_FMA.
 *
 * The polynomials are the classic Cephes single-precision minimax fits (public domain);
 * accuracy is ~1 ulp, checked against float64 libm in tests/test_mz_math.py.
 *
 * Functions of the reference these serve: jax.nn.softmax / jnp.log / jnp.sqrt / jax.nn.elu as used by
 * muax/model.py:251-282, muax/nn.py:37-115, muax/utils.py:70-102 and the mctx selection formulas
 * restated in SURVEY.md Appendix A.5/A.6.
 */
#ifndef MZ_MATH_H_
#define MZ_MATH_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define MZ_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define MZ_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define MZ_MUL(a, b) __fmul_rn((a), (b))
#define MZ_ADD(a, b) __fadd_rn((a), (b))
#define MZ_SUB(a, b) __fsub_rn((a), (b))
#define MZ_DIV(a, b) __fdiv_rn((a), (b))
#define MZ_SQRT(a) __fsqrt_rn((a))
#define MZ_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define MZ_RINT(a) rintf((a))
#else
#define MZ_MUL(a, b) ((a) * (b))
#define MZ_ADD(a, b) ((a) + (b))
#define MZ_SUB(a, b) ((a) - (b))
#define MZ_DIV(a, b) ((a) / (b))
#define MZ_SQRT(a) sqrtf((a))
#define MZ_FMA(a, b, c) fmaf((a), (b), (c))
#define MZ_RINT(a) rintf((a))
#endif

#define MZ_F32_TINY 1.17549435e-38f /* finfo(float32).tiny */
#define MZ_F32_MAX 3.40282347e+38f  /* -MZ_F32_MAX == finfo(float32).min */

MZ_HD uint32_t mz_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

MZ_HD float mz_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

MZ_HD float mz_inf(void) { return mz_u2f(0x7f800000u); }
MZ_HD float mz_nan(void) { return mz_u2f(0x7fc00000u); }
MZ_HD float mz_fmax(float a, float b) { return a > b ? a : b; } /* callers never pass NaN */
MZ_HD float mz_fmin(float a, float b) { return a < b ? a : b; }
MZ_HD float mz_fabs(float a) { return mz_u2f(mz_f2u(a) & 0x7fffffffu); }

/* 2^k for k in [-126, 127]. */
MZ_HD float mz_pow2i(int k) { return mz_u2f((uint32_t)(k + 127) << 23); }

/* Cephes expf core on the reduced argument r in [-ln2/2, ln2/2]: returns (exp(r) - 1 - r) / r^2. */
MZ_HD float mz_exp_poly(float r) {
  float p = 1.9875691500e-4f;
  p = MZ_FMA(p, r, 1.3981999507e-3f);
  p = MZ_FMA(p, r, 8.3334519073e-3f);
  p = MZ_FMA(p, r, 4.1665795894e-2f);
  p = MZ_FMA(p, r, 1.6666665459e-1f);
  p = MZ_FMA(p, r, 5.0000001201e-1f);
  return p;
}

/* exp(x).  Written branch-free (selects only) so that a warp evaluating it on mixed arguments does not
 * serialise; the value for every input is the same as the textbook early-return formulation it replaced
 * (checked over all 2^32 inputs by tests/test_mz_math.py::test_branch_free_forms_are_exhaustively_identical). */
MZ_HD float mz_expf(float x) {
  const float xc = mz_fmin(mz_fmax(x, -104.0f), 89.0f); /* keeps n inside [-151, 129]; NaN handled below */
  const float nf = MZ_RINT(MZ_MUL(xc, 1.44269504088896341f));
  float r = MZ_FMA(nf, -0.693359375f, xc);   /* ln2 high part: 8 significant bits, nf*hi is exact */
  r = MZ_FMA(nf, 2.12194440e-4f, r);         /* minus the (negative) low part */
  const float p = mz_exp_poly(r);
  const float y = MZ_ADD(MZ_FMA(p, MZ_MUL(r, r), r), 1.0f);
  const int n = (int)nf;
  const int n1 = n / 2;
  const int n2 = n - n1;
  float v = MZ_MUL(MZ_MUL(y, mz_pow2i(n1)), mz_pow2i(n2));
  v = x > 88.72283905f ? mz_inf() : v;
  v = x < -103.972084f ? 0.0f : v;
  return x != x ? x : v;
}

/* expm1 for the ELU branch (x <= 0 in practice); accurate relative to x near 0.  Branch-free: for
 * |x| < ln2/2 the range reduction of mz_expf gives n = 0 and r = x, so both branches share r, p and
 * z = p*r^2 + r; the small branch returns z, the other one (z + 1) * 2^n - 1. */
MZ_HD float mz_expm1f(float x) {
  const float xc = mz_fmin(mz_fmax(x, -104.0f), 89.0f);
  const float nf = MZ_RINT(MZ_MUL(xc, 1.44269504088896341f));
  float r = MZ_FMA(nf, -0.693359375f, xc);
  r = MZ_FMA(nf, 2.12194440e-4f, r);
  const float p = mz_exp_poly(r);
  const float z = MZ_FMA(p, MZ_MUL(r, r), r);
  const int n = (int)nf;
  const int n1 = n / 2;
  const int n2 = n - n1;
  float e = MZ_MUL(MZ_MUL(MZ_ADD(z, 1.0f), mz_pow2i(n1)), mz_pow2i(n2));
  e = x > 88.72283905f ? mz_inf() : e;
  e = x < -103.972084f ? 0.0f : e;
  const float big = MZ_SUB(e, 1.0f);
  const float v = mz_fabs(x) < 0.34657359f ? z : big;
  return x != x ? x : v;
}

/* Early-return reference formulations (what the branch-free mz_expf / mz_expm1f must equal bit for bit);
 * compiled only into the CPU checker for the exhaustive equivalence test. */
#if !defined(__CUDACC__)
static inline float mz_expf_ref(float x) {
  if (x != x) return x;
  if (x > 88.72283905f) return mz_inf();
  if (x < -103.972084f) return 0.0f;
  float nf = MZ_RINT(MZ_MUL(x, 1.44269504088896341f));
  float r = MZ_FMA(nf, -0.693359375f, x);
  r = MZ_FMA(nf, 2.12194440e-4f, r);
  float p = mz_exp_poly(r);
  float y = MZ_ADD(MZ_FMA(p, MZ_MUL(r, r), r), 1.0f);
  int n = (int)nf;
  int n1 = n / 2;
  int n2 = n - n1;
  return MZ_MUL(MZ_MUL(y, mz_pow2i(n1)), mz_pow2i(n2));
}
static inline float mz_expm1f_ref(float x) {
  if (mz_fabs(x) < 0.34657359f) {
    float p = mz_exp_poly(x);
    return MZ_FMA(p, MZ_MUL(x, x), x);
  }
  return MZ_SUB(mz_expf_ref(x), 1.0f);
}
#endif

MZ_HD float mz_logf(float x) {
  if (x != x) return x;
  if (x < 0.0f) return mz_nan();
  if (x == 0.0f) return -mz_inf();
  uint32_t u = mz_f2u(x);
  if (u == 0x7f800000u) return x;
  int e = 0;
  if (u < 0x00800000u) { /* subnormal: scale up by 2^23 (exact) */
    x = MZ_MUL(x, 8388608.0f);
    u = mz_f2u(x);
    e = -23;
  }
  e += (int)(u >> 23) - 126;                              /* x = m * 2^e, m in [0.5, 1) */
  float m = mz_u2f((u & 0x007fffffu) | 0x3f000000u);
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = MZ_SUB(MZ_ADD(m, m), 1.0f);
  } else {
    m = MZ_SUB(m, 1.0f);
  }
  float z = MZ_MUL(m, m);
  float y = 7.0376836292e-2f;
  y = MZ_FMA(y, m, -1.1514610310e-1f);
  y = MZ_FMA(y, m, 1.1676998740e-1f);
  y = MZ_FMA(y, m, -1.2420140846e-1f);
  y = MZ_FMA(y, m, 1.4249322787e-1f);
  y = MZ_FMA(y, m, -1.6668057665e-1f);
  y = MZ_FMA(y, m, 2.0000714765e-1f);
  y = MZ_FMA(y, m, -2.4999993993e-1f);
  y = MZ_FMA(y, m, 3.3333331174e-1f);
  y = MZ_MUL(MZ_MUL(y, m), z);
  float fe = (float)e;
  y = MZ_FMA(fe, -2.12194440e-4f, y);
  y = MZ_FMA(z, -0.5f, y);
  float r = MZ_ADD(m, y);
  r = MZ_FMA(fe, 0.693359375f, r);
  return r;
}

/* jax.random.uniform bit trick: 23 random mantissa bits -> [0, 1). */
MZ_HD float mz_bits_to_unit(uint32_t bits) { return MZ_SUB(mz_u2f((bits >> 9) | 0x3f800000u), 1.0f); }

/* -log(-log(u)), u = max(tiny, unit(bits)) : jax.random.gumbel on one 32-bit draw. */
MZ_HD float mz_bits_to_gumbel(uint32_t bits) {
  float u = mz_fmax(MZ_F32_TINY, mz_bits_to_unit(bits));
  return -mz_logf(-mz_logf(u));
}

/* jax.nn.elu with alpha = 1 (muax/nn.py:78,82,98,102). */
MZ_HD float mz_elu(float x) { return x > 0.0f ? x : mz_expm1f(x); }

/* muax/utils.py:70-76 _inv_scaling, eps = 1e-3, op order exactly as written there. */
MZ_HD float mz_inv_scaling(float x) {
  const float eps = 1e-3f;
  const float four_eps = 0.004f; /* python evaluates 4 * eps in double, then casts */
  const float two_eps = 0.002f;
  float t = MZ_ADD(MZ_ADD(mz_fabs(x), 1.0f), eps);
  t = MZ_ADD(1.0f, MZ_MUL(four_eps, t));
  t = MZ_DIV(MZ_SUB(MZ_SQRT(t), 1.0f), two_eps);
  t = MZ_SUB(MZ_MUL(t, t), 1.0f);
  float s = x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f);
  return MZ_MUL(s, t);
}

#endif /* MZ_MATH_H_ */

// fp64 exp and complementary normal CDF for the shared-class kernel.
//
// CUDA's exp() / normcdf() materialise every polynomial coefficient with two
// uniform moves (UMOV) per DFMA: 38 % of the instructions of the first
// class-kernel version (profiles/r1_v1_class_kernel_details.csv) were such moves.
// Here the coefficients live in constant memory, so each Horner step is a single
// DFMA with a constant-bank operand.  Coefficients: scripts/gen_math_coeffs.py.
//
// Accuracy (tests/test_gpu_exact.py::test_fast_math_accuracy): fm_exp_neg relative
// error <= 4e-16 on [-708, 0]; fm_half_erfc relative error <= 5e-15 for y <= 5 and
// absolute error <= 1e-17 beyond.  The leaf functions they replace are
// utilities/useful_functions.py:417 (exp in gaussian_pdf) and :509 (ndtr).
#pragma once

__constant__ double FM_EXP[13] = {
    1.0,
    0.99999999999999999847,
    0.49999999999999999989,
    0.16666666666666702412,
    0.041666666666666692193,
    0.0083333333333095270794,
    0.001388888888887188822,
    0.00019841269909217101366,
    0.00002480158735011079562,
    2.7557224958392622686e-6,
    2.7557251904174912593e-7,
    2.5114869528657764941e-8,
    2.0921579964950300073e-9,
};
__constant__ double FM_ERFC[23] = {
    1.232995118625552509,
    -0.13962111684055700664,
    0.015379652102618402432,
    0.068097054254206726434,
    -0.10103906603618063989,
    0.093732835011056907418,
    -0.066330365815416512383,
    0.037167515383604771047,
    -0.016197734020451596114,
    0.005031971017234253862,
    -0.00075777352573824577731,
    -0.00019926100185442791047,
    0.0001506225713704314352,
    -0.00002438769126430604539,
    -0.000011224122281497543974,
    5.6894090899650726899e-6,
    2.968409172531307117e-7,
    -8.0191746843271911164e-7,
    7.2025175848360047549e-8,
    9.2907728730322144612e-8,
    -1.6738127974372159983e-8,
    -6.9840211001585419858e-9,
    1.6403183696752708229e-9,
};

#define FM_L2E 1.4426950408889634074
#define FM_LN2_HI 6.93147180369123816490e-01
#define FM_LN2_LO 1.90821492927058770002e-10
#define FM_MAGIC 6755399441055744.0   // 1.5 * 2^52: rounds to nearest integer in the low word

// e^x for x <= 0 (0 below -708: the subnormal range is never needed here)
__device__ __forceinline__ double fm_exp_neg(double x) {
  const double t = fma(x, FM_L2E, FM_MAGIC);
  const double nd = t - FM_MAGIC;
  double r = fma(nd, -FM_LN2_HI, x);
  r = fma(nd, -FM_LN2_LO, r);
  double p = FM_EXP[12];
#pragma unroll
  for (int q = 11; q >= 0; --q) p = fma(p, r, FM_EXP[q]);
  const int n = __double2loint(t);
  const double v = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return x < -708. ? 0. : v;
}

// 0.5 erfc(y) = 1 - Phi(sqrt(2) y), any y.
//   a = |y|, t = (a - 4)/(a + 4):  erfc(a) = e^{-a^2} P(t) / (1 + 2a)
// one reciprocal serves both t and 1/(1 + 2a).
__device__ __forceinline__ double fm_half_erfc(double y) {
  const double a = fabs(y);
  const double u = fma(2., a, 1.);            // 1 + 2a
  const double den = (a + 4.) * u;
  const double rc = 1. / den;
  const double t = ((a - 4.) * u) * rc;
  const double w = (a + 4.) * rc;             // 1 / (1 + 2a)
  double p = FM_ERFC[22];
#pragma unroll
  for (int q = 21; q >= 0; --q) p = fma(p, t, FM_ERFC[q]);
  const double e = fm_exp_neg(-a * a);
  const double r = (0.5 * e) * (p * w);
  return y < 0. ? 1. - r : r;
}

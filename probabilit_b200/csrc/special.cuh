// Device special functions behind the inverse CDFs of the graph kernel (graph.cu).
//
// The reference evaluates  getattr(scipy.stats, distr)(*args, **kwargs).ppf(q)
// (src/probabilit/modeling.py:795-812).  SciPy 1.18's sources for the special functions are not
// on the box (compiled xsf / cdflib / Boost), so these are restatements of the *published*
// algorithms SciPy uses, evaluated in the same operation order where that matters:
//   gamma.ppf   -> gammaincinv(a, q): Cephes/xsf `igami`/`igamci` = an initial guess followed by
//                  Halley steps on P(a, x) - p  (DiDonato & Morris 1986 starting values);
//                  P, Q: Cephes `igam`/`igamc` (power series DLMF 8.11.4, series DLMF 8.7.3,
//                  continued fraction DLMF 8.9.2) with the Boost Lanczos(13, g=6.0247) prefactor.
//   poisson.ppf -> min{k : pdtr(k, mu) >= q}   (scipy/stats/_discrete_distns.py:1015-1019)
//   binom.ppf   -> min{k : cdf(k; n, p) >= q}  (Boost quantile, _discrete_distns.py:100-101)
// Accuracy contract (tests/test_graph_gpu.py): poisson/binom are exact integers; gamma is
// reported as a ulp distribution against scipy (scipy itself is up to 18 ulp from the truth).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "ndtri.cuh"

namespace pbl {

constexpr double kMachEp = 1.11022302462515654042E-16;
constexpr double kMaxLog = 7.09782712893383996843E2;
constexpr double kInf = __builtin_huge_val();
#define PBL_NAN __longlong_as_double(0x7FF8000000000000LL)

// Lanczos sum (N = 13, g = 6.024680040776729583740234375), exp(g)-scaled: Boost lanczos13m53,
// the form SciPy's igam_fac uses.  Rational function evaluated in x (|x| <= 1) or 1/x.
__device__ __noinline__ double lanczos_sum_expg_scaled(double x) {
  const double num[13] = {0.006061842346248906525783753964555936883222,
                          0.5098416655656676188125178644804694509993,
                          19.51992788247617482847860966235652136208,
                          449.9445569063168119446858607650988409623,
                          6955.999602515376140356310115515198987526,
                          75999.29304014542649875303443598909137092,
                          601859.6171681098786670226533699352302507,
                          3481712.15498064590882071018964774556468,
                          14605578.08768506808414169982791359218571,
                          43338889.32467613834773723740590533316085,
                          86363131.28813859145546927288977868422342,
                          103794043.1163445451906271053616070238554,
                          56906521.91347156388090791033559122686859};
  const double den[13] = {1., 66., 1925., 32670., 357423., 2637558., 13339535., 45995730.,
                          105258076., 150917976., 120543840., 39916800., 0.};
  double na, da;
  if (fabs(x) > 1.0) {
    const double y = 1.0 / x;
    na = num[12];
    da = den[12];
#pragma unroll
    for (int i = 11; i >= 0; --i) {
      na = na * y + num[i];
      da = da * y + den[i];
    }
  } else {
    na = num[0];
    da = den[0];
#pragma unroll
    for (int i = 1; i < 13; ++i) {
      na = na * x + num[i];
      da = da * x + den[i];
    }
  }
  return na / da;
}

// x^a e^-x / Gamma(a)
__device__ __noinline__ double igam_fac(double a, double x) {
  const double g = 6.024680040776729583740234375;
  if (fabs(a - x) > 0.4 * fabs(a)) {
    const double ax = a * log(x) - x - lgamma(a);
    if (ax < -kMaxLog) return 0.0;
    return exp(ax);
  }
  const double fac = a + g - 0.5;
  double res = sqrt(fac / 2.718281828459045) / lanczos_sum_expg_scaled(a);
  if (a < 200.0 && x < 200.0) {
    res *= exp(a - x) * pow(x / fac, a);
  } else {
    const double num = x - a - g + 0.5;
    const double numfac = num / fac;
    res *= exp(a * (log1p(numfac) - numfac) + x * (0.5 - g) / fac);
  }
  return res;
}

// P(a, x) by the power series x^a e^-x / Gamma(a+1) * sum_n x^n / ((a+1)...(a+n))
__device__ __noinline__ double igam_series(double a, double x) {
  const double ax = igam_fac(a, x);
  if (ax == 0.0) return 0.0;
  double r = a, c = 1.0, ans = 1.0;
  for (int i = 0; i < 4000; ++i) {
    r += 1.0;
    c *= x / r;
    ans += c;
    if (c <= kMachEp * ans) break;
  }
  return ans * ax / a;
}

// Q(a, x) for small x without cancellation (DLMF 8.7.3)
__device__ __noinline__ double igamc_series(double a, double x) {
  double fac = 1.0, sum = 0.0;
  for (int n = 1; n < 2000; ++n) {
    fac *= -x / n;
    const double term = fac / (a + n);
    sum += term;
    if (fabs(term) <= kMachEp * fabs(sum)) break;
  }
  const double logx = log(x);
  const double term = -expm1(a * logx - lgamma(1.0 + a));
  return term - exp(a * logx - lgamma(a)) * sum;
}

// Q(a, x) by the continued fraction (DLMF 8.9.2)
__device__ __noinline__ double igamc_cf(double a, double x) {
  const double big = 4.503599627370496e15, biginv = 2.22044604925031308085e-16;
  const double ax = igam_fac(a, x);
  if (ax == 0.0) return 0.0;
  double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
  double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
  double ans = pkm1 / qkm1;
  for (int i = 0; i < 4000; ++i) {
    c += 1.0;
    y += 1.0;
    z += 2.0;
    const double yc = y * c;
    const double pk = pkm1 * z - pkm2 * yc;
    const double qk = qkm1 * z - qkm2 * yc;
    double t = 1.0;
    if (qk != 0.0) {
      const double r = pk / qk;
      t = fabs((ans - r) / r);
      ans = r;
    }
    pkm2 = pkm1;
    pkm1 = pk;
    qkm2 = qkm1;
    qkm1 = qk;
    if (fabs(pk) > big) {
      pkm2 *= biginv;
      pkm1 *= biginv;
      qkm2 *= biginv;
      qkm1 *= biginv;
    }
    if (t <= kMachEp) break;
  }
  return ans * ax;
}

__device__ double igamc(double a, double x);

// regularised lower incomplete gamma P(a, x)
__device__ __noinline__ double igam(double a, double x) {
  if (x < 0.0 || a < 0.0 || a != a || x != x) return PBL_NAN;
  if (a == 0.0) return x > 0.0 ? 1.0 : PBL_NAN;
  if (x == 0.0) return 0.0;
  if (isinf(a)) return isinf(x) ? PBL_NAN : 0.0;
  if (isinf(x)) return 1.0;
  if (x > 1.0 && x > a) return 1.0 - igamc(a, x);
  return igam_series(a, x);
}

// regularised upper incomplete gamma Q(a, x)
__device__ __noinline__ double igamc(double a, double x) {
  if (x < 0.0 || a < 0.0 || a != a || x != x) return PBL_NAN;
  if (a == 0.0) return x > 0.0 ? 0.0 : PBL_NAN;
  if (x == 0.0) return 1.0;
  if (isinf(a)) return isinf(x) ? PBL_NAN : 1.0;
  if (isinf(x)) return 0.0;
  if (x > 1.1) return (x < a) ? 1.0 - igam_series(a, x) : igamc_cf(a, x);
  if (x <= 0.5) return (-0.4 / log(x) < a) ? 1.0 - igam_series(a, x) : igamc_series(a, x);
  return (x * 1.1 < a) ? 1.0 - igam_series(a, x) : igamc_series(a, x);
}

// Starting value for the inverse of P(a, .) at p (q = 1 - p): Cephes/xsf `find_inverse_gamma` (a port of
// Boost's), i.e. DiDonato & Morris (1986), "Computation of the incomplete gamma function ratios and their
// inverse", eqs. 21-25 and 31-36, in the published operation order.  SciPy's result after its fixed
// three Halley steps depends on this value, so the branches and series are restated exactly.
__device__ __forceinline__ double didonato_eq25(double a, double y) {
  const double c1 = (a - 1.0) * log(y);
  const double c1_2 = c1 * c1, c1_3 = c1_2 * c1, c1_4 = c1_2 * c1_2, a_2 = a * a, a_3 = a_2 * a;
  const double c2 = (a - 1.0) * (1.0 + c1);
  const double c3 = (a - 1.0) * (-(c1_2 / 2.0) + (a - 2.0) * c1 + (3.0 * a - 5.0) / 2.0);
  const double c4 = (a - 1.0) * ((c1_3 / 3.0) - (3.0 * a - 5.0) * c1_2 / 2.0 + (a_2 - 6.0 * a + 7.0) * c1 +
                                 (11.0 * a_2 - 46.0 * a + 47.0) / 6.0);
  const double c5 = (a - 1.0) * (-(c1_4 / 4.0) + (11.0 * a - 17.0) * c1_3 / 6.0 + (-3.0 * a_2 + 13.0 * a - 13.0) * c1_2 +
                                 (2.0 * a_3 - 25.0 * a_2 + 72.0 * a - 61.0) * c1 / 2.0 +
                                 (25.0 * a_3 - 195.0 * a_2 + 477.0 * a - 379.0) / 12.0);
  const double y_2 = y * y, y_3 = y_2 * y, y_4 = y_2 * y_2;
  return y + c1 + (c2 / y) + (c3 / y_2) + (c4 / y_3) + (c5 / y_4);
}

__device__ __noinline__ double igami_start(double a, double p, double q) {
  const double euler = 0.5772156649015328606;
  if (a == 1.0) return q > 0.9 ? -log1p(-p) : -log(q);
  if (a < 1.0) {
    const double g = tgamma(a);
    const double b = q * g;
    if (b > 0.6 || (b >= 0.45 && a >= 0.3)) {  // eq. 21
      const double u = (b * q > 1e-8 && q > 1e-5) ? pow(p * g * a, 1.0 / a) : exp((-q / a) - euler);
      return u / (1.0 - (u / (a + 1.0)));
    }
    if (a < 0.3 && b >= 0.35) {  // eq. 22
      const double t = exp(-euler - b);
      const double u = t * exp(t);
      return t * exp(u);
    }
    const double y = -log(b);
    if (b > 0.15 || a >= 0.3) {  // eq. 23
      const double u = y - (1.0 - a) * log(y);
      return y - (1.0 - a) * log(u) - log(1.0 + (1.0 - a) / (1.0 + u));
    }
    if (b > 0.1) {  // eq. 24
      const double u = y - (1.0 - a) * log(y);
      return y - (1.0 - a) * log(u) -
             log((u * u + 2.0 * (3.0 - a) * u + (2.0 - a) * (3.0 - a)) / (u * u + (5.0 - a) * u + 2.0));
    }
    return didonato_eq25(a, y);
  }
  // a > 1: eq. 31 around the normal quantile of eq. 32 (a rational approximation, not ndtri)
  double s;
  {
    const double t = sqrt(-2.0 * log(p < 0.5 ? p : q));
    const double num = ((0.213623493715853 * t + 4.28342155967104) * t + 11.6616720288968) * t + 3.31125922108741;
    const double den = (((0.3611708101884203e-1 * t + 1.27364489782223) * t + 6.40691597760039) * t + 6.61053765625462) * t + 1.0;
    s = t - num / den;
    if (p < 0.5) s = -s;
  }
  const double s_2 = s * s, s_3 = s_2 * s, s_4 = s_2 * s_2, s_5 = s_4 * s, ra = sqrt(a);
  double w = a + s * ra + (s_2 - 1.0) / 3.0;
  w += (s_3 - 7.0 * s) / (36.0 * ra);
  w -= (3.0 * s_4 + 7.0 * s_2 - 16.0) / (810.0 * a);
  w += (9.0 * s_5 + 256.0 * s_3 - 433.0 * s) / (38880.0 * a * ra);
  if (a >= 500.0 && fabs(1.0 - w / a) < 1e-6) return w;
  if (p > 0.5) {
    if (w < 3.0 * a) return w;
    const double D = fmax(2.0, a * (a - 1.0));
    const double lb = log(q) + lgamma(a);
    if (lb < -D * 2.3) return didonato_eq25(a, -lb);
    const double u = -lb + (a - 1.0) * log(w) - log(1.0 + (1.0 - a) / (1.0 + w));  // eq. 33
    return -lb + (a - 1.0) * log(u) - log(1.0 + (1.0 - a) / (1.0 + u));
  }
  double z = w;
  const double ap1 = a + 1.0, ap2 = a + 2.0;
  if (w < 0.15 * ap1) {  // eq. 35
    const double v = log(p) + lgamma(ap1);
    z = exp((v + w) / a);
    double t = log1p(z / ap1 * (1.0 + z / ap2));
    z = exp((v + z - t) / a);
    t = log1p(z / ap1 * (1.0 + z / ap2));
    z = exp((v + z - t) / a);
    t = log1p(z / ap1 * (1.0 + z / ap2 * (1.0 + z / (a + 3.0))));
    z = exp((v + z - t) / a);
  }
  if (z <= 0.01 * ap1 || z > 0.7 * ap1) return z;
  // eq. 36
  double sum = 1.0, partial = z / (a + 1.0);
  sum += partial;
  for (int i = 2; i <= 100; ++i) {
    partial *= z / (a + i);
    sum += partial;
    if (partial < 1e-4) break;
  }
  const double ls = log(sum);
  const double v = log(p) + lgamma(ap1);
  z = exp((v + z - ls) / a);
  return z * (1.0 - (a * log(z) - z - v + ls) / (a - z));
}

// scipy.special.gammaincinv(a, p): x with P(a, x) = p.  Cephes/xsf igami / igamci: the starting value above
// followed by EXACTLY three Halley steps (on P for p <= 0.9, on Q = 1 - P beyond): what SciPy returns is the
// third iterate, converged or not, so iterating to a fixed point would agree with the true root but not
// with SciPy (gamma(a=0.5): 76 % within 4 ulp of SciPy with a fixed-point iteration, > 99 % this way).
__device__ __noinline__ double igami(double a, double p) {
  if (a != a || p != p) return PBL_NAN;
  if (a < 0.0 || p < 0.0 || p > 1.0) return PBL_NAN;
  if (p == 0.0) return 0.0;
  if (p == 1.0) return kInf;
  if (a == 0.0) return 0.0;
  const bool upper = p > 0.9;          // igami -> igamci(a, 1 - p)
  const double q = 1.0 - p;
  // (igamci hands q' = 1 - p back to igami when q' > 0.9, i.e. never for p > 0.9)
  double x = igami_start(a, upper ? 1.0 - q : p, q);
#pragma unroll 1
  for (int it = 0; it < 3; ++it) {
    const double fac = igam_fac(a, x);
    if (fac == 0.0) return x;
    const double f_fp = upper ? (igamc(a, x) - q) * x / (-fac) : (igam(a, x) - p) * x / fac;
    const double fpp_fp = -1.0 + (a - 1.0) / x;
    x = isinf(fpp_fp) ? x - f_fp : x - f_fp / (1.0 - 0.5 * f_fp * fpp_fp);  // Newton if the ratio overflows
  }
  return x;
}

// ---- regularised incomplete beta I_x(a, b) and its inverse (beta.ppf -> PERT,
//      reference src/probabilit/distributions.py:79-94; scipy: Boost ibeta_inv) ----
// Continued fraction of DLMF 8.17.22 evaluated with the modified Lentz algorithm.
__device__ __noinline__ double betacf(double a, double b, double x) {
  const double tiny = 1e-300;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 2000; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) <= 2.0 * kMachEp) break;
  }
  return h;
}
// returns I_x(a, b); *comp receives 1 - I_x(a, b) computed without cancellation
__device__ __noinline__ double ibeta(double a, double b, double x, double* comp) {
  if (x <= 0.0) { *comp = 1.0; return 0.0; }
  if (x >= 1.0) { *comp = 0.0; return 1.0; }
  const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
  const double front = exp(a * log(x) + b * log1p(-x) - lbeta);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    const double v = front * betacf(a, b, x) / a;
    *comp = 1.0 - v;
    return v;
  }
  const double v = front * betacf(b, a, 1.0 - x) / b;
  *comp = v;
  return 1.0 - v;
}
// scipy.special.betaincinv(a, b, p): safeguarded Newton / Halley on I_x(a, b) - p inside a bracket
// (on the complement for p > 1/2), starting from the Abramowitz-Stegun 26.5.22 approximation.
__device__ __noinline__ double ibeta_inv(double a, double b, double p) {
  if (a != a || b != b || p != p || !(a > 0.0) || !(b > 0.0) || p < 0.0 || p > 1.0) return PBL_NAN;
  if (p == 0.0) return 0.0;
  if (p == 1.0) return 1.0;
  double x;
  if (a >= 1.0 && b >= 1.0) {
    const double pp = p < 0.5 ? p : 1.0 - p;
    const double t = sqrt(-2.0 * log(pp));
    double y = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
    if (p < 0.5) y = -y;
    const double al = (y * y - 3.0) / 6.0;
    const double h = 2.0 / (1.0 / (2.0 * a - 1.0) + 1.0 / (2.0 * b - 1.0));
    const double w = y * sqrt(al + h) / h - (1.0 / (2.0 * b - 1.0) - 1.0 / (2.0 * a - 1.0)) * (al + 5.0 / 6.0 - 2.0 / (3.0 * h));
    x = a / (a + b * exp(2.0 * w));
  } else {
    const double lna = log(a / (a + b)), lnb = log(b / (a + b));
    const double t = exp(a * lna) / a, u = exp(b * lnb) / b;
    const double w = t + u;
    x = p < t / w ? pow(a * w * p, 1.0 / a) : 1.0 - pow(b * w * (1.0 - p), 1.0 / b);
  }
  if (!(x > 0.0)) x = 1e-300;
  if (!(x < 1.0)) x = 1.0 - kMachEp;
  const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
  const bool upper = p > 0.5;
  const double q = 1.0 - p;
  double lo = 0.0, hi = 1.0;
  for (int it = 0; it < 60; ++it) {
    double comp;
    const double v = ibeta(a, b, x, &comp);
    const double err = upper ? (q - comp) : (v - p);  // > 0: x too large
    if (err == 0.0) break;
    if (err > 0.0) hi = x; else lo = x;
    const double pdf = exp((a - 1.0) * log(x) + (b - 1.0) * log1p(-x) - lbeta);
    double xn = x;
    if (pdf > 0.0 && !isinf(pdf)) {
      const double u = err / pdf;
      const double g = (a - 1.0) / x - (b - 1.0) / (1.0 - x);  // pdf' / pdf
      double den = 1.0 - 0.5 * u * g;
      if (!(den > 0.5)) den = 0.5;  // keep the Halley correction bounded
      xn = x - u / den;
    }
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);  // left the bracket: bisect
    const double dx = fabs(xn - x);
    x = xn;
    if (dx <= 2.0 * kMachEp * x) break;
  }
  return x;
}

// ---- truncated normal (truncnorm.ppf, reference distributions.py:17-29) ----
// Phi and 1 - Phi through erfc (accurate in relative terms in both tails); the quantile is taken on
// the side where the truncated mass is computed without cancellation.
__device__ __noinline__ double truncnorm_ppf_core(double q, double a, double b) {
  const double rs2 = 0.70710678118654752440;
  if (a < 0.0) {  // work with the lower tails
    const double Fa = 0.5 * erfc(-a * rs2), Fb = 0.5 * erfc(-b * rs2);
    return ndtri(Fa + q * (Fb - Fa));
  }
  const double Sa = 0.5 * erfc(a * rs2), Sb = 0.5 * erfc(b * rs2);  // upper tails
  return -ndtri(Sb + (1.0 - q) * (Sa - Sb));
}

// log of the Poisson pmf at integer k
__device__ __forceinline__ double poisson_logpmf(double k, double mu) {
  return k * log(mu) - mu - lgamma(k + 1.0);
}

// poisson(mu)._ppf(q) for 0 < q < 1, mu >= 0:  min{k >= 0 : pdtr(k, mu) >= q}
__device__ __noinline__ double poisson_ppf_core(double q, double mu) {
  if (mu == 0.0) return 0.0;
  if (mu <= 32.0) {
    double pk = exp(-mu), F = pk, k = 0.0;
    const double kmax = 400.0;
    while (F < q && k < kmax) {
      k += 1.0;
      pk *= mu / k;
      F += pk;
    }
    return k;
  }
  // start at the Cornish-Fisher guess, get the CDF there from Q(k+1, mu), then walk
  const double z = ndtri(q), sd = sqrt(mu);
  double k = floor(mu + z * sd + (z * z - 1.0) / 6.0);
  if (k < 0.0) k = 0.0;
  double F = igamc(k + 1.0, mu);
  double pk = exp(poisson_logpmf(k, mu));
  if (F >= q) {
    while (k > 0.0) {
      const double Fm = F - pk;  // cdf(k - 1)
      if (!(Fm >= q)) break;
      F = Fm;
      pk *= k / mu;
      k -= 1.0;
    }
    return k;
  }
  const double kmax = mu + 40.0 * sd + 40.0;
  while (F < q && k < kmax) {
    k += 1.0;
    pk *= mu / k;
    F += pk;
  }
  return k;
}

// ---- binomial pmf after C. Loader, "Fast and accurate computation of binomial probabilities"
// (2000): saddle-point form with the Stirling error and the deviance term, accurate to a few ulp
// for any n (a plain lgamma difference loses log10(n log n) digits).
__device__ __forceinline__ double stirlerr(double n) {
  // log(n!) - log(sqrt(2 pi n) (n/e)^n)
  if (n <= 15.0) return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - 0.918938533204672741780329736406;
  const double nn = n * n;
  if (n > 500.0) return (1.0 / 12.0 - (1.0 / 360.0) / nn) / n;
  if (n > 80.0) return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0) / nn) / nn) / n;
  if (n > 35.0) return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0) / nn) / nn) / nn) / n;
  return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0) / nn) / nn) / nn) / nn) / n;
}
__device__ __forceinline__ double bd0(double x, double np) {
  // x log(x/np) + np - x, without cancellation when x ~ np
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np);
    double s = (x - np) * v;
    double ej = 2.0 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; ++j) {
      ej *= v;
      const double s1 = s + ej / (double)(2 * j + 1);
      if (s1 == s) return s1;
      s = s1;
    }
    return s;
  }
  return x * log(x / np) + np - x;
}
__device__ __noinline__ double binom_pmf(double x, double n, double p) {
  const double q = 1.0 - p;
  if (x == 0.0) return exp(n * log1p(-p));
  if (x == n) return exp(n * log(p));
  const double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
  return exp(lc) * sqrt(n / (6.283185307179586476925286766559 * x * (n - x)));
}

// binom(n, p)._ppf(q) for 0 < q < 1, integer n >= 0, 0 <= p <= 1:  min{k : cdf(k) >= q}.
// Two-sided walk from the mode: towards the relevant tail until the pmf no longer matters next to
// the target mass, then back, accumulating the tail mass from its small end (no cancellation in
// either tail; the upper tail works on 1 - q, which is exact for q >= 0.5).
__device__ __noinline__ double binom_ppf_core(double q, double n, double p) {
  if (n == 0.0 || p == 0.0) return 0.0;
  if (p == 1.0) return n;
  const double odds = p / (1.0 - p);
  double k = floor((n + 1.0) * p);
  if (k > n) k = n;
  double pk = binom_pmf(k, n, p);
  if (q <= 0.5) {
    const double stop = fmax(q * 1e-20, 1e-300);
    while (k > 0.0 && pk > stop) {  // pmf(k-1) = pmf(k) * k / (n-k+1) / odds
      pk *= k / (n - k + 1.0) / odds;
      k -= 1.0;
    }
    double F = pk;
    while (F < q && k < n) {
      pk *= (n - k) / (k + 1.0) * odds;
      k += 1.0;
      F += pk;
    }
    return k;
  }
  const double s = 1.0 - q;  // want min k with P(X > k) <= s
  const double stop = fmax(s * 1e-20, 1e-300);
  while (k < n && pk > stop) {
    pk *= (n - k) / (k + 1.0) * odds;
    k += 1.0;
  }
  double S = 0.0;  // P(X > k), the mass above the walk's end is negligible next to s
  while (k > 0.0) {
    const double Sn = S + pk;  // P(X > k-1)
    if (!(Sn <= s)) break;
    S = Sn;
    pk *= k / (n - k + 1.0) / odds;
    k -= 1.0;
  }
  return k;
}

}  // namespace pbl
